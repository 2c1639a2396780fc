"""CPU-side tests: the C-ABI library loads and exports every symbol include/bcp_b200.h declares, the drop-in
modules keep the reference's state_dict keys / parameter order, host logic (box drawing, arenas, packs)."""
import os

import numpy as np
import pytest
import torch

from oracle import bcp_oracle as O


def test_library_exports_every_declared_symbol():
    from bcp_b200._native import LIB, parse_header
    protos = parse_header()
    assert len(protos) >= 40
    lib = LIB.load()
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.bcp_abi_version() == 1
    assert isinstance(lib.bcp_last_error(), bytes)


def test_argument_errors_do_not_abort():
    from bcp_b200._native import LIB
    lib = LIB.load()
    rc = lib.bcp_mask_mix(None, None, None, 1, 1, 4, 4, 4, None, None)
    assert rc == -1 and b"null" in lib.bcp_last_error()
    assert lib.bcp_mix_loss_ctx_floats(2, 2) == 6 + 2 * 2 * 2 * 3
    assert lib.bcp_norm_chunks(2, 16, 10) == 1


def test_product_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bcp_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.mask_mix(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4), (0, 0, 0, 1, 1, 1))
    from bcp_b200.networks.VNet import VNet
    net = VNet(1, 2, 16, "batchnorm", False)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 16, 16, 16))


def test_state_dict_keys_match_reference_layout():
    from bcp_b200.networks.VNet import VNet
    from bcp_b200.networks.unet import UNet_2d, UNet
    from bcp_b200.pancreas.Vnet import VNet as PanVNet
    pairs = [(VNet(1, 2, 16, "batchnorm", True), O.OracleVNet(1, 2, 16, "batchnorm", True), 259),
             (UNet_2d(1, 4), O.OracleUNet2d(1, 4), 226), (UNet(1, 4), O.OracleUNet2d(1, 4, True), 226),
             (PanVNet(), O.OraclePanVNet(), 60)]
    for a, b, nkeys in pairs:
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa.keys()) == list(sb.keys()) and len(sa) == nkeys
        assert [tuple(v.shape) for v in sa.values()] == [tuple(v.shape) for v in sb.values()]
        assert [v.dtype for v in sa.values()] == [v.dtype for v in sb.values()]
        assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]


def test_shipped_checkpoints_load():
    root = "/root/reference/models"
    if not os.path.isdir(root):
        pytest.skip("reference checkpoints not present on this box")
    from bcp_b200.networks.VNet import VNet
    from bcp_b200.networks.unet import UNet_2d
    for f in ("LA/LA_5.pth", "LA/LA_10.pth"):
        VNet(1, 2, 16, "batchnorm", True).load_state_dict(torch.load(os.path.join(root, f), map_location="cpu"))
    for f in ("ACDC/ACDC_5.pth", "ACDC/ACDC_10.pth"):
        UNet_2d(1, 4).load_state_dict(torch.load(os.path.join(root, f), map_location="cpu"))


def test_flat_arena_and_packs():
    from bcp_b200.networks.VNet import VNet
    net = VNet(1, 2, 16, "batchnorm", True)
    ref = {k: v.clone() for k, v in net.state_dict().items()}
    rt = net.runtime
    rt.flatten_()
    assert rt.is_flat()
    assert rt.n_train == 9448866 and rt.n_param == 9457318            # SURVEY.md section 6 parameter counts
    for k, v in net.state_dict().items():
        assert torch.equal(v, ref[k])
    # load_state_dict keeps the views; a parameter that was re-allocated is detected
    net.load_state_dict(ref)
    assert rt.is_flat()
    net.encoder.block_one.conv[0].weight.data = net.encoder.block_one.conv[0].weight.data.clone()
    assert not rt.is_flat()
    rt.flatten_()
    g = rt.ensure_grad_arena()
    assert g.numel() == rt.n_train
    assert net.decoder.out_conv.weight.grad.data_ptr() >= g.data_ptr()
    assert rt.njobs == 64       # 20 'same' convs x (fwd, dgrad) + 8 stride-2 convs x 3 packs; block_one/out_conv read fp32
    # the EMA teacher never runs a 3x3x3 data gradient: no kind-1 (flipped / transposed) packs, everything else unchanged
    from bcp_b200.networks.runtime import _JOB
    ema = VNet(1, 2, 16, "batchnorm", True)
    for p in ema.parameters():
        p.detach_()
    ema.runtime.flatten_()
    assert ema.runtime.njobs == 64 - 20
    kinds = np.frombuffer(ema.runtime.jobs.cpu().numpy().tobytes(), dtype=_JOB)["kind"]
    assert 1 not in set(kinds.tolist()) and {0, 2, 3} <= set(kinds.tolist())


def test_launch_count_rules():
    """gpu_launches in bench.py is counted from C-ABI calls: the per-call kernel counts follow the dispatch in norm.cu"""
    from bcp_b200._native import KERNELS_PER_CALL, _norm_bwd_kernels
    base = [0] * 8 + [1, 1] + [0, 0, 0]
    assert _norm_bwd_kernels(base + [4, 256, 245, 2, 0.0, 1, 0, 0]) == 1          # tiny layer: reduce+apply in one launch
    assert _norm_bwd_kernels(base + [4, 16, 1003520, 2, 0.0, 1, 0, 0]) == 2        # reduce pass + apply pass
    assert _norm_bwd_kernels([0] * 8 + [None, None] + [0, 0, 0] + [4, 16, 1003520, 2, 0.0, 0, 0, 0]) == 1   # no statistics gradient
    assert KERNELS_PER_CALL["bcp_largest_cc"] == 5
    import ctypes
    wg = KERNELS_PER_CALL["bcp_conv_tc_wgrad"]          # finalises in-kernel: one launch per z-window (<= 254 columns: one)
    args = lambda z: [0] * 8 + [(ctypes.c_int * 3)(1, 256, z)]
    assert wg(args(80)) == 1 and wg(args(254)) == 1 and wg(args(256)) == 2 and wg(args(260)) == 3


def test_box_draw_order_matches_reference():
    from bcp_b200.utils.BCP_utils import context_box
    from bcp_b200.step import _acdc_box, _pan_box
    np.random.seed(1337)
    box = context_box((2, 1, 112, 112, 80), 2 / 3)
    _, _, obox = O.context_mask_la(torch.zeros(2, 1, 112, 112, 80), 2 / 3, np.random.RandomState(1337))
    assert box == obox
    np.random.seed(7)
    b2 = _acdc_box((6, 1, 256, 256))
    assert b2 == O.generate_mask_acdc(torch.zeros(6, 1, 256, 256), np.random.RandomState(7))[2]
    np.random.seed(9)
    assert _pan_box(64) == O.generate_mask_pan(torch.zeros(2, 1, 96, 96, 96), 64, np.random.RandomState(9))[2]


def test_z_window_weight_gradient_decomposition():
    """The identity the z-windowed tcgen05 weight gradient (csrc/conv_tc.cu conv_tc_wgrad_impl) rests on: with the output
    gradient cut into windows of <= 128 columns and the layer input read with its REAL neighbour columns as halo (zero only
    outside the tensor), the per-window weight gradients sum to the full one.  Window width = the C planner's rule."""
    import torch.nn.functional as F

    def tile_width(Z):                                   # wg_ztiles / wg_tile_width
        nt = 1 if Z + 2 <= 256 else (Z + 127) // 128
        return Z if nt == 1 else ((Z + nt - 1) // nt + 15) // 16 * 16, nt

    assert tile_width(80) == (80, 1) and tile_width(254) == (254, 1) and tile_width(256) == (128, 2) and tile_width(260) == (96, 3)
    torch.manual_seed(4)
    for dims, kernel in (((1, 6, 260), (1, 3, 3)), ((3, 4, 272), (3, 3, 3))):
        cin, cout, Z = 3, 2, dims[2]
        pad = tuple(k // 2 for k in kernel)
        a = torch.randn(2, cin, *dims, dtype=torch.float64)
        g = torch.randn(2, cout, *dims, dtype=torch.float64)
        w = torch.zeros(cout, cin, *kernel, dtype=torch.float64, requires_grad=True)
        F.conv3d(a, w, None, padding=pad).backward(g)
        tw, nt = tile_width(Z)
        total = torch.zeros_like(w)
        for t in range(nt):
            z0, z1 = t * tw, min(Z, (t + 1) * tw)
            lo, hi = max(z0 - 1, 0), min(z1 + 1, Z)
            aw = F.pad(a[..., lo:hi], (1 if z0 == 0 else 0, 1 if z1 == Z else 0))       # zero halo only at the tensor's edges
            ww = torch.zeros_like(w).requires_grad_(True)
            F.conv3d(aw, ww, None, padding=(pad[0], pad[1], 0)).backward(g[..., z0:z1].contiguous())
            total += ww.grad
        assert torch.allclose(total, w.grad, rtol=1e-12, atol=1e-12)
