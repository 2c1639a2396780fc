"""Every convolution shape the three entry points actually launch (SURVEY.md section 8d layer table), checked through
the C ABI against fp32 torch on IDENTICAL bf16-rounded operands: forward, data gradient and weight gradient.

  tolerance fwd   6e-3 relative RMS  (= the bf16 rounding of the stored output, 2^-9 ~ 2e-3, with margin)
            dgrad 1e-2               (same, the upstream gradient is bf16 as well)
            wgrad 1e-4 (3e-4 when one output element sums more than 2^21 products: fp32 accumulation order; torch's own
                        fp32 cuDNN result moves by the same amount against float64)

The second half replays the layers of a REAL train-mode forward/backward of the fp32 oracle V-Net (activations and
upstream gradients captured with hooks), so every kernel also sees realistic statistics (post-ReLU sparsity, the
common-mode offsets BatchNorm removes) -- kernel error in isolation, without BatchNorm's amplification of bf16 noise.
"""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import bcp_oracle as O
from tests.test_gpu_primitives import _packs
from tests.util import cb8_from_planar, planar_from_cb8, rel_rms, record, wgrad_fp64

pytestmark = pytest.mark.gpu

FWD_TOL, DGRAD_TOL = 6e-3, 1e-2


BIG = 1 << 18      # voxels per weight-gradient sum above which the reference is float64 (see check_wgrad)


def wgrad_tol(nvox):
    return 1e-4 if nvox <= (1 << 21) else 3e-4


def check_wgrad(tag, got, torch_fp32, a, dy, kernel, nvox):
    """Weight-gradient criterion.  Up to 2^18 voxels per sum: <= 1e-4 against fp32 torch.  Beyond that fp32 torch (cuDNN)
    itself is > 1e-4 away from the exact result (summation order over millions of products), so the reference becomes
    float64 and the requirement is: within 1e-4 of fp64, or at least no further from fp64 than 1.5x stock fp32 torch."""
    if nvox <= BIG or a is None:
        e = rel_rms(got, torch_fp32)
        record(tag, e)
        assert e <= wgrad_tol(nvox), (tag, e)
        return e
    ref = wgrad_fp64(a, dy, kernel, tuple(k // 2 for k in kernel))
    e, e_t = rel_rms(got, ref), rel_rms(torch_fp32, ref)
    record(tag + "_vs_fp64", e)
    record(tag + "_torch_fp32_vs_fp64", e_t)
    assert e <= max(1e-4, 1.5 * e_t), (tag, e, e_t)
    return e


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from bcp_b200 import ops as _ops
    return _ops


# (n, cin, cout, dims, kernel): LA V-Net (networks/VNet.py:151-209) at 112x112x80 batch 4, Pancreas at 96^3 batch 4,
# ACDC U-Net (networks/unet.py) at 256x256 batch 12 (1x3x3 and 1x1x1 layers)
SAME = [
    (4, 16, 16, (112, 112, 80), (3, 3, 3)), (4, 32, 32, (56, 56, 40), (3, 3, 3)), (4, 64, 64, (28, 28, 20), (3, 3, 3)),
    (4, 128, 128, (14, 14, 10), (3, 3, 3)), (4, 256, 256, (7, 7, 5), (3, 3, 3)),
    (4, 16, 16, (96, 96, 96), (3, 3, 3)), (4, 32, 32, (48, 48, 48), (3, 3, 3)), (4, 256, 256, (6, 6, 6), (3, 3, 3)),
    (12, 16, 16, (1, 256, 256), (1, 3, 3)), (12, 16, 32, (1, 128, 128), (1, 3, 3)), (12, 32, 32, (1, 128, 128), (1, 3, 3)),
    (12, 32, 64, (1, 64, 64), (1, 3, 3)), (12, 64, 64, (1, 64, 64), (1, 3, 3)), (12, 64, 128, (1, 32, 32), (1, 3, 3)),
    (12, 128, 128, (1, 32, 32), (1, 3, 3)), (12, 128, 256, (1, 16, 16), (1, 3, 3)), (12, 256, 256, (1, 16, 16), (1, 3, 3)),
    (12, 256, 128, (1, 32, 32), (1, 3, 3)), (12, 128, 64, (1, 64, 64), (1, 3, 3)), (12, 64, 32, (1, 128, 128), (1, 3, 3)),
    (12, 32, 16, (1, 256, 256), (1, 3, 3)),
    (12, 256, 128, (1, 16, 16), (1, 1, 1)), (12, 128, 64, (1, 32, 32), (1, 1, 1)), (12, 64, 32, (1, 64, 64), (1, 1, 1)),
    (12, 32, 16, (1, 128, 128), (1, 1, 1)),
]
# (n, c_full, c_half, half_dims): every stride-2 / transposed layer of the V-Nets
S2 = [(4, 16, 32, (56, 56, 40)), (4, 32, 64, (28, 28, 20)), (4, 64, 128, (14, 14, 10)), (4, 128, 256, (7, 7, 5)),
      (4, 16, 32, (48, 48, 48)), (4, 128, 256, (6, 6, 6))]


def _check(tag, got, ref, tol):
    e = rel_rms(got, ref)
    record(tag, e)
    assert e <= tol, (tag, e, tol)


@pytest.mark.parametrize("n,cin,cout,dims,kernel", SAME)
def test_conv_same_production_shape(ops, dev, n, cin, cout, dims, kernel):
    torch.manual_seed(cin * 31 + cout + dims[1])
    x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, *kernel, device=dev) / np.sqrt(cin * np.prod(kernel))).to(torch.bfloat16).float().requires_grad_(True)
    b = (0.1 * torch.randn(cout, device=dev)).requires_grad_(True)
    pack = _packs(ops, dev, w, (0, 1))
    xcb = cb8_from_planar(x).requires_grad_(True)
    y = ops.ConvSame.apply(xcb, w, b, pack, kernel)
    xr = x.clone().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br, padding=tuple(k // 2 for k in kernel))
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    tag = "conv_c%d_%d_%s_k%d" % (cin, cout, "x".join(map(str, dims)), kernel[0] * kernel[1])
    _check(tag + "_fwd", planar_from_cb8(y.detach(), cout), yr.detach(), FWD_TOL)
    _check(tag + "_dgrad", planar_from_cb8(xcb.grad, cin), xr.grad, DGRAD_TOL)
    check_wgrad(tag + "_wgrad", w.grad, wr.grad, x, g, kernel, n * int(np.prod(dims)))
    _check(tag + "_bgrad", b.grad, br.grad, 1e-3)


@pytest.mark.parametrize("n,c_full,c_half,half", S2)
def test_conv_stride2_production_shape(ops, dev, n, c_full, c_half, half):
    torch.manual_seed(c_full + c_half)
    full = tuple(2 * h for h in half)
    nvox = n * int(np.prod(half))
    # down: nn.Conv3d(c_full, c_half, 2, stride=2)
    x = torch.randn(n, c_full, *full, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(c_half, c_full, 2, 2, 2, device=dev) / np.sqrt(8 * c_full)).to(torch.bfloat16).float().requires_grad_(True)
    b = (0.1 * torch.randn(c_half, device=dev)).requires_grad_(True)
    xcb = cb8_from_planar(x).requires_grad_(True)
    y = ops.ConvDown2.apply(xcb, w, b, _packs(ops, dev, w, (0, 2, 3)))
    xr, wr, br = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br, stride=2)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    tag = "s2_c%d_%d_%s" % (c_full, c_half, "x".join(map(str, half)))
    _check(tag + "_down_fwd", planar_from_cb8(y.detach(), c_half), yr.detach(), FWD_TOL)
    _check(tag + "_down_dgrad", planar_from_cb8(xcb.grad, c_full), xr.grad, DGRAD_TOL)
    _check(tag + "_down_wgrad", w.grad, wr.grad, wgrad_tol(nvox))
    # up: nn.ConvTranspose3d(c_half, c_full, 2, stride=2)
    x = torch.randn(n, c_half, *half, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(c_half, c_full, 2, 2, 2, device=dev) / np.sqrt(c_half)).to(torch.bfloat16).float().requires_grad_(True)
    b = (0.1 * torch.randn(c_full, device=dev)).requires_grad_(True)
    xcb = cb8_from_planar(x).requires_grad_(True)
    y = ops.ConvUp2.apply(xcb, w, b, _packs(ops, dev, w, (0, 2, 3)))
    xr, wr, br = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv_transpose3d(xr, wr, br, stride=2)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    _check(tag + "_up_fwd", planar_from_cb8(y.detach(), c_full), yr.detach(), FWD_TOL)
    _check(tag + "_up_dgrad", planar_from_cb8(xcb.grad, c_half), xr.grad, DGRAD_TOL)
    _check(tag + "_up_wgrad", w.grad, wr.grad, wgrad_tol(nvox))


def test_first_layer_and_head_production_shape(ops, dev):
    torch.manual_seed(5)
    n, dims = 4, (112, 112, 80)
    x = torch.randn(n, 1, *dims, device=dev)
    w = (torch.randn(16, 1, 3, 3, 3, device=dev) / 5).requires_grad_(True)
    b = (0.1 * torch.randn(16, device=dev)).requires_grad_(True)
    y = ops.ConvFirst.apply(x, w, b)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv3d(x, wr, br, padding=1)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    _check("first_fwd", planar_from_cb8(y.detach(), 16), yr.detach(), FWD_TOL)
    check_wgrad("first_wgrad", w.grad, wr.grad, x, g, (3, 3, 3), n * int(np.prod(dims)))
    a = torch.randn(n, 16, *dims, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(2, 16, 1, 1, 1, device=dev) / 4).requires_grad_(True)
    b = (0.1 * torch.randn(2, device=dev)).requires_grad_(True)
    acb = cb8_from_planar(a).requires_grad_(True)
    lo = ops.Head.apply(acb, w, b, False)
    ar, wr, br = a.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    lr = F.conv3d(ar, wr, br)
    g = torch.randn_like(lr)
    lr.backward(g)
    lo.backward(g)
    _check("head_fwd", lo.detach(), lr.detach(), 1e-5)
    _check("head_dgrad", planar_from_cb8(acb.grad, 16), ar.grad, DGRAD_TOL)
    check_wgrad("head_wgrad", w.grad, wr.grad, a, g, (1, 1, 1), n * int(np.prod(dims)))
    _check("head_bgrad", b.grad, br.grad, 3e-4)


# ------------------------------------------------------------------------------------------------------------
# layer-by-layer replay of a real train-mode step of the fp32 oracle network
# ------------------------------------------------------------------------------------------------------------
def test_vnet_layers_on_oracle_activations(ops, dev):
    """Train-mode fp32 oracle V-Net (LA size, batch 2) forward/backward on the GPU with hooks on every conv and norm
    layer; each native kernel is then fed the ORACLE's input activation and upstream gradient (rounded to bf16) and
    compared with fp32 torch on the same rounded operands.  This is the gradient check that can fail: a wrong tap,
    channel or split shows up as O(1) error in exactly one layer, while BatchNorm's amplification of bf16 rounding (the
    reason end-to-end train-mode comparisons need loose budgets) does not enter."""
    shape = (2, 1, 112, 112, 80)
    x = O.synthetic_volume(shape, 77).to(dev)
    ref = O.OracleVNet(1, 2, 16, "batchnorm", False)
    O.fill_state_dict_(ref, 23)
    ref = ref.to(dev).train()
    cap = {}

    def fwd_hook(name):
        def h(m, inp, out):
            cap[name] = [inp[0].detach(), None]
            # tensor hook (the in-place ReLU that follows rules out module backward hooks): gradient w.r.t. this output
            out.register_hook(lambda g, name=name: cap[name].__setitem__(1, g.detach()))
        return h
    layers = [(n, m) for n, m in ref.named_modules() if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d, nn.BatchNorm3d))]
    for n, m in layers:
        m.register_forward_hook(fwd_hook(n))
    lo, _ = ref(x)
    (lo * O.synthetic_volume(tuple(lo.shape), 78).to(dev)).sum().backward()
    worst = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0, "norm_fwd": 0.0, "norm_bwd": 0.0}
    checked = 0
    for name, m in layers:
        a, dy = cap[name]
        if dy is None:
            continue
        nvox = a.shape[0] * int(np.prod(a.shape[2:]))
        if isinstance(m, nn.BatchNorm3d):
            c = a.shape[1]
            yb = a.to(torch.bfloat16).float()
            gb = dy.to(torch.bfloat16).float()
            gamma, beta = m.weight.detach().clone().requires_grad_(True), m.bias.detach().clone().requires_grad_(True)
            ycb = cb8_from_planar(yb).requires_grad_(True)
            out = ops.NormAct.apply(ycb, gamma, beta, None, None, None, "batch", a.shape[0], m.eps, 0.1, 0.0, None, None, 1.0, None)
            yr = yb.clone().requires_grad_(True)
            g2, b2 = m.weight.detach().clone().requires_grad_(True), m.bias.detach().clone().requires_grad_(True)
            oref = F.relu(F.batch_norm(yr, None, None, g2, b2, True, 0.1, m.eps))
            oref.backward(gb)
            out.backward(cb8_from_planar(gb))
            e1, e2 = rel_rms(planar_from_cb8(out.detach(), c), oref.detach()), rel_rms(planar_from_cb8(ycb.grad, c), yr.grad)
            worst["norm_fwd"], worst["norm_bwd"] = max(worst["norm_fwd"], e1), max(worst["norm_bwd"], e2)
            assert e1 <= 4e-3 and e2 <= 8e-3, (name, e1, e2)
            assert rel_rms(gamma.grad, g2.grad) <= 2e-3 and rel_rms(beta.grad, b2.grad) <= 2e-3, name
            checked += 1
            continue
        cin, cout = m.in_channels, m.out_channels
        wq = m.weight.detach().to(torch.bfloat16).float()
        ab, gb = a.to(torch.bfloat16).float(), dy.to(torch.bfloat16).float()
        w = wq.clone().requires_grad_(True)
        wr = wq.clone().requires_grad_(True)
        ar = ab.clone().requires_grad_(True)
        if isinstance(m, nn.ConvTranspose3d):
            yr = F.conv_transpose3d(ar, wr, None, stride=2)
            acb = cb8_from_planar(ab).requires_grad_(True)
            y = ops.ConvUp2.apply(acb, w, None, _packs(ops, dev, w, (0, 2, 3)))
        elif m.kernel_size == (2, 2, 2):
            yr = F.conv3d(ar, wr, None, stride=2)
            acb = cb8_from_planar(ab).requires_grad_(True)
            y = ops.ConvDown2.apply(acb, w, None, _packs(ops, dev, w, (0, 2, 3)))
        elif cin == 1:
            ab = a                                        # network input stays fp32
            ar = ab.clone().requires_grad_(True)
            w = m.weight.detach().clone().requires_grad_(True)
            wr = m.weight.detach().clone().requires_grad_(True)
            yr = F.conv3d(ar, wr, None, padding=1)
            acb = None
            y = ops.ConvFirst.apply(ab, w, None)
        elif m.kernel_size == (1, 1, 1):
            w = m.weight.detach().clone().requires_grad_(True)
            wr = m.weight.detach().clone().requires_grad_(True)
            gb = dy                                       # logits gradient is fp32
            yr = F.conv3d(ar, wr, None)
            acb = cb8_from_planar(ab).requires_grad_(True)
            y = ops.Head.apply(acb, w, None, False)
        else:
            yr = F.conv3d(ar, wr, None, padding=1)
            acb = cb8_from_planar(ab).requires_grad_(True)
            y = ops.ConvSame.apply(acb, w, None, _packs(ops, dev, w, (0, 1)), (3, 3, 3))
        yr.backward(gb)
        head = m.kernel_size == (1, 1, 1)
        y.backward(gb if head else cb8_from_planar(gb))
        e_f = rel_rms(y.detach() if head else planar_from_cb8(y.detach(), cout), yr.detach())
        stride1 = isinstance(m, nn.Conv3d) and m.stride == (1, 1, 1)
        e_w = check_wgrad("layerwise_" + name + "_wgrad", w.grad, wr.grad, ab if stride1 else None, gb, m.kernel_size, nvox)
        worst["fwd"], worst["wgrad"] = max(worst["fwd"], e_f), max(worst["wgrad"], e_w)
        assert e_f <= (1e-5 if head else FWD_TOL), (name, "fwd", e_f)
        if acb is not None:
            e_d = rel_rms(planar_from_cb8(acb.grad, cin), ar.grad)
            worst["dgrad"] = max(worst["dgrad"], e_d)
            assert e_d <= DGRAD_TOL, (name, "dgrad", e_d)
        checked += 1
        del y, yr, acb, ar
    for k, v in worst.items():
        record("vnet_layerwise_worst_" + k, v)
    assert checked >= 55, checked            # 30 conv + 29 norm layers of the V-Net
