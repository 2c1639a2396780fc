"""GPU parity tests of the C-ABI kernels (through bcp_b200.ops) against the golden vectors minted from the
reference, the CPU oracle, and plain PyTorch fp32 references of the same op.  Integer/byte/index work is
bit-exact; floating point uses the tolerance written next to each check."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bcp_oracle as O
from tests.util import G, load_golden, T, cb8_from_planar, planar_from_cb8, rel_rms, record

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    torch.backends.cudnn.allow_tf32 = False          # the torch references below must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from bcp_b200 import ops as _ops
    return _ops


def bits_equal_nan_aware(a: np.ndarray, b: np.ndarray):
    na, nb = np.isnan(a), np.isnan(b)
    assert (na == nb).all()
    assert a[~na].tobytes() == b[~nb].tobytes()


# ------------------------------------------------------------------------------------------- mask mix
def test_mask_mix_bit_exact(ops, dev):
    g = load_golden("functions")
    a, b = T(g["mix_a"]).to(dev), T(g["mix_b"]).to(dev)
    out = ops.mask_mix(a, b, (0, 0, 0, 7, 5, 6)).cpu().numpy()
    bits_equal_nan_aware(out, g["mix_out"])      # NaN payloads differ between x86 and sm_100 (canonical NaN); rest bit-exact
    # against the reference's tensor expression evaluated on the same device, random boxes, full LA size
    rs = np.random.RandomState(0)
    x = torch.randn(2, 1, 112, 112, 80, device=dev)
    y = torch.randn(2, 1, 112, 112, 80, device=dev)
    for _ in range(3):
        w, h, z = rs.randint(0, 38), rs.randint(0, 38), rs.randint(0, 27)
        m = torch.ones(112, 112, 80, device=dev)
        m[w:w + 74, h:h + 74, z:z + 53] = 0
        m = m.long()
        ref = x * m + y * (1 - m)
        got = ops.mask_mix(x, y, (w, h, z, 74, 74, 53))
        assert torch.equal(ref, got)
    # 2-D and clipping
    x2, y2 = torch.randn(3, 1, 64, 48, device=dev), torch.randn(3, 1, 64, 48, device=dev)
    m = torch.ones(64, 48, device=dev)
    m[50:50 + 42, 10:10 + 32] = 0
    assert torch.equal(x2 * m.long() + y2 * (1 - m.long()), ops.mask_mix(x2, y2, (50, 10, 42, 32)))
    la, lb = torch.randint(0, 4, (2, 20, 18, 16), device=dev, dtype=torch.uint8), torch.randint(0, 4, (2, 20, 18, 16), device=dev, dtype=torch.uint8)
    mm = torch.ones(20, 18, 16, device=dev, dtype=torch.uint8)
    mm[3:9, 2:11, 5:16] = 0
    assert torch.equal(la * mm + lb * (1 - mm), ops.label_mix(la, lb, (3, 2, 5, 6, 9, 11)))


# ------------------------------------------------------------------------------------------- pseudo labels
def test_pseudo_label_bit_exact(ops, dev):
    g = load_golden("functions")
    for key, cut in (("pl_logits", "pl_cut"), ("pl_logits2", "pl_cut2")):
        lg = T(g[key]).to(dev)
        got = ops.pseudo_label(lg, "thresh", 0.5)
        mism = int((got.cpu().numpy() != g[cut]).sum())
        record("pseudo_label_mismatch_vs_cpu_" + key, mism)
        assert mism == 0                      # bit-exact vs the reference's CPU path, including forced near-ties
        ref_dev = (F.softmax(lg, 1) >= 0.5).long()[:, 1]
        record("pseudo_label_mismatch_vs_torch_gpu_" + key, int((got.long() != ref_dev).sum()))
    lg = T(g["acdc_pl_logits"]).to(dev)
    got = ops.pseudo_label(lg, "argmax")
    assert torch.equal(got.long(), torch.max(F.softmax(lg, 1), 1)[1])
    assert (got.cpu().numpy() == g["acdc_pl_argmax"]).all()
    # ordinary logits: identical to torch's softmax on the same device as well
    x = torch.randn(2, 2, 40, 40, 40, device=dev) * 3
    assert torch.equal(ops.pseudo_label(x).long(), (F.softmax(x, 1) >= 0.5).long()[:, 1])
    # near ties against the CPU reference expression
    x = torch.randn(2, 2, 24, 24, 24)
    x[:, 1] = x[:, 0] + torch.randn(2, 24, 24, 24) * 1e-7
    assert torch.equal(ops.pseudo_label(x.to(dev)).long().cpu(), (F.softmax(x, 1) >= 0.5).long()[:, 1])


# ------------------------------------------------------------------------------------------- connected components
def test_largest_cc_matches_reference(ops, dev):
    g = load_golden("functions")
    cut = T(g["pl_cut2"]).to(torch.uint8).to(dev)
    assert (ops.largest_cc(cut, out_float=True).cpu().numpy() == g["pl_cc2"]).all()
    assert (ops.largest_cc(cut, connectivity=2).cpu().numpy() == g["pl_cc2_conn2"]).all()
    assert (ops.largest_cc(cut, connectivity=1).cpu().numpy() == g["pl_cc2_conn1"]).all()
    empty = torch.zeros(1, 8, 8, 8, dtype=torch.uint8, device=dev)
    assert (ops.largest_cc(empty).cpu().numpy() == g["pl_cc_empty"]).all()
    am = T(g["acdc_pl_argmax"]).to(torch.uint8).to(dev)
    assert (ops.largest_cc(am).cpu().numpy() == g["acdc_pl_cc"]).all()
    # bigger random volumes against the CPU oracle (scipy labelling), all connectivities; blobby + salt noise
    for seed, shape in ((1, (2, 48, 40, 36)), (2, (1, 112, 112, 80))):
        lab = O.synthetic_labels(shape, seed)
        noise = torch.from_numpy((np.random.RandomState(seed).random_sample(shape) > 0.97).astype(np.int64))
        seg = ((lab + noise) > 0).long()
        for conn in (1, 2, 3):
            ref = O.largest_cc(seg, conn).numpy()
            got = ops.largest_cc(seg.to(torch.uint8).to(dev), connectivity=conn).cpu().numpy()
            assert (got == ref).all(), (seed, conn)
    # equal-size components: the first in raster order wins
    t = torch.zeros(1, 4, 8, 8, dtype=torch.uint8)
    t[0, 0, 0:2, 0:2] = 1
    t[0, 3, 5:7, 5:7] = 1
    assert (ops.largest_cc(t.to(dev)).cpu() == torch.from_numpy(O.largest_cc(t.long()).numpy().astype(np.uint8))).all()
    seg4 = O.synthetic_labels((4, 96, 80), 7, n_classes=4)
    assert (ops.largest_cc(seg4.to(torch.uint8).to(dev)).cpu().numpy() == O.acdc_2d_largest_cc(seg4).numpy()).all()


# ------------------------------------------------------------------------------------------- losses
def test_mix_loss_vs_golden(ops, dev):
    g = load_golden("functions")
    lg = T(g["la_logits"]).to(dev)
    la, lb = T(g["la_lab_a"]).to(torch.uint8).to(dev), T(g["la_lab_b"]).to(torch.uint8).to(dev)
    box = (2, 3, 1, 7, 5, 5)
    x = lg.clone().requires_grad_(True)
    r = ops.MixLoss.apply(x, la, lb, box, None, 0, 1.0, 0.5)
    r[0].backward()
    assert abs(float(r[0]) - float(g["la_mix_loss"])) <= 1e-5 * abs(float(g["la_mix_loss"]))       # rel 1e-5
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["la_mix_grad"], rtol=1e-4, atol=1e-7)
    x = lg.clone().requires_grad_(True)
    r = ops.MixLoss.apply(x, lb, la, box, None, 0, 0.5, 1.0)
    r[0].backward()
    assert abs(float(r[0]) - float(g["la_mix_loss_unlab"])) <= 1e-5 * abs(float(g["la_mix_loss_unlab"]))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["la_mix_grad_unlab"], rtol=1e-4, atol=1e-7)
    # explicit mask tensor path == box path
    m8 = T(g["la_mask"]).to(torch.uint8).to(dev)
    r2 = ops.MixLoss.apply(lg, lb, la, None, m8, 0, 0.5, 1.0)
    assert torch.allclose(r2, r.detach(), rtol=1e-6, atol=0)
    # pre-train loss = empty box
    x = lg.clone().requires_grad_(True)
    r = ops.MixLoss.apply(x, la, la, (0, 0, 0, 0, 0, 0), None, 0, 1.0, 0.0)
    r[0].backward()
    assert abs(float(r[0]) - float(g["la_pre_loss"])) <= 1e-5 * abs(float(g["la_pre_loss"]))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["la_pre_grad"], rtol=1e-4, atol=1e-7)
    assert abs(float(r[1]) - float(g["la_dice_unmasked"])) <= 1e-5
    # ACDC form
    lg4 = T(g["acdc_logits"]).to(dev)
    ta, tb = T(g["acdc_lab_a"]).to(dev), T(g["acdc_lab_b"]).to(torch.uint8).to(dev)
    x = lg4.clone().requires_grad_(True)
    r = ops.MixLoss.apply(x, ta, tb, (3, 2, 10, 8), None, 1, 1.0, 0.5)
    ((r[1] + r[2]) / 2).backward()
    assert abs(float(r[1]) - float(g["acdc_mix_dice"])) <= 1e-5 * abs(float(g["acdc_mix_dice"]))
    assert abs(float(r[2]) - float(g["acdc_mix_ce"])) <= 1e-5 * abs(float(g["acdc_mix_ce"]))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["acdc_mix_grad"], rtol=1e-4, atol=1e-7)
    r = ops.MixLoss.apply(lg4, tb, ta, (3, 2, 10, 8), None, 1, 0.5, 1.0)
    assert abs(float(r[1]) - float(g["acdc_mix_dice_unlab"])) <= 1e-5 * abs(float(g["acdc_mix_dice_unlab"]))


def test_mix_loss_full_size_vs_oracle(ops, dev):
    torch.manual_seed(0)
    lg = torch.randn(2, 2, 112, 112, 80) * 2
    la, lb = O.synthetic_labels((2, 112, 112, 80), 3), O.synthetic_labels((2, 112, 112, 80), 4)
    mask, lmask, box = O.context_mask_la(lg, 2 / 3, np.random.RandomState(5))
    xo = lg.clone().requires_grad_(True)
    lo = O.mix_loss_la(xo, la, lb, lmask, u_weight=0.5)
    lo.backward()
    x = lg.to(dev).requires_grad_(True)
    r = ops.MixLoss.apply(x, la.to(torch.uint8).to(dev), lb.to(torch.uint8).to(dev), box, None, 0, 1.0, 0.5)
    r[0].backward()
    rel = abs(float(r[0]) - float(lo)) / abs(float(lo))
    record("mix_loss_full_rel_err", rel)
    assert rel <= 2e-5
    assert rel_rms(x.grad.cpu(), xo.grad) <= 1e-4


# ------------------------------------------------------------------------------------------- optimiser / EMA
def test_sgd_ema_vs_torch(dev):
    from bcp_b200._native import LIB, ptr, stream
    n = 100003
    torch.manual_seed(1)
    p0, g1, g2, e0 = torch.randn(n), torch.randn(n), torch.randn(n), torch.randn(n)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([p], lr=0.01, momentum=0.9, weight_decay=1e-4)
    e = e0.clone()
    pd, ed, buf = p0.to(dev).clone(), e0.to(dev).clone(), torch.zeros(n, device=dev)
    hyper = torch.tensor([0.01, 0.9, 1e-4, 0.99, 1.0, 1.0 - 0.99], dtype=torch.float32, device=dev)
    for gstep in (g1, g2):
        p.grad = gstep.clone()
        opt.step()
        e.mul_(0.99).add_((1 - 0.99) * p.data)
        LIB.call("bcp_sgd_ema_step", ptr(pd), ptr(gstep.to(dev)), ptr(buf), ptr(ed), ptr(hyper), n - 7, n, stream())
    assert torch.allclose(pd.cpu()[:n - 7], p.data[:n - 7], rtol=1e-6, atol=1e-7)
    assert torch.equal(pd.cpu()[n - 7:], p0[n - 7:])                       # EMA-only tail is not stepped
    assert torch.allclose(ed.cpu()[:n - 7], e[:n - 7], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------- layout
def test_layout_roundtrip(ops, dev):
    x = torch.randn(2, 24, 5, 6, 7, device=dev)
    a = ops.PlanarToCB8.apply(x)
    assert torch.equal(a, cb8_from_planar(x))
    assert torch.equal(ops.CB8ToPlanar.apply(a, 24, False), x.to(torch.bfloat16).float())


# ------------------------------------------------------------------------------------------- norm + act
@pytest.mark.parametrize("c,spg,slope,use_drop,use_res", [(16, 2, 0.0, True, False), (32, 2, 0.0, False, True), (16, 1, 0.0, False, False), (64, 3, 0.01, False, False)])
def test_norm_act_vs_torch(ops, dev, c, spg, slope, use_drop, use_res):
    torch.manual_seed(c + spg)
    n = 2 * spg
    y = (torch.randn(n, c, 6, 10, 12, device=dev) * 2 + 0.5).to(torch.bfloat16).float()
    gamma, beta = (1 + 0.1 * torch.randn(c, device=dev)).requires_grad_(True), (0.1 * torch.randn(c, device=dev)).requires_grad_(True)
    rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    cs = (torch.rand(n, c, device=dev) > 0.5).float() * 2 if use_drop else None
    res = torch.randn(n, c, 6, 10, 12, device=dev).to(torch.bfloat16).float() if use_res else None
    ycb = cb8_from_planar(y).requires_grad_(True)
    rescb = cb8_from_planar(res).requires_grad_(True) if use_res else None
    out = ops.NormAct.apply(ycb, gamma, beta, rm, rv, nbt, "batch", spg, 1e-5, 0.1, slope, cs, None, 1.0, rescb)
    # fp32 torch reference: one F.batch_norm call per group (updates running stats sequentially, like the reference's
    # two forward calls)
    yr = y.clone().requires_grad_(True)
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    outs = []
    for gi in range(n // spg):
        o = F.batch_norm(yr[gi * spg:(gi + 1) * spg], rm2, rv2, g2, b2, True, 0.1, 1e-5)
        outs.append(F.leaky_relu(o, slope) if slope else F.relu(o))
    ref = torch.cat(outs)
    if use_drop:
        ref = ref * cs[:, :, None, None, None]
    if use_res:
        ref = ref + res
    got = planar_from_cb8(out.detach(), c)
    assert rel_rms(got, ref.detach()) <= 4e-3                 # bf16 output rounding (2^-9 relative)
    assert torch.allclose(rm, rm2, rtol=1e-5, atol=1e-6) and torch.allclose(rv, rv2, rtol=1e-4, atol=1e-6)
    assert int(nbt) == n // spg
    w = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(w)
    out.backward(cb8_from_planar(w))
    assert rel_rms(planar_from_cb8(ycb.grad, c), yr.grad) <= 6e-3
    assert rel_rms(gamma.grad, g2.grad) <= 2e-3 and rel_rms(beta.grad, b2.grad) <= 2e-3
    if use_res:
        assert torch.equal(planar_from_cb8(rescb.grad, c), w)


@pytest.mark.parametrize("n,c,dims,spg,drop,res", [(4, 64, (28, 28, 20), 2, False, False), (4, 128, (14, 14, 10), 2, True, False),
                                                   (4, 256, (7, 7, 5), 2, False, True), (12, 128, (1, 32, 32), 6, False, False),
                                                   (3, 64, (24, 24, 24), 1, False, False), (2, 24, (5, 9, 7), 1, True, True),
                                                   (4, 32, (28, 28, 40), 2, False, False), (4, 32, (40, 40, 40), 1, True, True)])
def test_norm_fused_cluster_vs_streaming(ops, dev, n, c, dims, spg, drop, res):
    """csrc/norm_fused.cu (one cluster kernel per direction) against csrc/norm.cu (statistics / apply / reduce / apply
    launches) on the mid-size and deep layer shapes it serves: same statistics to fp32 rounding, same running-stat updates,
    outputs and gradients equal up to rare 1-ulp bf16 flips; and a batched call of G groups is BIT-identical to G calls."""
    from bcp_b200._native import LIB
    assert LIB.query("bcp_norm_fused_supported", n, c, int(np.prod(dims)), spg) == 1
    torch.manual_seed(n * c + spg)
    y = cb8_from_planar((torch.randn(n, c, *dims, device=dev) * 1.5 + 0.3).to(torch.bfloat16).float())
    w = cb8_from_planar(torch.randn(n, c, *dims, device=dev).to(torch.bfloat16).float())
    cs = (torch.rand(n, c, device=dev) > 0.5).float() * 2 if drop else None
    r0 = cb8_from_planar(torch.randn(n, c, *dims, device=dev).to(torch.bfloat16).float()) if res else None
    runs = {}
    for fused in (True, False):
        old = ops._NORM_FUSED
        ops._NORM_FUSED = fused
        try:
            gamma = (1 + 0.1 * torch.arange(c, device=dev).float().sin()).requires_grad_(True)
            beta = (0.1 * torch.arange(c, device=dev).float().cos()).requires_grad_(True)
            rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
            ycb = y.clone().requires_grad_(True)
            rcb = r0.clone().requires_grad_(True) if res else None
            l0 = LIB.launches
            out = ops.NormAct.apply(ycb, gamma, beta, rm, rv, nbt, "batch", spg, 1e-5, 0.1, 0.01, cs, None, 1.0, rcb)
            fwd_launches = LIB.launches - l0
            out.backward(w)
            torch.cuda.synchronize()
            runs[fused] = dict(out=out.detach().float(), dy=ycb.grad.float(), dg=gamma.grad.clone(), db=beta.grad.clone(), rm=rm, rv=rv,
                               nbt=int(nbt), launches=fwd_launches)
        finally:
            ops._NORM_FUSED = old
    a, b = runs[True], runs[False]
    assert a["launches"] == 1 and b["launches"] == 2
    assert a["nbt"] == b["nbt"] == n // spg
    assert torch.allclose(a["rm"], b["rm"], rtol=1e-6, atol=1e-7) and torch.allclose(a["rv"], b["rv"], rtol=1e-6, atol=1e-7)
    assert rel_rms(a["out"], b["out"]) <= 2e-4 and rel_rms(a["dy"], b["dy"]) <= 2e-4
    assert rel_rms(a["dg"], b["dg"]) <= 1e-5 and rel_rms(a["db"], b["db"]) <= 1e-5
    if n // spg > 1:            # batched groups == separate calls, bit for bit (forward, running statistics, input gradient)
        gamma = (1 + 0.1 * torch.arange(c, device=dev).float().sin())
        beta = (0.1 * torch.arange(c, device=dev).float().cos())
        rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
        outs, dys = [], []
        for g in range(n // spg):
            sl = slice(g * spg, (g + 1) * spg)
            yg = y[sl].clone().requires_grad_(True)
            o = ops.NormAct.apply(yg, gamma, beta, rm, rv, nbt, "batch", spg, 1e-5, 0.1, 0.01, cs[sl].contiguous() if drop else None,
                                  None, 1.0, r0[sl].clone() if res else None)
            o.backward(w[sl].contiguous())
            outs.append(o.detach().float())
            dys.append(yg.grad.float())
        assert torch.equal(torch.cat(outs), a["out"]) and torch.equal(torch.cat(dys), a["dy"])
        assert torch.equal(rm, a["rm"]) and torch.equal(rv, a["rv"])


def test_norm_eval_and_instance(ops, dev):
    torch.manual_seed(3)
    c, n = 16, 2
    y = torch.randn(n, c, 4, 6, 8, device=dev).to(torch.bfloat16).float()
    gamma, beta = 1 + 0.1 * torch.randn(c, device=dev), 0.1 * torch.randn(c, device=dev)
    rm, rv = 0.1 * torch.randn(c, device=dev), 1 + 0.1 * torch.rand(c, device=dev)
    out = ops.NormAct.apply(cb8_from_planar(y), gamma, beta, rm, rv, None, "eval", n, 1e-5, 0.0, 0.0, None, None, 1.0, None)
    ref = F.relu(F.batch_norm(y, rm, rv, gamma, beta, False, 0.0, 1e-5))
    assert rel_rms(planar_from_cb8(out, c), ref) <= 4e-3
    out = ops.NormAct.apply(cb8_from_planar(y), None, None, None, None, None, "batch", 1, 1e-5, 0.0, 0.0, None, None, 1.0, None)
    ref = F.relu(F.instance_norm(y, eps=1e-5))
    assert rel_rms(planar_from_cb8(out, c), ref) <= 4e-3


def test_element_dropout_mask(ops, dev):
    torch.manual_seed(4)
    c, n = 16, 2
    y = torch.randn(n, c, 1, 12, 10, device=dev).to(torch.bfloat16).float()
    keep_planar = (torch.rand(n, c, 1, 12, 10, device=dev) > 0.3)
    keep = keep_planar.to(torch.uint8).reshape(n, c // 8, 8, 1, 12, 10).permute(0, 1, 3, 4, 5, 2).contiguous()
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    ycb = cb8_from_planar(y).requires_grad_(True)
    out = ops.NormAct.apply(ycb, gamma, beta, None, None, None, "batch", n, 1e-5, 0.1, 0.01, None, keep, 1 / 0.7, None)
    yr = y.clone().requires_grad_(True)
    ref = F.leaky_relu(F.batch_norm(yr, None, None, gamma, beta, True, 0.1, 1e-5), 0.01) * keep_planar / 0.7
    assert rel_rms(planar_from_cb8(out.detach(), c), ref.detach()) <= 4e-3
    w = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(w)
    out.backward(cb8_from_planar(w))
    assert rel_rms(planar_from_cb8(ycb.grad, c), yr.grad) <= 6e-3


# ------------------------------------------------------------------------------------------- convolutions
def _packs(ops, dev, weight, kinds):
    """Build operand packs for one weight through the real repack kernel."""
    from bcp_b200.networks.runtime import _JOB
    from bcp_b200._native import LIB, ptr, stream
    a, b = weight.shape[0], weight.shape[1]
    taps = int(np.prod(weight.shape[2:]))
    jobs, views, dst = [], [], 0
    for kind in kinds:
        n = taps * ((b + 7) // 8) * a * 8 if kind == 0 else taps * ((a + 7) // 8) * b * 8
        jobs.append((0, dst, a, b, taps, kind))
        views.append((dst, n))
        dst += (n + 127) // 128 * 128
    packed = torch.zeros(dst, dtype=torch.bfloat16, device=dev)
    jt = torch.from_numpy(np.array(jobs, dtype=_JOB).view(np.uint8).copy()).to(dev)
    w = weight.detach().contiguous()
    LIB.call("bcp_weights_repack", ptr(w), ptr(packed), ptr(jt), len(jobs), stream())
    return ops.ConvPack({kind: packed[o:o + n] for kind, (o, n) in zip(kinds, views)})


def _check_conv(got_y, ref_y, grads_got, grads_ref, tol_fwd=6e-3, tol_bwd=1e-2):
    assert rel_rms(got_y, ref_y) <= tol_fwd, rel_rms(got_y, ref_y)
    for name in grads_ref:
        e = rel_rms(grads_got[name], grads_ref[name])
        assert e <= tol_bwd, (name, e)


@pytest.mark.parametrize("cin,cout,dims,kernel", [
    (16, 16, (6, 10, 12), (3, 3, 3)), (32, 16, (5, 7, 9), (3, 3, 3)), (16, 32, (1, 12, 14), (1, 3, 3)),
    (64, 32, (1, 9, 8), (1, 1, 1)), (8, 24, (4, 5, 6), (3, 3, 3))])
def test_conv_same_direct(ops, dev, cin, cout, dims, kernel):
    torch.manual_seed(cin * 7 + cout)
    n = 2
    x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, *kernel, device=dev) / np.sqrt(cin * np.prod(kernel))).to(torch.bfloat16).float().requires_grad_(True)
    b = (0.1 * torch.randn(cout, device=dev)).requires_grad_(True)
    pack = _packs(ops, dev, w, (0, 1))
    xcb = cb8_from_planar(x).requires_grad_(True)
    import bcp_b200.ops as O2
    y = O2.ConvSame.apply(xcb, w, b, pack, kernel)
    xr = x.clone().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br, padding=tuple(k // 2 for k in kernel))
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    _check_conv(planar_from_cb8(y.detach(), cout), yr.detach(),
                dict(x=planar_from_cb8(xcb.grad, cin), w=w.grad, b=b.grad), dict(x=xr.grad, w=wr.grad, b=br.grad))


@pytest.mark.parametrize("cin,cout,dims,kernel", [(16, 16, (1, 12, 260), (1, 3, 3)),       # three windows of 96 columns, ragged last one
                                                  (32, 16, (1, 6, 256), (1, 3, 3)),        # the ACDC full-resolution decoder layer
                                                  (16, 32, (3, 5, 272), (3, 3, 3))])       # 3-D, dx-folded plan per window
def test_conv_wgrad_z_windows(ops, dev, cin, cout, dims, kernel):
    """tcgen05 weight gradient on z-lines longer than a TMA box (> 254 columns): one launch per window of <= 128 columns,
    dy map based at the window, `a` map shifted by WgParams::zoff so the halo columns are the real neighbours
    (csrc/conv_tc.cu conv_tc_wgrad_impl; autograd's dW of networks/unet.py:20,24 at 256 x 256)."""
    from bcp_b200._native import LIB, i3
    assert LIB.query("bcp_conv_tc_wgrad_supported", cin, cout, i3(*dims), i3(*kernel)) == 1
    torch.manual_seed(cin + cout + dims[2])
    n = 2
    x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
    g = torch.randn(n, cout, *dims, device=dev).to(torch.bfloat16).float()
    w = torch.zeros(cout, cin, *kernel, device=dev, requires_grad=True)
    pad = tuple(k // 2 for k in kernel)
    F.conv3d(x, w, None, padding=pad).backward(g)
    a, dy = cb8_from_planar(x), cb8_from_planar(g)
    d_t = ops._wgrad(a, dy, cin, cout, dims, kernel, (1, 1, 1), pad, w.shape, allow_tc=True)
    d_d = ops._wgrad(a, dy, cin, cout, dims, kernel, (1, 1, 1), pad, w.shape, allow_tc=False)
    # accumulate mode: a second call adds the same gradient again
    acc = d_t.clone()
    ops._wgrad(a, dy, cin, cout, dims, kernel, (1, 1, 1), pad, w.shape, allow_tc=True, into=acc)
    e_t, e_d, e_a = rel_rms(d_t, w.grad), rel_rms(d_d, w.grad), rel_rms(acc, 2 * w.grad)
    record("wgrad_zwin_%d_%d_%s" % (cin, cout, "x".join(map(str, dims))), e_t)
    assert e_t <= 1e-4 and e_d <= 1e-4 and e_a <= 1e-4, (e_t, e_d, e_a)


def test_conv_stride2_family(ops, dev):
    torch.manual_seed(11)
    n, cin, cout = 2, 16, 32
    x = torch.randn(n, cin, 8, 6, 10, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, 2, 2, 2, device=dev) / np.sqrt(cin * 8)).to(torch.bfloat16).float().requires_grad_(True)
    b = (0.1 * torch.randn(cout, device=dev)).requires_grad_(True)
    xcb = cb8_from_planar(x).requires_grad_(True)
    y = ops.ConvDown2.apply(xcb, w, b, _packs(ops, dev, w, (0, 2, 3)))
    xr, wr, br = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br, stride=2)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    _check_conv(planar_from_cb8(y.detach(), cout), yr.detach(),
                dict(x=planar_from_cb8(xcb.grad, cin), w=w.grad, b=b.grad), dict(x=xr.grad, w=wr.grad, b=br.grad))
    # transposed
    cin, cout = 32, 16
    x = torch.randn(n, cin, 4, 3, 5, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(cin, cout, 2, 2, 2, device=dev) / np.sqrt(cin)).to(torch.bfloat16).float().requires_grad_(True)
    b = (0.1 * torch.randn(cout, device=dev)).requires_grad_(True)
    xcb = cb8_from_planar(x).requires_grad_(True)
    y = ops.ConvUp2.apply(xcb, w, b, _packs(ops, dev, w, (0, 2, 3)))
    xr, wr, br = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv_transpose3d(xr, wr, br, stride=2)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    _check_conv(planar_from_cb8(y.detach(), cout), yr.detach(),
                dict(x=planar_from_cb8(xcb.grad, cin), w=w.grad, b=b.grad), dict(x=xr.grad, w=wr.grad, b=br.grad))


def test_conv_first_and_head(ops, dev):
    torch.manual_seed(12)
    x = torch.randn(2, 1, 10, 12, 14, device=dev)
    w = (torch.randn(16, 1, 3, 3, 3, device=dev) / 5).requires_grad_(True)
    b = (0.1 * torch.randn(16, device=dev)).requires_grad_(True)
    y = ops.ConvFirst.apply(x, w, b)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.conv3d(x, wr, br, padding=1)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g))
    _check_conv(planar_from_cb8(y.detach(), 16), yr.detach(), dict(w=w.grad, b=b.grad), dict(w=wr.grad, b=br.grad))
    x2 = torch.randn(3, 1, 20, 16, device=dev)
    w2 = (torch.randn(16, 1, 3, 3, device=dev) / 3).requires_grad_(True)
    y2 = ops.ConvFirst.apply(x2, w2, None)
    assert rel_rms(planar_from_cb8(y2.detach(), 16)[:, :, 0], F.conv2d(x2, w2.detach(), padding=1)) <= 6e-3
    # heads: 1x1x1 16->2 and 2-D 3x3 16->4
    for two_d, ncls, k, sp in ((False, 2, (1, 1, 1), (6, 8, 10)), (True, 4, (3, 3), (14, 18))):
        a = torch.randn(2, 16, *sp, device=dev).to(torch.bfloat16).float()
        w = (torch.randn(ncls, 16, *k, device=dev) / 4).requires_grad_(True)
        b = (0.1 * torch.randn(ncls, device=dev)).requires_grad_(True)
        a5 = a.unsqueeze(2) if two_d else a
        acb = cb8_from_planar(a5).requires_grad_(True)
        lo = ops.Head.apply(acb, w, b, two_d)
        ar, wr, br = a.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        lr = F.conv2d(ar, wr, br, padding=1) if two_d else F.conv3d(ar, wr, br)
        g = torch.randn_like(lr)
        lr.backward(g)
        lo.backward(g)
        assert rel_rms(lo.detach(), lr.detach()) <= 1e-5
        ga = planar_from_cb8(acb.grad, 16)
        ga = ga[:, :, 0] if two_d else ga
        assert rel_rms(ga, ar.grad) <= 6e-3 and rel_rms(w.grad, wr.grad) <= 1e-4 and rel_rms(b.grad, br.grad) <= 1e-4


@pytest.mark.parametrize("shape,kernel", [((2, 1, 10, 12, 16), (3, 3, 3)),      # one z tile, merged dy map, ragged last y tile
                                          ((1, 1, 5, 9, 132), (3, 3, 3)),       # z run > 128: un-merged 5-D dy map
                                          ((3, 1, 24, 260), (3, 3)),            # 2-D, three z tiles, ragged last one
                                          ((2, 1, 40, 16), (3, 3))])            # 2-D, Cob*kx = 2 -> six splits per octet
def test_conv_first_wgrad_tma_path(ops, dev, shape, kernel):
    """The TMA-staged first-layer weight gradient (csrc/conv_first_tma.cu) on shapes that exercise its tiling: box loads
    with out-of-bounds zero fill on every side, several z tiles, both dy tensor-map forms (autograd's dW of
    networks/VNet.py:151 block_one / networks/unet.py:72 in_conv)."""
    from bcp_b200._native import LIB, i3
    torch.manual_seed(31)
    two_d = len(shape) == 4
    dims = (1,) + tuple(shape[2:]) if two_d else tuple(shape[2:])
    k3 = (1,) + tuple(kernel) if two_d else tuple(kernel)
    assert LIB.query("bcp_conv_first_wgrad_tma_chunks", shape[0], 16, i3(*dims), i3(*k3)) > 0
    x = torch.randn(*shape, device=dev)
    w = (torch.randn(16, 1, *kernel, device=dev) / 4).requires_grad_(True)
    y = ops.ConvFirst.apply(x, w, None)
    wr = w.detach().clone().requires_grad_(True)
    yr = F.conv2d(x, wr, padding=1) if two_d else F.conv3d(x, wr, padding=1)
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    g5 = g.unsqueeze(2) if two_d else g
    y.backward(cb8_from_planar(g5))
    err = rel_rms(w.grad, wr.grad)
    record("first_wgrad_tma_%s" % "x".join(map(str, shape)), err)
    assert err <= 1e-5, err


def test_pool_and_upsample(ops, dev):
    torch.manual_seed(13)
    x = torch.randn(2, 16, 12, 20, device=dev).to(torch.bfloat16).float()
    xcb = cb8_from_planar(x.unsqueeze(2)).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y, yr = ops.MaxPool2.apply(xcb), F.max_pool2d(xr, 2)
    assert torch.equal(planar_from_cb8(y.detach(), 16)[:, :, 0], yr.detach())
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g.unsqueeze(2)))
    assert torch.equal(planar_from_cb8(xcb.grad, 16)[:, :, 0], xr.grad)
    xcb = cb8_from_planar(x.unsqueeze(2)).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y, yr = ops.Upsample2.apply(xcb), F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    assert rel_rms(planar_from_cb8(y.detach(), 16)[:, :, 0], yr.detach()) <= 4e-3
    g = torch.randn_like(yr).to(torch.bfloat16).float()
    yr.backward(g)
    y.backward(cb8_from_planar(g.unsqueeze(2)))
    assert rel_rms(planar_from_cb8(xcb.grad, 16)[:, :, 0], xr.grad) <= 6e-3
    x5 = torch.randn(2, 32, 7, 7, 5, device=dev).to(torch.bfloat16).float()
    assert torch.equal(ops.maxpool3d_k3s2(cb8_from_planar(x5), 32), F.max_pool3d(x5, 3, stride=2))


@pytest.mark.parametrize("cin,cout,dims,kernel,spg", [(16, 16, (24, 24, 16), (3, 3, 3), 2), (32, 64, (20, 22, 20), (3, 3, 3), 2),
                                                        (16, 32, (1, 96, 96), (1, 3, 3), 3)])
def test_conv_fused_norm_statistics(ops, dev, cin, cout, dims, kernel, spg):
    """bcp_conv_tc_fwd_stats: the statistics produced by the conv epilogue equal those of the separate pass over the
    stored tensor (same stat/coef up to fp32 rounding of a double sum, identical running statistics semantics)."""
    torch.manual_seed(cin + cout)
    n = 2 * spg
    x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, *kernel, device=dev) / np.sqrt(cin * np.prod(kernel))).to(torch.bfloat16).float()
    b = 0.1 * torch.randn(cout, device=dev)
    pack = _packs(ops, dev, w, (0, 1))
    a = cb8_from_planar(x)
    gamma, beta = 1 + 0.1 * torch.randn(cout, device=dev), 0.1 * torch.randn(cout, device=dev)

    def run(fused):
        rm, rv = torch.zeros(cout, device=dev), torch.ones(cout, device=dev)
        nbt = torch.zeros((), dtype=torch.int64, device=dev)
        req = dict(gamma=gamma, beta=beta, running_mean=rm, running_var=rv, nbt=nbt, spg=spg, eps=1e-5, momentum=0.1)
        old, old_fold = ops._FUSE_STATS, ops._TC_FOLD
        ops._FUSE_STATS, ops._TC_FOLD = fused, False      # same (unfolded) kernel on both sides: y must be bit-identical
        try:
            y = ops.ConvSame.apply(a, w, b, pack, kernel, False, req)
        finally:
            ops._FUSE_STATS, ops._TC_FOLD = old, old_fold
        assert ("out" in req) == fused, "layer expected to be eligible for the fused-statistics epilogue"
        out = ops.NormAct.apply(y, gamma, beta, rm, rv, nbt, "batch", spg, 1e-5, 0.1, 0.0, None, None, 1.0, None, req.get("out"))
        torch.cuda.synchronize()
        return y, out, rm, rv, int(nbt), req.get("out")

    y0, o0, rm0, rv0, nbt0, _ = run(False)
    y1, o1, rm1, rv1, nbt1, (stat, coef) = run(True)
    assert torch.equal(y0, y1)
    assert nbt0 == nbt1 == n // spg
    assert torch.allclose(rm0, rm1, rtol=1e-5, atol=1e-7) and torch.allclose(rv0, rv1, rtol=1e-5, atol=1e-7)
    # reference statistics of the stored tensor in double
    yp = planar_from_cb8(y1, cout).double().reshape(n // spg, spg, cout, -1)
    mean = yp.mean(dim=(1, 3))
    var = yp.var(dim=(1, 3), unbiased=False)
    assert torch.allclose(stat[..., 0].double(), mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(stat[..., 1].double(), 1 / torch.sqrt(var + 1e-5), rtol=1e-5)
    e = rel_rms(planar_from_cb8(o1, cout), planar_from_cb8(o0, cout))
    record("fused_stats_out_rel_rms_c%d" % cout, e)
    assert e <= 1e-3            # a 1-ulp difference of scale/shift flips a handful of bf16 roundings


@pytest.mark.parametrize("n,cin,cout,dims,kernel", [(1, 16, 16, (4, 6, 8), (3, 3, 3)), (2, 16, 16, (8, 12, 20), (3, 3, 3)),
                                                    (2, 32, 32, (6, 10, 12), (3, 3, 3)), (2, 16, 32, (9, 7, 11), (3, 3, 3)),
                                                    (3, 16, 16, (1, 64, 64), (1, 3, 3)), (1, 16, 16, (112, 112, 80), (3, 3, 3))])
def test_conv_tc_fold_vs_torch(ops, dev, n, cin, cout, dims, kernel):
    """conv_tc_fold_kernel (dz folded into N, SBO = 96 B row groups) against torch on bf16-rounded operands."""
    torch.manual_seed(n + cin + cout)
    x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, *kernel, device=dev) / np.sqrt(cin * np.prod(kernel))).to(torch.bfloat16).float()
    b = 0.1 * torch.randn(cout, device=dev)
    pack = _packs(ops, dev, w, (0, 1))
    old = ops._TC_FOLD
    ys = {}
    try:
        for fold in (True, False):
            ops._TC_FOLD = fold
            ys[fold] = ops._conv_same(cb8_from_planar(x), pack.k[0], b, cout, kernel)
    finally:
        ops._TC_FOLD = old
    ref = F.conv3d(x, w, b, padding=tuple(k // 2 for k in kernel))
    assert rel_rms(planar_from_cb8(ys[True], cout), ref) <= 6e-3
    assert rel_rms(planar_from_cb8(ys[False], cout), ref) <= 6e-3
    # the two kernels sum the 27 taps in different orders (fp32 accumulate): equal up to the last bf16 ulp
    assert rel_rms(planar_from_cb8(ys[True], cout), planar_from_cb8(ys[False], cout)) <= 3e-3
