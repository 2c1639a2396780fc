"""Pins oracle/bcp_oracle.py against golden vectors minted from the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bcp_oracle as O
from tests.golden.golden_common import inject_dropout, digest_named, tensor_digest

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=1e-5, atol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.allclose(a, b, rtol=rtol, atol=atol), float(np.abs(a - b).max())


def digests_close(a, b, rtol=2e-4):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = np.abs(b[:, 1:2]) + 1e-12           # abs-sum of each tensor
    err = np.abs(a - b)
    assert (err[:, :2] <= rtol * scale + 1e-7).all(), float((err[:, :2] / scale).max())


def test_function_vectors():
    g = load("functions")
    lg, la, lb, m = T(g["la_logits"]), T(g["la_lab_a"]), T(g["la_lab_b"]), T(g["la_mask"])
    close(O.mask_dice_loss(lg, la, m), g["la_dice_masked"])
    close(O.mask_dice_loss(lg, la), g["la_dice_unmasked"])
    x = lg.clone().requires_grad_(True)
    l = O.mix_loss_la(x, la, lb, m, u_weight=0.5)
    l.backward()
    close(l.detach(), g["la_mix_loss"])
    close(x.grad, g["la_mix_grad"], atol=1e-8)
    x = lg.clone().requires_grad_(True)
    l = O.mix_loss_la(x, lb, la, m, u_weight=0.5, unlab=True)
    l.backward()
    close(l.detach(), g["la_mix_loss_unlab"])
    close(x.grad, g["la_mix_grad_unlab"], atol=1e-8)
    close(O.mix_loss_la(lg, la, lb.long(), m).detach(), g["pan_mix_loss"])
    x = lg.clone().requires_grad_(True)
    lp = (F.cross_entropy(x, la) + O.mask_dice_loss(x, la)) / 2
    lp.backward()
    close(lp.detach(), g["la_pre_loss"])
    close(x.grad, g["la_pre_grad"], atol=1e-8)

    lg4, ta, tb, m2 = T(g["acdc_logits"]), T(g["acdc_lab_a"]), T(g["acdc_lab_b"]), T(g["acdc_mask"])
    x = lg4.clone().requires_grad_(True)
    d, c = O.mix_loss_acdc(x, ta, tb, m2, u_weight=0.5)
    ((d + c) / 2).backward()
    close(d.detach(), g["acdc_mix_dice"])
    close(c.detach(), g["acdc_mix_ce"])
    close(x.grad, g["acdc_mix_grad"], atol=1e-8)
    d, c = O.mix_loss_acdc(lg4, tb, ta, m2, u_weight=0.5, unlab=True)
    close(d, g["acdc_mix_dice_unlab"])
    close(c, g["acdc_mix_ce_unlab"])


def _bbox(mask):
    return np.array([[int(i.min()), int(i.max()) + 1] for i in torch.nonzero(mask == 0, as_tuple=True)])


def test_mask_vectors():
    g = load("functions")
    rs = np.random.RandomState(1337)
    mk, lm, box = O.context_mask_la(torch.zeros(2, 1, 112, 112, 80), 2 / 3, rs)
    assert (_bbox(mk) == g["la_ctx_mask_zero_bbox"]).all()
    assert int(mk.sum()) == int(g["la_ctx_mask_sum"]) and int(lm.sum()) == int(g["la_ctx_lmask_sum"])
    assert box[3:] == (74, 74, 53)
    rs = np.random.RandomState(1337)
    mk, lm, box = O.generate_mask_acdc(torch.zeros(6, 1, 256, 256), rs)
    assert (_bbox(mk) == g["acdc_mask_zero_bbox"]).all() and int(mk.sum()) == int(g["acdc_mask_sum"])
    rs = np.random.RandomState(2020)
    mk, lm, box = O.generate_mask_pan(torch.zeros(2, 1, 96, 96, 96), 64, rs)
    assert (_bbox(mk) == g["pan_mask_zero_bbox"]).all()
    out = O.mask_mix(T(g["mix_a"]), T(g["mix_b"]), T(g["mix_mask"]))
    assert out.numpy().tobytes() == g["mix_out"].tobytes()          # bit-exact incl. NaN payload / -0.0


def test_pseudo_label_vectors():
    g = load("functions")
    assert (O.get_cut_mask(T(g["pl_logits"]), nms=0).numpy() == g["pl_cut"]).all()
    assert (O.get_cut_mask(T(g["pl_logits2"]), nms=0).numpy() == g["pl_cut2"]).all()
    assert (O.get_cut_mask(T(g["pl_logits2"]), nms=1).numpy() == g["pl_cc2"]).all()
    assert (O.get_cut_mask(T(g["pl_logits2"]), nms=1, connectivity=2).numpy() == g["pl_cc2_conn2"]).all()
    assert (O.get_cut_mask(T(g["pl_logits2"]), nms=1, connectivity=1).numpy() == g["pl_cc2_conn1"]).all()
    e = torch.zeros(1, 2, 8, 8, 8)
    e[:, 0] = 5
    assert (O.get_cut_mask(e, nms=1).numpy() == g["pl_cc_empty"]).all()
    assert (O.get_acdc_masks(T(g["acdc_pl_logits"]), nms=0).numpy() == g["acdc_pl_argmax"]).all()
    assert (O.get_acdc_masks(T(g["acdc_pl_logits"]), nms=1).numpy() == g["acdc_pl_cc"]).all()


def test_ema_vectors():
    g = load("functions")
    m1, m2 = O.OracleVNet(1, 2, 4, "batchnorm", True), O.OracleVNet(1, 2, 4, "batchnorm", True)
    O.fill_state_dict_(m1, 3)
    O.fill_state_dict_(m2, 4)
    O.update_ema_variables(m1, m2, 0.99)
    close(digest_named(m2.state_dict()), g["ema_la_digest"], rtol=1e-7, atol=0)
    u1, u2 = O.OracleUNet2d(1, 4), O.OracleUNet2d(1, 4)
    O.fill_state_dict_(u1, 5)
    O.fill_state_dict_(u2, 6)
    for k, v in u1.state_dict().items():
        if k.endswith("num_batches_tracked"):
            v.fill_(7)
    for k, v in u2.state_dict().items():
        if k.endswith("num_batches_tracked"):
            v.fill_(3)
    O.update_model_ema(u1, u2, 0.99)
    close(digest_named(u2.state_dict()), g["ema_acdc_digest"], rtol=1e-7, atol=0)
    assert int(u2.state_dict()["encoder.in_conv.conv_conv.1.num_batches_tracked"]) == int(g["ema_acdc_nbt"])


def test_network_vectors():
    g = load("networks")
    net = O.OracleVNet(1, 2, 16, "batchnorm", False)
    O.fill_state_dict_(net, 21)
    net.eval()
    with torch.no_grad():
        lo, feat = net(O.synthetic_volume((1, 1, 48, 48, 48), 22))
    close(lo, g["vnet_eval_logits"], rtol=1e-4, atol=1e-4)
    close(feat, g["vnet_eval_feat"], rtol=1e-4, atol=1e-4)

    net = O.net_factory("VNet", 1, 2, "train")
    O.fill_state_dict_(net, 23)
    net.train()
    inject_dropout(net, seed=24)
    lo, _ = net(O.synthetic_volume((2, 1, 48, 48, 48), 25))
    close(lo.detach(), g["vnet_train_logits"], rtol=1e-4, atol=1e-4)
    (lo * O.synthetic_volume(tuple(lo.shape), 26)).sum().backward()
    digests_close(digest_named({n: p.grad for n, p in net.named_parameters() if p.grad is not None}), g["vnet_train_grad_digest"], rtol=1e-3)
    close(net.encoder.block_one.conv[0].weight.grad, g["vnet_train_grad_first"], rtol=1e-3, atol=1e-3)

    un = O.OracleUNet2d(1, 4)
    O.fill_state_dict_(un, 31)
    un.eval()
    x = O.synthetic_volume((2, 1, 64, 48), 32, "rand")
    with torch.no_grad():
        close(un(x), g["unet_eval_logits"], rtol=1e-4, atol=1e-4)
    un.train()
    inject_dropout(un, seed=33)
    lo = un(x)
    close(lo.detach(), g["unet_train_logits"], rtol=1e-4, atol=1e-4)

    pn = O.OraclePanVNet()
    O.fill_state_dict_(pn, 41)
    pn.train()
    lo = pn(O.synthetic_volume((2, 1, 32, 16, 32), 42))[0]
    close(lo.detach(), g["pan_train_logits"], rtol=1e-4, atol=1e-4)


def _la_models(seed_w, seed_d1, seed_d2):
    model, ema = O.net_factory("VNet", 1, 2, "train"), O.net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(model, seed_w)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    inject_dropout(model, seed=seed_d1)
    inject_dropout(ema, seed=seed_d2)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    return model, ema, opt


def _check_la_step(g, nsteps, shape, sub):
    model, ema, opt = _la_models(51, 52, 53)
    rs = np.random.RandomState(int(g["box_seed"]))
    for it in range(nsteps):
        vol = O.synthetic_volume((8, 1) + shape, 60 + it)
        lab = O.synthetic_labels((8,) + shape, 70 + it)
        r = O.la_self_train_step(model, ema, opt, vol, lab, rng=rs)
        for k in ("loss", "loss_l", "loss_u"):
            close(r[k], g[f"s{it}_{k}"], rtol=2e-5)
        assert float(r["plab_a"].sum()) == float(g[f"s{it}_plab_a_sum"])
        close(r["out_l"][..., ::sub, ::sub, ::sub], g[f"s{it}_out_l"], rtol=1e-3, atol=1e-3)
        close(tensor_digest(r["mixl"]), g[f"s{it}_mixl_digest"], rtol=1e-9, atol=0)
        digests_close(digest_named(model.state_dict()), g[f"s{it}_model_digest"])
        digests_close(digest_named(ema.state_dict()), g[f"s{it}_ema_digest"])


def test_la_step_small():
    _check_la_step(load("la_step_small"), 2, (48, 48, 48), 2)


def test_la_pre_step():
    g = load("la_pre_step")
    model = O.net_factory("VNet", 1, 2, "train")
    O.fill_state_dict_(model, 81)
    model.train()
    inject_dropout(model, seed=82)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    shape = (48, 48, 48)
    r = O.la_pre_train_step(model, opt, O.synthetic_volume((4, 1) + shape, 83), O.synthetic_labels((4,) + shape, 84),
                            rng=np.random.RandomState(int(g["box_seed"])))
    close(r["loss"], g["loss"], rtol=2e-5)
    close(r["out"], g["out"], rtol=1e-3, atol=1e-3)
    digests_close(digest_named(model.state_dict()), g["model_digest"])


def test_acdc_step():
    g = load("acdc_step")
    model, ema = O.BCP_net(1, 4), O.BCP_net(1, 4, ema=True)
    O.fill_state_dict_(model, 91)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    inject_dropout(model, seed=92)
    inject_dropout(ema, seed=93)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    rs = np.random.RandomState(1337)
    for it in range(2):
        vol = O.synthetic_volume((8, 1, 64, 64), 100 + it, "rand")
        lab = O.synthetic_labels((8, 64, 64), 110 + it, n_classes=4).to(torch.uint8)
        r = O.acdc_self_train_step(model, ema, opt, vol, lab, labeled_bs=4, rng=rs)
        close(r["loss"], g[f"s{it}_loss"], rtol=2e-5)
        close(r["loss_dice"], g[f"s{it}_loss_dice"], rtol=2e-5)
        assert (r["plab_a"].numpy() == g[f"s{it}_plab_a"]).all()
        close(r["out_l"], g[f"s{it}_out_l"], rtol=1e-3, atol=1e-3)
        digests_close(digest_named(model.state_dict()), g[f"s{it}_model_digest"])
        digests_close(digest_named(ema.state_dict()), g[f"s{it}_ema_digest"])


@pytest.mark.slow
def test_la_step_full():
    if not os.path.exists(os.path.join(G, "la_step_full.npz")):
        pytest.skip("full-size fixture not generated")
    _check_la_step(load("la_step_full"), 1, (112, 112, 80), 4)


@pytest.mark.slow
def test_pan_step():
    if not os.path.exists(os.path.join(G, "pan_step.npz")):
        pytest.skip("fixture not generated")
    g = load("pan_step")
    net, ema = O.OraclePanVNet(), O.OraclePanVNet()
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(net, 121)
    ema.load_state_dict(net.state_dict())
    net.train()
    ema.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    rs = np.random.RandomState(2020)
    S = (96, 96, 96)
    v = O.synthetic_volume((8, 1) + S, 130)
    l = O.synthetic_labels((8,) + S, 140)
    r = O.pan_self_train_step(net, ema, opt, v[0:2], l[0:2], v[2:4], l[2:4], v[4:6], v[6:8], rng=rs)
    close(r["loss"], g["s0_loss"], rtol=2e-5)
    assert float(r["plab_a"].sum()) == float(g["s0_plab_a_sum"])
    close(r["out_1"][..., ::4, ::4, ::4], g["s0_out_1"], rtol=1e-3, atol=1e-3)
    digests_close(digest_named(net.state_dict()), g["s0_model_digest"])
    digests_close(digest_named(ema.state_dict()), g["s0_ema_digest"])


def test_acdc_pre_step():
    """ACDC pre-training step (ACDC_BCP_train.py:237-255) of the oracle against the reference-generated fixture."""
    g = load("acdc_pre_step")
    model = O.BCP_net(1, 4)
    O.fill_state_dict_(model, 151)
    model.train()
    inject_dropout(model, seed=152)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    vol = O.synthetic_volume((4, 1, 64, 64), 153, "rand")
    lab = O.synthetic_labels((4, 64, 64), 154, n_classes=4).to(torch.uint8)
    r = O.acdc_pre_train_step(model, opt, vol, lab, labeled_bs=4, rng=np.random.RandomState(int(g["seed"])))
    close(r["loss"], g["loss"], rtol=2e-5)
    close(r["loss_dice"], g["loss_dice"], rtol=2e-5)
    close(r["loss_ce"], g["loss_ce"], rtol=2e-5)
    close(r["out"], g["out"], rtol=1e-3, atol=1e-3)
    digests_close(digest_named(model.state_dict()), g["model_digest"])


@pytest.mark.slow
def test_pan_pre_step():
    """Pancreas pre-training step (pancreas/train_pancreas.py:82-99)."""
    g = load("pan_pre_step")
    net = O.OraclePanVNet()
    O.fill_state_dict_(net, 161)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    S = (96, 96, 96)
    v = O.synthetic_volume((2, 1) + S, 162)
    l = O.synthetic_labels((2,) + S, 163)
    r = O.pan_pre_train_step(net, opt, v[0:1], l[0:1], v[1:2], l[1:2], 64, rng=np.random.RandomState(int(g["seed"])))
    close(r["loss"], g["loss"], rtol=2e-5)
    close(r["loss_ce"], g["loss_ce"], rtol=2e-5)
    close(r["loss_dice"], g["loss_dice"], rtol=2e-5)
    close(r["out"][..., ::4, ::4, ::4], g["out"], rtol=1e-3, atol=1e-3)
    digests_close(digest_named(net.state_dict()), g["model_digest"], rtol=2e-3)


def test_sliding_window_validation():
    """utils/test_3d_patch.py:82-141 (test_single_case) -- clamped last window and pad/crop."""
    g = load("sliding_window")
    model = O.net_factory("VNet", 1, 2, "test")
    O.fill_state_dict_(model, 171)
    model.eval()
    for tag, seed in (("a", 172), ("b", 173)):
        shape = tuple(int(v) for v in g[tag + "_shape"])
        img = O.synthetic_volume(shape, seed).numpy()
        label, score = O.sliding_window_predict(model, img, 18, 4, (48, 48, 48), num_classes=2)
        assert label.shape == shape and score.shape == (2,) + shape
        close(score[0, ::2, ::2, ::2], g[tag + "_score"], rtol=1e-4, atol=1e-5)
        # voxels whose mean probability sits within float noise of the 0.5 threshold may flip; everything else is exact
        sure = np.abs(score[0] - 0.5) > 1e-4
        assert np.array_equal(label[sure], g[tag + "_label"][sure].astype(np.int64))
        assert (~sure).mean() < 1e-3


# ---- fixtures on the reference's shipped checkpoints (LA_10 / ACDC_10 rounded to bf16) ---------------------------------
def test_ckpt_weight_fixture_roundtrip():
    """weights_*_bf16.npz unpack to fp32 tensors that are exactly bf16-representable and load into the oracle nets."""
    from tests.golden.golden_common import unpack_weights_bf16
    sd = unpack_weights_bf16(load("weights_acdc10_bf16"))
    net = O.BCP_net(1, 4)
    net.load_state_dict(sd)
    w = sd["encoder.in_conv.conv_conv.0.weight"]
    assert torch.equal(w, w.to(torch.bfloat16).float())
    sd = unpack_weights_bf16(load("weights_la10_bf16"))
    O.net_factory("VNet", 1, 2, "train").load_state_dict(sd)
    assert len(sd) == 259


@pytest.mark.slow
def test_la_ckpt_step():
    """The oracle's restatement of LA_BCP_train.py:234-270 from the shipped LA_10 weights at BASELINE configs[1]."""
    from tests.golden.golden_common import unpack_weights_bf16, synthetic_scene, unpackbits
    g = load("la_ckpt_step")
    shape = tuple(int(v) for v in g["shape"])
    model, ema = O.net_factory("VNet", 1, 2, "train"), O.net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    sd = unpack_weights_bf16(load("weights_la10_bf16"))
    model.load_state_dict(sd)
    ema.load_state_dict(sd)
    model.train()
    ema.train()
    inject_dropout(model, seed=252)
    inject_dropout(ema, seed=253)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    rs = np.random.RandomState(int(g["box_seed"]))
    vol, lab = synthetic_scene(8, shape, 260)
    r = O.la_self_train_step(model, ema, opt, vol, lab, rng=rs)
    for k in ("loss", "loss_l", "loss_u"):
        close(r[k], g[f"s0_{k}"], rtol=2e-5)
    plab = torch.cat([r["plab_a"], r["plab_b"]]).numpy().astype(np.uint8)
    assert np.array_equal(plab, unpackbits(g["s0_plab"], plab.shape))
    close(tensor_digest(r["mixl"]), g["s0_mixl_digest"], rtol=1e-9, atol=0)
    digests_close(digest_named(model.state_dict()), g["s0_model_digest"])
    digests_close(digest_named(ema.state_dict()), g["s0_ema_digest"])


def test_acdc_ckpt_step():
    from tests.golden.golden_common import unpack_weights_bf16, synthetic_scene
    g = load("acdc_ckpt_step")
    H, W = (int(v) for v in g["shape"])
    B, labeled_bs = int(g["B"]), int(g["labeled_bs"])
    model, ema = O.BCP_net(1, 4), O.BCP_net(1, 4, ema=True)
    sd = unpack_weights_bf16(load("weights_acdc10_bf16"))
    model.load_state_dict(sd)
    ema.load_state_dict(sd)
    model.train()
    ema.train()
    inject_dropout(model, seed=292)
    inject_dropout(ema, seed=293)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    rs = np.random.RandomState(1337)
    vol, lab = synthetic_scene(B, (H, W), 300, n_classes=4, kind="rand")
    r = O.acdc_self_train_step(model, ema, opt, vol, lab.to(torch.uint8), labeled_bs=labeled_bs, rng=rs)
    for k in ("loss", "loss_dice", "loss_ce"):
        close(r[k], g[f"s0_{k}"], rtol=2e-5)
    assert np.array_equal(torch.cat([r["plab_a"], r["plab_b"]]).numpy().astype(np.uint8), g["s0_plab"])
    digests_close(digest_named(model.state_dict()), g["s0_model_digest"])
    digests_close(digest_named(ema.state_dict()), g["s0_ema_digest"])


def test_dataset_pipeline():
    """oracle/dataset_oracle.py against tests/golden/dataset.npz (the reference's LAHeart + RandomRotFlip + RandomCrop +
    ToTensor + TwoStreamBatchSampler under a single-process DataLoader): same np.random stream, bit-identical batches."""
    from oracle import dataset_oracle as D
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset.npz"))
    vols = D.synthetic_la_volumes(int(g["nvol"]), 4242)
    patch = tuple(int(v) for v in g["patch"])
    lab_n, bs, lbs = int(g["labeled"]), int(g["batch_size"]), int(g["labeled_bs"])
    np.random.seed(7)
    idx = [list(t) for _ in range(int(g["epochs"])) for t in D.two_stream_batches(list(range(lab_n)), list(range(lab_n, len(vols))), bs, bs - lbs)]
    assert np.array_equal(np.array(idx), g["sampler_indices"])
    np.random.seed(int(g["seed"]))
    b = 0
    for _ in range(int(g["epochs"])):
        for indices in D.two_stream_batches(list(range(lab_n)), list(range(lab_n, len(vols))), bs, bs - lbs):
            im, lb = D.la_batch(vols, indices, patch)
            assert np.array_equal(im, g[f"b{b}_image"]) and np.array_equal(lb.astype(np.uint8), g[f"b{b}_label"])
            b += 1
    assert b == int(g["nbatches"])
    assert any(v[0].shape[0] <= patch[0] or v[0].shape[1] <= patch[1] or v[0].shape[2] <= patch[2] for v in vols), "padding branch not exercised"


def test_dataset_host_draws_match_oracle():
    """bcp_b200.dataloaders.dataset (host side: sampler + parameter draws, no kernels) consumes np.random exactly like the
    reference pipeline: same index tuples, and the drawn (k, axis, pad, origin) reproduce the golden batches when applied
    with numpy."""
    from bcp_b200.dataloaders import dataset as P
    from oracle import dataset_oracle as D
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset.npz"))
    vols = D.synthetic_la_volumes(int(g["nvol"]), 4242)
    patch = tuple(int(v) for v in g["patch"])
    lab_n, bs, lbs = int(g["labeled"]), int(g["batch_size"]), int(g["labeled_bs"])
    sampler = P.TwoStreamBatchSampler(list(range(lab_n)), list(range(lab_n, len(vols))), bs, bs - lbs)
    assert len(sampler) == lab_n // lbs
    np.random.seed(7)
    idx = [[int(v) for v in t] for _ in range(int(g["epochs"])) for t in sampler]
    assert np.array_equal(np.array(idx), g["sampler_indices"])
    np.random.seed(int(g["seed"]))
    b = 0
    for _ in range(int(g["epochs"])):
        for indices in sampler:
            for j, i in enumerate(indices):
                im, lb = vols[int(i)]
                p = P.draw_rotflip_crop(im.shape, patch)
                r = np.flip(np.rot90(im, p["k"]), axis=p["axis"])
                pw, ph, pd = p["pad"]
                r = np.pad(r, [(pw, pw), (ph, ph), (pd, pd)], mode="constant")
                w1, h1, d1 = p["origin"]
                assert np.array_equal(r[w1:w1 + patch[0], h1:h1 + patch[1], d1:d1 + patch[2]], g[f"b{b}_image"][j, 0])
            b += 1


def test_metric_closed_forms():
    """oracle/metrics_oracle.py (medpy restatement, parity unpinned) on cases with known answers, and the product's own
    host-side metrics (bcp_b200/utils/val_2d.py) against it on random masks."""
    from oracle import metrics_oracle as M
    a = np.zeros((20, 20), bool)
    b = np.zeros((20, 20), bool)
    a[5:10, 5:10] = True
    b[5:10, 8:13] = True                       # same square shifted by 3 columns: overlap 5x2
    assert abs(M.dc(a, b) - 2 * 10 / 50) < 1e-12
    assert M.dc(a, a) == 1.0 and M.dc(np.zeros(4, bool), np.zeros(4, bool)) == 0.0
    assert abs(M.hd95(a, b) - 3.0) < 1e-9      # every border pixel is <= 3 away, the far edges exactly 3
    assert M.hd95(a, a) == 0.0
    from bcp_b200.utils import val_2d as V
    rs = np.random.RandomState(0)
    for _ in range(5):
        p, q = rs.random_sample((12, 30, 28)) > 0.6, rs.random_sample((12, 30, 28)) > 0.5
        assert V.dc(p, q) == M.dc(p, q) and V.hd95(p, q) == M.hd95(p, q)
        assert V.calculate_metric_percase(p.astype(np.int64) * 3, q.astype(np.int64)) == (M.dc(p, q), M.hd95(p, q))
    assert V.calculate_metric_percase(np.zeros((3, 3)), np.ones((3, 3))) == (0, 0)


@pytest.mark.parametrize("prefetch,workers", [(False, 0), (True, 0), (True, 2)])
def test_acdc_pipeline_matches_reference_batches(prefetch, workers):
    """bcp_b200.dataloaders.dataset BaseDataSets + RandomGenerator + TwoStreamBatchSampler + SliceLoader against
    tests/golden/acdc_dataset.npz: the batches the reference's own classes (dataloaders/dataset.py:15-88,280-307) produced
    under a single-process DataLoader with the same ``np.random`` / ``random`` seeds.  Host-side mirror (scipy nearest-neighbour
    rotate / zoom, as the reference's workers run it): bit-exact, labels uint8, with and without the prefetch thread and with the
    resampling farmed out to worker processes (draws stay on one thread, in the reference's order)."""
    import random
    from bcp_b200.dataloaders import dataset as P
    from oracle import dataset_oracle as D
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "acdc_dataset.npz"))
    slices = D.synthetic_acdc_slices(int(g["nslices"]), int(g["slice_seed"]))
    patch = tuple(int(v) for v in g["patch"])
    db = P.BaseDataSets(split="train", transform=P.RandomGenerator(patch), slices=slices)
    nl, bs, lbs = int(g["labeled"]), int(g["batch_size"]), int(g["labeled_bs"])
    sampler = P.TwoStreamBatchSampler(list(range(nl)), list(range(nl, len(slices))), bs, bs - lbs)
    loader = P.SliceLoader(db, sampler, pin=False, prefetch=prefetch, workers=workers)
    np.random.seed(int(g["np_seed"]))
    random.seed(int(g["py_seed"]))
    b = 0
    for _ in range(int(g["epochs"])):
        for batch in loader:
            assert batch["image"].dtype == torch.float32 and batch["label"].dtype == torch.uint8
            assert batch["image"].shape == (bs, 1) + patch and batch["label"].shape == (bs,) + patch
            assert np.array_equal(batch["image"].numpy(), g[f"b{b}_image"]), b
            assert np.array_equal(batch["label"].numpy(), g[f"b{b}_label"]), b
            b += 1
    loader.close()
    assert b == int(g["nbatches"]) and b >= 6
    assert P.patients_to_slices("ACDC", 7) == 136 and P.patients_to_slices("/data/ACDC", 140) == 1312
