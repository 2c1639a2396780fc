"""Golden-vector generator: runs the UNMODIFIED reference (/root/reference/code) on seeded inputs.

Run once in the build container (``python tests/golden/make_golden.py``); outputs ``*.npz`` next
to this file.  The reference's own functions are used wherever they can be imported
(networks.*, utils.BCP_utils, utils.losses, pancreas/Vnet.py, pancreas/losses.py) and the
script-local helpers of LA_BCP_train.py / ACDC_BCP_train.py / pancreas_utils.py are exec'd
verbatim from their source (oracle/ref_shims.extract_defs).  Only the *loop bodies*
(LA_BCP_train.py:234-270, ACDC_BCP_train.py:354-390, train_pancreas.py:144-174) are re-typed
here because they live inside functions that build data loaders.

Weights/inputs come from numpy RandomState streams (oracle.bcp_oracle.fill_state_dict_ etc.) so
the GPU box can regenerate them without the reference tree.  Dropout is made reproducible by
swapping the reference model's Dropout modules for ``InjectedDropout`` (same math, mask from a
numpy stream) -- the product accepts the same mask provider.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_shims  # noqa: E402
from oracle import bcp_oracle as O  # noqa: E402
from tests.golden.golden_common import (InjectedDropout, inject_dropout, tensor_digest, digest_named,  # noqa: E402
                                        find_box_seed_la, pack_weights_bf16, unpack_weights_bf16, synthetic_scene, packbits)

torch.set_num_threads(os.cpu_count() or 8)
R = ref_shims.load()
LA = ref_shims.extract_defs(os.path.join(ref_shims.REF_CODE, "LA_BCP_train.py"),
                            ["get_cut_mask", "LargestCC_pancreas"])[0]
ACDC_ns, acdc_glb = ref_shims.extract_defs(
    os.path.join(ref_shims.REF_CODE, "ACDC_BCP_train.py"),
    ["get_ACDC_2DLargestCC", "get_ACDC_masks", "update_model_ema", "generate_mask", "mix_loss"])
acdc_glb["dice_loss"] = R.losses.DiceLoss(n_classes=4)          # ACDC_BCP_train.py:57
PANU, panu_glb = ref_shims.extract_defs(os.path.join(ref_shims.REF_CODE, "pancreas", "pancreas_utils.py"),
                                        ["generate_mask", "get_cut_mask", "LargestCC_pancreas", "update_ema_variables"])


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


# ------------------------------------------------------------------ function-level vectors
def gen_functions():
    out = {}
    rs = np.random.RandomState(11)
    # LA masked dice / mix_loss (utils/losses.py:47-77, utils/BCP_utils.py:58-69)
    logits = torch.from_numpy(rs.standard_normal((2, 2, 12, 10, 8)).astype(np.float32)) * 3
    la = torch.from_numpy((rs.random_sample((2, 12, 10, 8)) > 0.7).astype(np.int64))
    lb = torch.from_numpy((rs.random_sample((2, 12, 10, 8)) > 0.5).astype(np.float32))   # plab is float32 in ref
    m = torch.ones(2, 12, 10, 8, dtype=torch.int64)
    m[:, 2:9, 3:8, 1:6] = 0
    out.update(la_logits=logits, la_lab_a=la, la_lab_b=lb, la_mask=m)
    dice = R.losses.mask_DiceLoss(nclass=2)
    out["la_dice_masked"] = dice(logits, la, m)
    out["la_dice_unmasked"] = dice(logits, la)
    lg = logits.clone().requires_grad_(True)
    l1 = R.BCP_utils.mix_loss(lg, la, lb, m, u_weight=0.5)
    l1.backward()
    out["la_mix_loss"], out["la_mix_grad"] = l1, lg.grad.clone()
    lg = logits.clone().requires_grad_(True)
    l2 = R.BCP_utils.mix_loss(lg, lb, la, m, u_weight=0.5, unlab=True)
    l2.backward()
    out["la_mix_loss_unlab"], out["la_mix_grad_unlab"] = l2, lg.grad.clone()
    lg = logits.clone().requires_grad_(True)
    l3 = R.pan_losses.mix_loss(lg, la, lb.long(), m)
    out["pan_mix_loss"] = l3
    # pre-train loss LA_BCP_train.py:159-161
    lg = logits.clone().requires_grad_(True)
    lp = (F.cross_entropy(lg, la) + dice(lg, la)) / 2
    lp.backward()
    out["la_pre_loss"], out["la_pre_grad"] = lp, lg.grad.clone()

    # ACDC dice / mix_loss (utils/losses.py:102-134, ACDC_BCP_train.py:167-179)
    lg4 = torch.from_numpy(rs.standard_normal((3, 4, 16, 12)).astype(np.float32)) * 2
    ta = torch.from_numpy(rs.randint(0, 4, (3, 16, 12)).astype(np.uint8))
    tb = torch.from_numpy(rs.randint(0, 4, (3, 16, 12)).astype(np.float32))
    m2 = torch.ones(3, 16, 12, dtype=torch.int64)
    m2[:, 3:13, 2:10] = 0
    out.update(acdc_logits=lg4, acdc_lab_a=ta, acdc_lab_b=tb, acdc_mask=m2)
    x = lg4.clone().requires_grad_(True)
    d, c = ACDC_ns.mix_loss(x, ta, tb, m2, u_weight=0.5)
    ((d + c) / 2).backward()
    out["acdc_mix_dice"], out["acdc_mix_ce"], out["acdc_mix_grad"] = d, c, x.grad.clone()
    x = lg4.clone().requires_grad_(True)
    d, c = ACDC_ns.mix_loss(x, tb, ta, m2, u_weight=0.5, unlab=True)
    out["acdc_mix_dice_unlab"], out["acdc_mix_ce_unlab"] = d, c

    # masks: utils/BCP_utils.py:18-28, ACDC_BCP_train.py:131-140, pancreas_utils.py:187-200
    np.random.seed(1337)
    mk, lm = R.BCP_utils.context_mask(torch.zeros(2, 1, 112, 112, 80), 2 / 3)
    out["la_ctx_mask_zero_bbox"] = np.array([[int(i.min()), int(i.max()) + 1] for i in torch.nonzero(mk == 0, as_tuple=True)])
    out["la_ctx_mask_sum"], out["la_ctx_lmask_sum"] = mk.sum(), lm.sum()
    np.random.seed(1337)
    mk, lm = ACDC_ns.generate_mask(torch.zeros(6, 1, 256, 256))
    out["acdc_mask_zero_bbox"] = np.array([[int(i.min()), int(i.max()) + 1] for i in torch.nonzero(mk == 0, as_tuple=True)])
    out["acdc_mask_sum"] = mk.sum()
    np.random.seed(2020)
    mk, lm = PANU.generate_mask(torch.zeros(2, 1, 96, 96, 96), 64)
    out["pan_mask_zero_bbox"] = np.array([[int(i.min()), int(i.max()) + 1] for i in torch.nonzero(mk == 0, as_tuple=True)])

    # mask-mix (LA_BCP_train.py:248-249) incl. special values: -0.0, inf, nan
    a = torch.from_numpy(rs.standard_normal((2, 1, 12, 10, 8)).astype(np.float32))
    b = torch.from_numpy(rs.standard_normal((2, 1, 12, 10, 8)).astype(np.float32))
    a.view(-1)[:6] = torch.tensor([-0.0, 0.0, float("inf"), float("-inf"), float("nan"), -1.5])
    b.view(-1)[:6] = torch.tensor([0.0, -0.0, 1.0, float("inf"), 2.0, float("nan")])
    a.view(-1)[-6:] = torch.tensor([-0.0, 0.0, float("inf"), float("-inf"), float("nan"), -1.5])
    b.view(-1)[-6:] = torch.tensor([0.0, -0.0, 1.0, float("inf"), 2.0, float("nan")])
    mm = torch.ones(12, 10, 8, dtype=torch.int64)
    mm[0:7, 0:5, 0:6] = 0
    out.update(mix_a=a, mix_b=b, mix_mask=mm, mix_out=a * mm + b * (1 - mm))

    # pseudo labels: LA_BCP_train.py:57-77 ; near-tie logits exercise the fp32 softmax>=0.5 rule
    pl = torch.from_numpy(rs.standard_normal((2, 2, 16, 14, 12)).astype(np.float32))
    d = torch.from_numpy(rs.standard_normal((2, 16, 14, 12)).astype(np.float32))
    pl[:, 1] = pl[:, 0] + d * torch.tensor([1e-8, 3e-8, 6e-8, 1e-7]).repeat(3)[None, None, None, :]
    pl[0, :, :4] *= 30.0
    sm = O.synthetic_labels((2, 16, 14, 12), 5).float() * 4 - 2
    pl2 = torch.stack([-sm, sm], 1) + 0.3 * torch.from_numpy(rs.standard_normal((2, 2, 16, 14, 12)).astype(np.float32))
    out.update(pl_logits=pl, pl_cut=LA.get_cut_mask(pl, nms=0), pl_logits2=pl2,
               pl_cut2=LA.get_cut_mask(pl2, nms=0), pl_cc2=LA.get_cut_mask(pl2, nms=1),
               pl_cc2_conn2=PANU.get_cut_mask(pl2, nms=True, connect_mode=2),
               pl_cc2_conn1=PANU.get_cut_mask(pl2, nms=True, connect_mode=1))
    empty = torch.full((1, 2, 8, 8, 8), 0.0)
    empty[:, 0] = 5
    out["pl_cc_empty"] = LA.get_cut_mask(empty, nms=1)
    # ACDC argmax + per-class 2-D CC: ACDC_BCP_train.py:89-117
    a4 = O.synthetic_labels((3, 32, 28), 9, n_classes=4)
    lg = F.one_hot(a4, 4).permute(0, 3, 1, 2).float() * 2 + 0.8 * torch.from_numpy(rs.standard_normal((3, 4, 32, 28)).astype(np.float32))
    lg[0, 1, :4] = lg[0, 2, :4]          # exact ties -> lowest index
    out.update(acdc_pl_logits=lg, acdc_pl_argmax=ACDC_ns.get_ACDC_masks(lg, nms=0), acdc_pl_cc=ACDC_ns.get_ACDC_masks(lg, nms=1))

    # EMA: utils/BCP_utils.py:78-81, ACDC_BCP_train.py:123-129
    m1, m2_ = R.VNet.VNet(1, 2, 4, "batchnorm", True), R.VNet.VNet(1, 2, 4, "batchnorm", True)
    O.fill_state_dict_(m1, 3)
    O.fill_state_dict_(m2_, 4)
    R.BCP_utils.update_ema_variables(m1, m2_, 0.99)
    out["ema_la_digest"] = digest_named(m2_.state_dict())
    u1, u2 = R.unet.UNet_2d(1, 4), R.unet.UNet_2d(1, 4)
    O.fill_state_dict_(u1, 5)
    O.fill_state_dict_(u2, 6)
    for k, v in u1.state_dict().items():
        if k.endswith("num_batches_tracked"):
            v.fill_(7)
    for k, v in u2.state_dict().items():
        if k.endswith("num_batches_tracked"):
            v.fill_(3)
    ACDC_ns.update_model_ema(u1, u2, 0.99)
    out["ema_acdc_digest"] = digest_named(u2.state_dict())
    out["ema_acdc_nbt"] = u2.state_dict()["encoder.in_conv.conv_conv.1.num_batches_tracked"]
    save("functions", **out)


# ------------------------------------------------------------------ network-level vectors
def gen_networks():
    out = {}
    # eval-mode (running-stat BN, no dropout): deterministic logits
    net = R.VNet.VNet(1, 2, 16, "batchnorm", False)
    O.fill_state_dict_(net, 21)
    net.eval()
    x = O.synthetic_volume((1, 1, 48, 48, 48), 22)
    with torch.no_grad():
        lo, feat = net(x)
    out["vnet_eval_logits"], out["vnet_eval_feat"] = lo, feat
    # train-mode, dropout injected, fwd+bwd
    net = R.net_factory.net_factory("VNet", 1, 2, "train")
    O.fill_state_dict_(net, 23)
    net.train()
    inject_dropout(net, seed=24)
    x = O.synthetic_volume((2, 1, 48, 48, 48), 25)
    lo, feat = net(x)
    out["vnet_train_logits"] = lo
    w = O.synthetic_volume(tuple(lo.shape), 26)
    (lo * w).sum().backward()
    out["vnet_train_grad_digest"] = digest_named({n: p.grad for n, p in net.named_parameters() if p.grad is not None})
    out["vnet_train_grad_first"] = net.encoder.block_one.conv[0].weight.grad
    out["vnet_train_grad_mid"] = net.encoder.block_three.conv[3].weight.grad[:8, :8]
    out["vnet_train_bn_state"] = digest_named({k: v for k, v in net.state_dict().items() if "running" in k})

    un = R.unet.UNet_2d(1, 4)
    O.fill_state_dict_(un, 31)
    un.eval()
    x = O.synthetic_volume((2, 1, 64, 48), 32, "rand")
    with torch.no_grad():
        out["unet_eval_logits"] = un(x)
    un.train()
    inject_dropout(un, seed=33)
    lo = un(x)
    out["unet_train_logits"] = lo
    w = O.synthetic_volume(tuple(lo.shape), 34)
    (lo * w).sum().backward()
    out["unet_train_grad_digest"] = digest_named({n: p.grad for n, p in un.named_parameters() if p.grad is not None})

    pn = R.pan_Vnet.VNet()
    O.fill_state_dict_(pn, 41)
    pn.train()
    x = O.synthetic_volume((2, 1, 32, 16, 32), 42)
    lo = pn(x)[0]
    out["pan_train_logits"] = lo
    w = O.synthetic_volume(tuple(lo.shape), 43)
    (lo * w).sum().backward()
    out["pan_train_grad_digest"] = digest_named({n: p.grad for n, p in pn.named_parameters() if p.grad is not None})
    save("networks", **out)


# ------------------------------------------------------------------ step-level vectors
def la_step_reference(model, ema, opt, volume, label, labeled_bs=4, mask_ratio=2 / 3, u_weight=0.5):
    """LA_BCP_train.py:234-270 with the reference's own functions."""
    sub_bs = labeled_bs // 2
    img_a, img_b = volume[:sub_bs], volume[sub_bs:labeled_bs]
    lab_a, lab_b = label[:sub_bs], label[sub_bs:labeled_bs]
    unimg_a, unimg_b = volume[labeled_bs:labeled_bs + sub_bs], volume[labeled_bs + sub_bs:]
    with torch.no_grad():
        unoutput_a, _ = ema(unimg_a)
        unoutput_b, _ = ema(unimg_b)
        plab_a = LA.get_cut_mask(unoutput_a, nms=1)
        plab_b = LA.get_cut_mask(unoutput_b, nms=1)
        img_mask, loss_mask = R.BCP_utils.context_mask(img_a, mask_ratio)
    mixl_img = img_a * img_mask + unimg_a * (1 - img_mask)
    mixu_img = unimg_b * img_mask + img_b * (1 - img_mask)
    outputs_l, _ = model(mixl_img)
    outputs_u, _ = model(mixu_img)
    loss_l = R.BCP_utils.mix_loss(outputs_l, lab_a, plab_a, loss_mask, u_weight=u_weight)
    loss_u = R.BCP_utils.mix_loss(outputs_u, plab_b, lab_b, loss_mask, u_weight=u_weight, unlab=True)
    loss = loss_l + loss_u
    opt.zero_grad()
    loss.backward()
    opt.step()
    R.BCP_utils.update_ema_variables(model, ema, 0.99)
    return dict(loss=loss.detach(), loss_l=loss_l.detach(), loss_u=loss_u.detach(), plab_a=plab_a, plab_b=plab_b,
                out_l=outputs_l.detach(), out_u=outputs_u.detach(), mixl=mixl_img, mixu=mixu_img)


def gen_la_step(tag, shape, nsteps, box_seed, sub=1):
    t0 = time.time()
    model = R.net_factory.net_factory("VNet", 1, 2, "train")
    ema = R.net_factory.net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(model, 51)
    ema.load_state_dict(model.state_dict())           # both start from the pre-trained net (:220-222)
    model.train()
    ema.train()
    inject_dropout(model, seed=52)
    inject_dropout(ema, seed=53)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    out = dict(shape=np.array(shape), box_seed=box_seed, nsteps=nsteps, sub=sub)
    np.random.seed(box_seed)
    for it in range(nsteps):
        vol = O.synthetic_volume((8, 1) + tuple(shape), 60 + it)
        lab = O.synthetic_labels((8,) + tuple(shape), 70 + it)
        r = la_step_reference(model, ema, opt, vol, lab)
        for k in ("loss", "loss_l", "loss_u"):
            out[f"s{it}_{k}"] = r[k]
        out[f"s{it}_plab_a_sum"], out[f"s{it}_plab_b_sum"] = r["plab_a"].sum(), r["plab_b"].sum()
        out[f"s{it}_plab"] = torch.cat([r["plab_a"], r["plab_b"]]).to(torch.uint8)
        out[f"s{it}_out_l"] = r["out_l"][..., ::sub, ::sub, ::sub]
        out[f"s{it}_out_u"] = r["out_u"][..., ::sub, ::sub, ::sub]
        out[f"s{it}_mixl_digest"] = tensor_digest(r["mixl"])
        out[f"s{it}_grad_digest"] = digest_named({n: p.grad for n, p in model.named_parameters() if p.grad is not None})
        out[f"s{it}_model_digest"] = digest_named(model.state_dict())
        out[f"s{it}_ema_digest"] = digest_named(ema.state_dict())
        print(tag, "step", it, "loss", float(r["loss"]), "t=%.1fs" % (time.time() - t0))
    save(tag, **out)


def gen_la_pre_step():
    model = R.net_factory.net_factory("VNet", 1, 2, "train")
    O.fill_state_dict_(model, 81)
    model.train()
    inject_dropout(model, seed=82)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    DICE = R.losses.mask_DiceLoss(nclass=2)
    shape = (48, 48, 48)
    bs = find_box_seed_la(shape)
    np.random.seed(bs)
    vol = O.synthetic_volume((4, 1) + shape, 83)
    lab = O.synthetic_labels((4,) + shape, 84)
    img_a, img_b, lab_a, lab_b = vol[:2], vol[2:], lab[:2], lab[2:]
    with torch.no_grad():
        img_mask, loss_mask = R.BCP_utils.context_mask(img_a, 2 / 3)
    volume_batch = img_a * img_mask + img_b * (1 - img_mask)
    label_batch = lab_a * img_mask + lab_b * (1 - img_mask)
    outputs, _ = model(volume_batch)
    loss_ce = F.cross_entropy(outputs, label_batch)
    loss_dice = DICE(outputs, label_batch)
    loss = (loss_ce + loss_dice) / 2
    opt.zero_grad()
    loss.backward()
    opt.step()
    save("la_pre_step", box_seed=bs, loss=loss.detach(), loss_ce=loss_ce.detach(), loss_dice=loss_dice.detach(),
         out=outputs.detach(), model_digest=digest_named(model.state_dict()))


def gen_acdc_step():
    model = R.net_factory.BCP_net(1, 4)
    ema = R.net_factory.BCP_net(1, 4, ema=True)
    O.fill_state_dict_(model, 91)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    inject_dropout(model, seed=92)
    inject_dropout(ema, seed=93)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    H, W, B, labeled_bs = 64, 64, 8, 4
    out = dict(shape=np.array([H, W]), B=B, labeled_bs=labeled_bs)
    np.random.seed(1337)
    for it in range(2):
        volume_batch = O.synthetic_volume((B, 1, H, W), 100 + it, "rand")
        label_batch = O.synthetic_labels((B, H, W), 110 + it, n_classes=4).to(torch.uint8)
        ls, us = labeled_bs // 2, (B - labeled_bs) // 2
        img_a, img_b = volume_batch[:ls], volume_batch[ls:labeled_bs]
        uimg_a, uimg_b = volume_batch[labeled_bs:labeled_bs + us], volume_batch[labeled_bs + us:]
        lab_a, lab_b = label_batch[:ls], label_batch[ls:labeled_bs]
        with torch.no_grad():
            pre_a, pre_b = ema(uimg_a), ema(uimg_b)
            plab_a = ACDC_ns.get_ACDC_masks(pre_a, nms=1)
            plab_b = ACDC_ns.get_ACDC_masks(pre_b, nms=1)
            img_mask, loss_mask = ACDC_ns.generate_mask(img_a)
        net_input_unl = uimg_a * img_mask + img_a * (1 - img_mask)
        net_input_l = img_b * img_mask + uimg_b * (1 - img_mask)
        out_unl, out_l = model(net_input_unl), model(net_input_l)
        unl_dice, unl_ce = ACDC_ns.mix_loss(out_unl, plab_a, lab_a, loss_mask, u_weight=0.5, unlab=True)
        l_dice, l_ce = ACDC_ns.mix_loss(out_l, lab_b, plab_b, loss_mask, u_weight=0.5)
        loss_ce, loss_dice = unl_ce + l_ce, unl_dice + l_dice
        loss = (loss_dice + loss_ce) / 2
        opt.zero_grad()
        loss.backward()
        opt.step()
        ACDC_ns.update_model_ema(model, ema, 0.99)
        out.update({f"s{it}_loss": loss.detach(), f"s{it}_loss_dice": loss_dice.detach(), f"s{it}_loss_ce": loss_ce.detach(),
                    f"s{it}_plab_a": plab_a, f"s{it}_plab": torch.cat([plab_a, plab_b]).to(torch.uint8), f"s{it}_out_unl": out_unl.detach(), f"s{it}_out_l": out_l.detach(),
                    f"s{it}_grad_digest": digest_named({n: p.grad for n, p in model.named_parameters() if p.grad is not None}),
                    f"s{it}_model_digest": digest_named(model.state_dict()), f"s{it}_ema_digest": digest_named(ema.state_dict())})
        print("acdc step", it, float(loss))
    save("acdc_step", **out)


def gen_acdc_pre_step():
    """ACDC_BCP_train.py:237-255 with the reference's own modules and script-local helpers."""
    model = R.net_factory.BCP_net(1, 4)
    O.fill_state_dict_(model, 151)
    model.train()
    inject_dropout(model, seed=152)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    H, W, labeled_bs = 64, 64, 4
    np.random.seed(4242)
    volume_batch = O.synthetic_volume((labeled_bs, 1, H, W), 153, "rand")
    label_batch = O.synthetic_labels((labeled_bs, H, W), 154, n_classes=4).to(torch.uint8)
    sub = labeled_bs // 2
    img_a, img_b = volume_batch[:sub], volume_batch[sub:labeled_bs]
    lab_a, lab_b = label_batch[:sub], label_batch[sub:labeled_bs]
    img_mask, loss_mask = ACDC_ns.generate_mask(img_a)
    net_input = img_a * img_mask + img_b * (1 - img_mask)
    out_mixl = model(net_input)
    loss_dice, loss_ce = ACDC_ns.mix_loss(out_mixl, lab_a, lab_b, loss_mask, u_weight=1.0, unlab=True)
    loss = (loss_dice + loss_ce) / 2
    opt.zero_grad()
    loss.backward()
    opt.step()
    save("acdc_pre_step", shape=np.array([H, W]), labeled_bs=labeled_bs, seed=4242, loss=loss.detach(), loss_dice=loss_dice.detach(),
         loss_ce=loss_ce.detach(), out=out_mixl.detach(), net_input=net_input,
         grad_digest=digest_named({n: p.grad for n, p in model.named_parameters() if p.grad is not None}),
         model_digest=digest_named(model.state_dict()))
    print("acdc pre step", float(loss))


def gen_pan_pre_step():
    """pancreas/train_pancreas.py:82-99."""
    net = R.pan_Vnet.VNet()
    O.fill_state_dict_(net, 161)
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    DICE = R.pan_losses.DiceLoss(nclass=2)
    np.random.seed(777)
    S = (96, 96, 96)
    v = O.synthetic_volume((2, 1) + S, 162)
    l = O.synthetic_labels((2,) + S, 163)
    img_a, img_b, lab_a, lab_b = v[0:1], v[1:2], l[0:1], l[1:2]
    img_mask, loss_mask = PANU.generate_mask(img_a, 64)
    img = img_a * img_mask + img_b * (1 - img_mask)
    lab = lab_a * img_mask + lab_b * (1 - img_mask)
    out = net(img)[0]
    ce = F.cross_entropy(out, lab)
    dice = DICE(out, lab)
    loss = (ce + dice) / 2
    opt.zero_grad()
    loss.backward()
    opt.step()
    save("pan_pre_step", seed=777, loss=loss.detach(), loss_ce=ce.detach(), loss_dice=dice.detach(), out=out.detach()[..., ::4, ::4, ::4],
         grad_digest=digest_named({n: p.grad for n, p in net.named_parameters() if p.grad is not None}),
         model_digest=digest_named(net.state_dict()))
    print("pan pre step", float(loss))


def gen_sliding_window():
    """utils/test_3d_patch.py:82-141 (test_single_case), function source taken verbatim from the reference file; the two
    cases exercise the clamped last window and the pad-then-crop branch."""
    import math
    ns, glb = ref_shims.extract_defs(os.path.join(ref_shims.REF_CODE, "utils", "test_3d_patch.py"), ["test_single_case"])
    glb["math"] = math
    had = hasattr(np, "int")
    if not had:
        np.int = int                       # the reference predates numpy 1.24 (uses np.int); restored below
    try:
        model = R.net_factory.net_factory("VNet", 1, 2, "test")
        O.fill_state_dict_(model, 171)
        model.eval()
        out = {}
        for tag, shape in (("a", (60, 56, 52)), ("b", (40, 50, 48))):
            img = O.synthetic_volume(shape, 172 if tag == "a" else 173).numpy()
            label, score = ns.test_single_case(model, img, 18, 4, (48, 48, 48), num_classes=2)
            out[tag + "_label"] = label.astype(np.uint8)
            out[tag + "_score"] = score.astype(np.float32)[0, ::2, ::2, ::2]
            out[tag + "_shape"] = np.array(shape)
            print("sliding window", tag, shape, "positive fraction", float(label.mean()))
    finally:
        if not had:
            del np.int
    save("sliding_window", **out)


def gen_val_2d():
    """utils/val_2d.py:10-41 (calculate_metric_percase, test_single_volume), sources taken verbatim from the reference file,
    run on the shipped ACDC_10 weights; ``medpy`` (absent) is stubbed by oracle/metrics_oracle.py."""
    import types
    from scipy.ndimage import zoom
    from oracle import metrics_oracle as M
    metric = types.SimpleNamespace(binary=types.SimpleNamespace(dc=M.dc, hd95=M.hd95))
    ns, glb = ref_shims.extract_defs(os.path.join(ref_shims.REF_CODE, "utils", "val_2d.py"), ["calculate_metric_percase", "test_single_volume"])
    glb.update(metric=metric, zoom=zoom)
    model = R.net_factory.BCP_net(1, 4)
    model.load_state_dict(unpack_weights_bf16(np.load(os.path.join(HERE, "weights_acdc10_bf16.npz"))))
    # the shipped running statistics describe real MR slices; on the synthetic scenes eval mode predicts background only.
    # Re-estimate the BatchNorm buffers on synthetic scenes (train-mode forwards, no weight update) and ship them.
    model.train()
    with torch.no_grad():
        for i in range(24):
            model(synthetic_scene(12, (256, 256), 710 + i, n_classes=4, kind="rand")[0])
    model.eval()
    vol, lab = synthetic_scene(6, (200, 180), 700, n_classes=4, kind="rand")          # [6,1,200,180], [6,200,180]
    image, label = vol[:, 0].unsqueeze(0), lab.unsqueeze(0)                             # DataLoader batch of one volume
    res = ns.test_single_volume(image, label, model, classes=4)
    out = dict(metrics=np.array(res, dtype=np.float64), depth=6, hw=np.array([200, 180]), seed=700)
    for k, v in model.state_dict().items():
        if "running_" in k or "num_batches" in k:
            out["buf." + k] = v.numpy()
    print("val_2d metrics", res)
    save("val_2d", **out)


# ------------------------------------------------------------------ fixtures on the SHIPPED checkpoints
CKPT = {"la10": os.path.join(ref_shims.REF_ROOT, "models", "LA", "LA_10.pth"),
        "acdc10": os.path.join(ref_shims.REF_ROOT, "models", "ACDC", "ACDC_10.pth")}
SURE_MARGIN = 1.0         # |logit margin| above which a pseudo label is called "sure": ~50x the RMS bf16 noise of the teacher's
                          # logits (0.02), whose error distribution has heavy tails (13 of 4M voxels flipped at margin 0.25)


def gen_ckpt_weights():
    """models/LA/LA_10.pth and models/ACDC/ACDC_10.pth rounded to bf16 (tests/golden/weights_*_bf16.npz): the trained
    weights the step fixtures below -- and bench.py -- run on (non-degenerate pseudo labels)."""
    for tag, path in CKPT.items():
        sd = torch.load(path, map_location="cpu")
        np.savez_compressed(os.path.join(HERE, "weights_%s_bf16.npz" % tag), **pack_weights_bf16(sd))
        print("wrote weights", tag, os.path.getsize(os.path.join(HERE, "weights_%s_bf16.npz" % tag)) // 1024, "KiB")


def _ckpt_sd(tag):
    return unpack_weights_bf16(np.load(os.path.join(HERE, "weights_%s_bf16.npz" % tag)))


def gen_la_ckpt_step(nsteps=2):
    """LA_BCP_train.py:234-270 at BASELINE configs[1] (8 volumes 112x112x80) from the shipped LA_10 weights: the teacher
    is a trained network, so its pseudo labels are decided (no plab_override in the parity test)."""
    t0 = time.time()
    shape = (112, 112, 80)
    model = R.net_factory.net_factory("VNet", 1, 2, "train")
    ema = R.net_factory.net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    sd = _ckpt_sd("la10")
    model.load_state_dict(sd)
    ema.load_state_dict(sd)
    model.train()
    ema.train()
    inject_dropout(model, seed=252)
    inject_dropout(ema, seed=253)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    out = dict(shape=np.array(shape), box_seed=1337, nsteps=nsteps, sub=4, sure_margin=SURE_MARGIN)
    np.random.seed(1337)
    for it in range(nsteps):
        vol, lab = synthetic_scene(8, shape, 260 + 10 * it)
        # teacher logits of this step (before the step changes the EMA net): same dropout-stream position as the step's own
        # calls, so run the step and capture the teacher outputs through a forward hook
        cap = []
        h = ema.register_forward_hook(lambda m, i, o: cap.append(o[0].detach().clone()))
        r = la_step_reference(model, ema, opt, vol, lab)
        h.remove()
        t_logits = torch.cat(cap[:2])
        margin = t_logits[:, 1] - t_logits[:, 0]
        for k in ("loss", "loss_l", "loss_u"):
            out[f"s{it}_{k}"] = r[k]
        out[f"s{it}_plab"] = packbits(torch.cat([r["plab_a"], r["plab_b"]]))
        out[f"s{it}_plab_raw"] = packbits(margin >= 0)           # softmax >= 0.5 before largest-CC (informative; the
        out[f"s{it}_sure"] = packbits(margin.abs() >= SURE_MARGIN)  # bit-exact kernel check lives in test_gpu_primitives)
        out[f"s{it}_teacher_logits"] = t_logits[..., ::4, ::4, ::4]
        out[f"s{it}_out_l"] = r["out_l"][..., ::4, ::4, ::4]
        out[f"s{it}_out_u"] = r["out_u"][..., ::4, ::4, ::4]
        out[f"s{it}_mixl_digest"] = tensor_digest(r["mixl"])
        out[f"s{it}_mixu_digest"] = tensor_digest(r["mixu"])
        out[f"s{it}_grad_digest"] = digest_named({n: p.grad for n, p in model.named_parameters() if p.grad is not None})
        out[f"s{it}_model_digest"] = digest_named(model.state_dict())
        out[f"s{it}_ema_digest"] = digest_named(ema.state_dict())
        print("la_ckpt step", it, "loss", float(r["loss"]), "plab fg", float(torch.cat([r["plab_a"], r["plab_b"]]).float().mean()),
              "sure", float((margin.abs() >= SURE_MARGIN).float().mean()), "t=%.1fs" % (time.time() - t0))
    save("la_ckpt_step", **out)


def gen_acdc_ckpt_step(nsteps=2):
    """ACDC_BCP_train.py:354-390 at BASELINE configs[2] (batch 24 / labeled 12 of 256x256) from the shipped ACDC_10."""
    model = R.net_factory.BCP_net(1, 4)
    ema = R.net_factory.BCP_net(1, 4, ema=True)
    sd = _ckpt_sd("acdc10")
    model.load_state_dict(sd)
    ema.load_state_dict(sd)
    model.train()
    ema.train()
    inject_dropout(model, seed=292)
    inject_dropout(ema, seed=293)
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    H, W, B, labeled_bs = 256, 256, 24, 12
    out = dict(shape=np.array([H, W]), B=B, labeled_bs=labeled_bs, sure_margin=SURE_MARGIN, nsteps=nsteps)
    np.random.seed(1337)
    for it in range(nsteps):
        volume_batch, label_batch = synthetic_scene(B, (H, W), 300 + 10 * it, n_classes=4, kind="rand")
        label_batch = label_batch.to(torch.uint8)
        ls, us = labeled_bs // 2, (B - labeled_bs) // 2
        img_a, img_b = volume_batch[:ls], volume_batch[ls:labeled_bs]
        uimg_a, uimg_b = volume_batch[labeled_bs:labeled_bs + us], volume_batch[labeled_bs + us:]
        lab_a, lab_b = label_batch[:ls], label_batch[ls:labeled_bs]
        with torch.no_grad():
            pre_a, pre_b = ema(uimg_a), ema(uimg_b)
            plab_a = ACDC_ns.get_ACDC_masks(pre_a, nms=1)
            plab_b = ACDC_ns.get_ACDC_masks(pre_b, nms=1)
            img_mask, loss_mask = ACDC_ns.generate_mask(img_a)
        net_input_unl = uimg_a * img_mask + img_a * (1 - img_mask)
        net_input_l = img_b * img_mask + uimg_b * (1 - img_mask)
        out_unl, out_l = model(net_input_unl), model(net_input_l)
        unl_dice, unl_ce = ACDC_ns.mix_loss(out_unl, plab_a, lab_a, loss_mask, u_weight=0.5, unlab=True)
        l_dice, l_ce = ACDC_ns.mix_loss(out_l, lab_b, plab_b, loss_mask, u_weight=0.5)
        loss_ce, loss_dice = unl_ce + l_ce, unl_dice + l_dice
        loss = (loss_dice + loss_ce) / 2
        opt.zero_grad()
        loss.backward()
        opt.step()
        ACDC_ns.update_model_ema(model, ema, 0.99)
        pre = torch.cat([pre_a, pre_b])
        top2 = pre.topk(2, dim=1).values
        out.update({f"s{it}_loss": loss.detach(), f"s{it}_loss_dice": loss_dice.detach(), f"s{it}_loss_ce": loss_ce.detach(),
                    f"s{it}_plab": torch.cat([plab_a, plab_b]).to(torch.uint8), f"s{it}_plab_raw": pre.argmax(1).to(torch.uint8),
                    f"s{it}_sure": packbits((top2[:, 0] - top2[:, 1]) >= SURE_MARGIN),
                    f"s{it}_teacher_logits": pre[..., ::4, ::4], f"s{it}_out_unl": out_unl.detach()[..., ::4, ::4],
                    f"s{it}_out_l": out_l.detach()[..., ::4, ::4],
                    f"s{it}_grad_digest": digest_named({n: p.grad for n, p in model.named_parameters() if p.grad is not None}),
                    f"s{it}_model_digest": digest_named(model.state_dict()), f"s{it}_ema_digest": digest_named(ema.state_dict())})
        print("acdc_ckpt step", it, float(loss), "plab classes", np.bincount(torch.cat([plab_a, plab_b]).long().flatten().numpy(), minlength=4),
              "sure", float(((top2[:, 0] - top2[:, 1]) >= SURE_MARGIN).float().mean()))
    save("acdc_ckpt_step", **out)


def gen_pan_step():
    t0 = time.time()
    net, ema = R.pan_Vnet.VNet(), R.pan_Vnet.VNet()
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(net, 121)
    ema.load_state_dict(net.state_dict())
    net.train()
    ema.train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    np.random.seed(2020)
    S = (96, 96, 96)
    out = {}
    for it in range(1):
        v = O.synthetic_volume((8, 1) + S, 130 + it)
        l = O.synthetic_labels((8,) + S, 140 + it)
        img_a, img_b, unimg_a, unimg_b = v[0:2], v[2:4], v[4:6], v[6:8]
        lab_a, lab_b = l[0:2], l[2:4]
        with torch.no_grad():
            oa, ob = ema(unimg_a)[0], ema(unimg_b)[0]
            plab_a = PANU.get_cut_mask(oa, nms=True, connect_mode=2)
            plab_b = PANU.get_cut_mask(ob, nms=True, connect_mode=2)
            img_mask, loss_mask = PANU.generate_mask(img_a, 64)
        in_l = unimg_a * img_mask + img_b * (1 - img_mask)
        in_u = img_a * img_mask + unimg_b * (1 - img_mask)
        o1 = net(in_l)[0]
        loss_1 = R.pan_losses.mix_loss(o1, plab_a.long(), lab_b, loss_mask, unlab=True)
        o2 = net(in_u)[0]
        loss_2 = R.pan_losses.mix_loss(o2, lab_a, plab_b.long(), loss_mask)
        loss = loss_1 + loss_2
        opt.zero_grad()
        loss.backward()
        opt.step()
        PANU.update_ema_variables(net, ema, 0.99)
        out.update({f"s{it}_loss": loss.detach(), f"s{it}_loss_1": loss_1.detach(), f"s{it}_loss_2": loss_2.detach(),
                    f"s{it}_plab_a_sum": plab_a.sum(), f"s{it}_plab": torch.cat([plab_a, plab_b]).to(torch.uint8), f"s{it}_out_1": o1.detach()[..., ::4, ::4, ::4],
                    f"s{it}_grad_digest": digest_named({n: p.grad for n, p in net.named_parameters() if p.grad is not None}),
                    f"s{it}_model_digest": digest_named(net.state_dict()), f"s{it}_ema_digest": digest_named(ema.state_dict())})
        print("pan step", it, float(loss), "t=%.1fs" % (time.time() - t0))
    save("pan_step", **out)


def gen_dataset():
    """dataloaders/dataset.py: the reference's LAHeart + Compose([RandomRotFlip, RandomCrop, ToTensor]) + TwoStreamBatchSampler
    driven by a single-process torch DataLoader (num_workers=0: one np.random stream, sampler permutations interleaved
    with the per-sample transform draws) over seeded in-memory 'scans'.  h5py is absent here: a stub ``h5py.File`` serves
    the volumes by case name, every line of the reference module runs unmodified."""
    import importlib.util
    import tempfile
    import types
    from oracle import dataset_oracle as D
    vols = D.synthetic_la_volumes(6, 4242)
    names = ["case%02d" % i for i in range(len(vols))]

    class _File(dict):
        def __init__(self, path, mode="r"):
            name = os.path.basename(os.path.dirname(path))
            im, lb = vols[names.index(name)]
            super().__init__(image=im, label=lb)
    h5 = types.ModuleType("h5py")
    h5.File = _File
    sys.modules["h5py"] = h5
    ref_shims._install_stubs()
    if not hasattr(sys.modules["skimage"], "transform"):
        tr = types.ModuleType("skimage.transform")
        sys.modules["skimage"].transform = tr
        sys.modules["skimage.transform"] = tr
    spec = importlib.util.spec_from_file_location("ref_dataloaders_dataset", os.path.join(ref_shims.REF_CODE, "dataloaders", "dataset.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)
    from torchvision import transforms as TV
    patch = (24, 20, 16)
    out = dict(patch=np.array(patch), nvol=len(vols), seed=99, labeled=2, batch_size=4, labeled_bs=2, epochs=3)
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "train.list"), "w") as f:
            f.write("\n".join(names) + "\n")
        db = ds.LAHeart(base_dir=d, split="train", transform=TV.Compose([ds.RandomRotFlip(), ds.RandomCrop(patch), ds.ToTensor()]))
        sampler = ds.TwoStreamBatchSampler(list(range(2)), list(range(2, 6)), 4, 4 - 2)
        loader = torch.utils.data.DataLoader(db, batch_sampler=sampler, num_workers=0)
        np.random.seed(99)
        b = 0
        for epoch in range(3):
            for batch in loader:
                out[f"b{b}_image"] = batch["image"].numpy().astype(np.float32)
                out[f"b{b}_label"] = batch["label"].numpy().astype(np.uint8)
                b += 1
        out["nbatches"] = b
        # the index stream alone (same seed, sampler only)
        np.random.seed(7)
        out["sampler_indices"] = np.array([list(t) for _ in range(3) for t in sampler], dtype=np.int64)
    save("dataset", **out)


def gen_acdc_dataset():
    """dataloaders/dataset.py:15-88: the reference's BaseDataSets + RandomGenerator + TwoStreamBatchSampler driven by a
    single-process torch DataLoader over seeded in-memory slices of varying size (stub ``h5py.File`` keyed by case name;
    every line of the reference module runs unmodified, scipy's rotate / zoom included)."""
    import importlib.util
    import random
    import tempfile
    import types
    from oracle import dataset_oracle as D
    slices = D.synthetic_acdc_slices(10, 777)
    names = ["patient%03d_slice_%d" % (i // 3, i % 3) for i in range(len(slices))]

    class _File(dict):
        def __init__(self, path, mode="r"):
            name = os.path.splitext(os.path.basename(path))[0]
            im, lb = slices[names.index(name)]
            super().__init__(image=im, label=lb)
    h5 = types.ModuleType("h5py")
    h5.File = _File
    sys.modules["h5py"] = h5
    ref_shims._install_stubs()
    if not hasattr(sys.modules["skimage"], "transform"):
        tr = types.ModuleType("skimage.transform")
        sys.modules["skimage"].transform = tr
        sys.modules["skimage.transform"] = tr
    spec = importlib.util.spec_from_file_location("ref_dataloaders_dataset_acdc", os.path.join(ref_shims.REF_CODE, "dataloaders", "dataset.py"))
    ds = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ds)
    from torchvision import transforms as TV
    patch = (48, 40)
    out = dict(patch=np.array(patch), nslices=len(slices), slice_seed=777, np_seed=31, py_seed=5, labeled=4, batch_size=4, labeled_bs=2, epochs=3)
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "train_slices.list"), "w") as f:
            f.write("\n".join(names) + "\n")
        db = ds.BaseDataSets(base_dir=d, split="train", num=None, transform=TV.Compose([ds.RandomGenerator(patch)]))
        sampler = ds.TwoStreamBatchSampler(list(range(4)), list(range(4, 10)), 4, 4 - 2)
        loader = torch.utils.data.DataLoader(db, batch_sampler=sampler, num_workers=0)
        np.random.seed(31)
        random.seed(5)
        b = 0
        for epoch in range(3):
            for batch in loader:
                out[f"b{b}_image"] = batch["image"].numpy().astype(np.float32)
                out[f"b{b}_label"] = batch["label"].numpy().astype(np.uint8)
                b += 1
        out["nbatches"] = b
    save("acdc_dataset", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["functions", "networks", "la_small", "la_pre", "acdc", "la_full", "pan", "acdc_pre", "pan_pre", "sliding",
                             "ckpt_weights", "la_ckpt", "acdc_ckpt", "dataset", "val_2d", "acdc_dataset"]
    if "dataset" in which:
        gen_dataset()
    if "val_2d" in which:
        gen_val_2d()
    if "acdc_dataset" in which:
        gen_acdc_dataset()
    if "ckpt_weights" in which:
        gen_ckpt_weights()
    if "la_ckpt" in which:
        gen_la_ckpt_step()
    if "acdc_ckpt" in which:
        gen_acdc_ckpt_step()
    if "sliding" in which:
        gen_sliding_window()
    if "acdc_pre" in which:
        gen_acdc_pre_step()
    if "pan_pre" in which:
        gen_pan_pre_step()
    if "functions" in which:
        gen_functions()
    if "networks" in which:
        gen_networks()
    if "la_small" in which:
        gen_la_step("la_step_small", (48, 48, 48), 2, find_box_seed_la((48, 48, 48)), sub=2)
    if "la_pre" in which:
        gen_la_pre_step()
    if "acdc" in which:
        gen_acdc_step()
    if "la_full" in which:
        gen_la_step("la_step_full", (112, 112, 80), 1, 1337, sub=4)
    if "pan" in which:
        gen_pan_step()
