"""The module-level drop-in surface INTEGRATION.md section 1 promises (utils.BCP_utils.*, utils.losses.*, the optimiser
wrappers), exercised the way the reference calls it, plus the optimiser kernels against torch.optim."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bcp_oracle as O
from tests.golden.golden_common import inject_dropout, digest_named
from tests.util import load_golden, T, rel_rms, record

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _rel(a, b):
    return abs(float(a) - float(b)) / abs(float(b))


# ------------------------------------------------------------------------------------------- optimisers
def test_adam_ema_vs_torch(dev):
    """bcp_adam_tick + bcp_adam_ema_step against torch.optim.Adam (pancreas/dataloaders.py:182) + the parameter EMA of
    pancreas_utils.py:299-302 for three steps; the bias corrections come from the DEVICE step counter."""
    from bcp_b200._native import LIB, ptr, stream
    n = 200003
    torch.manual_seed(2)
    p0, e0 = torch.randn(n), torch.randn(n)
    grads = [torch.randn(n) * s for s in (1.0, 0.3, 2.0)]
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-3)
    e = e0.clone()
    pd, ed = p0.to(dev).clone(), e0.to(dev).clone()
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    hyper = torch.zeros(12, dtype=torch.float32, device=dev)
    hyper[:6] = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.99, 1.0])
    hyper[8] = 1.0 - 0.99
    hyper[10], hyper[11] = 1.0 - 0.9, 1.0 - 0.999
    step = torch.zeros(1, dtype=torch.int64, device=dev)
    for gstep in grads:
        p.grad = gstep.clone()
        opt.step()
        e.mul_(0.99).add_((1 - 0.99) * p.data)
        LIB.call("bcp_adam_tick", ptr(hyper), ptr(step), stream())
        LIB.call("bcp_adam_ema_step", ptr(pd), ptr(gstep.to(dev)), ptr(m), ptr(v), ptr(ed), ptr(hyper), n - 5, n, stream())
    assert int(step) == 3
    st = opt.state[p]
    # torch's CPU kernels and this kernel contract multiply-adds differently (1-ulp differences per step)
    assert torch.allclose(m.cpu()[:n - 5], st["exp_avg"][:n - 5], rtol=1e-5, atol=1e-7)
    assert torch.allclose(v.cpu()[:n - 5], st["exp_avg_sq"][:n - 5], rtol=1e-5, atol=1e-9)
    err = (pd.cpu()[:n - 5] - p.data[:n - 5]).abs().max()
    record("adam_param_abs_err_max_3_steps", float(err))
    assert err <= 1e-6                                       # updates are ~1e-3 per step: 1e-3 of one update (a wrong
                                                             # bias correction or step size would be off by >= 10 %)
    assert torch.equal(pd.cpu()[n - 5:], p0[n - 5:])         # EMA-only tail is not stepped
    assert torch.allclose(ed.cpu()[:n - 5], e[:n - 5], rtol=1e-6, atol=1e-7)


def test_pan_step_post_update_state(dev):
    """FusedAdam_EMA inside the Pancreas self-training step: parameters and EMA teacher AFTER optimizer.step() against
    the reference's (pan_step fixture: torch.optim.Adam + update_ema_variables), given the reference's pseudo labels."""
    from bcp_b200.pancreas.Vnet import VNet
    from bcp_b200.optim import FusedAdam_EMA
    from bcp_b200.step import pan_self_train_step
    g = load_golden("pan_step")
    net, ema = VNet().to(dev), VNet().to(dev)
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(net, 121)
    ema.load_state_dict(net.state_dict())
    net.train()
    ema.train()
    w0 = {k: v.clone() for k, v in net.state_dict().items()}
    opt = FusedAdam_EMA(net, ema, lr=1e-3, ema_alpha=0.99)
    np.random.seed(2020)
    S = (96, 96, 96)
    v = O.synthetic_volume((8, 1) + S, 130).to(dev)
    l = O.synthetic_labels((8,) + S, 140).to(torch.uint8).to(dev)
    pan_self_train_step(net, ema, opt, v[0:2], l[0:2], v[2:4], l[2:4], v[4:6], v[6:8], plab_override=T(g["s0_plab"]).to(dev))
    # Adam's first step moves every weight by lr * sign(g) (|m|/sqrt(v) = 1): the digests are insensitive to gradient
    # noise but pin step size, bias correction and the EMA blend
    dm, ref = digest_named(net.state_dict()), g["s0_model_digest"]
    big = ref[:, 1] > 1e-6
    rel = np.abs(dm[big, 1] - ref[big, 1]) / ref[big, 1]
    record("pan_post_step_model_abs_sum_rel_max", float(rel.max()))
    assert rel.max() <= 2e-2            # measured 6e-3: elements whose tiny gradient changes sign under bf16 noise move +lr instead of -lr
    de, refe = digest_named(ema.state_dict()), g["s0_ema_digest"]
    bige = refe[:, 1] > 1e-6
    assert (np.abs(de[bige, 1] - refe[bige, 1]) / refe[bige, 1]).max() <= 5e-3
    moved = max(float((net.state_dict()[k] - w0[k]).abs().max()) for k in w0 if w0[k].dtype == torch.float32)
    assert 0.5e-3 <= moved <= 1.5e-3                         # first Adam step = lr per element


def test_graphed_step_follows_lr_decay(dev):
    """a captured step must see optimizer.param_groups changes (LA_BCP_train.py:273-276 decays the LR every
    2500 iterations).  Replays with lr=0 must leave the weights untouched; restoring the LR must move them again."""
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    from bcp_b200.graph import GraphedStep
    shape = (32, 32, 16)
    model, ema = net_factory("VNet", 1, 2, "train"), net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    O.fill_state_dict_(model, 7)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    opt = FusedSGD_EMA(model, ema, lr=0.01, momentum=0.0, weight_decay=0.0, ema_alpha=0.99)
    before_capture = model.runtime.arena.clone() if model.runtime.arena is not None else None
    gs = GraphedStep("la", model, ema, opt, (8, 1) + shape, labeled_bs=4)
    w_init = model.runtime.arena.clone()
    if before_capture is not None:
        assert torch.equal(w_init, before_capture), "graph warm-up must not change the training state"
    assert opt.step_count == 0
    vol = O.synthetic_volume((8, 1) + shape, 3).pin_memory()
    lab = O.synthetic_labels((8,) + shape, 4).to(torch.uint8).pin_memory()
    box = (3, 5, 2, 21, 21, 10)
    gs(vol, lab, box=box)
    torch.cuda.synchronize()
    w1 = model.runtime.arena.clone()
    assert not torch.equal(w1, w_init) and opt.step_count == 1
    opt.param_groups[0]["lr"] = 0.0
    gs(vol, lab, box=box)
    torch.cuda.synchronize()
    assert torch.equal(model.runtime.arena[:model.runtime.n_train], w1[:model.runtime.n_train]), "lr=0 replay moved the weights"
    opt.param_groups[0]["lr"] = 0.001
    gs(vol, lab, box=box)
    torch.cuda.synchronize()
    d_small = (model.runtime.arena - w1)[:model.runtime.n_train].abs().max()
    d_big = (w1 - w_init)[:model.runtime.n_train].abs().max()
    assert 0 < float(d_small) < 0.5 * float(d_big)
    assert opt.step_count == 3


def test_graphed_step_equals_eager(dev):
    """One CUDA-graph replay produces the same loss and weights as the eager step on the same inputs and box."""
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    from bcp_b200.graph import GraphedStep
    from bcp_b200.step import la_self_train_step
    shape = (32, 32, 16)
    vol = O.synthetic_volume((8, 1) + shape, 3)
    lab = O.synthetic_labels((8,) + shape, 4).to(torch.uint8)
    box = (3, 5, 2, 21, 21, 10)
    res = []
    for graphed in (False, True):
        # mode="test" builds the V-Net without Dropout3d (random masks differ between an eager run and a replay); train()
        # keeps batch-statistic BatchNorm, the running-stat updates and everything else of the training step
        model, ema = net_factory("VNet", 1, 2, "test"), net_factory("VNet", 1, 2, "test")
        for p in ema.parameters():
            p.detach_()
        O.fill_state_dict_(model, 7)
        ema.load_state_dict(model.state_dict())
        model.train()
        ema.train()
        opt = FusedSGD_EMA(model, ema, lr=0.01, momentum=0.9, weight_decay=1e-4, ema_alpha=0.99)
        if graphed:
            gs = GraphedStep("la", model, ema, opt, (8, 1) + shape, labeled_bs=4)
            r = gs(vol.pin_memory(), lab.pin_memory(), box=box)
        else:
            r = la_self_train_step(model, ema, opt, vol.to(dev), lab.to(dev), box=box)
        torch.cuda.synchronize()
        res.append((float(r["loss"]), model.runtime.arena.clone(), ema.runtime.arena.clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


# ------------------------------------------------------------------------------------------- utils.* drop-ins
def test_bcp_utils_dropin_functions(dev):
    """context_mask / mix_loss / mask_mix / update_ema_variables with the reference's signatures and call patterns
    (LA_BCP_train.py:150-160,245-262) against the function-level golden vectors."""
    from bcp_b200.utils import BCP_utils as BU
    from bcp_b200.utils import losses as L
    g = load_golden("functions")
    # context_mask: same randint order as utils/BCP_utils.py:22-25
    np.random.seed(1337)
    img = torch.zeros(2, 1, 112, 112, 80, device=dev)
    mask, loss_mask = BU.context_mask(img, 2 / 3)
    assert mask.dtype == torch.int64 and tuple(mask.shape) == (112, 112, 80) and tuple(loss_mask.shape) == (2, 112, 112, 80)
    bb = np.array([[int(i.min()), int(i.max()) + 1] for i in torch.nonzero(mask == 0, as_tuple=True)])
    assert np.array_equal(bb, g["la_ctx_mask_zero_bbox"])
    assert int(mask.sum()) == int(g["la_ctx_mask_sum"]) and int(loss_mask.sum()) == int(g["la_ctx_lmask_sum"])
    a, b = torch.randn(2, 1, 112, 112, 80, device=dev), torch.randn(2, 1, 112, 112, 80, device=dev)
    assert torch.equal(BU.mask_mix(a, b, mask), a * mask + b * (1 - mask))
    # mix_loss with a materialised mask tensor (no box attached) and with the box-carrying mask: both = reference value
    lg = T(g["la_logits"]).to(dev)
    la, lb = T(g["la_lab_a"]).to(dev), T(g["la_lab_b"]).to(dev)          # int64 / float32 like the reference's tensors
    m = T(g["la_mask"]).to(dev)
    x = lg.clone().requires_grad_(True)
    loss = BU.mix_loss(x, la, lb, m, u_weight=0.5)
    loss.backward()
    assert _rel(loss, g["la_mix_loss"]) <= 1e-5
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["la_mix_grad"], rtol=1e-4, atol=1e-7)
    loss = BU.mix_loss(lg, lb, la, m, u_weight=0.5, unlab=True)
    assert _rel(loss, g["la_mix_loss_unlab"]) <= 1e-5
    _, lm = BU.box_to_masks((2, 3, 1, 7, 5, 5), 2, (12, 10, 8), dev)
    assert torch.equal(lm, m)
    assert _rel(BU.mix_loss(lg, la, lb, lm, u_weight=0.5), g["la_mix_loss"]) <= 1e-5
    # an in-place edit after context_mask() must invalidate the attached box: loss follows the tensor
    lm2 = lm.clone()
    lm2.box, lm2._bcp_box_version = lm.box, lm2._version
    lm2[:, :2] = 0
    ref = O.mix_loss_la(lg.cpu(), la.cpu(), lb.cpu(), lm2.cpu(), u_weight=0.5)
    assert _rel(BU.mix_loss(lg, la, lb, lm2, u_weight=0.5), ref) <= 1e-5
    # mask_DiceLoss (utils/losses.py:47-77): masked, unmasked
    dice = L.mask_DiceLoss(nclass=2)
    assert abs(float(dice(lg, la, m)) - float(g["la_dice_masked"])) <= 1e-5
    assert abs(float(dice(lg, la)) - float(g["la_dice_unmasked"])) <= 1e-5
    # update_ema_variables (utils/BCP_utils.py:78-81)
    from bcp_b200.networks.VNet import VNet
    m1, m2 = VNet(1, 2, 4, "batchnorm", True), VNet(1, 2, 4, "batchnorm", True)
    O.fill_state_dict_(m1, 3)
    O.fill_state_dict_(m2, 4)
    m1, m2 = m1.to(dev), m2.to(dev)
    BU.update_ema_variables(m1, m2, 0.99)
    d = digest_named(m2.state_dict())
    assert np.allclose(d, g["ema_la_digest"], rtol=1e-6, atol=1e-9)


def test_acdc_diceloss_on_probabilities(dev):
    """utils.losses.DiceLoss called exactly like ACDC_BCP_train.py:167-179 does (probabilities in, masks [N,1,H,W]):
    the reference's own mix_loss body, verbatim, on top of the drop-in class -- values and gradient vs the golden."""
    from bcp_b200.utils import losses as L
    import torch.nn as nn
    g = load_golden("functions")
    dice_loss = L.DiceLoss(n_classes=4)

    def mix_loss(output, img_l, patch_l, mask, l_weight=1.0, u_weight=0.5, unlab=False):      # ACDC_BCP_train.py:167-179
        CE = nn.CrossEntropyLoss(reduction='none')
        img_l, patch_l = img_l.type(torch.int64), patch_l.type(torch.int64)
        output_soft = F.softmax(output, dim=1)
        image_weight, patch_weight = l_weight, u_weight
        if unlab:
            image_weight, patch_weight = u_weight, l_weight
        patch_mask = 1 - mask
        loss_dice = dice_loss(output_soft, img_l.unsqueeze(1), mask.unsqueeze(1)) * image_weight
        loss_dice += dice_loss(output_soft, patch_l.unsqueeze(1), patch_mask.unsqueeze(1)) * patch_weight
        loss_ce = image_weight * (CE(output, img_l) * mask).sum() / (mask.sum() + 1e-16)
        loss_ce += patch_weight * (CE(output, patch_l) * patch_mask).sum() / (patch_mask.sum() + 1e-16)
        return loss_dice, loss_ce

    lg4 = T(g["acdc_logits"]).to(dev)
    ta, tb, m2 = T(g["acdc_lab_a"]).to(dev), T(g["acdc_lab_b"]).to(dev), T(g["acdc_mask"]).to(dev)
    x = lg4.clone().requires_grad_(True)
    d, c = mix_loss(x, ta, tb, m2, u_weight=0.5)
    ((d + c) / 2).backward()
    assert _rel(d, g["acdc_mix_dice"]) <= 1e-5 and _rel(c, g["acdc_mix_ce"]) <= 1e-5
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["acdc_mix_grad"], rtol=1e-4, atol=1e-7)
    d, c = mix_loss(lg4, tb, ta, m2, u_weight=0.5, unlab=True)
    assert _rel(d, g["acdc_mix_dice_unlab"]) <= 1e-5 and _rel(c, g["acdc_mix_ce_unlab"]) <= 1e-5
    # softmax=True entry and the unmasked form against the oracle
    s = F.softmax(lg4, 1)
    ref = O.acdc_dice_loss(s.cpu(), ta.cpu().long().unsqueeze(1), torch.ones_like(ta.cpu()).unsqueeze(1).float())
    assert _rel(dice_loss(lg4, ta.unsqueeze(1), softmax=True), ref) <= 1e-5
    assert _rel(dice_loss(s, ta.unsqueeze(1)), ref) <= 1e-5


def test_reference_loop_body_with_swapped_imports(dev):
    """The reference's self-training loop body (LA_BCP_train.py:236-270), typed here line by line, running on the
    drop-in modules with ONLY the imports of INTEGRATION.md section 1 swapped (and .cuda() calls already satisfied:
    tensors live on the GPU).  Checked against the la_step_small fixture produced by the unmodified reference."""
    from bcp_b200.networks.net_factory import net_factory                                    # networks.net_factory
    from bcp_b200.utils.BCP_utils import context_mask, mix_loss, update_ema_variables        # utils.BCP_utils
    from bcp_b200 import ops
    import torch.optim as optim
    g = load_golden("la_step_small")
    shape = (48, 48, 48)
    model = net_factory(net_type="VNet", in_chns=1, class_num=2, mode="train")
    ema_model = net_factory(net_type="VNet", in_chns=1, class_num=2, mode="train")
    for param in ema_model.parameters():
        param.detach_()
    O.fill_state_dict_(model, 51)
    ema_model.load_state_dict(model.state_dict())
    inject_dropout(model, seed=52)
    inject_dropout(ema_model, seed=53)
    optimizer = optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)    # LA_BCP_train.py:218, stock torch
    model.train()
    ema_model.train()
    labeled_bs, sub_bs = 4, 2
    np.random.seed(int(g["box_seed"]))

    def get_cut_mask(out, thres=0.5, nms=0):                                                  # LA_BCP_train.py:57-63
        masks = ops.pseudo_label(out, "thresh", thres)
        return ops.largest_cc(masks, out_float=True) if nms == 1 else masks

    volume_batch = O.synthetic_volume((8, 1) + shape, 60).to(dev)
    label_batch = O.synthetic_labels((8,) + shape, 70).to(dev)                                # int64, like the loader's
    # ---- LA_BCP_train.py:237-270 ------------------------------------------------------------------------------------
    img_a, img_b = volume_batch[:sub_bs], volume_batch[sub_bs:labeled_bs]
    lab_a, lab_b = label_batch[:sub_bs], label_batch[sub_bs:labeled_bs]
    unimg_a, unimg_b = volume_batch[labeled_bs:labeled_bs + sub_bs], volume_batch[labeled_bs + sub_bs:]
    with torch.no_grad():
        unoutput_a, _ = ema_model(unimg_a)
        unoutput_b, _ = ema_model(unimg_b)
        plab_a = get_cut_mask(unoutput_a, nms=1)
        plab_b = get_cut_mask(unoutput_b, nms=1)
        img_mask, loss_mask = context_mask(img_a, 2 / 3)
    # the reference's pseudo labels for the student (random-weight fixture: its teacher sits at p ~ 0.5)
    gp = T(g["s0_plab"]).to(dev).float()
    record("loop_body_plab_mismatch_frac", float((torch.cat([plab_a, plab_b]) != gp).float().mean()))
    plab_a, plab_b = gp[:2], gp[2:]
    mixl_img = img_a * img_mask + unimg_a * (1 - img_mask)
    mixu_img = unimg_b * img_mask + img_b * (1 - img_mask)
    outputs_l, _ = model(mixl_img)
    outputs_u, _ = model(mixu_img)
    loss_l = mix_loss(outputs_l, lab_a, plab_a, loss_mask, u_weight=0.5)
    loss_u = mix_loss(outputs_u, plab_b, lab_b, loss_mask, u_weight=0.5, unlab=True)
    loss = loss_l + loss_u
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_ema_variables(model, ema_model, 0.99)
    # ---------------------------------------------------------------------------------------------------------------------
    for k, v in (("loss", loss), ("loss_l", loss_l), ("loss_u", loss_u)):
        rel = _rel(v, g[f"s0_{k}"])
        record(f"loop_body_{k}_rel_err", rel)
        assert rel <= 2e-3, (k, rel)
    assert all(p.grad is not None for n, p in model.named_parameters() if n.startswith(("encoder.", "decoder.")))
    dm, ref = digest_named(model.state_dict()), g["s0_model_digest"]
    big = ref[:, 1] > 1e-6
    assert (np.abs(dm[big, 1] - ref[big, 1]) / ref[big, 1]).max() <= 5e-2
