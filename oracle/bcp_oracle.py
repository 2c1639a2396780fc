"""CPU/fp32 ORACLE for the BCP training-step hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch fp32 (no custom kernels), the algorithm of the
reference repository DeepMed-Lab-ECNU/BCP for the hot path named in BASELINE.json.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import it; the product package ``bcp_b200`` never does.

Arithmetic lives in a third-party dependency of the reference (PyTorch; the reference
pins nothing, README names torch 1.8.0) -- we run torch 2.11.0 here and anchor parity on
the reference's own call sites.  PINNING: the reference ships no tests / golden vectors
(SURVEY.md section 4), so the pins are minted by running the *unmodified reference
modules* (imported from /root/reference/code through ``oracle/ref_shims.py``) on seeded
inputs -- see ``tests/golden/make_golden.py`` (generator, committed) and
``tests/test_oracle_golden.py`` (this file == those vectors).

Every function cites the reference file:line it follows (paths relative to
/root/reference/code).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------

_NORMS = {
    "batchnorm": lambda c: nn.BatchNorm3d(c),
    "groupnorm": lambda c: nn.GroupNorm(num_groups=16, num_channels=c),
    "instancenorm": lambda c: nn.InstanceNorm3d(c),
}


def _stack3d(kind: str, n_stages: int, cin: int, cout: int, norm: str) -> nn.Sequential:
    """Sequential [conv, (norm), relu] * n_stages with the reference's child indices.

    kind='same'  -> Conv3d k3 p1            (networks/VNet.py:6-32, pancreas/Vnet.py:8-35)
    kind='down'  -> Conv3d k2 s2            (networks/VNet.py:68-92, pancreas/Vnet.py:38-62)
    kind='up'    -> ConvTranspose3d k2 s2   (networks/VNet.py:95-119, pancreas/Vnet.py:65-90)
    """
    layers = []
    for i in range(n_stages):
        c_in = cin if i == 0 else cout
        if kind == "same":
            layers.append(nn.Conv3d(c_in, cout, 3, padding=1))
        elif kind == "down":
            layers.append(nn.Conv3d(c_in, cout, 2, stride=2, padding=0))
        elif kind == "up":
            layers.append(nn.ConvTranspose3d(c_in, cout, 2, stride=2, padding=0))
        else:  # pragma: no cover
            raise ValueError(kind)
        if norm != "none":
            layers.append(_NORMS[norm](cout))
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class _Wrap(nn.Module):
    """Holds a Sequential under the attribute name ``conv`` (reference key prefix)."""

    def __init__(self, seq: nn.Sequential):
        super().__init__()
        self.conv = seq

    def forward(self, x):
        return self.conv(x)


_ENC_PLAN = [  # (name, kind, stages, cin_mult, cout_mult)   networks/VNet.py:151-163
    ("block_one", "same", 1, None, 1), ("block_one_dw", "down", 1, 1, 2),
    ("block_two", "same", 2, 2, 2), ("block_two_dw", "down", 1, 2, 4),
    ("block_three", "same", 3, 4, 4), ("block_three_dw", "down", 1, 4, 8),
    ("block_four", "same", 3, 8, 8), ("block_four_dw", "down", 1, 8, 16),
    ("block_five", "same", 3, 16, 16),
]
_DEC_PLAN = [  # networks/VNet.py:198-209
    ("block_five_up", "up", 1, 16, 8), ("block_six", "same", 3, 8, 8),
    ("block_six_up", "up", 1, 8, 4), ("block_seven", "same", 3, 4, 4),
    ("block_seven_up", "up", 1, 4, 2), ("block_eight", "same", 2, 2, 2),
    ("block_eight_up", "up", 1, 2, 1), ("block_nine", "same", 1, 1, 1),
]


def _mlp_heads(owner: nn.Module, n_sel: int):
    """The never-trained projection/prediction/selector heads (networks/VNet.py:250-278,
    networks/unet.py:160-190) -- present only so state_dict keys/param order match."""
    owner.projection_head = nn.Sequential(nn.Linear(16, 32), nn.BatchNorm1d(32), nn.ReLU(inplace=True), nn.Linear(32, 32))
    owner.prediction_head = nn.Sequential(nn.Linear(32, 32), nn.BatchNorm1d(32), nn.ReLU(inplace=True), nn.Linear(32, 32))
    for stem in ("contrastive_class_selector_", "contrastive_class_selector_memory"):
        for c in range(n_sel):
            setattr(owner, stem + str(c), nn.Sequential(
                nn.Linear(32, 32), nn.BatchNorm1d(32), nn.LeakyReLU(negative_slope=0.2, inplace=True), nn.Linear(32, 1)))


class OracleVNet(nn.Module):
    """LA V-Net.  Follows networks/VNet.py:145-290 (Encoder/Decoder/VNet)."""

    def __init__(self, n_channels=3, n_classes=2, n_filters=16, normalization="none",
                 has_dropout=False, has_residual=False):
        super().__init__()
        assert not has_residual, "entry points never enable has_residual"
        f = n_filters
        self.encoder = nn.Module()
        self.decoder = nn.Module()
        for name, kind, st, ci, co in _ENC_PLAN:
            cin = n_channels if ci is None else ci * f
            setattr(self.encoder, name, _Wrap(_stack3d(kind, st, cin, co * f, normalization)))
        self.encoder.dropout = nn.Dropout3d(p=0.5, inplace=False)
        for name, kind, st, ci, co in _DEC_PLAN:
            setattr(self.decoder, name, _Wrap(_stack3d(kind, st, ci * f, co * f, normalization)))
        self.decoder.out_conv = nn.Conv3d(f, n_classes, 1, padding=0)
        self.decoder.dropout = nn.Dropout3d(p=0.5, inplace=False)
        self.has_dropout = has_dropout
        self.pool = nn.MaxPool3d(3, stride=2)
        _mlp_heads(self, 2)

    def forward(self, x):
        e, d = self.encoder, self.decoder
        x1 = e.block_one(x)                                   # VNet.py:167-186
        x2 = e.block_two(e.block_one_dw(x1))
        x3 = e.block_three(e.block_two_dw(x2))
        x4 = e.block_four(e.block_three_dw(x3))
        x5 = e.block_five(e.block_four_dw(x4))
        if self.has_dropout:
            x5 = e.dropout(x5)
        u = d.block_five_up(x5) + x4                          # VNet.py:213-239
        u = d.block_six_up(d.block_six(u)) + x3
        u = d.block_seven_up(d.block_seven(u)) + x2
        u = d.block_eight_up(d.block_eight(u)) + x1
        x9 = d.block_nine(u)
        if self.has_dropout:
            x9 = d.dropout(x9)
        return d.out_conv(x9), self.pool(x5)                  # VNet.py:286-290


class OraclePanVNet(nn.Module):
    """Pancreas V-Net.  Follows pancreas/Vnet.py:92-194 (flat blocks, list output)."""

    def __init__(self, n_channels=1, n_classes=2, n_filters=16, normalization="instancenorm", has_dropout=False):
        super().__init__()
        f = n_filters
        self.has_dropout = has_dropout
        for name, kind, st, ci, co in _ENC_PLAN + _DEC_PLAN[:-1]:
            cin = n_channels if ci is None else ci * f
            setattr(self, name, _Wrap(_stack3d(kind, st, cin, co * f, normalization)))
        if has_dropout:
            self.dropout = nn.Dropout3d(p=0.5)
        mods = [_Wrap(_stack3d("same", 1, f, f, normalization))]
        if has_dropout:
            mods.append(nn.Dropout3d(p=0.5))
        mods.append(nn.Conv3d(f, n_classes, 1, padding=0))
        self.branchs = nn.ModuleList([nn.Sequential(*mods)])

    def forward(self, x, turnoff_drop=False):
        drop = self.has_dropout and not turnoff_drop
        x1 = self.block_one(x)
        x2 = self.block_two(self.block_one_dw(x1))
        x3 = self.block_three(self.block_two_dw(x2))
        x4 = self.block_four(self.block_three_dw(x3))
        x5 = self.block_five(self.block_four_dw(x4))
        if drop:
            x5 = self.dropout(x5)
        u = self.block_five_up(x5) + x4
        u = self.block_six_up(self.block_six(u)) + x3
        u = self.block_seven_up(self.block_seven(u)) + x2
        u = self.block_eight_up(self.block_eight(u)) + x1
        return [b(u) for b in self.branchs]


def _unet_convblock(cin, cout, p):
    """networks/unet.py:15-30: conv-bn-lrelu-dropout-conv-bn-lrelu under ``conv_conv``."""
    m = nn.Module()
    m.conv_conv = nn.Sequential(
        nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU(), nn.Dropout(p),
        nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.LeakyReLU())
    return m


class OracleUNet2d(nn.Module):
    """ACDC U-Net.  Follows networks/unet.py:60-116 (Encoder/Decoder) and :203-257."""
    FT = [16, 32, 64, 128, 256]
    DROP = [0.05, 0.1, 0.2, 0.3, 0.5]

    def __init__(self, in_chns, class_num, return_features=False):
        super().__init__()
        ft, dp = self.FT, self.DROP
        self.encoder = nn.Module()
        self.decoder = nn.Module()
        self.encoder.in_conv = _unet_convblock(in_chns, ft[0], dp[0])
        for i in range(1, 5):
            blk = nn.Module()
            blk.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), _unet_convblock(ft[i - 1], ft[i], dp[i]))
            setattr(self.encoder, f"down{i}", blk)
        for i in range(1, 5):
            c1, c2 = ft[5 - i], ft[4 - i]
            up = nn.Module()
            up.conv1x1 = nn.Conv2d(c1, c2, kernel_size=1)
            up.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            up.conv = _unet_convblock(c2 * 2, c2, 0.0)
            setattr(self.decoder, f"up{i}", up)
        self.decoder.out_conv = nn.Conv2d(ft[0], class_num, kernel_size=3, padding=1)
        self.return_features = return_features  # UNet (:198-201) returns (out, x_last); UNet_2d (:254-257) returns out
        _mlp_heads(self, 4)

    def forward(self, x):
        e, d = self.encoder, self.decoder
        feats = [e.in_conv.conv_conv(x)]
        for i in range(1, 5):
            feats.append(getattr(e, f"down{i}").maxpool_conv[1].conv_conv(F.max_pool2d(feats[-1], 2)))
        y = feats[4]
        for i in range(1, 5):
            up = getattr(d, f"up{i}")
            y = up.up(up.conv1x1(y))
            y = up.conv.conv_conv(torch.cat([feats[4 - i], y], dim=1))
        out = d.out_conv(y)
        return (out, y) if self.return_features else out


# The U-Net's child modules are anonymous nn.Module holders; give the maxpool_conv.1 child the
# ``conv_conv`` attribute path the reference has (maxpool_conv.1.conv_conv.*): already true above.


def net_factory(net_type="unet", in_chns=1, class_num=2, mode="train", tsne=0):
    """networks/net_factory.py:5-12 (without .cuda(): the oracle is a CPU object)."""
    if net_type == "unet" and mode == "train":
        return OracleUNet2d(in_chns, class_num, return_features=True)
    if net_type == "VNet" and tsne == 0:
        return OracleVNet(in_chns, class_num, normalization="batchnorm", has_dropout=(mode == "train"))
    raise ValueError((net_type, mode))


def BCP_net(in_chns=1, class_num=2, ema=False):
    """networks/net_factory.py:14-19."""
    net = OracleUNet2d(in_chns, class_num)
    if ema:
        for p in net.parameters():
            p.detach_()
    return net


# --------------------------------------------------------------------------------------
# masks, mixing, pseudo labels
# --------------------------------------------------------------------------------------

def context_mask_la(img, mask_ratio, rng=np.random):
    """utils/BCP_utils.py:18-28.  Box (int(X*r),int(Y*r),int(Z*r)); origin drawn w,h,z with the
    reference's hard-coded 112/112/80 bounds.  Returns (mask[X,Y,Z] i64, loss_mask[B,X,Y,Z] i64, box)."""
    b, _, X, Y, Z = img.shape
    px, py, pz = int(X * mask_ratio), int(Y * mask_ratio), int(Z * mask_ratio)
    w = rng.randint(0, 112 - px)
    h = rng.randint(0, 112 - py)
    z = rng.randint(0, 80 - pz)
    mask = torch.ones(X, Y, Z, device=img.device)          # the reference does .cuda() here
    mask[w:w + px, h:h + py, z:z + pz] = 0
    return mask.long(), mask.long().unsqueeze(0).repeat(b, 1, 1, 1), (w, h, z, px, py, pz)


def generate_mask_acdc(img, rng=np.random):
    """ACDC_BCP_train.py:131-140."""
    b, _, X, Y = img.shape
    px, py = int(X * 2 / 3), int(Y * 2 / 3)
    w = rng.randint(0, X - px)
    h = rng.randint(0, Y - py)
    mask = torch.ones(X, Y, device=img.device)
    mask[w:w + px, h:h + py] = 0
    return mask.long(), mask.long().unsqueeze(0).repeat(b, 1, 1), (w, h, px, py)


def generate_mask_pan(img, patch_size, rng=np.random):
    """pancreas/pancreas_utils.py:187-200 (hard-coded 96^3)."""
    b = img.shape[0]
    w = rng.randint(0, 96 - patch_size)
    h = rng.randint(0, 96 - patch_size)
    z = rng.randint(0, 96 - patch_size)
    mask = torch.ones(96, 96, 96, device=img.device)
    mask[w:w + patch_size, h:h + patch_size, z:z + patch_size] = 0
    return mask.long(), mask.long().unsqueeze(0).repeat(b, 1, 1, 1), (w, h, z, patch_size, patch_size, patch_size)


def mask_mix(a, b, m):
    """LA_BCP_train.py:248-251 / ACDC_BCP_train.py:372-373: a*M + b*(1-M), M int64 broadcast."""
    return a * m + b * (1 - m)


def _label_components(vol: np.ndarray, connectivity: int | None):
    """skimage.measure.label(vol, connectivity=c) on a binary array, via scipy.ndimage.label with
    the full/partial structuring element (both number components by first voxel in C raster order)."""
    from scipy import ndimage
    c = vol.ndim if connectivity is None else connectivity
    st = ndimage.generate_binary_structure(vol.ndim, c)
    lab, _ = ndimage.label(vol != 0, structure=st)
    return lab


def largest_cc(seg: torch.Tensor, connectivity: int | None = None) -> torch.Tensor:
    """LA_BCP_train.py:65-77 / pancreas_utils.py:284-296: per sample keep the largest component
    (ties -> lowest label), empty volumes unchanged; result float32."""
    out = []
    for n in range(seg.shape[0]):
        v = seg[n].detach().cpu().numpy()
        lab = _label_components(v, connectivity)
        if lab.max() != 0:
            out.append((lab == np.argmax(np.bincount(lab.flat)[1:]) + 1))
        else:
            out.append(v)
    return torch.from_numpy(np.stack(out).astype(np.float32)).to(seg.device)     # torch.Tensor(list).cuda() in the reference


def get_cut_mask(out, thres=0.5, nms=0, connectivity=None):
    """LA_BCP_train.py:57-63 / pancreas_utils.py:275-281."""
    probs = F.softmax(out, 1)
    masks = (probs >= thres).type(torch.int64)[:, 1].contiguous()
    if nms:
        masks = largest_cc(masks, connectivity)
    return masks


def acdc_2d_largest_cc(seg: torch.Tensor) -> torch.Tensor:
    """ACDC_BCP_train.py:89-109: per slice, per class 1..3 keep the largest 8-connected component."""
    out = []
    for i in range(seg.shape[0]):
        acc = None
        for c in range(1, 4):
            v = (seg[i] == c).cpu().numpy().astype(np.int64)
            lab = _label_components(v, None)
            if lab.max() != 0:
                keep = (lab == np.argmax(np.bincount(lab.flat)[1:]) + 1) * c
            else:
                keep = v
            acc = keep if acc is None else acc + keep
        out.append(acc)
    return torch.from_numpy(np.stack(out).astype(np.float32)).to(seg.device)


def get_acdc_masks(output, nms=0):
    """ACDC_BCP_train.py:112-117."""
    probs = F.softmax(output, dim=1)
    _, idx = torch.max(probs, dim=1)
    return acdc_2d_largest_cc(idx) if nms else idx


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------

def mask_dice_loss(logits, target, mask=None, smooth=1e-5):
    """utils/losses.py:47-77 (== pancreas/losses.py:82-112): per-(N,C) masked soft Dice."""
    n, c = logits.shape[:2]
    p = F.softmax(logits.reshape(n, c, -1), dim=1)
    oh = torch.zeros_like(p).scatter_(1, target.reshape(n, 1, -1).long(), 1.0)
    inter, union = p * oh, p + oh
    if mask is not None:
        m = mask.reshape(n, 1, -1)
        inter, union = inter * m, union * m
    inter, union = inter.sum(2), union.sum(2)
    return 1 - ((2 * inter + smooth) / (union + smooth)).mean()


def mix_loss_la(out, img_l, patch_l, mask, l_weight=1.0, u_weight=0.5, unlab=False):
    """utils/BCP_utils.py:58-69 (== pancreas/losses.py:129-141)."""
    img_l, patch_l = img_l.long(), patch_l.long()
    wi, wp = (u_weight, l_weight) if unlab else (l_weight, u_weight)
    pm = 1 - mask
    dice = mask_dice_loss(out, img_l, mask) * wi + mask_dice_loss(out, patch_l, pm) * wp
    ce = wi * (F.cross_entropy(out, img_l, reduction="none") * mask).sum() / (mask.sum() + 1e-16)
    ce = ce + wp * (F.cross_entropy(out, patch_l, reduction="none") * pm).sum() / (pm.sum() + 1e-16)
    return (dice + ce) / 2


def acdc_dice_loss(probs, target, mask, n_classes=4):
    """utils/losses.py:102-134: batch-global per-class masked Dice on probabilities."""
    tot = 0.0
    m = mask.float()
    for i in range(n_classes):
        s = probs[:, i]
        t = (target[:, 0] == i).float()
        inter = torch.sum(s * t * m[:, 0])
        y = torch.sum(t * t * m[:, 0])
        z = torch.sum(s * s * m[:, 0])
        tot = tot + (1 - (2 * inter + 1e-10) / (z + y + 1e-10))
    return tot / n_classes


def mix_loss_acdc(output, img_l, patch_l, mask, l_weight=1.0, u_weight=0.5, unlab=False):
    """ACDC_BCP_train.py:167-179; returns (dice, ce)."""
    img_l, patch_l = img_l.long(), patch_l.long()
    soft = F.softmax(output, dim=1)
    wi, wp = (u_weight, l_weight) if unlab else (l_weight, u_weight)
    pm = 1 - mask
    dice = acdc_dice_loss(soft, img_l.unsqueeze(1), mask.unsqueeze(1)) * wi
    dice = dice + acdc_dice_loss(soft, patch_l.unsqueeze(1), pm.unsqueeze(1)) * wp
    ce = wi * (F.cross_entropy(output, img_l, reduction="none") * mask).sum() / (mask.sum() + 1e-16)
    ce = ce + wp * (F.cross_entropy(output, patch_l, reduction="none") * pm).sum() / (pm.sum() + 1e-16)
    return dice, ce


# --------------------------------------------------------------------------------------
# EMA
# --------------------------------------------------------------------------------------

@torch.no_grad()
def update_ema_variables(model, ema_model, alpha):
    """utils/BCP_utils.py:78-81 (parameters only)."""
    for e, p in zip(ema_model.parameters(), model.parameters()):
        e.data.mul_(alpha).add_((1 - alpha) * p.data)


@torch.no_grad()
def update_model_ema(model, ema_model, alpha):
    """ACDC_BCP_train.py:123-129 (whole state_dict incl. buffers; ints truncated by load)."""
    ms, es = model.state_dict(), ema_model.state_dict()
    ema_model.load_state_dict({k: alpha * es[k] + (1 - alpha) * ms[k] for k in ms})


# --------------------------------------------------------------------------------------
# step bodies
# --------------------------------------------------------------------------------------

def la_self_train_step(model, ema_model, optimizer, volume, label, labeled_bs=4, mask_ratio=2 / 3,
                       u_weight=0.5, alpha=0.99, rng=np.random, nms=1):
    """LA_BCP_train.py:234-270.  volume [B,1,X,Y,Z] f32, label [B,X,Y,Z] i64.  Returns dict of scalars/tensors."""
    sub = labeled_bs // 2
    img_a, img_b = volume[:sub], volume[sub:labeled_bs]
    lab_a, lab_b = label[:sub], label[sub:labeled_bs]
    un_a, un_b = volume[labeled_bs:labeled_bs + sub], volume[labeled_bs + sub:]
    with torch.no_grad():
        oa, _ = ema_model(un_a)
        ob, _ = ema_model(un_b)
        plab_a = get_cut_mask(oa, nms=nms)
        plab_b = get_cut_mask(ob, nms=nms)
        img_mask, loss_mask, box = context_mask_la(img_a, mask_ratio, rng)
    mixl = mask_mix(img_a, un_a, img_mask)
    mixu = mask_mix(un_b, img_b, img_mask)
    out_l, _ = model(mixl)
    out_u, _ = model(mixu)
    loss_l = mix_loss_la(out_l, lab_a, plab_a, loss_mask, u_weight=u_weight)
    loss_u = mix_loss_la(out_u, plab_b, lab_b, loss_mask, u_weight=u_weight, unlab=True)
    loss = loss_l + loss_u
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_ema_variables(model, ema_model, alpha)
    return dict(loss=loss.detach(), loss_l=loss_l.detach(), loss_u=loss_u.detach(), box=box,
                plab_a=plab_a, plab_b=plab_b, mixl=mixl, mixu=mixu, out_l=out_l.detach(), out_u=out_u.detach())


def la_pre_train_step(model, optimizer, volume, label, labeled_bs=4, mask_ratio=2 / 3, rng=np.random):
    """LA_BCP_train.py:146-171."""
    sub = labeled_bs // 2
    v, l = volume[:labeled_bs], label[:labeled_bs]
    img_a, img_b, lab_a, lab_b = v[:sub], v[sub:], l[:sub], l[sub:]
    img_mask, _, box = context_mask_la(img_a, mask_ratio, rng)
    vol = mask_mix(img_a, img_b, img_mask)
    lab = mask_mix(lab_a, lab_b, img_mask)
    out, _ = model(vol)
    loss_ce = F.cross_entropy(out, lab)
    loss_dice = mask_dice_loss(out, lab)
    loss = (loss_ce + loss_dice) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_ce=loss_ce.detach(), loss_dice=loss_dice.detach(), box=box, out=out.detach())


def acdc_self_train_step(model, ema_model, optimizer, volume, label, labeled_bs=12, u_weight=0.5,
                         alpha=0.99, rng=np.random, nms=1):
    """ACDC_BCP_train.py:354-390.  volume [B,1,H,W] f32, label [B,H,W] (uint8/long)."""
    B = volume.shape[0]
    ls, us = labeled_bs // 2, (B - labeled_bs) // 2
    img_a, img_b = volume[:ls], volume[ls:labeled_bs]
    uimg_a, uimg_b = volume[labeled_bs:labeled_bs + us], volume[labeled_bs + us:]
    lab_a, lab_b = label[:ls], label[ls:labeled_bs]
    with torch.no_grad():
        pre_a, pre_b = ema_model(uimg_a), ema_model(uimg_b)
        plab_a, plab_b = get_acdc_masks(pre_a, nms=nms), get_acdc_masks(pre_b, nms=nms)
        img_mask, loss_mask, box = generate_mask_acdc(img_a, rng)
    in_unl = mask_mix(uimg_a, img_a, img_mask)
    in_l = mask_mix(img_b, uimg_b, img_mask)
    out_unl, out_l = model(in_unl), model(in_l)
    unl_dice, unl_ce = mix_loss_acdc(out_unl, plab_a, lab_a, loss_mask, u_weight=u_weight, unlab=True)
    l_dice, l_ce = mix_loss_acdc(out_l, lab_b, plab_b, loss_mask, u_weight=u_weight)
    loss_ce, loss_dice = unl_ce + l_ce, unl_dice + l_dice
    loss = (loss_dice + loss_ce) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_model_ema(model, ema_model, alpha)
    return dict(loss=loss.detach(), loss_dice=loss_dice.detach(), loss_ce=loss_ce.detach(), box=box,
                plab_a=plab_a, plab_b=plab_b, in_unl=in_unl, in_l=in_l, out_unl=out_unl.detach(), out_l=out_l.detach())


def acdc_pre_train_step(model, optimizer, volume, label, labeled_bs=12, rng=np.random):
    """ACDC_BCP_train.py:237-255: two labeled sub-batches mixed through the box, mix_loss(u_weight=1.0, unlab=True)."""
    sub = labeled_bs // 2
    img_a, img_b = volume[:sub], volume[sub:labeled_bs]
    lab_a, lab_b = label[:sub], label[sub:labeled_bs]
    img_mask, loss_mask, box = generate_mask_acdc(img_a, rng)
    net_input = mask_mix(img_a, img_b, img_mask)
    out = model(net_input)
    loss_dice, loss_ce = mix_loss_acdc(out, lab_a, lab_b, loss_mask, u_weight=1.0, unlab=True)
    loss = (loss_dice + loss_ce) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_dice=loss_dice.detach(), loss_ce=loss_ce.detach(), box=box, out=out.detach())


def pan_pre_train_step(net, optimizer, img_a, lab_a, img_b, lab_b, patch_size=64, rng=np.random):
    """pancreas/train_pancreas.py:82-99: image AND label mixed through the box, unmasked CE + Dice."""
    img_mask, _, box = generate_mask_pan(img_a, patch_size, rng)
    img = mask_mix(img_a, img_b, img_mask)
    lab = mask_mix(lab_a, lab_b, img_mask)
    out = net(img)[0]
    loss_ce = F.cross_entropy(out, lab)
    loss_dice = mask_dice_loss(out, lab)
    loss = (loss_ce + loss_dice) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_ce=loss_ce.detach(), loss_dice=loss_dice.detach(), box=box, out=out.detach())


def pan_self_train_step(net, ema_net, optimizer, img_a, lab_a, img_b, lab_b, unimg_a, unimg_b,
                        patch_size=64, alpha=0.99, connect_mode=2, rng=np.random):
    """pancreas/train_pancreas.py:144-174."""
    with torch.no_grad():
        oa, ob = ema_net(unimg_a)[0], ema_net(unimg_b)[0]
        plab_a = get_cut_mask(oa, nms=True, connectivity=connect_mode)
        plab_b = get_cut_mask(ob, nms=True, connectivity=connect_mode)
        img_mask, loss_mask, box = generate_mask_pan(img_a, patch_size, rng)
    in_l = mask_mix(unimg_a, img_b, img_mask)
    in_u = mask_mix(img_a, unimg_b, img_mask)
    out_1 = net(in_l)[0]
    loss_1 = mix_loss_la(out_1, plab_a.long(), lab_b, loss_mask, unlab=True)
    out_2 = net(in_u)[0]
    loss_2 = mix_loss_la(out_2, lab_a, plab_b.long(), loss_mask)
    loss = loss_1 + loss_2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    update_ema_variables(net, ema_net, alpha)
    return dict(loss=loss.detach(), loss_1=loss_1.detach(), loss_2=loss_2.detach(), box=box,
                plab_a=plab_a, plab_b=plab_b, out_1=out_1.detach(), out_2=out_2.detach())


# --------------------------------------------------------------------------------------
# sliding-window validation (utils/test_3d_patch.py:82-141)
# --------------------------------------------------------------------------------------
def sliding_window_predict(model, image: np.ndarray, stride_xy: int, stride_z: int, patch_size, num_classes: int = 1):
    """test_single_case: zero-pad the volume up to the patch size (split evenly, extra voxel on the high side), visit
    windows on a stride grid whose last window is clamped to the border, add the class-1 softmax probability of every
    window into a score map (broadcast over `num_classes` planes) and a visit count, threshold the mean at 0.5.
    Returns (label_map int64 [w,h,d], score_map float32 [num_classes,w,h,d]) cropped back to the input size."""
    w, h, d = image.shape
    pads = [max(patch_size[i] - image.shape[i], 0) for i in range(3)]
    lo = [p // 2 for p in pads]
    if any(pads):
        image = np.pad(image, [(lo[i], pads[i] - lo[i]) for i in range(3)], mode="constant", constant_values=0)
    ww, hh, dd = image.shape
    nx = math.ceil((ww - patch_size[0]) / stride_xy) + 1
    ny = math.ceil((hh - patch_size[1]) / stride_xy) + 1
    nz = math.ceil((dd - patch_size[2]) / stride_z) + 1
    score = np.zeros((num_classes,) + image.shape, np.float32)
    cnt = np.zeros(image.shape, np.float32)
    for ixw in range(nx):
        xs = min(stride_xy * ixw, ww - patch_size[0])
        for iyw in range(ny):
            ys = min(stride_xy * iyw, hh - patch_size[1])
            for izw in range(nz):
                zs = min(stride_z * izw, dd - patch_size[2])
                patch = image[xs:xs + patch_size[0], ys:ys + patch_size[1], zs:zs + patch_size[2]].astype(np.float32)
                with torch.no_grad():
                    logits, _ = model(torch.from_numpy(patch[None, None]))
                    prob = F.softmax(logits, dim=1)[0, 1].numpy()
                sl = (slice(None), slice(xs, xs + patch_size[0]), slice(ys, ys + patch_size[1]), slice(zs, zs + patch_size[2]))
                score[sl] = score[sl] + prob
                cnt[sl[1:]] = cnt[sl[1:]] + 1
    score = score / cnt[None]
    label = (score[0] > 0.5).astype(np.int64)
    if any(pads):
        label = label[lo[0]:lo[0] + w, lo[1]:lo[1] + h, lo[2]:lo[2] + d]
        score = score[:, lo[0]:lo[0] + w, lo[1]:lo[1] + h, lo[2]:lo[2] + d]
    return label, score


# --------------------------------------------------------------------------------------
# deterministic, torch-version-independent weights and inputs for fixtures / parity tests
# --------------------------------------------------------------------------------------

def fill_state_dict_(module: nn.Module, seed: int) -> None:
    """Overwrite every entry of ``module.state_dict()`` with values from numpy RandomState streams
    (stable across torch versions/machines).  Conv/linear weights ~ N(0, 2/fan_in), norm weight
    ~ 1+0.1N, biases ~ 0.1N, running_mean ~ 0.1N, running_var ~ 1+0.1|N|, counters 0."""
    sd = module.state_dict()
    for i, (k, v) in enumerate(sd.items()):
        rs = np.random.RandomState(seed * 1000 + i)
        if k.endswith("num_batches_tracked"):
            v.zero_()
            continue
        x = rs.standard_normal(tuple(v.shape)).astype(np.float32)
        if k.endswith("running_var"):
            x = 1.0 + 0.1 * np.abs(x)
        elif k.endswith("running_mean") or k.endswith("bias"):
            x = 0.1 * x
        elif v.dim() == 1:            # norm weight
            x = 1.0 + 0.1 * x
        else:
            fan_in = int(np.prod(v.shape[1:]))
            x = x * np.sqrt(2.0 / max(fan_in, 1))
        v.copy_(torch.from_numpy(x))


def synthetic_volume(shape, seed: int, kind: str = "randn") -> torch.Tensor:
    rs = np.random.RandomState(seed)
    x = rs.standard_normal(shape) if kind == "randn" else rs.random_sample(shape)
    return torch.from_numpy(x.astype(np.float32))


def synthetic_labels(shape, seed: int, n_classes: int = 2) -> torch.Tensor:
    """Blobby labels: thresholded box-filtered noise, so largest-CC / Dice are non-degenerate."""
    rs = np.random.RandomState(seed)
    x = torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
    nd = len(shape) - 1
    k = 5
    xx = x.unsqueeze(1)
    pool = F.avg_pool3d if nd == 3 else F.avg_pool2d
    sm = pool(xx, k, stride=1, padding=k // 2, count_include_pad=True)[:, 0]
    sm = sm / sm.std()
    if n_classes == 2:
        return (sm > 1.0).long()
    edges = torch.tensor([0.6, 1.0, 1.5])
    return torch.bucketize(sm, edges).clamp_(max=n_classes - 1).long()
