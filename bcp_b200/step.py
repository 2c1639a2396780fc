"""BCP training-step bodies on the sm_100a kernels.

la_self_train_step / la_pre_train_step     <- LA_BCP_train.py:234-270 / :146-171
acdc_self_train_step / acdc_pre_train_step <- ACDC_BCP_train.py:354-390 / :237-255
pan_self_train_step / pan_pre_train_step   <- pancreas/train_pancreas.py:144-174 / :82-99

Differences from the reference that do not change the arithmetic: the two teacher forwards (and the two student
forwards) of a step run as ONE batched call with per-call BatchNorm groups; the box mask is never materialised;
pseudo labels, largest-CC, mixing, losses, optimiser and EMA never leave the device and never synchronise.
Labels are uint8 on the device (the reference moves int64 labels: 8x the bytes).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .utils.BCP_utils import context_box


# teacher pass concurrent with the student's forward (see la_self_train_step); module switch for A/B measurements and tests
TEACHER_ON_SIDE_STREAM = True
_SIDE = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    s = _SIDE.get(key)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _SIDE[key] = s
    return s


class _TeacherPass:
    """``with _TeacherPass(dev) as tp:`` runs its body on the side stream (forked from the current one); ``tp.join(*tensors)``
    makes the current stream wait for it and hands the tensors over."""

    def __init__(self, dev):
        self.dev = dev
        self.cur = torch.cuda.current_stream(dev)
        self.side = _side_stream(dev) if TEACHER_ON_SIDE_STREAM else self.cur

    def __enter__(self):
        self.side.wait_stream(self.cur)
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)

    def join(self, *tensors):
        self.cur.wait_stream(self.side)
        for t in tensors:
            t.record_stream(self.cur)


def _acdc_box(shape):
    """ACDC_BCP_train.py:131-140: 2/3 box, origin from two np.random.randint calls (w then h)."""
    X, Y = shape[-2], shape[-1]
    px, py = int(X * 2 / 3), int(Y * 2 / 3)
    w = np.random.randint(0, X - px)
    h = np.random.randint(0, Y - py)
    return (w, h, px, py)


def _pan_box(patch_size):
    """pancreas/pancreas_utils.py:187-200 (hard-coded 96)."""
    w = np.random.randint(0, 96 - patch_size)
    h = np.random.randint(0, 96 - patch_size)
    z = np.random.randint(0, 96 - patch_size)
    return (w, h, z, patch_size, patch_size, patch_size)


def la_self_train_step(model, ema_model, optimizer, volume, label, labeled_bs=4, mask_ratio=2 / 3, u_weight=0.5,
                       nms=1, box=None, plab_override=None):
    """volume [B,1,X,Y,Z] fp32 (labeled first), label [B,X,Y,Z] uint8 -- both on the GPU.  Returns device scalars.
    ``plab_override`` (uint8 [2*sub,...]) replaces the teacher's pseudo labels AFTER the teacher pass has run; it exists
    for parity tests, where a random-weight teacher sits at p~0.5 and bf16 noise flips labels."""
    sub = labeled_bs // 2
    label = ops.to_u8_labels(label)
    img_a, img_b = volume[:sub], volume[sub:labeled_bs]
    lab_a, lab_b = label[:sub], label[sub:labeled_bs]
    un = volume[labeled_bs:]
    un_a, un_b = un[:sub], un[sub:]
    # The student's INPUT needs only the images and the box; the teacher's pseudo labels enter at the loss.  So the teacher
    # pass (forward -> pseudo labels -> largest-CC) runs on a side stream next to the student's forward: the deep layers and
    # the small normalisation kernels of either network leave SMs idle that the other one fills.
    with torch.no_grad(), _TeacherPass(volume.device) as tp:
        t_out, _ = ema_model(un, groups=2, with_features=False)            # ema_model(unimg_a), ema_model(unimg_b)
        plab = plab_raw = ops.pseudo_label(t_out, "thresh", 0.5)           # get_cut_mask
        if nms:
            plab = ops.largest_cc(plab)                                    # LargestCC_pancreas, 26-connectivity
    with torch.no_grad():
        if box is None:
            box = context_box(img_a.shape, mask_ratio)
        mixed = torch.empty((2 * sub,) + tuple(volume.shape[1:]), dtype=torch.float32, device=volume.device)
        ops.mask_mix(img_a, un_a, box, out=mixed[:sub])                    # mixl_img
        ops.mask_mix(un_b, img_b, box, out=mixed[sub:])                    # mixu_img
    out, _ = model(mixed, groups=2, with_features=False)
    tp.join(t_out, plab, plab_raw)
    plab_own = plab
    if plab_override is not None:
        plab = ops.to_u8_labels(plab_override).contiguous()
    plab_a, plab_b = plab[:sub], plab[sub:]
    loss_l = ops.MixLoss.apply(out[:sub], lab_a, plab_a, box, None, 0, 1.0, u_weight)[0]
    loss_u = ops.MixLoss.apply(out[sub:], plab_b, lab_b, box, None, 0, u_weight, 1.0)[0]
    loss = loss_l + loss_u
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()                                                        # SGD + EMA teacher update, fused
    return dict(loss=loss.detach(), loss_l=loss_l.detach(), loss_u=loss_u.detach(), box=box, plab=plab_own, plab_raw=plab_raw,
                teacher_out=t_out, mixed=mixed, out=out.detach())


def la_pre_train_step(model, optimizer, volume, label, labeled_bs=4, mask_ratio=2 / 3, box=None):
    sub = labeled_bs // 2
    label = ops.to_u8_labels(label)
    v, l = volume[:labeled_bs], label[:labeled_bs]
    img_a, img_b, lab_a, lab_b = v[:sub], v[sub:], l[:sub], l[sub:]
    if box is None:
        box = context_box(img_a.shape, mask_ratio)
    with torch.no_grad():
        vol = ops.mask_mix(img_a, img_b, box)
    out, _ = model(vol, with_features=False)
    # label_batch = lab_a*M + lab_b*(1-M) and unmasked CE/Dice == mix_loss with both weights 1 over the two regions
    # evaluated as ONE region: use the mixed label as a single target with an empty box.
    lab = ops.label_mix(lab_a, lab_b, box)
    r = ops.MixLoss.apply(out, lab, lab, (0, 0, 0, 0, 0, 0), None, 0, 1.0, 0.0)
    loss = r[0]
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_dice=r[1].detach(), loss_ce=r[2].detach(), box=box, out=out.detach())


def acdc_pre_train_step(model, optimizer, volume, label, labeled_bs=12, box=None):
    """ACDC_BCP_train.py:237-255: the two labeled sub-batches mixed through the box; mix_loss(u_weight=1.0, unlab=True),
    i.e. both regions weighted 1.  volume [B,1,H,W] fp32, label [B,H,W] uint8 on the GPU."""
    sub = labeled_bs // 2
    label = ops.to_u8_labels(label)
    vol, lab = volume[:labeled_bs], label[:labeled_bs]
    img_a, img_b, lab_a, lab_b = vol[:sub], vol[sub:], lab[:sub], lab[sub:]
    if box is None:
        box = _acdc_box(img_a.shape)
    with torch.no_grad():
        net_input = ops.mask_mix(img_a, img_b, box)                        # :244
    out = model(net_input)
    r = ops.MixLoss.apply(out, lab_a, lab_b, box, None, 1, 1.0, 1.0)        # :249
    loss = (r[1] + r[2]) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_dice=r[1].detach(), loss_ce=r[2].detach(), box=box, out=out.detach())


def acdc_self_train_step(model, ema_model, optimizer, volume, label, labeled_bs=12, u_weight=0.5, nms=1, box=None,
                         plab_override=None):
    """volume [B,1,H,W] fp32, label [B,H,W] uint8 on the GPU."""
    B = volume.shape[0]
    ls, us = labeled_bs // 2, (B - labeled_bs) // 2
    label = ops.to_u8_labels(label)
    img_a, img_b = volume[:ls], volume[ls:labeled_bs]
    un = volume[labeled_bs:]
    uimg_a, uimg_b = un[:us], un[us:]
    lab_a, lab_b = label[:ls], label[ls:labeled_bs]
    with torch.no_grad(), _TeacherPass(volume.device) as tp:              # teacher pass next to the student's forward
        pre = ema_model(un, groups=2)
        plab = plab_raw = ops.pseudo_label(pre, "argmax")
        if nms:
            plab = ops.largest_cc(plab)                                    # per-class 8-connected largest component
    with torch.no_grad():
        if box is None:
            box = _acdc_box(img_a.shape)
        mixed = torch.empty((ls + us,) + tuple(volume.shape[1:]), dtype=torch.float32, device=volume.device)
        ops.mask_mix(uimg_a, img_a, box, out=mixed[:us])                   # net_input_unl
        ops.mask_mix(img_b, uimg_b, box, out=mixed[us:])                   # net_input_l
    out = model(mixed, groups=2)
    tp.join(pre, plab, plab_raw)
    plab_own = plab
    if plab_override is not None:
        plab = ops.to_u8_labels(plab_override).contiguous()
    plab_a, plab_b = plab[:us], plab[us:]
    r_unl = ops.MixLoss.apply(out[:us], plab_a, lab_a, box, None, 1, u_weight, 1.0)     # unlab=True
    r_l = ops.MixLoss.apply(out[us:], lab_b, plab_b, box, None, 1, 1.0, u_weight)
    loss_dice, loss_ce = r_unl[1] + r_l[1], r_unl[2] + r_l[2]
    loss = (loss_dice + loss_ce) / 2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()                                                        # SGD + state_dict EMA, fused
    return dict(loss=loss.detach(), loss_dice=loss_dice.detach(), loss_ce=loss_ce.detach(), box=box, plab=plab_own,
                plab_raw=plab_raw, teacher_out=pre, mixed=mixed, out=out.detach())


def pan_pre_train_step(net, optimizer, img_a, lab_a, img_b, lab_b, patch_size=64, box=None):
    """pancreas/train_pancreas.py:82-99: image and label mixed through the box, unmasked CE + Dice on the mixed label."""
    lab_a, lab_b = ops.to_u8_labels(lab_a), ops.to_u8_labels(lab_b)
    if box is None:
        box = _pan_box(patch_size)
    with torch.no_grad():
        img = ops.mask_mix(img_a, img_b, box)
    out = net(img)[0]
    lab = ops.label_mix(lab_a, lab_b, box)
    # one region (empty box), both terms on the mixed label: the same reduction la_pre_train_step uses
    r = ops.MixLoss.apply(out, lab, lab, (0, 0, 0, 0, 0, 0), None, 0, 1.0, 0.0)
    loss = r[0]
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_dice=r[1].detach(), loss_ce=r[2].detach(), box=box, out=out.detach())


def pan_self_train_step(net, ema_net, optimizer, img_a, lab_a, img_b, lab_b, unimg_a, unimg_b, patch_size=64,
                        connect_mode=2, box=None, plab_override=None):
    lab_a, lab_b = ops.to_u8_labels(lab_a), ops.to_u8_labels(lab_b)
    n = img_a.shape[0]
    with torch.no_grad(), _TeacherPass(img_a.device) as tp:               # teacher pass next to the student's forward
        un = torch.cat([unimg_a, unimg_b])
        t_out = ema_net(un)[0]
        plab = ops.largest_cc(ops.pseudo_label(t_out, "thresh", 0.5), connectivity=connect_mode)
    with torch.no_grad():
        if box is None:
            box = _pan_box(patch_size)
        mixed = torch.empty((2 * n,) + tuple(img_a.shape[1:]), dtype=torch.float32, device=img_a.device)
        ops.mask_mix(unimg_a, img_b, box, out=mixed[:n])                   # net3_input_l
        ops.mask_mix(img_a, unimg_b, box, out=mixed[n:])                   # net3_input_unlab
    out = net(mixed)[0]
    tp.join(t_out, plab)
    plab_own = plab
    if plab_override is not None:
        plab = ops.to_u8_labels(plab_override).contiguous()
    plab_a, plab_b = plab[:n], plab[n:]
    loss_1 = ops.MixLoss.apply(out[:n], plab_a, lab_b, box, None, 0, 0.5, 1.0)[0]       # unlab=True, u_weight default .5
    loss_2 = ops.MixLoss.apply(out[n:], lab_a, plab_b, box, None, 0, 1.0, 0.5)[0]
    loss = loss_1 + loss_2
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return dict(loss=loss.detach(), loss_1=loss_1.detach(), loss_2=loss_2.detach(), box=box, plab=plab_own, out=out.detach())
