"""Drop-in for the reference's ``utils/BCP_utils.py`` hot-path functions (context_mask, mix_loss,
update_ema_variables; /root/reference/code/utils/BCP_utils.py:18-28,58-69,78-81) on the sm_100a kernels.

``context_mask`` still returns the two int64 mask tensors the reference returns, but they carry the box
coordinates as a ``.box`` attribute so ``mix_loss`` / ``mask_mix`` never read the mask from HBM.
"""
import numpy as np
import torch

from .. import ops
from .._native import LIB, ptr, stream


def context_box(shape, mask_ratio):
    """Box (w,h,z,px,py,pz) drawn exactly like the reference: three np.random.randint calls in the order
    w,h,z with the hard-coded 112/112/80 bounds (utils/BCP_utils.py:22-25)."""
    img_x, img_y, img_z = shape[-3], shape[-2], shape[-1]
    px, py, pz = int(img_x * mask_ratio), int(img_y * mask_ratio), int(img_z * mask_ratio)
    w = np.random.randint(0, 112 - px)
    h = np.random.randint(0, 112 - py)
    z = np.random.randint(0, 80 - pz)
    return (w, h, z, px, py, pz)


def box_to_masks(box, batch_size, spatial, device):
    """Materialise (mask[X,Y,Z], loss_mask[B,X,Y,Z]) int64 like the reference (API compatibility only)."""
    mask = torch.ones(tuple(spatial), dtype=torch.int64, device=device)
    if len(box) == 6:
        w, h, z, px, py, pz = box
        mask[w:w + px, h:h + py, z:z + pz] = 0
    else:
        w, h, px, py = box
        mask[w:w + px, h:h + py] = 0
    loss_mask = mask.unsqueeze(0).repeat(batch_size, *([1] * len(spatial)))
    for m in (mask, loss_mask):
        m.box = tuple(box)
        m._bcp_box_version = m._version        # an in-place edit of the mask afterwards invalidates the attached box
    return mask, loss_mask


def context_mask(img, mask_ratio):
    box = context_box(img.shape, mask_ratio)
    return box_to_masks(box, img.shape[0], img.shape[2:], img.device)


def _box_of(mask):
    """The box a mask tensor was built from, or None when there is none / the tensor was modified in place since
    (then the callers fall back to reading the mask itself)."""
    if isinstance(mask, (tuple, list)):
        return tuple(mask)
    box = getattr(mask, "box", None)
    if box is not None and getattr(mask, "_bcp_box_version", None) != mask._version:
        return None
    return box


def mix_loss(net3_output, img_l, patch_l, mask, l_weight=1.0, u_weight=0.5, unlab=False):
    """(masked Dice + masked CE) / 2 with the image/patch weights of utils/BCP_utils.py:58-69, one fused kernel.
    ``mask``: the loss_mask from context_mask (box attached), a box tuple, or any 0/1 tensor [N,X,Y,Z]."""
    image_weight, patch_weight = (u_weight, l_weight) if unlab else (l_weight, u_weight)
    box = _box_of(mask)
    mask_u8 = None if box is not None else (mask != 0).to(torch.uint8)
    out3 = ops.MixLoss.apply(net3_output, ops.to_u8_labels(img_l), ops.to_u8_labels(patch_l), box, mask_u8, 0,
                             image_weight, patch_weight)
    return out3[0]


def mask_mix(a, b, mask):
    """a*M + b*(1-M) for a mask produced by context_mask (or a box tuple)."""
    box = _box_of(mask)
    if box is None:
        raise ValueError("mask_mix needs an unmodified mask from context_mask()/generate_mask() (carrying .box) or a box tuple")
    return ops.mask_mix(a, b, box)


@torch.no_grad()
def update_ema_variables(model, ema_model, alpha):
    """ema = ema*alpha + (1-alpha)*param for every parameter (buffers untouched), one kernel over the flat arenas
    (utils/BCP_utils.py:78-81)."""
    rt, ert = model.runtime, ema_model.runtime
    for r in (rt, ert):
        if not r.is_flat():
            r.flatten_()
    n = rt.n_param
    assert ert.n_param == n
    hyper = torch.tensor([0.0, 0.0, 0.0, float(alpha), 1.0, 1.0 - float(alpha)], dtype=torch.float32, device=rt.arena.device)
    LIB.call("bcp_sgd_ema_step", ptr(rt.arena), ptr(rt.arena), ptr(rt.arena), ptr(ert.arena), ptr(hyper), 0, n, stream())
    ert.dirty = True
