"""Sliding-window validation on the device -- host mirror of the reference's ``utils/test_3d_patch.py:82-141``
(``test_single_case``): same arguments, same returned ``(label_map, score_map)`` numpy arrays.

Differences that do not change the arithmetic: windows go through the network ``batch`` at a time (eval-mode BatchNorm uses
running statistics, so batching cannot change a window's logits); the score / count maps live on the GPU and are updated
by ``bcp_window_accumulate`` window by window in the reference's order; one D2H copy at the end instead of one per window.
EXPERIMENTAL: first GPU run pending."""
from __future__ import annotations

import math

import numpy as np
import torch

from .._native import LIB, i3, ptr, stream


def window_origins(shape, patch_size, stride_xy, stride_z):
    """Window origins in the reference's loop order (x outer, z inner); the last window of an axis is clamped to the border."""
    ww, hh, dd = shape
    nx = math.ceil((ww - patch_size[0]) / stride_xy) + 1
    ny = math.ceil((hh - patch_size[1]) / stride_xy) + 1
    nz = math.ceil((dd - patch_size[2]) / stride_z) + 1
    out = []
    for a in range(nx):
        xs = min(stride_xy * a, ww - patch_size[0])
        for b in range(ny):
            ys = min(stride_xy * b, hh - patch_size[1])
            for c in range(nz):
                out.append((xs, ys, min(stride_z * c, dd - patch_size[2])))
    return out


def test_single_case(model, image, stride_xy, stride_z, patch_size, num_classes=1, batch=4):
    if not torch.cuda.is_available():
        raise RuntimeError("bcp_b200.utils.test_3d_patch needs a CUDA device (no CPU fallback)")
    dev = next(model.parameters()).device
    image = np.asarray(image, dtype=np.float32)
    w, h, d = image.shape
    pads = [max(patch_size[i] - image.shape[i], 0) for i in range(3)]
    lo = [p // 2 for p in pads]
    if any(pads):
        image = np.pad(image, [(lo[i], pads[i] - lo[i]) for i in range(3)], mode="constant", constant_values=0)
    vol = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
    shape = tuple(vol.shape)
    score = torch.zeros(shape, dtype=torch.float32, device=dev)
    count = torch.zeros(shape, dtype=torch.float32, device=dev)
    label = torch.empty(shape, dtype=torch.uint8, device=dev)
    origins = window_origins(shape, patch_size, stride_xy, stride_z)
    px, py, pz = patch_size
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            for k in range(0, len(origins), batch):
                chunk = origins[k:k + batch]
                patches = torch.stack([vol[x:x + px, y:y + py, z:z + pz] for x, y, z in chunk]).unsqueeze(1).contiguous()
                out = model(patches)
                logits = (out[0] if isinstance(out, (tuple, list)) else out).contiguous()       # [B, C, px, py, pz] fp32 planar
                c = logits.shape[1]
                for j, org in enumerate(chunk):
                    LIB.call("bcp_window_accumulate", ptr(logits[j]), ptr(score), ptr(count), c, 1, i3(px, py, pz), i3(*shape),
                             i3(*org), stream())
            LIB.call("bcp_window_finalize", ptr(score), ptr(count), ptr(label), score.numel(), 0.5, stream())
    finally:
        model.train(was_training)
    score_np = score.cpu().numpy()
    label_np = label.cpu().numpy().astype(np.int64)
    if any(pads):
        label_np = label_np[lo[0]:lo[0] + w, lo[1]:lo[1] + h, lo[2]:lo[2] + d]
        score_np = score_np[lo[0]:lo[0] + w, lo[1]:lo[1] + h, lo[2]:lo[2] + d]
    # the reference keeps `num_classes` identical planes (it adds the class-1 probability to every plane)
    return label_np, np.broadcast_to(score_np[None], (num_classes,) + score_np.shape).copy()
