"""Sliding-window validation on the device -- host mirror of the reference's ``utils/test_3d_patch.py:82-141``
(``test_single_case``): same arguments, same returned ``(label_map, score_map)`` numpy arrays.

Differences that do not change the arithmetic: windows go through the network ``batch`` at a time (eval-mode BatchNorm uses
running statistics, so batching cannot change a window's logits); the score / count maps live on the GPU and are updated
by ``bcp_window_accumulate`` window by window in the reference's order; one D2H copy at the end instead of one per window.
Parity: tests/test_gpu_networks.py::test_sliding_window_validation (fixture minted from the reference function)."""
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


def _read_case(root_path, name):
    import os
    h5 = os.path.join(root_path, "2018LA_Seg_Training Set", name, "mri_norm2.h5")
    if os.path.exists(h5):
        import h5py
        with h5py.File(h5, "r") as f:
            return f["image"][:], f["label"][:]
    z = np.load(os.path.join(root_path, name + ".npz"))
    return z["image"], z["label"]


def dice_binary(pred, gt):
    """medpy.metric.binary.dc: 2|A & B| / (|A| + |B|), 0 when both are empty."""
    a, b = np.asarray(pred) > 0, np.asarray(gt) > 0
    den = int(a.sum()) + int(b.sum())
    return 2.0 * int((a & b).sum()) / den if den else 0.0


def var_all_case_LA(model, num_classes, patch_size=(112, 112, 80), stride_xy=18, stride_z=4, root_path="/data/byh_data/SSNet_data/LA",
                    cases=None):
    """utils/test_3d_patch.py:20-39: mean Dice of the sliding-window prediction over the test list.  ``cases``: optional
    [(image, label)] already in memory (tests)."""
    if cases is None:
        import os
        with open(os.path.join(root_path, "test.list"), "r") as f:
            names = [l.replace("\n", "") for l in f.readlines()]
        cases = (_read_case(root_path, n) for n in names)
    total, count = 0.0, 0
    for image, label in cases:
        prediction, _ = test_single_case(model, image, stride_xy, stride_z, patch_size, num_classes=num_classes)
        total += 0.0 if np.sum(prediction) == 0 else dice_binary(prediction, label)
        count += 1
    avg = total / max(count, 1)
    print("average metric is {}".format(avg))
    return avg
