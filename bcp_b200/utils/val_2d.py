"""Per-slice validation of the 2-D (ACDC) network -- mirror of the reference's ``utils/val_2d.py:10-41``
(``calculate_metric_percase``, ``test_single_volume``), called every 200 iterations by ACDC_BCP_train.py:270-283,411-424.

Same arguments and return value (``[(dice, hd95)]`` for classes 1..classes-1).  What moves to the device: all slices of
the volume go through the network in batches (eval-mode BatchNorm uses running statistics, so batching cannot change a
slice's logits) and ``argmax(softmax)`` is the pseudo-label kernel (``bcp_pseudo_label`` mode 1; softmax is monotone, the
tie rule is the first maximum like torch.argmax).  The nearest-neighbour resampling (``scipy.ndimage.zoom(order=0)``) and
the metrics stay on the host exactly as in the reference.  ``medpy`` is not a dependency: ``dc`` and ``hd95`` restate
medpy 0.4's ``metric.binary`` (Dice = 2|A&B|/(|A|+|B|); HD95 = 95th percentile of the two directed surface-distance sets,
surfaces = object minus its erosion with the connectivity-1 structuring element, distances by Euclidean distance transform).
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure, zoom

from ..ops import pseudo_label


def dc(result, reference):
    result, reference = np.atleast_1d(np.asarray(result).astype(bool)), np.atleast_1d(np.asarray(reference).astype(bool))
    inter = np.count_nonzero(result & reference)
    den = np.count_nonzero(result) + np.count_nonzero(reference)
    return 2.0 * inter / float(den) if den else 0.0


def _surface_distances(result, reference, voxelspacing=None, connectivity=1):
    result, reference = np.atleast_1d(np.asarray(result).astype(bool)), np.atleast_1d(np.asarray(reference).astype(bool))
    footprint = generate_binary_structure(result.ndim, connectivity)
    if not np.count_nonzero(result):
        raise RuntimeError("The first supplied array does not contain any binary object.")
    if not np.count_nonzero(reference):
        raise RuntimeError("The second supplied array does not contain any binary object.")
    result_border = result ^ binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ binary_erosion(reference, structure=footprint, iterations=1)
    dt = distance_transform_edt(~reference_border, sampling=voxelspacing)
    return dt[result_border]


def hd95(result, reference, voxelspacing=None, connectivity=1):
    hd1 = _surface_distances(result, reference, voxelspacing, connectivity)
    hd2 = _surface_distances(reference, result, voxelspacing, connectivity)
    return np.percentile(np.hstack((hd1, hd2)), 95)


def calculate_metric_percase(pred, gt):
    pred, gt = np.asarray(pred).copy(), np.asarray(gt).copy()
    pred[pred > 0] = 1
    gt[gt > 0] = 1
    if pred.sum() > 0:
        return dc(pred, gt), hd95(pred, gt)
    return 0, 0


def test_single_volume(image, label, model, classes, patch_size=(256, 256), batch=16):
    if not torch.cuda.is_available():
        raise RuntimeError("bcp_b200.utils.val_2d needs a CUDA device (no CPU fallback)")
    dev = next(model.parameters()).device
    image = image.squeeze(0).cpu().detach().numpy() if torch.is_tensor(image) else np.asarray(image)
    label = label.squeeze(0).cpu().detach().numpy() if torch.is_tensor(label) else np.asarray(label)
    prediction = np.zeros_like(label)
    d, x, y = image.shape
    slices = np.stack([zoom(image[i], (patch_size[0] / x, patch_size[1] / y), order=0) for i in range(d)]).astype(np.float32)
    was_training = model.training
    model.eval()
    try:
        with torch.no_grad():
            for k in range(0, d, batch):
                inp = torch.from_numpy(slices[k:k + batch]).unsqueeze(1).to(dev)
                out = model(inp)
                if isinstance(out, (tuple, list)):
                    out = out[0]
                lab = pseudo_label(out, mode="argmax").cpu().numpy()
                for j in range(lab.shape[0]):
                    prediction[k + j] = zoom(lab[j], (x / patch_size[0], y / patch_size[1]), order=0)
    finally:
        model.train(was_training)
    return [calculate_metric_percase(prediction == i, label == i) for i in range(1, classes)]
