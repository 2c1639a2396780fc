"""CPU restatement of the two medpy metrics the reference's validation calls -- TEST INFRASTRUCTURE.

medpy (``from medpy import metric``: utils/val_2d.py:3, utils/test_3d_patch.py:4) is a third-party dependency that is NOT
in /root/reference nor in this image (requirements: medpy 0.4.0).  PARITY UNPINNED for these two functions: they restate the
published algorithm of ``medpy.metric.binary.dc`` / ``hd95`` (medpy 0.4.0, metric/binary.py) and are anchored only on
closed-form cases (tests/test_oracle_golden.py::test_metric_closed_forms).  tests/golden/make_golden.py installs them as the
``medpy`` stub when it executes the reference's own ``val_2d.test_single_volume``.
"""
import numpy as np
from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure


def dc(result, reference):
    result = np.atleast_1d(result.astype(bool))
    reference = np.atleast_1d(reference.astype(bool))
    intersection = int(np.count_nonzero(result & reference))      # Python ints: 0/0 raises (-> 0.0) as under the numpy medpy 0.4 targets
    size_i1 = int(np.count_nonzero(result))
    size_i2 = int(np.count_nonzero(reference))
    try:
        return 2. * intersection / float(size_i1 + size_i2)
    except ZeroDivisionError:
        return 0.0


def surface_distances(result, reference, voxelspacing=None, connectivity=1):
    result = np.atleast_1d(result.astype(bool))
    reference = np.atleast_1d(reference.astype(bool))
    footprint = generate_binary_structure(result.ndim, connectivity)
    if 0 == np.count_nonzero(result):
        raise RuntimeError('The first supplied array does not contain any binary object.')
    if 0 == np.count_nonzero(reference):
        raise RuntimeError('The second supplied array does not contain any binary object.')
    result_border = result ^ binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ binary_erosion(reference, structure=footprint, iterations=1)
    dt = distance_transform_edt(~reference_border, sampling=voxelspacing)
    return dt[result_border]


def hd95(result, reference, voxelspacing=None, connectivity=1):
    hd1 = surface_distances(result, reference, voxelspacing, connectivity)
    hd2 = surface_distances(reference, result, voxelspacing, connectivity)
    return np.percentile(np.hstack((hd1, hd2)), 95)
