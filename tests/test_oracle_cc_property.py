"""The oracle's largest-connected-component step rests on scipy.ndimage.label standing in for skimage.measure.label
(scikit-image is not installed here).  skimage numbers components by their first voxel in C raster order and the reference
keeps `np.argmax(np.bincount(labels.flat)[1:]) + 1`, i.e. the largest component, the FIRST one in raster order on a tie.
This test restates exactly that rule with a plain flood fill and checks the oracle against it on random small volumes,
with forced ties, for every connectivity the reference uses (3-D: 26 / 18 / 6 neighbours, 2-D: 8)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import bcp_oracle as O


def _flood_label(vol, connectivity):
    """Components numbered 1.. in the raster order of their first voxel; neighbours differ in <= `connectivity` axes."""
    nd = vol.ndim
    offs = [d for d in itertools.product((-1, 0, 1), repeat=nd) if any(d) and sum(abs(x) for x in d) <= connectivity]
    lab = np.zeros(vol.shape, np.int64)
    cur = 0
    for idx in np.ndindex(*vol.shape):
        if vol[idx] == 0 or lab[idx]:
            continue
        cur += 1
        lab[idx] = cur
        stack = [idx]
        while stack:
            p = stack.pop()
            for d in offs:
                q = tuple(p[i] + d[i] for i in range(nd))
                if all(0 <= q[i] < vol.shape[i] for i in range(nd)) and vol[q] != 0 and not lab[q]:
                    lab[q] = cur
                    stack.append(q)
    return lab


def _largest_first(vol, connectivity):
    lab = _flood_label(vol, connectivity)
    if lab.max() == 0:
        return vol.astype(np.float32)
    sizes = np.bincount(lab.ravel())[1:]
    return (lab == int(np.argmax(sizes)) + 1).astype(np.float32)


@pytest.mark.parametrize("connectivity", [1, 2, 3])
def test_largest_cc_3d_matches_flood_fill(connectivity):
    rng = np.random.RandomState(100 + connectivity)
    vols = []
    for density in (0.15, 0.3, 0.5):
        vols.append((rng.rand(6, 7, 5) < density).astype(np.int64))
    vols.append(np.zeros((6, 7, 5), np.int64))                       # empty volume: returned unchanged
    tie = np.zeros((6, 7, 5), np.int64)                              # two components of equal size: the first in raster order wins
    tie[0, 0, 0:3] = 1
    tie[5, 6, 2:5] = 1
    vols.append(tie)
    seg = torch.from_numpy(np.stack(vols))
    got = O.largest_cc(seg, connectivity).numpy()
    for i, v in enumerate(vols):
        assert np.array_equal(got[i], _largest_first(v, connectivity)), (connectivity, i)
    assert got[-1][0, 0, 0] == 1 and got[-1][5, 6, 4] == 0


def test_acdc_per_class_largest_cc_matches_flood_fill():
    rng = np.random.RandomState(7)
    seg = torch.from_numpy(rng.randint(0, 4, size=(3, 12, 11)).astype(np.int64))
    got = O.acdc_2d_largest_cc(seg).numpy()
    for i in range(seg.shape[0]):
        want = np.zeros(seg.shape[1:], np.float32)
        for c in (1, 2, 3):
            want += _largest_first((seg[i].numpy() == c).astype(np.int64), 2) * c      # 2-D default connectivity = 8 neighbours
        assert np.array_equal(got[i], want), i
