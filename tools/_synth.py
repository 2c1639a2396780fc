"""Seeded stand-in 2-D slices for the loader tools (no dependency on oracle/ or tests/)."""
import numpy as np


def acdc_slices(n, seed, lo=(200, 200), hi=(260, 260)):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        h, w = int(rs.randint(lo[0], hi[0])), int(rs.randint(lo[1], hi[1]))
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        img = (np.sin(yy / rs.uniform(3, 9)) * np.cos(xx / rs.uniform(3, 9)) * 0.5 + 0.5).astype(np.float32)
        d = np.sqrt((yy - 0.5 * h) ** 2 + (xx - 0.5 * w) ** 2)
        r = rs.uniform(0.15, 0.3) * min(h, w)
        lab = np.zeros((h, w), np.uint8)
        lab[d < r] = 1
        lab[d < 0.66 * r] = 2
        lab[d < 0.33 * r] = 3
        out.append((img, lab))
    return out
