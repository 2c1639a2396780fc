"""Numpy model of the round-2 forward-conv plan: dz folded into the MMA N dimension, A tiles whose
8-row core-matrix groups start every 6 rows (descriptor SBO = 96 B), outputs combined with two intra-group lane shifts.
Checks the index algebra against a direct 3x3x3 'same' convolution on one brick; no GPU involved."""
import numpy as np

rng = np.random.default_rng(0)
BX, BY, BZ, C, NS = 2, 5, 14, 16, 16
HX, HY, HZ = BX + 2, BY + 2, BZ + 2
frame = rng.standard_normal((HX, HY, HZ, C)).astype(np.float32)            # halo'd brick (zero padding already inside)
W = rng.standard_normal((27, C, NS)).astype(np.float32)
rows = frame.reshape(-1, C)                                                 # halo-frame rows, 16 B each in shared memory
nrows = rows.shape[0]
pad = np.zeros((4096, C), np.float32)                                       # reads past the brick hit other smem: garbage x dropped
rows_p = np.concatenate([rows, pad])

# reference: out[x,y,z] = sum_taps frame[x+dx, y+dy, z+dz] . W[tap]   (x,y,z brick-local, frame offset by the halo)
ref = np.zeros((BX, BY, BZ, NS), np.float32)
for dx in range(3):
    for dy in range(3):
        for dz in range(3):
            ref += frame[dx:dx + BX, dy:dy + BY, dz:dz + BZ] @ W[(dx * 3 + dy) * 3 + dz]

# model: tile t covers frame rows [96 t, 96 t + 98); tile row i -> frame row 96 t + 6 (i // 8) + (i % 8)
out = np.full((nrows, NS), np.nan, np.float32)
lmax = ((BX - 1) * HY + (BY - 1)) * HZ + BZ                                  # one past the last output row (halo-frame index)
ntiles = (lmax + 95) // 96
mma = 0
for t in range(ntiles):
    i = np.arange(128)
    fr = 96 * t + 6 * (i // 8) + (i % 8)
    D = np.zeros((128, 3 * NS), np.float32)                                  # TMEM: 128 lanes x 3*Ns columns
    for dx in range(3):
        for dy in range(3):
            A = rows_p[fr + (dx * HY + dy) * HZ]                              # ONE descriptor start offset per (dx,dy)
            B = np.concatenate([W[(dx * 3 + dy) * 3 + dz] for dz in range(3)], axis=1)   # [C][3*Ns]
            D += A @ B
            mma += 1
    # epilogue: lane i takes column group dz from lane i+dz of its own 8-lane group (shfl_down, width 8)
    for lane in range(128):
        if lane % 8 >= 6:
            continue
        o = D[lane, :NS] + D[lane + 1, NS:2 * NS] + D[lane + 2, 2 * NS:]
        out[fr[lane]] = o
# gather valid outputs: output voxel (ix,iy,iz) sits at frame row (ix*HY+iy)*HZ+iz
err = 0.0
for ix in range(BX):
    for iy in range(BY):
        for iz in range(BZ):
            r = (ix * HY + iy) * HZ + iz
            assert not np.isnan(out[r]).any(), (ix, iy, iz)
            err = max(err, float(np.abs(out[r] - ref[ix, iy, iz]).max()))
print("tiles", ntiles, "MMAs (N=%d)" % (3 * NS), mma, "vs unfolded", 27 * ((lmax + 127) // 128), "(N=%d)" % NS, "max abs err", err)
assert err < 1e-3
