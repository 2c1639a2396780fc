"""Host throughput of the ACDC SliceLoader (24 slices of ~230x230 -> 256x256 per batch) with 0 / 4 / 8 / 12 resampling workers."""
import os, sys, time, random
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _synth import acdc_slices
from bcp_b200.dataloaders import dataset as P

slices = acdc_slices(192, 11)
for workers in (0, 4, 8, 12):
    db = P.BaseDataSets(split="train", transform=P.RandomGenerator((256, 256)), slices=slices)
    sampler = P.TwoStreamBatchSampler(list(range(48)), list(range(48, 192)), 24, 12)
    loader = P.SliceLoader(db, sampler, pin=False, prefetch=True, workers=workers)
    np.random.seed(1); random.seed(1)
    n, t0 = 0, None
    for _ in range(6):
        for batch in loader:
            if t0 is None:
                t0 = time.time()          # first batch = pool start-up
            else:
                n += 1
    dt = time.time() - t0
    loader.close()
    print(f"workers={workers:2d}: {n / dt:6.1f} batches/s = {24 * n / dt:7.0f} slices/s", flush=True)
