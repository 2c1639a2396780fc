import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import bcp_oracle as O
from bcp_b200 import ops
dev = torch.device("cuda:0")
for seed, shape in ((1, (2, 48, 40, 36)), (2, (1, 112, 112, 80)), (3, (4, 48, 48, 48))):
    lab = O.synthetic_labels(shape, seed)
    noise = torch.from_numpy((np.random.RandomState(seed).random_sample(shape) > 0.97).astype(np.int64))
    seg = ((lab + noise) > 0).long()
    for conn in (1, 2, 3):
        ref = O.largest_cc(seg, conn).numpy()
        for rep in range(3):
            got = ops.largest_cc(seg.to(torch.uint8).to(dev), connectivity=conn).cpu().numpy().astype(np.float32)
            mism = int((got != ref).sum())
            print(f"seed {seed} shape {shape} conn {conn} rep {rep}: ref vox {int(ref.sum())} got vox {int(got.sum())} mismatch {mism} "
                  f"got-not-ref {int(((got == 1) & (ref == 0)).sum())} ref-not-got {int(((got == 0) & (ref == 1)).sum())}", flush=True)
