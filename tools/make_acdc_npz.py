"""A stand-in ACDC tree of seeded .npz slices (image fp32 / label uint8, varying sizes) for exercising
``code/ACDC_BCP_train.py --synthetic 0`` where the real h5 data is absent: <root>/train_slices.list, <root>/data/slices/*.npz."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _synth import acdc_slices

root, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 96
os.makedirs(os.path.join(root, "data", "slices"), exist_ok=True)
names = ["patient%03d_frame01_slice_%d" % (i // 8, i % 8) for i in range(n)]
for name, (im, lb) in zip(names, acdc_slices(n, 11)):
    np.savez(os.path.join(root, "data", "slices", name + ".npz"), image=im, label=lb)
with open(os.path.join(root, "train_slices.list"), "w") as f:
    f.write("\n".join(names) + "\n")
print("wrote", n, "slices under", root)
