"""Launch one forward conv layer a few times (for `ncu -k regex:<kernel> --launch-skip 2 -c 1` captures).
usage: python tools/run_conv_once.py C N X Y Z [fold|std|wgrad]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bcp_b200 import ops  # noqa: E402
from tests.test_gpu_primitives import _packs  # noqa: E402
from tests.util import cb8_from_planar  # noqa: E402

c, n, X, Y, Z = (int(v) for v in sys.argv[1:6])
mode = sys.argv[6] if len(sys.argv) > 6 else "fold"
dev = torch.device("cuda:0")
torch.manual_seed(0)
x = torch.randn(n, c, X, Y, Z, device=dev).to(torch.bfloat16).float()
w = (torch.randn(c, c, 3, 3, 3, device=dev) / np.sqrt(c * 27)).to(torch.bfloat16).float()
b = 0.1 * torch.randn(c, device=dev)
pack = _packs(ops, dev, w, (0, 1))
a = cb8_from_planar(x)
ops._TC_FOLD = mode == "fold"
for _ in range(4):
    if mode == "wgrad":
        ops._wgrad(a, a, c, c, (X, Y, Z), (3, 3, 3), (1, 1, 1), (1, 1, 1), w.shape)
    else:
        ops._conv_same(a, pack.k[0], b, c, (3, 3, 3))
torch.cuda.synchronize()
print("done", mode, c, n, (X, Y, Z))
