"""Stand-alone timing of the first-layer kernels at the LA production shape (CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from bcp_b200._native import LIB, ptr, i3, stream
from bcp_b200.ops import cb8_shape

dev = torch.device("cuda:0")
n, dims, k = 4, (112, 112, 80), (3, 3, 3)
x = torch.randn(n, 1, *dims, device=dev)
w = torch.randn(16, 1, 3, 3, 3, device=dev) / 5
b = torch.randn(16, device=dev) / 10
out = torch.empty(cb8_shape(n, 16, *dims), dtype=torch.bfloat16, device=dev)
dy = torch.randn(cb8_shape(n, 16, *dims), device=dev).to(torch.bfloat16)
dw = torch.zeros_like(w)
ws = torch.empty(LIB.query("bcp_conv_first_wgrad_workspace_floats", n, 16, i3(*dims), i3(*k)), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("tma chunks", LIB.query("bcp_conv_first_wgrad_tma_chunks", n, 16, i3(*dims), i3(*k)))


def timeit(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


fwd = lambda: LIB.call("bcp_conv_first_fwd", ptr(x), ptr(w), ptr(b), ptr(out), n, 16, i3(*dims), i3(*k), stream())
wg = lambda: LIB.call("bcp_conv_first_wgrad", ptr(x), ptr(dy), ptr(dw), ptr(ws), n, 16, i3(*dims), i3(*k), 0, stream())
for name, fn in (("first_fwd", fwd), ("first_wgrad(+finalize)", wg)):
    fn(); torch.cuda.synchronize()
    med, best = timeit(fn)
    print(f"{name:28s} median {med:7.1f} us  best {best:7.1f} us")
