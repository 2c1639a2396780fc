"""All-reduce of the flat gradient arena (9.45 M fp32) under different NCCL settings: time per call, CUDA events, max over
ranks.  torchrun --nproc-per-node N tools/nccl_probe.py [tag]"""
import os
import sys
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 9_450_000
g = torch.randn(n, device="cuda")
for _ in range(5):
    dist.all_reduce(g)
torch.cuda.synchronize()
res = {}
for label, graphed in (("eager", False), ("graph", True)):
    if graphed:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                dist.all_reduce(g)
        torch.cuda.synchronize()
        run = gr.replay
    else:
        run = lambda: dist.all_reduce(g)
    for _ in range(3):
        run()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
        g.mul_(0.5)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[label] = float(t)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    g.mul_(0.5)
e1.record(); torch.cuda.synchronize()
base = e0.elapsed_time(e1) / 20
if rank == 0:
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    print(f"NCCLPROBE world={world} {tag:40s} eager {res['eager'] - base:.3f} ms  graph {res['graph'] - base:.3f} ms  "
          f"(mul alone {base:.3f})  algbw {n * 4 / (res['graph'] - base) / 1e6:.0f} GB/s", flush=True)
if 'graph' in res:
    del gr
dist.destroy_process_group()
