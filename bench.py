#!/usr/bin/env python
"""bench.py -- LA V-Net BCP self-training step throughput (patches/s) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # the reference algorithm's CPU path (oracle port)

A "step" is one BCP self-training step (LA_BCP_train.py:234-270): teacher forward on 4 unlabeled volumes, pseudo
labels + largest-CC, bidirectional copy-paste mix, student forward/backward on the 4 mixed 112x112x80 patches,
masked Dice+CE, SGD and the EMA teacher update.  "patches" = mixed volumes through the student (4 per step per GPU).
Prints ONE JSON line (rank 0).  See the task contract in DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = (112, 112, 80)
FLOP_PER_STEP = 1280.0e9          # 4 teacher fwd + 4 student fwd+bwd, conv MACs x2
PATCHES_PER_STEP = 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def synthetic_batch(rank, gen_device):
    g = torch.Generator(device="cpu").manual_seed(1337 + rank)
    vol = torch.randn((8, 1) + SHAPE, generator=g, dtype=torch.float32)
    # blobby labels: threshold a box-filtered noise field
    noise = torch.randn((8, 1) + SHAPE, generator=g)
    sm = torch.nn.functional.avg_pool3d(noise, 5, stride=1, padding=2)
    lab = (sm[:, 0] > sm.std()).to(torch.uint8)
    return vol, lab


def build_native(dev):
    from bcp_b200.networks.net_factory import net_factory
    from bcp_b200.optim import FusedSGD_EMA
    torch.manual_seed(1337)
    model = net_factory("VNet", 1, 2, "train")
    ema = net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    opt = FusedSGD_EMA(model, ema, lr=0.01, momentum=0.9, weight_decay=1e-4, ema_alpha=0.99, ema_mode="params")
    return model, ema, opt


def run_native(args):
    import torch.distributed as dist
    from bcp_b200._native import LIB
    from bcp_b200.step import la_self_train_step
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, ema, opt = build_native(dev)
    vol_h, lab_h = synthetic_batch(rank, dev)
    vol_h, lab_h = vol_h.pin_memory(), lab_h.pin_memory()
    vol_d, lab_d = vol_h.to(dev), lab_h.to(dev)
    np.random.seed(1337 + rank)

    graphed, graph_note = None, "eager (BENCH_NO_GRAPH=1)"
    if os.environ.get("BENCH_NO_GRAPH", "0") != "1" and not args.profile:
        try:
            from bcp_b200.graph import GraphedLAStep
            graphed = GraphedLAStep(model, ema, opt, (8, 1) + SHAPE)
            graphed.load(vol_d, lab_d)
            graph_note = "whole step captured in one CUDA graph (%d kernels per replay)" % graphed.kernels_per_replay
        except Exception as exc:                 # capture problems must not hide a measurement: fall back to eager launches
            graphed, graph_note = None, "eager (graph capture failed: %s)" % (str(exc).splitlines()[0][:120])
            torch.cuda.synchronize()

    def step_resident():
        if graphed is not None:
            return graphed.replay_resident()
        return la_self_train_step(model, ema, opt, vol_d, lab_d)

    def step_e2e():
        if graphed is not None:
            r = graphed(vol_h, lab_h)            # H2D of the step's inputs from pinned memory, then one graph replay
        else:
            r = la_self_train_step(model, ema, opt, vol_h.to(dev, non_blocking=True), lab_h.to(dev, non_blocking=True))
        return float(r["loss"].cpu())            # D2H read of the step's result

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        h0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        host_ms[0] = (time.perf_counter() - h0) * 1e3 / steps        # CPU time to ENQUEUE a step (no sync inside)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        barrier()
        return ms, out

    if args.profile:          # under ncu: one warm-up + one step, nothing else (numbers printed here are NOT bench values)
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("profiled_step")
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    for _ in range(max(args.warmup, 3)):
        r = step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = LIB.launches
    ms, r = timed(step_resident, args.steps)
    launches = (LIB.launches - l0) if graphed is None else graphed.kernels_per_replay * args.steps
    host_enqueue_ms = host_ms[0]
    clocks = sampler.stop() if rank == 0 else None
    loss = float(r["loss"])
    ms_e2e, _ = timed(step_e2e, args.steps)

    # roofline of the dominant kernel family (convolutions): one instrumented extra step with CUDA events around
    # every conv launch on the launching stream; algorithmic FLOPs = 2*MACs of that launch.
    roof = conv_roofline(LIB, lambda: la_self_train_step(model, ema, opt, vol_d, lab_d))
    pk, pk_src = peaks()
    if roof is not None:
        peak = pk["bf16_tflops_sustained"]
        roof.update({"bound": "tensor", "peak": peak, "unit": "TFLOP/s", "frac": roof["achieved"] / peak, "peak_source": pk_src + " (sustained)"})
    value = world * PATCHES_PER_STEP * args.steps / (ms / 1e3)
    e2e = world * PATCHES_PER_STEP * args.steps / (ms_e2e / 1e3)
    line = {
        "metric": "LA V-Net BCP train-step patches/sec", "value": value, "unit": "patches/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "LA V-Net BCP self-train step: per GPU 8 loaded volumes 112x112x80 (4 labeled + 4 unlabeled), "
                               "4 mixed student patches (BASELINE configs[1]); random-init weights",
                   "parallelism": "dp%d" % world, "per_gpu_student_patches": 4, "launch": graph_note,
                   "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2; no explicit flush"},
        "loss": loss, "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks,
        "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": vol_h.numel() * 4 + lab_h.numel(),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "step_flops": FLOP_PER_STEP, "step_tflops": world * FLOP_PER_STEP * args.steps / (ms / 1e3) / 1e12,
        "roofline": roof,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        # Leave without tearing the communicator down: destroy_process_group() on a communicator whose all-reduce sits
        # inside a live CUDA graph never returned on the 2-GPU box (the JSON line had long been printed).  Everything
        # measured is already reported; synchronise, meet at a barrier, and exit the process directly.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def conv_roofline(LIB, step_fn):
    names = ("bcp_conv_tc_fwd", "bcp_conv_tc_fold_fwd", "bcp_conv_tc_wgrad")
    recs = []
    orig = LIB.call

    def wrapped(name, *a):
        if name in names:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig(name, *a)
            e1.record()
            # args: (in, wpack, bias, out, n, cin, cout, dims, kernel, stream) for fwd; wgrad has (.., ws, counter, n, ...)
            o = 5 if name == "bcp_conv_tc_wgrad" else 4
            n, cin, cout, dims, kernel = a[o], a[o + 1], a[o + 2], a[o + 3], a[o + 4]
            vox = n * dims[0] * dims[1] * dims[2]
            taps = kernel[0] * kernel[1] * kernel[2]
            # algorithmic bytes: both activation tensors once (bf16) + the weights (bf16 pack or fp32 gradient)
            ab = 2.0 * vox * (cin + cout) + taps * cin * cout * (4.0 if name == "bcp_conv_tc_wgrad" else 2.0)
            recs.append((name, e0, e1, 2.0 * vox * taps * cin * cout, cin, cout, ab))
        else:
            orig(name, *a)
    LIB.call = wrapped
    try:
        step_fn()
        torch.cuda.synchronize()
    finally:
        LIB.call = orig
    if not recs:
        return None
    tot_ms = sum(r[1].elapsed_time(r[2]) for r in recs)
    tot_fl = sum(r[3] for r in recs)
    per = {}
    for name, e0, e1, f, cin, cout, _ in recs:
        k = "%s_c%d_%d" % (name.replace("bcp_conv_", "").replace("tc_fold_fwd", "tc_fwd"), cin, cout)
        d = per.setdefault(k, [0.0, 0.0, 0])
        d[0] += e0.elapsed_time(e1)
        d[1] += f
        d[2] += 1
    # DRAM bytes per launch of the same kernels from the committed ncu capture (profiles/, tools/traffic_from_ncu.py);
    # null when no capture has been committed for this build
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        ks = [t[k] for k in ("conv_tc_kernel", "conv_tc_wgrad_kernel") if k in t]
        if ks:
            traffic = sum(k["dram_bytes_per_launch"] * k["launches"] for k in ks) / sum(k["launches"] for k in ks)
            tsrc = "profiles/conv_traffic.json (ncu dram__bytes_read+write, mean over the conv_tc + conv_tc_wgrad launches of one step)"
    alg_bytes = sum(r[6] for r in recs) / len(recs)
    return {"kernel": "conv_tc + conv_tc_wgrad (tcgen05 implicit GEMM: fwd, dgrad, wgrad)", "launches": len(recs),
            "avg_launch_ms": tot_ms / len(recs),
            "achieved": tot_fl / (tot_ms / 1e3) / 1e12, "traffic": traffic, "traffic_source": tsrc,
            "algorithmic_bytes_per_launch": alg_bytes,
            "per_shape_tflops": {k: round(v[1] / (v[0] / 1e3) / 1e12, 1) for k, v in per.items()},
            "kernel_ms_per_step": tot_ms}


def set_cpu_threads():
    """All the host threads the process may run on (its CPU affinity mask, which honours cpusets; os.cpu_count() counts
    the machine's logical CPUs and oversubscribed the 128-thread GPU box); BCP_CPU_THREADS overrides."""
    n = int(os.environ.get("BCP_CPU_THREADS", "0"))
    if n <= 0:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        n = max(1, min(n, 64))        # PyTorch's conv/BN kernels stop scaling (and start thrashing) long before 128 threads
    torch.set_num_threads(n)


def oracle_step_runner():
    """The reference algorithm's CPU path (oracle/bcp_oracle.py: fp32 PyTorch restatement pinned to the reference)."""
    from oracle import bcp_oracle as O
    set_cpu_threads()
    torch.manual_seed(1337)
    model, ema = O.net_factory("VNet", 1, 2, "train"), O.net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    vol, lab = synthetic_batch(0, None)
    rs = np.random.RandomState(1337)

    def step():
        return O.la_self_train_step(model, ema, opt, vol, lab.long(), rng=rs)
    return step


def cpu_baseline_sample():
    """Bounded sample of the same workload: ONE self-training step at half the batch (labeled_bs 2: 4 loaded volumes,
    2 mixed student patches instead of 8 / 4), so the default bench run stays within a few minutes on the host CPU."""
    from oracle import bcp_oracle as O
    set_cpu_threads()
    torch.manual_seed(1337)
    model, ema = O.net_factory("VNet", 1, 2, "train"), O.net_factory("VNet", 1, 2, "train")
    for p in ema.parameters():
        p.detach_()
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    vol, lab = synthetic_batch(0, None)
    idx = [0, 2, 4, 6]                               # one volume of each of the four roles (img_a, img_b, unimg_a, unimg_b)
    vol, lab = vol[idx].contiguous(), lab[idx].long().contiguous()
    rs = np.random.RandomState(1337)
    t0 = time.time()
    O.la_self_train_step(model, ema, opt, vol, lab, labeled_bs=2, rng=rs)
    dt = time.time() - t0
    return {"value": 2.0 / dt, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "1 LA self-train step at half batch (4 volumes 112x112x80, 2 student patches), first call, fp32 PyTorch "
                      "CPU oracle, %.1f s" % dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    step = oracle_step_runner()
    t0 = time.time()
    step()                                      # warm-up (also sizes the sample)
    t_first = time.time() - t0
    budget = 150.0
    n_eff = max(1, min(args.steps, int(budget / max(t_first, 1e-3))))
    t0 = time.time()
    for _ in range(n_eff):
        step()
    dt = (time.time() - t0) / n_eff
    v = PATCHES_PER_STEP / dt
    line = {"impl": "reference", "metric": "LA V-Net BCP train-step patches/sec", "value": v, "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "LA V-Net BCP self-train step: 8 loaded volumes 112x112x80, 4 mixed student patches "
                                   "(BASELINE configs[1]) on the host CPU", "parallelism": "cpu"},
            "cpu_baseline": {"value": v, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d timed step(s) after 1 warm-up (time-bounded to ~150 s of the requested %d)" % (n_eff, args.steps)},
            "e2e": {"value": v, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="one warm-up + one step only (for ncu launch lists)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
