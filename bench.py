#!/usr/bin/env python
"""bench.py -- BCP self-training step throughput on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W                       # this repo's sm_100a path, LA V-Net (BASELINE configs[1])
    python bench.py --workload acdc | pancreas ...                      # BASELINE configs[2] / configs[4]
    python bench.py --impl reference --gpus N --steps K --warmup W      # the reference algorithm's CPU path (oracle port)
    python bench.py --impl cudnn ...                                    # the same reference modules on the GPU, stock PyTorch/cuDNN

A "step" is one BCP self-training step (LA_BCP_train.py:234-270 / ACDC_BCP_train.py:354-390 / train_pancreas.py:144-174):
teacher forward on the unlabeled half, pseudo labels + largest-CC, bidirectional copy-paste mix, student forward/backward
on the mixed samples, masked Dice+CE, SGD/Adam and the EMA teacher update.  "patches" ("slices") = mixed samples through
the student per step.  Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement".
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

# SURVEY.md section 8(d) / BASELINE.md section 2: conv MACs x2, teacher forward + student forward/backward per step
WORKLOADS = {
    "la": dict(metric="LA V-Net BCP train-step patches/sec", unit="patches/s", shape=(112, 112, 80), batch=8, labeled=4,
               units_per_step=4, flop_per_step=1280.0e9, weights="la10",
               desc="LA V-Net BCP self-train step: per GPU 8 loaded volumes 112x112x80 (4 labeled + 4 unlabeled), 4 mixed "
                    "student patches (BASELINE configs[1]); weights = shipped LA_10.pth rounded to bf16; SGD + EMA"),
    "acdc": dict(metric="ACDC U-Net BCP train-step slices/sec", unit="slices/s", shape=(256, 256), batch=24, labeled=12,
                 units_per_step=12, flop_per_step=282.9e9, weights="acdc10",
                 desc="ACDC 2-D U-Net BCP self-train step: per GPU 24 loaded slices 256x256 (12 labeled + 12 unlabeled), 12 "
                      "mixed student slices (BASELINE configs[2]); weights = shipped ACDC_10.pth rounded to bf16; SGD + "
                      "state_dict EMA"),
    "pancreas": dict(metric="Pancreas V-Net BCP train-step patches/sec", unit="patches/s", shape=(96, 96, 96), batch=8, labeled=4,
                     units_per_step=4, flop_per_step=1128.5e9, weights=None,
                     desc="Pancreas V-Net (InstanceNorm) BCP self-train step: per GPU 8 loaded volumes 96^3 (batch 2 per "
                          "stream), 4 mixed student patches (BASELINE configs[4]); default-init weights under "
                          "torch.manual_seed(2020); Adam + EMA"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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


# ------------------------------------------------------------------------------------------------------------
# synthetic data / weights (identical for every arm)
# ------------------------------------------------------------------------------------------------------------
def synthetic_batch(wl, rank):
    """Images that carry their labels (blobs + noise), numpy streams seeded per rank: volume fp32 [B,1,...], label uint8."""
    from tests.golden.golden_common import synthetic_scene
    w = WORKLOADS[wl]
    vol, lab = synthetic_scene(w["batch"], w["shape"], 9000 + 17 * rank, n_classes=4 if wl == "acdc" else 2,
                               kind="rand" if wl == "acdc" else "randn")
    return vol.contiguous(), lab.to(torch.uint8).contiguous()


def load_weights(wl):
    from tests.golden.golden_common import unpack_weights_bf16
    tag = WORKLOADS[wl]["weights"]
    if tag is None:
        return None
    return unpack_weights_bf16(np.load(os.path.join(ROOT, "tests", "golden", "weights_%s_bf16.npz" % tag)))


def build_native(wl, dev):
    from bcp_b200.networks.net_factory import net_factory, BCP_net
    from bcp_b200.optim import FusedSGD_EMA, FusedAdam_EMA
    sd = load_weights(wl)
    if wl == "la":
        model, ema = net_factory("VNet", 1, 2, "train"), net_factory("VNet", 1, 2, "train")
    elif wl == "acdc":
        model, ema = BCP_net(1, 4), BCP_net(1, 4, ema=True)
    else:
        from bcp_b200.pancreas.Vnet import VNet
        torch.manual_seed(2020)
        model, ema = VNet().to(dev), VNet().to(dev)
    for p in ema.parameters():
        p.detach_()
    if sd is not None:
        model.load_state_dict(sd)
    ema.load_state_dict(model.state_dict())
    model.train()
    ema.train()
    if wl == "pancreas":
        opt = FusedAdam_EMA(model, ema, lr=1e-3, ema_alpha=0.99)
    else:
        opt = FusedSGD_EMA(model, ema, lr=0.01, momentum=0.9, weight_decay=1e-4, ema_alpha=0.99,
                           ema_mode="params" if wl == "la" else "state_dict")
    return model, ema, opt


def eager_step(wl, model, ema, opt, vol, lab):
    from bcp_b200 import step as S
    if wl == "la":
        return S.la_self_train_step(model, ema, opt, vol, lab)
    if wl == "acdc":
        return S.acdc_self_train_step(model, ema, opt, vol, lab, labeled_bs=12)
    n = vol.shape[0] // 4
    return S.pan_self_train_step(model, ema, opt, vol[:n], lab[:n], vol[n:2 * n], lab[n:2 * n], vol[2 * n:3 * n], vol[3 * n:])


def build_oracle(wl, dev):
    """The reference's modules (oracle restatement, fp32) + stock torch optimiser, on `dev`."""
    from oracle import bcp_oracle as O
    sd = load_weights(wl)
    if wl == "la":
        model, ema = O.net_factory("VNet", 1, 2, "train"), O.net_factory("VNet", 1, 2, "train")
    elif wl == "acdc":
        model, ema = O.BCP_net(1, 4), O.BCP_net(1, 4, ema=True)
    else:
        torch.manual_seed(2020)
        model, ema = O.OraclePanVNet(), O.OraclePanVNet()
    for p in ema.parameters():
        p.detach_()
    if sd is not None:
        model.load_state_dict(sd)
    ema.load_state_dict(model.state_dict())
    model, ema = model.to(dev).train(), ema.to(dev).train()
    if wl == "pancreas":
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    else:
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=0.0001)
    rs = np.random.RandomState(1337)

    def step(vol, lab):
        if wl == "la":
            return O.la_self_train_step(model, ema, opt, vol, lab.long(), rng=rs)
        if wl == "acdc":
            return O.acdc_self_train_step(model, ema, opt, vol, lab, labeled_bs=12, rng=rs)
        n = vol.shape[0] // 4
        l = lab.long()
        return O.pan_self_train_step(model, ema, opt, vol[:n], l[:n], vol[n:2 * n], l[n:2 * n], vol[2 * n:3 * n], vol[3 * n:], rng=rs)
    return step


# ------------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch.distributed as dist
    from bcp_b200._native import LIB
    wl = args.workload
    W = WORKLOADS[wl]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, ema, opt = build_native(wl, dev)
    vol_h, lab_h = synthetic_batch(wl, rank)
    vol_h, lab_h = vol_h.pin_memory(), lab_h.pin_memory()
    vol_d, lab_d = vol_h.to(dev), lab_h.to(dev)
    np.random.seed(1337 + rank)
    warm = max(args.warmup, 3)

    if args.profile:          # under ncu: one warm-up + one eager step, nothing else (numbers printed here are NOT bench values)
        eager_step(wl, model, ema, opt, vol_d, lab_d)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("profiled_step")
        eager_step(wl, model, ema, opt, vol_d, lab_d)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return

    if os.environ.get("BENCH_SERIAL_TEACHER", "0") == "1":          # A/B switches: teacher pass / weight gradients on the main stream
        import bcp_b200.step as _S
        _S.TEACHER_ON_SIDE_STREAM = False
    if os.environ.get("BENCH_SERIAL_WGRAD", "0") == "1":
        import bcp_b200.ops as _O
        _O.WGRAD_ON_SIDE_STREAM = False
    graphed, graph_note = None, "eager launches (BENCH_NO_GRAPH=1)"
    if os.environ.get("BENCH_NO_GRAPH", "0") != "1":
        from bcp_b200.graph import GraphedStep
        kw = dict(labeled_bs=W["labeled"]) if wl != "pancreas" else {}
        graphed = GraphedStep({"la": "la", "acdc": "acdc", "pancreas": "pan"}[wl], model, ema, opt, (W["batch"], 1) + W["shape"], **kw)
        graphed.load(vol_d, lab_d)
        graph_note = "whole step = one CUDA graph replay (%d kernels)" % graphed.kernels_per_replay

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(loop):
        """CUDA-event time of loop() on the launching stream, max over ranks; barrier + synchronize on both sides."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        out = loop()
        host_ms[0] = (time.perf_counter() - h0) * 1e3
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        barrier()
        return ms, out

    def resident_loop(steps):
        def loop():
            r = None
            for _ in range(steps):
                r = graphed.step() if graphed is not None else eager_step(wl, model, ema, opt, vol_d, lab_d)
            return r
        return loop

    def e2e_loop(steps):
        """The call a user makes: every step's batch comes from pinned host memory (H2D inside the timed region; the
        NEXT step's copy is issued on the copy stream before this step's loss is read, so it overlaps the replay) and the
        step's loss is read back to the host (D2H + sync) every step."""
        def loop():
            loss = None
            if graphed is not None:
                graphed.load(vol_h, lab_h)
                for i in range(steps):
                    r = graphed.step()
                    if i + 1 < steps:
                        graphed.load(vol_h, lab_h)
                    loss = float(r["loss"])
            else:
                for i in range(steps):
                    r = eager_step(wl, model, ema, opt, vol_h.to(dev, non_blocking=True), lab_h.to(dev, non_blocking=True))
                    loss = float(r["loss"])
            return loss
        return loop

    timed(resident_loop(warm))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = LIB.launches
    ms, r = timed(resident_loop(args.steps))
    launches = (LIB.launches - l0) if graphed is None else graphed.kernels_per_replay * args.steps
    host_enqueue_ms = host_ms[0] / args.steps
    clocks = sampler.stop() if rank == 0 else None
    loss = float(r["loss"])
    ms_e2e, _ = timed(e2e_loop(args.steps))

    pk, pk_src = peaks()
    roof = None
    if wl != "acdc" or True:
        roof = conv_roofline(LIB, lambda: eager_step(wl, model, ema, opt, vol_d, lab_d))
    if roof is not None:
        roof.update({"bound": "tensor", "unit": "TFLOP/s", "peak": pk["bf16_tflops"], "frac": roof["achieved"] / pk["bf16_tflops"],
                     "peak_kind": "burst dense bf16 (kernels are event-timed one by one, not inside a power-capped long run)",
                     "peak_sustained": pk["bf16_tflops_sustained"], "frac_of_sustained": roof["achieved"] / pk["bf16_tflops_sustained"],
                     "peak_source": pk_src})
    units = world * W["units_per_step"] * args.steps
    value = units / (ms / 1e3)
    e2e = units / (ms_e2e / 1e3)
    step_tflops = world * W["flop_per_step"] * args.steps / (ms / 1e3) / 1e12
    line = {
        "metric": W["metric"], "value": value, "unit": W["unit"], "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": W["desc"], "parallelism": "dp%d" % world, "launch": graph_note,
                   "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; no explicit flush"},
        "loss": loss, "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks,
        "e2e": {"value": e2e, "unit": W["unit"], "h2d_bytes_per_step": vol_h.numel() * 4 + lab_h.numel(),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                "note": "pinned-host batch copied H2D every step on a copy stream (overlaps the previous replay), loss read back every step"},
        "step_flops": W["flop_per_step"],
        "step_tensor": {"achieved_tflops": step_tflops / world, "frac_of_burst": step_tflops / world / pk["bf16_tflops"],
                        "frac_of_sustained": step_tflops / world / pk["bf16_tflops_sustained"]},
        "roofline": roof,
    }
    hb = hbm_block(wl, ms / args.steps, pk)
    if hb is not None:
        line["hbm"] = hb
    if rank == 0:
        if world == 1 and not args.no_baselines:
            line["gpu_baseline"] = gpu_baseline_sample(wl, dev, vol_d, lab_d)
            line["cpu_baseline"] = cpu_baseline_sample(wl, vol_h, lab_h)
        print(json.dumps(line), flush=True)
    shutdown(graphed, world)


def shutdown(graphed, world):
    """Tear the communicator down properly: the captured graph holds the NCCL all-reduce node, so release it first.  A
    watchdog ends the process if the teardown does not return (it hung in round 1 with the graph still alive)."""
    if world <= 1:
        return
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(30.0, lambda: os._exit(0)).start()
    if graphed is not None:
        graphed.graph.reset()
        del graphed.graph
    torch.cuda.synchronize()
    dist.destroy_process_group()
    os._exit(0)          # timers / sampler threads must not keep the rank alive


def conv_roofline(LIB, step_fn):
    """One instrumented eager step with CUDA events around every tcgen05 conv launch (on the launching stream)."""
    names = ("bcp_conv_tc_fwd", "bcp_conv_tc_fold_fwd", "bcp_conv_tc_wgrad")
    recs = []
    orig = LIB.call

    def wrapped(name, *a):
        if name in names:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            orig(name, *a)
            e1.record()
            # fwd args: (in, wpack, bias, out, n, cin, cout, dims, kernel, stream); wgrad: (a, dy, dw, ws, counter, n, ...)
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
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp))
        ks = [t[k] for k in t if k.startswith("conv_tc") and isinstance(t[k], dict) and "dram_bytes_per_launch" in t[k]]
        if ks:
            traffic = sum(k["dram_bytes_per_launch"] * k["launches"] for k in ks) / sum(k["launches"] for k in ks)
            tsrc = t.get("_source", "profiles/conv_traffic.json (ncu dram__bytes_read+write, mean over the tcgen05 conv launches of one step)")
    alg_bytes = sum(r[6] for r in recs) / len(recs)
    return {"kernel": "tcgen05 implicit-GEMM conv family (conv_tc fwd/dgrad incl. dz-folded, conv_tc_wgrad)", "launches": len(recs),
            "avg_launch_ms": tot_ms / len(recs), "achieved": tot_fl / (tot_ms / 1e3) / 1e12, "traffic": traffic, "traffic_source": tsrc,
            "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flop": tot_fl,
            "per_shape_tflops": {k: round(v[1] / (v[0] / 1e3) / 1e12, 1) for k, v in per.items()},
            "kernel_ms_per_step": tot_ms}


def hbm_block(wl, ms_per_step, pk):
    """Whole-step DRAM traffic (ncu dram__bytes_read+write summed over one step's launches, committed under profiles/)
    divided by the measured step time, against the measured copy bandwidth."""
    p = os.path.join(ROOT, "profiles", "step_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(wl)
    if not t:
        return None
    gbs = t["dram_bytes_per_step"] / (ms_per_step / 1e3) / 1e9
    return {"bytes_per_step": t["dram_bytes_per_step"], "achieved_gbs": gbs, "peak_gbs": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"],
            "source": t.get("source")}


# ------------------------------------------------------------------------------------------------------------
# baselines: the reference's modules through stock PyTorch, on the GPU (cuDNN) and on the host CPU
# ------------------------------------------------------------------------------------------------------------
def gpu_baseline_sample(wl, dev, vol_d, lab_d):
    """the reference modules (oracle restatement, pinned to the reference) on the SAME B200 through
    stock PyTorch/cuDNN: fp32 as shipped (TF32 off) and bf16 autocast + cudnn.benchmark.  Bounded sample: 1 warm-up + 3
    timed steps each, same batch; the reference's CPU largest-CC (D2H, scipy/skimage, H2D) is part of its step."""
    W = WORKLOADS[wl]
    out = {"unit": W["unit"], "sample": "1 warm-up + 3 timed steps per mode, same synthetic batch and weights"}
    for mode in ("fp32_cudnn", "bf16_autocast_cudnn"):
        try:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.benchmark = (mode != "fp32_cudnn")
            step = build_oracle(wl, dev)

            def run():
                if mode == "fp32_cudnn":
                    return step(vol_d, lab_d)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return step(vol_d, lab_d)
            run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                r = run()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            out[mode] = {"value": W["units_per_step"] / dt, "ms_per_step": dt * 1e3, "loss": float(r["loss"])}
        except Exception as exc:            # a baseline must never take the measurement down
            out[mode] = {"error": str(exc).splitlines()[0][:160]}
        finally:
            torch.backends.cudnn.benchmark = False
            torch.cuda.empty_cache()
    return out


def set_cpu_threads():
    """All the host threads the process may run on (its CPU affinity mask, which honours cpusets); BCP_CPU_THREADS overrides."""
    n = int(os.environ.get("BCP_CPU_THREADS", "0"))
    if n <= 0:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        n = max(1, min(n, 64))        # PyTorch's conv/BN kernels stop scaling (and start thrashing) long before 128 threads
    torch.set_num_threads(n)


def cpu_baseline_sample(wl, vol_h, lab_h):
    """The reference algorithm's CPU path (oracle port, fp32) on the box's host cores: full batch, 1 warm-up + timed steps
    bounded to ~20 s."""
    W = WORKLOADS[wl]
    set_cpu_threads()
    step = build_oracle(wl, torch.device("cpu"))
    t0 = time.time()
    step(vol_h, lab_h)
    t_first = time.time() - t0
    n = max(1, min(3, int(20.0 / max(t_first, 1e-3))))
    t0 = time.time()
    for _ in range(n):
        step(vol_h, lab_h)
    dt = (time.time() - t0) / n
    return {"value": W["units_per_step"] / dt, "unit": W["unit"], "cores": torch.get_num_threads(), "kind": "port",
            "sample": "full batch, 1 warm-up + %d timed step(s) of the fp32 PyTorch CPU oracle, %.1f s/step" % (n, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    W = WORKLOADS[wl]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    set_cpu_threads()
    step = build_oracle(wl, torch.device("cpu"))
    vol, lab = synthetic_batch(wl, 0)
    budget = 170.0                              # seconds for warm-up + timed steps: the whole run stays within a few minutes
    t_start = time.time()
    t0 = time.time()
    step(vol, lab)
    t_first = time.time() - t0
    n_warm = 1
    while n_warm < args.warmup and (time.time() - t_start) + t_first < 0.25 * budget:
        step(vol, lab)
        n_warm += 1
    left = budget - (time.time() - t_start)
    n_eff = max(1, min(args.steps, int(left / max(t_first, 1e-3))))
    t0 = time.time()
    for _ in range(n_eff):
        step(vol, lab)
    dt = (time.time() - t0) / n_eff
    v = W["units_per_step"] / dt
    line = {"impl": "reference", "metric": W["metric"], "value": v, "unit": W["unit"],
            "n_gpus": world, "steps": n_eff, "warmup": n_warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": W["desc"], "parallelism": "cpu",
                       "launch": "the reference's algorithm (oracle port of its PyTorch modules) on the host CPU, rank 0 only"},
            "cpu_baseline": {"value": v, "unit": W["unit"], "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "%d timed step(s) after %d warm-up (time-bounded; %d/%d requested), full batch"
                                       % (n_eff, n_warm, args.steps, args.warmup)},
            "e2e": {"value": v, "unit": W["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_cudnn(args):
    """Stand-alone GPU comparator: the reference modules through stock PyTorch/cuDNN on one B200 (rank 0)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = args.workload
    W = WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    vol, lab = synthetic_batch(wl, 0)
    res = gpu_baseline_sample(wl, dev, vol.to(dev), lab.to(dev))
    best = max((v["value"] for v in res.values() if isinstance(v, dict) and "value" in v), default=None)
    print(json.dumps({"impl": "cudnn", "metric": W["metric"], "value": best, "unit": W["unit"], "n_gpus": 1, "higher_is_better": True,
                      "config": {"workload": W["desc"], "parallelism": "dp1", "launch": "stock PyTorch eager, cuDNN"},
                      "gpu_baseline": res}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "cudnn"])
    ap.add_argument("--workload", default="la", choices=sorted(WORKLOADS))
    ap.add_argument("--no-baselines", "--no-cpu-baseline", dest="no_baselines", action="store_true",
                    help="skip the cuDNN and CPU baseline samples (N=1 default run only)")
    ap.add_argument("--profile", action="store_true", help="one warm-up + one eager step only (for ncu launch lists)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "cudnn":
        run_cudnn(a)
    else:
        run_native(a)
