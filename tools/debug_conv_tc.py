"""Diagnostic harness for the tcgen05 conv kernel: compares against the CUDA-core kernel and torch on structured
inputs and prints where (which rows / channels / taps) they differ.  Run under `timeout` on the GPU box."""
import ctypes
import sys
import os
import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bcp_b200 import ops
from bcp_b200._native import LIB, i3
from tests.util import cb8_from_planar, planar_from_cb8, rel_rms
from tests.test_gpu_primitives import _packs

dev = torch.device("cuda:0")


def plan(n, cin, cout, dims, kernel):
    out = (ctypes.c_int * 10)()
    rc = LIB.query("bcp_conv_tc_plan", n, cin, cout, i3(*dims), i3(*kernel), out)
    return rc, list(out)


def run(n, cin, cout, dims, kernel, mode="rand", verbose=True):
    torch.manual_seed(0)
    if mode == "ones":
        x = torch.ones(n, cin, *dims, device=dev)
        w = torch.zeros(cout, cin, *kernel, device=dev)
        w[:, :, kernel[0] // 2, 1, 1] = 1.0 / cin
    elif mode == "tap":
        x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
        w = torch.zeros(cout, cin, *kernel, device=dev)
        for co in range(cout):
            w[co, co % cin, (co // 9) % kernel[0], (co // 3) % 3, co % 3] = 1.0
    else:
        x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
        w = (torch.randn(cout, cin, *kernel, device=dev) / np.sqrt(cin * np.prod(kernel))).to(torch.bfloat16).float()
    b = 0.1 * torch.randn(cout, device=dev)
    pack = _packs(ops, dev, w, (0, 1))
    a = cb8_from_planar(x)
    rc, pl = plan(n, cin, cout, dims, kernel)
    ref = F.conv3d(x, w, b, padding=tuple(k // 2 for k in kernel))
    y_d = planar_from_cb8(ops._conv_same(a, pack.k[0], b, cout, kernel, allow_tc=False), cout)
    y_t = planar_from_cb8(ops._conv_same(a, pack.k[0], b, cout, kernel, allow_tc=True), cout)
    torch.cuda.synchronize()
    e_d, e_t, e_td = rel_rms(y_d, ref), rel_rms(y_t, ref), rel_rms(y_t, y_d)
    print(f"[{mode}] n={n} cin={cin} cout={cout} dims={dims} k={kernel} plan(BX,BY,BZ,MT,SA,SB,AS,nb,cols,smem)={pl} "
          f"direct_vs_torch={e_d:.2e} tc_vs_torch={e_t:.2e} tc_vs_direct={e_td:.2e}", flush=True)
    if e_t > 1e-2 and verbose:
        bad = (y_t - ref).abs() > 0.05 * ref.abs().max()
        print("   bad fraction", float(bad.float().mean()))
        print("   bad by channel", bad.float().mean(dim=(0, 2, 3, 4)).cpu().numpy().round(2))
        print("   bad by x", bad.float().mean(dim=(0, 1, 3, 4)).cpu().numpy().round(2))
        print("   bad by y", bad.float().mean(dim=(0, 1, 2, 4)).cpu().numpy().round(2))
        print("   bad by z", bad.float().mean(dim=(0, 1, 2, 3)).cpu().numpy().round(2))
        print("   sample got", y_t[0, :4, 0, 0, :6].cpu().numpy().round(3))
        print("   sample ref", ref[0, :4, 0, 0, :6].cpu().numpy().round(3))
    return e_t


def run_prof(n, c, dims, kernel=(3, 3, 3), fold=False):
    """Per-role wait cycles of the instrumented forward kernels (bcp_conv_tc_fwd_profiled; fold=True: dz-folded kernel)."""
    from bcp_b200._native import ptr, stream
    torch.manual_seed(0)
    x = torch.randn(n, c, *dims, device=dev).to(torch.bfloat16).float()
    w = (torch.randn(c, c, *kernel, device=dev) / np.sqrt(c * 27)).to(torch.bfloat16).float()
    b = torch.zeros(c, device=dev)
    pack = _packs(ops, dev, w, (0, 1))
    a = cb8_from_planar(x)
    out = (ctypes.c_int * 10)()
    LIB.query("bcp_conv_tc_fold_plan" if fold else "bcp_conv_tc_plan", n, c, c, i3(*dims), i3(*kernel), out)
    pl = list(out)
    ops._TC_FOLD = bool(fold)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(3):
        ops._conv_same(a, pack.k[0], b, c, kernel, allow_tc=True)
    ev[0].record()
    for _ in range(10):
        ops._conv_same(a, pack.k[0], b, c, kernel, allow_tc=True)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 100.0
    flops = 2.0 * n * np.prod(dims) * c * c * 27
    buf = torch.zeros(160 * 16, dtype=torch.int64, device=dev)
    y = torch.empty(ops.cb8_shape(n, c, *dims), dtype=torch.bfloat16, device=dev)
    LIB.call("bcp_conv_tc_fwd_profiled", ptr(a), ptr(pack.k[0]), ptr(b), ptr(y), n, c, c, i3(*dims), i3(*kernel), int(fold),
             buf.data_ptr(), stream())
    torch.cuda.synchronize()
    p = buf.cpu().numpy().reshape(160, 16)
    p = p[p[:, 0] > 0]
    names = ["total", "prod_wait_emptyA", "prod_wait_emptyB", "mma_wait_fullA", "mma_wait_fullB", "mma_wait_tmem_empty",
             "epi_wait_tmem_full", "epi_work", "mma_loop_end", "items"]
    print(f"[prof{'-fold(epilogue warps ' + str(fold) + ')' if fold else ''}] n={n} c={c} dims={dims} plan(BX,BY,BZ,MT,SA,SB|NS*100+TG,AS,nb,cols,smem)={pl} {us:.1f} us "
          f"{flops / us * 1e-6:.1f} TF/s ctas={len(p)}", flush=True)
    print("       " + "  ".join(f"{nm}={p[:, i].mean():.0f}(max {p[:, i].max()})" for i, nm in enumerate(names)), flush=True)


def run_s2(n, c_full, c_half, half):
    """down conv (c_full -> c_half) and transposed conv (c_half -> c_full), fwd + all gradients, tc vs direct vs torch."""
    torch.manual_seed(2)
    full = tuple(2 * h for h in half)
    res = {}
    for tc in (False, True):
        ops._TC_FWD, ops._TC_WGRAD = tc, tc
        # down
        x = torch.randn(n, c_full, *full, device=dev).to(torch.bfloat16).float()
        w = (torch.randn(c_half, c_full, 2, 2, 2, device=dev) / np.sqrt(8 * c_full)).to(torch.bfloat16).float().requires_grad_(True)
        b = (0.1 * torch.randn(c_half, device=dev)).requires_grad_(True)
        xcb = cb8_from_planar(x).requires_grad_(True)
        y = ops.ConvDown2.apply(xcb, w, b, _packs(ops, dev, w, (0, 2, 3)))
        xr, wr, br = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        yr = F.conv3d(xr, wr, br, stride=2)
        g = torch.randn_like(yr).to(torch.bfloat16).float()
        yr.backward(g)
        y.backward(cb8_from_planar(g))
        torch.cuda.synchronize()
        res[("down", tc)] = (rel_rms(planar_from_cb8(y.detach(), c_half), yr.detach()), rel_rms(planar_from_cb8(xcb.grad, c_full), xr.grad),
                             rel_rms(w.grad, wr.grad))
        # up
        x = torch.randn(n, c_half, *half, device=dev).to(torch.bfloat16).float()
        w = (torch.randn(c_half, c_full, 2, 2, 2, device=dev) / np.sqrt(c_half)).to(torch.bfloat16).float().requires_grad_(True)
        b = (0.1 * torch.randn(c_full, device=dev)).requires_grad_(True)
        xcb = cb8_from_planar(x).requires_grad_(True)
        y = ops.ConvUp2.apply(xcb, w, b, _packs(ops, dev, w, (0, 2, 3)))
        xr, wr, br = x.clone().requires_grad_(True), w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
        yr = F.conv_transpose3d(xr, wr, br, stride=2)
        g = torch.randn_like(yr).to(torch.bfloat16).float()
        yr.backward(g)
        y.backward(cb8_from_planar(g))
        torch.cuda.synchronize()
        res[("up", tc)] = (rel_rms(planar_from_cb8(y.detach(), c_full), yr.detach()), rel_rms(planar_from_cb8(xcb.grad, c_half), xr.grad),
                           rel_rms(w.grad, wr.grad))
    ops._TC_FWD, ops._TC_WGRAD = True, True
    fmt = lambda t: "(y %.1e dx %.1e dw %.1e)" % t
    print(f"[s2] n={n} full_c={c_full} half_c={c_half} half={half} down direct{fmt(res[('down', False)])} tc{fmt(res[('down', True)])} "
          f"| up direct{fmt(res[('up', False)])} tc{fmt(res[('up', True)])}", flush=True)
    return max(max(res[("down", True)]), max(res[("up", True)]))


def run_wgrad(n, cin, cout, dims, kernel):
    torch.manual_seed(1)
    x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
    g = torch.randn(n, cout, *dims, device=dev).to(torch.bfloat16).float()
    a, dy = cb8_from_planar(x), cb8_from_planar(g)
    w = torch.zeros(cout, cin, *kernel, device=dev, requires_grad=True)
    F.conv3d(x, w, None, padding=tuple(k // 2 for k in kernel)).backward(g)
    pad = tuple(k // 2 for k in kernel)
    d_d = ops._wgrad(a, dy, cin, cout, dims, kernel, (1, 1, 1), pad, w.shape, allow_tc=False)
    d_t = ops._wgrad(a, dy, cin, cout, dims, kernel, (1, 1, 1), pad, w.shape, allow_tc=True)
    torch.cuda.synchronize()
    e_d, e_t = rel_rms(d_d, w.grad), rel_rms(d_t, w.grad)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        ops._wgrad(a, dy, cin, cout, dims, kernel, (1, 1, 1), pad, w.shape, allow_tc=True)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 200.0
    tf = 2.0 * n * np.prod(dims) * cin * cout * np.prod(kernel) / us * 1e-6
    print(f"[wgrad] n={n} cin={cin} cout={cout} dims={dims} k={kernel} direct_vs_torch={e_d:.2e} tc_vs_torch={e_t:.2e} "
          f"{us:.1f} us {tf:.1f} TF/s", flush=True)
    if e_t > 1e-2:
        bad = (d_t - w.grad).abs() > 0.05 * w.grad.abs().max()
        print("   bad frac", float(bad.float().mean()), "by tap", bad.float().mean(dim=(0, 1)).flatten().cpu().numpy().round(2))
        print("   by co", bad.float().mean(dim=(1, 2, 3, 4)).cpu().numpy().round(2)[:32])
        print("   by ci", bad.float().mean(dim=(0, 2, 3, 4)).cpu().numpy().round(2)[:32])
        print("   got", d_t[0, 0].flatten()[:9].cpu().numpy().round(2), "ref", w.grad[0, 0].flatten()[:9].cpu().numpy().round(2))
    return e_t


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    worst = 0.0
    worst = max(worst, run(1, 16, 16, (4, 6, 8), (3, 3, 3), "ones"))
    worst = max(worst, run(1, 16, 16, (4, 6, 8), (3, 3, 3), "tap"))
    for cfg in [(1, 16, 16, (4, 6, 8), (3, 3, 3)), (2, 16, 16, (8, 12, 20), (3, 3, 3)), (2, 32, 32, (6, 10, 12), (3, 3, 3)),
                (2, 64, 64, (28, 28, 20), (3, 3, 3)), (2, 128, 128, (14, 14, 10), (3, 3, 3)), (2, 256, 256, (7, 7, 5), (3, 3, 3)),
                (2, 32, 16, (9, 7, 11), (3, 3, 3)), (2, 16, 64, (5, 5, 5), (3, 3, 3)), (1, 16, 16, (112, 112, 80), (3, 3, 3)),
                (2, 32, 32, (56, 56, 40), (3, 3, 3)), (3, 16, 16, (1, 64, 64), (1, 3, 3)), (2, 32, 64, (1, 32, 32), (1, 3, 3)),
                (6, 16, 16, (1, 256, 256), (1, 3, 3)), (2, 256, 128, (1, 16, 16), (1, 3, 3))]:
        worst = max(worst, run(*cfg))
    print("WORST tc_vs_torch", worst, flush=True)
    if "--s2" in sys.argv:
        ws = 0.0
        for cfg in [(1, 16, 32, (4, 6, 8)), (2, 16, 32, (8, 6, 20)), (2, 32, 64, (14, 14, 10)), (2, 64, 128, (7, 7, 5)),
                    (2, 128, 256, (3, 4, 2)), (4, 16, 32, (56, 56, 40)), (4, 128, 256, (7, 7, 5))]:
            ws = max(ws, run_s2(*cfg))
        print("WORST s2 tc_vs_torch", ws, flush=True)
    if "--fold" in sys.argv:
        # experimental dz-folded kernel (BCP_TC_FOLD path) against the standard tcgen05 kernel: error vs torch and time
        for cfg in [(1, 16, 16, (4, 6, 8)), (2, 16, 16, (8, 12, 20)), (2, 32, 32, (6, 10, 12)), (4, 16, 16, (112, 112, 80)),
                    (4, 32, 32, (56, 56, 40))]:
            n, cin, cout, dims = cfg
            kernel = (3, 3, 3)
            torch.manual_seed(0)
            x = torch.randn(n, cin, *dims, device=dev).to(torch.bfloat16).float()
            w = (torch.randn(cout, cin, *kernel, device=dev) / np.sqrt(cin * 27)).to(torch.bfloat16).float()
            b = 0.1 * torch.randn(cout, device=dev)
            pack = _packs(ops, dev, w, (0, 1))
            a = cb8_from_planar(x)
            ref = F.conv3d(x, w, b, padding=1)
            res = {}
            for name, flag in (("std", False), ("fold", True)):
                ops._TC_FOLD = flag
                y = ops._conv_same(a, pack.k[0], b, cout, kernel)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ev[0].record()
                for _ in range(10):
                    ops._conv_same(a, pack.k[0], b, cout, kernel)
                ev[1].record()
                torch.cuda.synchronize()
                res[name] = (rel_rms(planar_from_cb8(y, cout), ref), ev[0].elapsed_time(ev[1]) * 100.0)
            ops._TC_FOLD = False
            print(f"[fold] n={n} c={cin}->{cout} dims={dims} std err {res['std'][0]:.2e} {res['std'][1]:.1f} us | "
                  f"fold err {res['fold'][0]:.2e} {res['fold'][1]:.1f} us", flush=True)
    if "--prof" in sys.argv:
        ops._TC_FOLD = False
        for cfg in [(4, 16, (112, 112, 80)), (4, 32, (56, 56, 40)), (4, 64, (28, 28, 20)), (4, 128, (14, 14, 10)), (4, 256, (7, 7, 5))]:
            run_prof(*cfg)
    if "--prof-fold" in sys.argv:
        for ew in (8, 12, 16):
            for cfg in [(4, 16, (112, 112, 80)), (4, 32, (56, 56, 40))]:
                run_prof(*cfg, fold=ew)
        ops._TC_FOLD = False
    if "--wgrad-splits" in sys.argv:
        from bcp_b200._native import LIB, i3, ptr, stream
        for n, c, dims in ((4, 64, (28, 28, 20)), (4, 128, (14, 14, 10)), (4, 256, (7, 7, 5)), (4, 32, (56, 56, 40))):
            x = torch.randn(n, c, *dims, device=dev).to(torch.bfloat16).float()
            a, dy = cb8_from_planar(x), cb8_from_planar(x.flip(1))
            dw = torch.zeros(c, c, 27, device=dev)
            ws = torch.empty(LIB.query("bcp_conv_tc_wgrad_workspace_floats", n, c, c, i3(*dims), i3(3, 3, 3)), device=dev)
            cnt = torch.zeros(4, dtype=torch.int32, device=dev)
            ref = None
            for ms in (0, -1):       # 0 = planner (per-tap compact mode where eligible), -1 = halo-brick plan
                args = (ptr(a), ptr(dy), ptr(dw), ptr(ws), ptr(cnt), n, c, c, i3(*dims), i3(3, 3, 3), 0, ms, stream())
                LIB.call("bcp_conv_tc_wgrad_capped", *args)
                torch.cuda.synchronize()
                if ref is None:
                    ref = dw.clone()
                err = rel_rms(dw, ref)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ev[0].record()
                for _ in range(5):
                    LIB.call("bcp_conv_tc_wgrad_capped", *args)
                ev[1].record()
                torch.cuda.synchronize()
                print(f"[wgrad-splits] c={c} dims={dims} max_splits={ms:3d}  {ev[0].elapsed_time(ev[1]) * 200.0:7.1f} us  vs planner's result {err:.1e}", flush=True)
    if "--wgrad" in sys.argv:
        ww = 0.0
        for cfg in [(1, 16, 16, (4, 6, 8), (3, 3, 3)), (2, 16, 16, (8, 12, 20), (3, 3, 3)), (2, 32, 32, (6, 10, 12), (3, 3, 3)),
                    (2, 64, 64, (28, 28, 20), (3, 3, 3)), (2, 128, 128, (14, 14, 10), (3, 3, 3)), (2, 256, 256, (7, 7, 5), (3, 3, 3)),
                    (2, 32, 16, (9, 7, 11), (3, 3, 3)), (2, 16, 64, (5, 5, 5), (3, 3, 3)), (1, 16, 16, (112, 112, 80), (3, 3, 3)),
                    (3, 16, 16, (1, 64, 64), (1, 3, 3)), (2, 256, 128, (1, 16, 16), (1, 3, 3)), (2, 64, 32, (6, 10, 12), (3, 3, 3)),
                    (4, 16, 16, (112, 112, 80), (3, 3, 3)), (4, 32, 32, (56, 56, 40), (3, 3, 3)), (4, 64, 64, (28, 28, 20), (3, 3, 3)),
                    (4, 128, 128, (14, 14, 10), (3, 3, 3)), (4, 256, 256, (7, 7, 5), (3, 3, 3))]:
            ww = max(ww, run_wgrad(*cfg))
        print("WORST wgrad tc_vs_torch", ww, flush=True)
