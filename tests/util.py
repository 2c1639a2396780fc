"""Helpers shared by the GPU parity tests."""
import json
import os

import numpy as np
import torch

G = os.path.join(os.path.dirname(__file__), "golden")
_METRICS = {}


def load_golden(name):
    return np.load(os.path.join(G, name + ".npz"))


def T(a):
    return torch.from_numpy(np.asarray(a))


def cb8_from_planar(x: torch.Tensor) -> torch.Tensor:
    """[N,C,X,Y,Z] float -> CB8 bf16 [N,C/8,X,Y,Z,8] with plain torch ops (test reference for the layout)."""
    n, c = x.shape[:2]
    sp = tuple(x.shape[2:])
    cb = (c + 7) // 8
    if cb * 8 != c:
        x = torch.cat([x, x.new_zeros((n, cb * 8 - c) + sp)], 1)
    return x.reshape(n, cb, 8, *sp).permute(0, 1, 3, 4, 5, 2).contiguous().to(torch.bfloat16)


def planar_from_cb8(a: torch.Tensor, c: int) -> torch.Tensor:
    n, cb, x, y, z, _ = a.shape
    return a.float().permute(0, 1, 5, 2, 3, 4).reshape(n, cb * 8, x, y, z)[:, :c].contiguous()


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    den = float(b.pow(2).mean().sqrt())
    return float((a - b).pow(2).mean().sqrt()) / (den + 1e-30)


def record(key, value):
    """Collect measured parity numbers; written to gpurun_out/test_metrics.json for DESIGN.md."""
    _METRICS[key] = float(value) if not isinstance(value, (list, dict)) else value
    out = os.path.join(os.path.dirname(__file__), "..", "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "test_metrics.json")
        cur = {}
        if os.path.exists(path):
            try:
                cur = json.load(open(path))
            except Exception:
                cur = {}
        cur.update(_METRICS)
        json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def wgrad_fp64(a: torch.Tensor, dy: torch.Tensor, kernel, pad) -> torch.Tensor:
    """float64 weight gradient of a stride-1 convolution, dW[co][ci][taps] = sum_{n,v} dy[n,co,v] * a[n,ci,v+tap-pad], as
    one fp64 matmul per tap (the exact reference for layers whose fp32 sums run over millions of voxels, where stock
    cuDNN fp32 itself differs from fp64 by >1e-4)."""
    import torch.nn.functional as F
    n, ci = a.shape[:2]
    co = dy.shape[1]
    X, Y, Z = dy.shape[2:]
    ap = F.pad(a.double(), (pad[2], pad[2], pad[1], pad[1], pad[0], pad[0]))
    g = dy.double().reshape(n, co, -1)
    out = torch.empty((co, ci) + tuple(kernel), dtype=torch.float64, device=a.device)
    for dx in range(kernel[0]):
        for dyy in range(kernel[1]):
            for dz in range(kernel[2]):
                sl = ap[:, :, dx:dx + X, dyy:dyy + Y, dz:dz + Z].reshape(n, ci, -1)
                out[:, :, dx, dyy, dz] = torch.einsum("nov,niv->oi", g, sl)
    return out
