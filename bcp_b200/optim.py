"""Fused optimiser + EMA-teacher update over the flat parameter arenas (one kernel per step).

FusedSGD_EMA  == torch.optim.SGD(momentum, weight_decay) (LA_BCP_train.py:218, ACDC_BCP_train.py:334) followed by
                 update_ema_variables (utils/BCP_utils.py:78-81, mode 'params') or update_model_ema
                 (ACDC_BCP_train.py:123-129, mode 'state_dict': parameters + BN buffers + int64 counters).
FusedAdam_EMA == torch.optim.Adam(lr) (pancreas/dataloaders.py:182) + pancreas/pancreas_utils.py:299-302.

Data parallel: when torch.distributed is initialised the flat gradient arena is all-reduced (sum) before the update and
the 1/world_size average is folded into the kernel (grad_scale).  The reduction is bucketed: ``notify_grad_ready(lo, hi)``
(called by the step body as backward finishes a block of layers) starts the all-reduce of that slice on a side stream so
it overlaps the rest of backward; ``step()`` reduces whatever is left and joins the side stream.  Replicas are made
identical at construction (rank 0's parameters / buffers / optimiser state are broadcast).

Hyper-parameters live in a small device vector so a captured CUDA graph sees changes: ``refresh_hyper()`` uploads it
(host side, outside any capture) and is what ``GraphedStep.__call__`` runs before each replay; Adam's bias corrections
are advanced on the device by ``bcp_adam_tick``.
"""
from __future__ import annotations

import math
import os

import torch

from ._native import LIB, ptr, stream


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size()
    return None, 1


def _capturing():
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()


class _FusedBase:
    def __init__(self, model, ema_model, ema_alpha, ema_mode):
        self.model, self.ema_model = model, ema_model
        self.rt = model.runtime
        self.ert = ema_model.runtime if ema_model is not None else None
        for r in (self.rt, self.ert):
            if r is not None and not r.is_flat():
                r.flatten_()
        self.ema_alpha, self.ema_mode = float(ema_alpha), ema_mode
        self.dev = self.rt.arena.device
        self.param_groups = [{"lr": None}]
        self.step_count = 0
        self._reduced_upto = None          # gradient-arena elements [_reduced_upto, n_train) are already being reduced
        self._comm_stream = None
        self._hyper_host = None
        # Bucketed overlap is opt-in (BCP_DP_OVERLAP=1): the conv kernels are persistent one-CTA-per-SM grids with static
        # work assignment, so a communication kernel that takes SMs away mid-backward stretches them; measured before use.
        if _world()[1] > 1 and os.environ.get("BCP_DP_OVERLAP", "0") == "1":
            self.rt.grad_ready_cb = self.notify_grad_ready

    # ---- data-parallel plumbing ---------------------------------------------------------------
    def broadcast_state(self, state_tensors=()):
        """Make every rank start from rank 0's parameters, buffers and optimiser state (replicas then stay bit-identical
        because they apply identical reduced gradients)."""
        dist, world = _world()
        if world <= 1:
            return
        for r in (self.rt, self.ert):
            if r is not None:
                dist.broadcast(r.arena, 0)
                for b in r.int_buffers:
                    dist.broadcast(b, 0)
                r.dirty = True
        for t in state_tensors:
            dist.broadcast(t, 0)

    def zero_grad(self, set_to_none: bool = False):
        g = self.rt.ensure_grad_arena()
        g.zero_()
        self._reduced_upto = None

    def notify_grad_ready(self, lo: int):
        """Backward has finished every gradient whose arena offset is >= ``lo`` (parameters are laid out in forward order,
        so backward completes the arena from the top down): start reducing [lo, previous lo) on the side stream."""
        dist, world = _world()
        if world <= 1:
            return
        from .ops import join_wgrad_stream
        join_wgrad_stream(self.dev)
        hi = self.rt.n_train if self._reduced_upto is None else self._reduced_upto
        lo = max(0, min(int(lo), hi))
        if lo >= hi:
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.dev)
        cur = torch.cuda.current_stream(self.dev)
        self._comm_stream.wait_stream(cur)
        with torch.cuda.stream(self._comm_stream):
            dist.all_reduce(self.rt.grad_arena[lo:hi])
        self._reduced_upto = lo

    def _allreduce(self):
        from .ops import join_wgrad_stream
        join_wgrad_stream(self.dev)                  # weight gradients run on their own stream (ops._wgrad_async)
        dist, world = _world()
        if world > 1:
            if self._reduced_upto is None:                             # nothing in flight: one all-reduce of the whole arena
                dist.all_reduce(self.rt.grad_arena)
            else:
                self.notify_grad_ready(0)                              # whatever has not been started yet
                torch.cuda.current_stream(self.dev).wait_stream(self._comm_stream)
                self._reduced_upto = None
        return world

    def _ema_extent(self):
        if self.ert is None:
            return self.rt.n_train
        return self.rt.n_param if self.ema_mode == "params" else self.rt.n_total

    def _after(self):
        if self.ert is not None and self.ema_mode == "state_dict":
            a = self.ema_alpha
            ea, ma = getattr(self.ert, "int_arena", None), getattr(self.rt, "int_arena", None)
            if ea is not None and ma is not None and ea.numel() == ma.numel():
                LIB.call("bcp_ema_i64", ptr(ea), ptr(ma), ea.numel(), a, 1.0 - a, stream())      # all counters, one launch
            else:
                for e, m in zip(self.ert.int_buffers, self.rt.int_buffers):
                    LIB.call("bcp_ema_i64", ptr(e), ptr(m), e.numel(), a, 1.0 - a, stream())
        self.rt.dirty = True
        if self.ert is not None:
            self.ert.dirty = True
        self.step_count += 1


class FusedSGD_EMA(_FusedBase):
    def __init__(self, model, ema_model=None, lr=0.01, momentum=0.9, weight_decay=1e-4, ema_alpha=0.99, ema_mode="params"):
        super().__init__(model, ema_model, ema_alpha, ema_mode)
        self.param_groups[0].update(lr=float(lr), momentum=float(momentum), weight_decay=float(weight_decay))
        self.buf = torch.zeros(self.rt.n_train, dtype=torch.float32, device=self.dev)   # buf=0 makes step 1 "buf = g"
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.dev)
        self.broadcast_state((self.buf,))
        self.refresh_hyper()

    def refresh_hyper(self):
        """Upload {lr, momentum, wd, alpha, 1/world, 1-alpha} when they changed (host side; never inside a capture)."""
        g = self.param_groups[0]
        vals = [g["lr"], g["momentum"], g["weight_decay"], self.ema_alpha, 1.0 / _world()[1], 1.0 - self.ema_alpha, 0.0, 0.0]
        if vals != self._hyper_host:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
            self._hyper_host = vals

    def step(self):
        self._allreduce()
        if not _capturing():
            self.refresh_hyper()
        rt = self.rt
        LIB.call("bcp_sgd_ema_step", ptr(rt.arena), ptr(rt.grad_arena), ptr(self.buf),
                 ptr(self.ert.arena) if self.ert is not None else None, ptr(self.hyper), rt.n_train, self._ema_extent(), stream())
        self._after()

    def state_dict(self):
        return {"momentum_buffer": self.buf, "param_groups": self.param_groups, "step": self.step_count}

    def load_state_dict(self, sd):
        self.buf.copy_(sd["momentum_buffer"])
        self.param_groups[0].update(sd["param_groups"][0])
        self.step_count = sd.get("step", 0)
        self.broadcast_state((self.buf,))
        self.refresh_hyper()


class FusedAdam_EMA(_FusedBase):
    def __init__(self, model, ema_model=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, ema_alpha=0.99, ema_mode="params"):
        super().__init__(model, ema_model, ema_alpha, ema_mode)
        self.param_groups[0].update(lr=float(lr), betas=tuple(betas), eps=float(eps))
        self.m = torch.zeros(self.rt.n_train, dtype=torch.float32, device=self.dev)
        self.v = torch.zeros(self.rt.n_train, dtype=torch.float32, device=self.dev)
        self.hyper = torch.zeros(12, dtype=torch.float32, device=self.dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)   # advanced by bcp_adam_tick (graph-replayable)
        self.broadcast_state((self.m, self.v, self.step_dev))
        self.refresh_hyper()

    def refresh_hyper(self):
        """Upload {lr, b1, b2, eps, alpha, 1/world, -, -, 1-alpha, -, 1-b1, 1-b2} when they changed; slots 6, 7, 9 (bias corrections and
        step size) belong to the device tick kernel and are preserved."""
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        vals = [g["lr"], b1, b2, g["eps"], self.ema_alpha, 1.0 / _world()[1], 1.0 - self.ema_alpha, 1.0 - b1, 1.0 - b2]
        if vals != self._hyper_host:
            h = torch.tensor(vals[:6], dtype=torch.float32)
            self.hyper[:6].copy_(h, non_blocking=False)
            self.hyper[8:9].copy_(torch.tensor(vals[6:7], dtype=torch.float32), non_blocking=False)
            self.hyper[10:12].copy_(torch.tensor(vals[7:9], dtype=torch.float32), non_blocking=False)
            self._hyper_host = vals

    def step(self):
        self._allreduce()
        if not _capturing():
            self.refresh_hyper()
        rt = self.rt
        LIB.call("bcp_adam_tick", ptr(self.hyper), ptr(self.step_dev), stream())
        LIB.call("bcp_adam_ema_step", ptr(rt.arena), ptr(rt.grad_arena), ptr(self.m), ptr(self.v),
                 ptr(self.ert.arena) if self.ert is not None else None, ptr(self.hyper), rt.n_train, self._ema_extent(), stream())
        self._after()

    def state_dict(self):
        return {"exp_avg": self.m, "exp_avg_sq": self.v, "param_groups": self.param_groups, "step": self.step_count}

    def load_state_dict(self, sd):
        self.m.copy_(sd["exp_avg"])
        self.v.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])
        self.step_count = sd.get("step", 0)
        self.step_dev.fill_(self.step_count)
        self.broadcast_state((self.m, self.v, self.step_dev))
        self.refresh_hyper()
