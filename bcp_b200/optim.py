"""Fused optimiser + EMA-teacher update over the flat parameter arenas (one kernel per step).

FusedSGD_EMA  == torch.optim.SGD(momentum, weight_decay) (LA_BCP_train.py:218, ACDC_BCP_train.py:334) followed by
                 update_ema_variables (utils/BCP_utils.py:78-81, mode 'params') or update_model_ema
                 (ACDC_BCP_train.py:123-129, mode 'state_dict': parameters + BN buffers + int64 counters).
FusedAdam_EMA == torch.optim.Adam(lr) (pancreas/dataloaders.py:182) + pancreas/pancreas_utils.py:299-302.

Data parallel: when torch.distributed is initialised the flat gradient arena is all-reduced (sum) with ONE NCCL
call before the update and the 1/world_size average is folded into the kernel (grad_scale).
"""
from __future__ import annotations

import math

import torch

from ._native import LIB, ptr, stream


def _world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_world_size()
    return None, 1


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

    def zero_grad(self, set_to_none: bool = False):
        g = self.rt.ensure_grad_arena()
        g.zero_()

    def _allreduce(self):
        dist, world = _world()
        if world > 1:
            dist.all_reduce(self.rt.grad_arena)
        return world

    def _ema_extent(self):
        if self.ert is None:
            return self.rt.n_train
        return self.rt.n_param if self.ema_mode == "params" else self.rt.n_total

    def _after(self):
        if self.ert is not None and self.ema_mode == "state_dict":
            a = self.ema_alpha
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
        self._hyper_host = None
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.dev)

    def _sync_hyper(self, world):
        g = self.param_groups[0]
        vals = [g["lr"], g["momentum"], g["weight_decay"], self.ema_alpha, 1.0 / world, 1.0 - self.ema_alpha, 0.0, 0.0]
        if vals != self._hyper_host:
            self.hyper.copy_(torch.tensor(vals, dtype=torch.float32), non_blocking=False)
            self._hyper_host = vals

    def step(self):
        world = self._allreduce()
        self._sync_hyper(world)
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


class FusedAdam_EMA(_FusedBase):
    def __init__(self, model, ema_model=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, ema_alpha=0.99, ema_mode="params"):
        super().__init__(model, ema_model, ema_alpha, ema_mode)
        self.param_groups[0].update(lr=float(lr), betas=tuple(betas), eps=float(eps))
        self.m = torch.zeros(self.rt.n_train, dtype=torch.float32, device=self.dev)
        self.v = torch.zeros(self.rt.n_train, dtype=torch.float32, device=self.dev)
        self.hyper = torch.zeros(12, dtype=torch.float32, device=self.dev)

    def step(self):
        world = self._allreduce()
        g = self.param_groups[0]
        t = self.step_count + 1
        b1, b2 = g["betas"]
        vals = [g["lr"], b1, b2, g["eps"], self.ema_alpha, 1.0 / world, 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t),
                1.0 - self.ema_alpha, 0.0, 0.0, 0.0]
        self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))
        rt = self.rt
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
