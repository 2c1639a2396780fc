"""Mid-run resume artefact.

The reference saves ``{'net', 'opt'}`` during pre-training and a bare ``state_dict`` during self-training
(LA_BCP_train.py:79-93,287-288) -- enough to start the next stage or to evaluate, not to continue an interrupted stage:
the EMA teacher, the iteration counter (learning-rate decay every 2500 iterations) and the random streams (box draws,
sampler permutations, dropout masks) are lost.  ``save_resume`` / ``load_resume`` keep the reference's two keys readable
by ``load_net`` (``torch.load(path)['net']``) and add the rest in the same file:

    {'net', 'opt',                       # as save_net_opt (LA_BCP_train.py:79-84)
     'ema': state_dict | None,           # the teacher
     'iter': int, 'stage': str,
     'rng': {'numpy', 'python', 'torch_cpu', 'torch_cuda'},
     'extra': {...}}                     # e.g. best_dice, sampler position

A run resumed from the artefact continues bit-identically to the uninterrupted run (tests/test_gpu_dropin.py::
test_resume_is_bit_identical): all kernels on the path are deterministic, so state + random streams determine the rest.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch


def rng_state(device=None):
    st = {"numpy": np.random.get_state(), "python": random.getstate(), "torch_cpu": torch.get_rng_state()}
    if torch.cuda.is_available():
        st["torch_cuda"] = torch.cuda.get_rng_state(device)
    return st


def set_rng_state(st, device=None):
    np.random.set_state(st["numpy"])
    random.setstate(st["python"])
    torch.set_rng_state(st["torch_cpu"])
    if "torch_cuda" in st and torch.cuda.is_available():
        torch.cuda.set_rng_state(st["torch_cuda"], device)


def save_resume(path, model, optimizer, ema_model=None, iteration=0, stage="", extra=None):
    dev = next(model.parameters()).device
    torch.cuda.synchronize(dev)
    sd = {"net": model.state_dict(), "opt": optimizer.state_dict(),
          "ema": ema_model.state_dict() if ema_model is not None else None,
          "iter": int(iteration), "stage": str(stage), "rng": rng_state(dev), "extra": dict(extra or {})}
    tmp = str(path) + ".tmp"
    torch.save(sd, tmp)
    os.replace(tmp, str(path))          # a crash mid-write never leaves a truncated artefact behind


def load_resume(path, model, optimizer, ema_model=None, restore_rng=True):
    """Restores everything ``save_resume`` stored; returns (iteration, stage, extra)."""
    dev = next(model.parameters()).device
    sd = torch.load(str(path), map_location="cpu", weights_only=False)      # RNG states must stay CPU byte tensors
    model.load_state_dict(sd["net"])
    if ema_model is not None:
        if sd.get("ema") is None:
            raise RuntimeError("%s holds no EMA teacher" % path)
        ema_model.load_state_dict(sd["ema"])
    optimizer.load_state_dict(sd["opt"])
    if restore_rng:
        set_rng_state(sd["rng"], dev)
    return int(sd["iter"]), sd.get("stage", ""), sd.get("extra", {})
