"""CUDA-graph capture of a whole BCP training step (teacher fwd -> pseudo labels -> largest-CC -> mix -> student
fwd/bwd -> loss -> [gradient all-reduce] -> SGD/Adam+EMA -> weight repack) for the three entry points.

A step is 200-350 kernel launches of a few microseconds to a few hundred microseconds each; enqueueing them from
Python costs ~10 ms, more than the GPU needs.  Capturing the step once and replaying it removes the host from the
loop.  Everything the host used to decide per step lives in device memory:
  * the random box: a 6-int device tensor (include/bcp_b200.h: box6_dev), refreshed from a ring of pinned slots;
  * the optimiser's hyper-parameters (learning-rate decay!): the device ``hyper`` vector, re-uploaded by
    ``optimizer.refresh_hyper()`` before each replay whenever ``param_groups`` changed; Adam's bias corrections are
    advanced on the device (bcp_adam_tick);
  * the step's inputs: static device buffers.  ``load()`` copies pinned host batches into one of two staging sets on a
    copy stream (overlapping the previous step), ``step()`` moves the staged set into the static buffers with a
    device-to-device copy (~10 us) and replays.
"""
from __future__ import annotations

import numpy as np
import torch

from . import step as S
from .utils.BCP_utils import context_box

_RING = 8


class GraphedStep:
    """kind: 'la' (LA_BCP_train.py:234-270), 'la_pre' (:146-171), 'acdc' (ACDC_BCP_train.py:354-390), 'acdc_pre'
    (:237-255), 'pan' (pancreas/train_pancreas.py:144-174).  ``volume_shape`` = [B,1,...] of the whole two-stream batch
    (pancreas: the 8 volumes img_a, img_b, unimg_a, unimg_b stacked, labels for the first 4)."""

    def __init__(self, kind, model, ema_model, optimizer, volume_shape, warmup=3, device=None, **kw):
        self.kind, self.model, self.ema_model, self.optimizer, self.kw = kind, model, ema_model, optimizer, kw
        dev = device or next(model.parameters()).device
        self.dev = dev
        vs = tuple(volume_shape)
        nlab = vs[0] if kind != "pan" else vs[0] // 2
        self.vol = torch.zeros(vs, dtype=torch.float32, device=dev)
        self.lab = torch.zeros((nlab,) + vs[2:], dtype=torch.uint8, device=dev)
        self.box = torch.zeros(6, dtype=torch.int32, device=dev)
        self._box_ring = torch.zeros(_RING, 6, dtype=torch.int32).pin_memory()
        self._box_ev = [None] * _RING
        self._box_i = 0
        self._stage = [(torch.empty_like(self.vol), torch.empty_like(self.lab)) for _ in range(2)]
        self._stage_ev = [None, None]
        self._stage_w = 0            # next staging set to fill
        self._stage_r = None         # staging set holding the batch of the next step (None: static buffers are current)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._set_box(self.draw_box())
        from ._native import LIB
        # warm-up on a side stream (allocator + lazy initialisation must settle before capture).  The warm-up steps run on
        # the all-zero static buffers, so the training state they touch is snapshotted first and restored after capture.
        snap = self._snapshot()
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = LIB.launches
        # capture on a HIGH-priority stream: the step forks side streams (teacher pass, weight gradients) of default priority,
        # and the chain on the capturing stream is the critical path -- its kernels should get free SMs first
        cap = torch.cuda.Stream(device=dev, priority=-1)           # measured: 6.88 vs 6.94 ms/step at default priority
        with torch.cuda.graph(self.graph, stream=cap):
            self.out = self._body()
        self.kernels_per_replay = LIB.launches - l0
        self._restore(snap)
        self.replays = 0

    def _state_tensors(self):
        ts = []
        for net in (self.model, self.ema_model):
            if net is not None:
                rt = net.runtime
                ts.append(rt.arena)
                ts.extend(rt.int_buffers)
        opt = self.optimizer
        for name in ("buf", "m", "v", "step_dev"):
            t = getattr(opt, name, None)
            if torch.is_tensor(t):
                ts.append(t)
        return ts

    def _snapshot(self):
        for net in (self.model, self.ema_model):
            if net is not None and not net.runtime.is_flat():
                net.runtime.flatten_()
        return [t.clone() for t in self._state_tensors()], self.optimizer.step_count

    def _restore(self, snap):
        saved, count = snap
        with torch.no_grad():
            for t, s in zip(self._state_tensors(), saved):
                t.copy_(s)
        self.optimizer.step_count = count
        for net in (self.model, self.ema_model):
            if net is not None:
                net.runtime.dirty = True

    # ---- box ----------------------------------------------------------------------------------
    def draw_box(self):
        """The reference's own np.random.randint call order for this entry point."""
        if self.kind in ("la", "la_pre"):
            return context_box(self.vol.shape, self.kw.get("mask_ratio", 2 / 3))
        if self.kind in ("acdc", "acdc_pre"):
            w, h, px, py = S._acdc_box(self.vol.shape)
            return (w, h, 0, px, py, 1)
        return S._pan_box(self.kw.get("patch_size", 64))

    def _set_box(self, box):
        box = tuple(int(v) for v in box)
        if len(box) == 4:
            box = (box[0], box[1], 0, box[2], box[3], 1)
        i = self._box_i
        if self._box_ev[i] is not None:
            self._box_ev[i].synchronize()          # the H2D copy that last read this pinned slot has finished
        self._box_ring[i].copy_(torch.tensor(box, dtype=torch.int32))
        self.box.copy_(self._box_ring[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._box_ev[i] = ev
        self._box_i = (i + 1) % _RING

    # ---- step body ------------------------------------------------------------------------------
    def _body(self):
        k, kw = self.kind, self.kw
        if k == "la":
            return S.la_self_train_step(self.model, self.ema_model, self.optimizer, self.vol, self.lab, box=self.box, **kw)
        if k == "la_pre":
            return S.la_pre_train_step(self.model, self.optimizer, self.vol, self.lab, box=self.box, **kw)
        if k == "acdc":
            return S.acdc_self_train_step(self.model, self.ema_model, self.optimizer, self.vol, self.lab, box=self.box, **kw)
        if k == "acdc_pre":
            return S.acdc_pre_train_step(self.model, self.optimizer, self.vol, self.lab, box=self.box, **kw)
        if k == "pan":
            n = self.vol.shape[0] // 4
            v, l = self.vol, self.lab
            return S.pan_self_train_step(self.model, self.ema_model, self.optimizer, v[:n], l[:n], v[n:2 * n], l[n:2 * n],
                                         v[2 * n:3 * n], v[3 * n:], box=self.box, **kw)
        raise ValueError(k)

    # ---- inputs ---------------------------------------------------------------------------------
    def load(self, volume, label):
        """Start copying the NEXT step's batch (pinned host or device tensors) into a staging set on the copy stream."""
        w = self._stage_w
        if self._stage_ev[w] is not None:
            self._stage_ev[w].synchronize()        # the D2D that last read this staging set has been enqueued and finished
        sv, sl = self._stage[w]
        cs = self._copy_stream
        with torch.cuda.stream(cs):
            sv.copy_(volume, non_blocking=True)
            sl.copy_(label[:sl.shape[0]], non_blocking=True)       # pancreas: labels of the labeled half only
            ev = torch.cuda.Event()
            ev.record(cs)
        self._stage_ev[w] = ev
        self._stage_r = w
        self._stage_w = 1 - w

    def step(self, box=None):
        """Replay one step on the most recently loaded batch (or on the batch already in the static buffers)."""
        cur = torch.cuda.current_stream(self.dev)
        if self._stage_r is not None:
            r = self._stage_r
            cur.wait_event(self._stage_ev[r])
            self.vol.copy_(self._stage[r][0], non_blocking=True)
            self.lab.copy_(self._stage[r][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            self._stage_ev[r] = ev
            self._stage_r = None
        self._set_box(box if box is not None else self.draw_box())
        self.optimizer.refresh_hyper()             # LR decay etc. reach the device vector outside the captured region
        self.graph.replay()
        self.optimizer.step_count += 1
        self.replays += 1
        return self.out

    def __call__(self, volume, label, box=None):
        self.load(volume, label)
        return self.step(box)

    def replay_resident(self, box=None):
        return self.step(box)


class GraphedLAStep(GraphedStep):
    """Round-1 name kept: the LA self-training step."""

    def __init__(self, model, ema_model, optimizer, volume_shape, labeled_bs=4, mask_ratio=2 / 3, u_weight=0.5, nms=1,
                 warmup=3, device=None):
        super().__init__("la", model, ema_model, optimizer, volume_shape, warmup=warmup, device=device,
                         labeled_bs=labeled_bs, mask_ratio=mask_ratio, u_weight=u_weight, nms=nms)
