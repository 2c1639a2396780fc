"""Per-network runtime: flat fp32 parameter arena, bf16 operand packs, norm-group size, dropout masks.

One ``NetRuntime`` is shared by all blocks of a network.  It owns
  * the flat fp32 arena that every parameter / float buffer is a view of (so the fused SGD/Adam+EMA
    kernel and the single-launch weight repack can address the whole network), laid out as
    [trainable | ema-only parameters | float buffers];
  * the bf16 operand packs the conv kernels read (rebuilt by ONE ``bcp_weights_repack`` launch
    whenever a parameter version changed);
  * ``spg`` -- samples per normalisation group for the current forward (a reference forward call of
    batch b is one group; the engine batches several calls as several groups).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .._native import LIB, ptr, stream
from ..ops import ConvPack

_JOB = np.dtype([("src_off", "<i8"), ("dst_off", "<i8"), ("dim_a", "<i4"), ("dim_b", "<i4"), ("taps", "<i4"), ("kind", "<i4")])


class NetRuntime:
    def __init__(self, owner: nn.Module, trainable_prefixes):
        self.owner_ref = [owner]          # list: keep the module out of nn.Module attribute registration
        self.trainable_prefixes = tuple(trainable_prefixes)
        self.spg = None                   # samples per norm group (None -> whole batch)
        self.arena = None
        self.packed = None
        self.jobs = None
        self.packs = {}                   # id(conv module) -> ConvPack
        self.layers = []                  # (conv module, kind_fwd, kind_bwd)
        self.n_train = 0
        self.n_param = 0
        self._versions = None
        self.dirty = True
        self.grad_arena = None
        self.grad_ready_cb = None         # set by the fused optimiser when data parallel: callable(lo_offset)

    def boundary(self, a, first_param_name):
        """Mark a point of the forward pass after which (in backward: before which) all parameters from
        ``first_param_name`` on have final gradients (ops.GradBoundary).  No-op unless a data-parallel optimiser listens."""
        if self.grad_ready_cb is None or not a.requires_grad:
            return a
        from ..ops import GradBoundary
        return GradBoundary.apply(a, self, self.offsets[first_param_name])

    # ---- registration -----------------------------------------------------------------------
    def register_conv(self, conv: nn.Module, kinds):
        """kinds: (fwd_kind, bwd_kind) repack kinds or None for layers that read fp32 weights directly."""
        self.layers.append((conv, kinds))

    # ---- flat arena ---------------------------------------------------------------------------
    def _named(self):
        owner = self.owner_ref[0]
        params = list(owner.named_parameters())
        train = [(n, p) for n, p in params if n.startswith(self.trainable_prefixes)]
        other = [(n, p) for n, p in params if not n.startswith(self.trainable_prefixes)]
        bufs = [(n, b) for n, b in owner.named_buffers() if b.dtype == torch.float32]
        ints = [(n, b) for n, b in owner.named_buffers() if b.dtype == torch.int64]
        return train, other, bufs, ints

    def is_flat(self):
        if self.arena is None:
            return False
        lo = self.arena.data_ptr()
        hi = lo + self.arena.numel() * 4
        owner = self.owner_ref[0]
        for p in owner.parameters():
            if not (lo <= p.data_ptr() < hi) or p.device != self.arena.device:
                return False
        return True

    def flatten_(self):
        """(Re)build the arena on the parameters' current device and make every parameter/buffer a view of it."""
        owner = self.owner_ref[0]
        train, other, bufs, ints = self._named()
        dev = next(owner.parameters()).device
        total = sum(t.numel() for _, t in train + other + bufs)
        arena = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = {}
        with torch.no_grad():
            for group in (train, other, bufs):
                for name, t in group:
                    n = t.numel()
                    view = arena[off:off + n].view(t.shape)
                    view.copy_(t.data)
                    t.data = view
                    self.offsets[name] = off
                    off += n
        self.arena = arena
        self.n_train = sum(t.numel() for _, t in train)
        self.n_param = self.n_train + sum(t.numel() for _, t in other)
        self.n_total = total
        self.int_buffers = [b for _, b in ints]
        # the int64 buffers (BatchNorm num_batches_tracked) as views of ONE flat buffer: the state_dict EMA of the ACDC entry
        # point (optim._after) is then one launch instead of one per normalisation layer
        self.int_arena = None
        if ints:
            ia = torch.empty(sum(b.numel() for _, b in ints), dtype=torch.int64, device=dev)
            ioff = 0
            with torch.no_grad():
                for _, b in ints:
                    n = b.numel()
                    view = ia[ioff:ioff + n].view(b.shape)
                    view.copy_(b.data)
                    b.data = view
                    ioff += n
            self.int_arena = ia
        self.grad_arena = None
        self._build_packs()
        self.dirty = True

    def ensure_grad_arena(self):
        """p.grad of every trainable parameter becomes a view of one flat fp32 buffer (same offsets as the arena)."""
        if self.grad_arena is None or self.grad_arena.device != self.arena.device:
            self.grad_arena = torch.zeros(self.n_train, dtype=torch.float32, device=self.arena.device)
        train, _, _, _ = self._named()
        for name, p in train:
            off = self.offsets[name]
            g = self.grad_arena[off:off + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g
            p._bcp_direct = True          # kernels accumulate straight into the arena view (ops._direct)
        return self.grad_arena

    # ---- operand packs ------------------------------------------------------------------------
    def _build_packs(self):
        jobs = []
        dst = 0
        dev = self.arena.device
        slots = []
        base = self.arena.data_ptr()
        for conv, kinds in self.layers:
            if kinds is None:
                continue
            w = conv.weight
            a, b = w.shape[0], w.shape[1]
            taps = int(np.prod(w.shape[2:]))
            src = (w.data_ptr() - base) // 4
            views = []
            # a network whose weights never receive gradients (the EMA teacher) never runs the 3x3x3 dgrad, the only user
            # of the flipped/transposed kind-1 pack (the stride-2 layers' packs serve both directions and stay)
            use = kinds if w.requires_grad else tuple(k for k in kinds if k != 1)
            for kind in use:
                if kind == 0:
                    n = taps * ((b + 7) // 8) * a * 8
                else:
                    n = taps * ((a + 7) // 8) * b * 8
                jobs.append((src, dst, a, b, taps, kind))
                views.append((kind, dst, n))
                dst += (n + 127) // 128 * 128          # keep every pack 256-byte aligned
            slots.append((conv, views))
        self.packed = torch.empty(max(dst, 128), dtype=torch.bfloat16, device=dev)
        arr = np.array(jobs, dtype=_JOB) if jobs else np.zeros(0, dtype=_JOB)
        self.njobs = len(jobs)
        self.jobs = torch.from_numpy(arr.view(np.uint8).copy()).to(dev) if jobs else None
        self.packs = {}
        for conv, views in slots:
            self.packs[id(conv)] = ConvPack({kind: self.packed[o:o + n] for kind, o, n in views})

    def pack(self, conv) -> ConvPack:
        return self.packs[id(conv)]

    def repack(self):
        if self.njobs:
            LIB.call("bcp_weights_repack", ptr(self.arena), ptr(self.packed), ptr(self.jobs), self.njobs, stream())
        self.dirty = False

    def prepare(self):
        """Called at the start of every forward: make sure the arena is flat and the packs are current."""
        owner = self.owner_ref[0]
        if not self.is_flat():
            self.flatten_()
        vers = [p._version for p in owner.parameters()]
        if self.dirty or vers != self._versions:
            self.repack()
            self._versions = vers

    # ---- dropout masks ------------------------------------------------------------------------
    @staticmethod
    def channel_dropout_scale(mod, n, c, device, spg=None):
        """Dropout3d: one Bernoulli per (n, c), value 0 or 1/(1-p) (networks/VNet.py:165,211).  With an injected mask
        provider (parity tests) one mask is drawn per reference forward call, i.e. per group of ``spg`` samples."""
        if hasattr(mod, "make_mask"):
            spg = spg or n
            parts = [mod.make_mask((spg, c, 1, 1, 1), device).reshape(spg, c).float() for _ in range(n // spg)]
            return torch.cat(parts).contiguous()
        p = float(mod.p)
        return torch.empty(n, c, dtype=torch.float32, device=device).bernoulli_(1.0 - p).div_(1.0 - p)

    @staticmethod
    def element_dropout_keep(mod, n, c, spatial, device, spg=None):
        """nn.Dropout: uint8 keep flags in CB8 order [N][C/8][X][Y][Z][8] plus the 1/(1-p) scale."""
        p = float(mod.p)
        if p <= 0.0:
            return None, 1.0
        x, y, z = spatial
        if hasattr(mod, "make_mask"):
            spg = spg or n
            shape = (spg, c, y, z) if x == 1 else (spg, c, x, y, z)
            full = torch.cat([mod.make_mask(shape, device) for _ in range(n // spg)])
            keep = (full > 0).to(torch.uint8).reshape(n, c // 8, 8, x, y, z)
            keep = keep.permute(0, 1, 3, 4, 5, 2).contiguous()
        else:
            keep = torch.empty((n, c // 8, x, y, z, 8), dtype=torch.uint8, device=device).bernoulli_(1.0 - p)
        return keep, 1.0 / (1.0 - p)
