"""Shared helpers for the golden-vector generator and the parity tests (test infrastructure)."""
import numpy as np
import torch
import torch.nn as nn


class InjectedDropout(nn.Module):
    """Dropout whose keep-mask comes from a numpy stream (reproducible across implementations).

    Same math as nn.Dropout / nn.Dropout3d in training mode: y = x * keep / (1-p); identity in eval.
    ``channelwise`` -> one Bernoulli per (n, c) (Dropout3d semantics, networks/VNet.py:165,211).
    The product's blocks call ``make_mask(shape)`` themselves (they fuse the multiply)."""

    def __init__(self, p: float, channelwise: bool, seed: int):
        super().__init__()
        self.p, self.channelwise, self.seed, self.calls = float(p), channelwise, int(seed), 0

    def make_mask(self, shape, device=None):
        rs = np.random.RandomState(self.seed * 100003 + self.calls)
        self.calls += 1
        mshape = tuple(shape[:2]) + (1,) * (len(shape) - 2) if self.channelwise else tuple(shape)
        if self.p <= 0.0:
            keep = np.ones(mshape, dtype=np.float32)
        else:
            keep = (rs.random_sample(mshape) >= self.p).astype(np.float32) / (1.0 - self.p)
        t = torch.from_numpy(keep)
        return t.to(device) if device is not None else t

    def forward(self, x):
        if not self.training:
            return x
        return x * self.make_mask(x.shape, x.device)


def inject_dropout(model: nn.Module, seed: int) -> int:
    """Replace every nn.Dropout / nn.Dropout3d (named_modules order) by InjectedDropout(seed+i)."""
    targets = [(n, m) for n, m in model.named_modules() if isinstance(m, (nn.Dropout, nn.Dropout3d))]
    for i, (name, mod) in enumerate(targets):
        parent = model
        parts = name.split(".")
        for part in parts[:-1]:
            parent = getattr(parent, part)
        new = InjectedDropout(mod.p, isinstance(mod, nn.Dropout3d), seed + i)
        new.train(mod.training)
        if parts[-1].isdigit() and isinstance(parent, nn.Sequential):
            parent[int(parts[-1])] = new
        else:
            setattr(parent, parts[-1], new)
    return len(targets)


def tensor_digest(t) -> np.ndarray:
    a = t.detach().cpu().double().reshape(-1).numpy() if torch.is_tensor(t) else np.asarray(t, dtype=np.float64).reshape(-1)
    if a.size == 0:
        return np.zeros(5)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum(), a[0], a[-1]], dtype=np.float64)


def digest_named(named) -> np.ndarray:
    """[n, 5] digests in iteration order of a dict name->tensor."""
    return np.stack([tensor_digest(v) for _, v in named.items()])


def find_box_seed_la(shape, start=0):
    """A numpy seed for which utils/BCP_utils.py:23-25's hard-coded randint bounds (112,112,80)
    put the 2/3 box fully inside a small test volume."""
    X, Y, Z = shape
    px, py, pz = int(X * 2 / 3), int(Y * 2 / 3), int(Z * 2 / 3)
    for s in range(start, start + 2000000):
        rs = np.random.RandomState(s)
        w, h, z = rs.randint(0, 112 - px), rs.randint(0, 112 - py), rs.randint(0, 80 - pz)
        if w + px <= X and h + py <= Y and z + pz <= Z:
            return s
    raise RuntimeError("no seed")


# ---------------------------------------------------------------------------------------------------------
# fixtures built on the reference's SHIPPED checkpoints (models/LA/LA_10.pth, models/ACDC/ACDC_10.pth)
# ---------------------------------------------------------------------------------------------------------
def bf16_round_np(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit patterns (uint16), round to nearest even (same rule as torch's .to(bfloat16))."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)


def pack_weights_bf16(sd) -> dict:
    """state_dict -> {key: uint16 bf16 bits (float tensors) | int64 (counters)}.  The checkpoint's values rounded to
    bf16 ARE the fixture's weights: the reference (fp32 arithmetic) and the native path both load exactly these numbers,
    so the weight rounding of the bf16 operand packs is not part of the measured difference."""
    out = {}
    for k, v in sd.items():
        a = v.detach().cpu().numpy()
        out[k] = a.astype(np.int64) if a.dtype.kind in "iu" else bf16_round_np(a)
    return out


def unpack_weights_bf16(npz) -> dict:
    sd = {}
    for k in npz.files:
        a = npz[k]
        if a.dtype == np.uint16:
            sd[k] = torch.from_numpy((a.astype(np.uint32) << 16).view(np.float32).copy())
        else:
            sd[k] = torch.from_numpy(a.copy())
    return sd


def synthetic_scene(n, shape, seed, n_classes=2, kind="randn"):
    """Images that carry their labels (blob + noise) so a trained network has something to segment:
    volume [n,1,*shape] fp32, label [n,*shape] int64.  numpy streams only (regenerable on the GPU box)."""
    from oracle import bcp_oracle as O
    lab = O.synthetic_labels((n,) + tuple(shape), seed + 1, n_classes=n_classes)
    noise = O.synthetic_volume((n, 1) + tuple(shape), seed, kind)
    if kind == "randn":
        vol = 0.6 * noise + 1.8 * (lab > 0).float().unsqueeze(1) - 0.3
    else:
        vol = (0.35 * noise + 0.2 * lab.float().unsqueeze(1)).clamp_(0.0, 1.0)
    return vol.contiguous(), lab


def packbits(t) -> np.ndarray:
    a = t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
    return np.packbits(a.astype(bool).reshape(-1))


def unpackbits(a, shape) -> np.ndarray:
    n = int(np.prod(shape))
    return np.unpackbits(np.asarray(a))[:n].reshape(shape)
