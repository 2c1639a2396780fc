"""torch.autograd.Function wrappers over the C-ABI kernels (bcp_b200/_native.py).

Hidden activations travel as channel-blocked bf16 tensors of shape [N, C/8, X, Y, Z, 8] ("CB8");
2-D networks use X == 1.  The network input and the logits are planar fp32 NC(D)HW exactly like
the reference modules' tensors.  PyTorch is used for memory and autograd plumbing only: every
arithmetic op on the path is a kernel of libbcp_b200.so, and a missing library is a hard error.
"""
from __future__ import annotations

import os

import torch
from torch.autograd import Function

from ._native import LIB, i3, i6, ptr, stream

BF16 = torch.bfloat16
# debugging switches (kernel selection only -- both settings run hand-written sm_100a kernels)
_TC_FWD = os.environ.get("BCP_DISABLE_TC", "0") != "1"
_TC_WGRAD = _TC_FWD and os.environ.get("BCP_DISABLE_TC_WGRAD", "0") != "1"
# conv epilogue produces the following norm's statistics (bcp_conv_tc_fwd_stats).  Opt-in: parity-tested, but on the LA step
# the fused epilogue + last-CTA finalize cost about what the (now 4-loads-in-flight) standalone statistics pass costs.
_FUSE_STATS = os.environ.get("BCP_FUSED_STATS", "0") == "1"
# dz-folded forward kernel for 16/32-channel layers (three dz taps ride in the MMA N dimension).
# Module switch for the parity tests / tools that compare it against the unfolded kernel.
_TC_FOLD = True
# single-launch cluster normalisation (csrc/norm_fused.cu) for layers of <= 64 Ki voxels per statistics group
_NORM_FUSED = os.environ.get("BCP_DISABLE_NORM_FUSED", "0") != "1"


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"bcp_b200.{what}: tensor is on {t.device}; the sm_100a kernels need a CUDA device "
                           "(there is no CPU fallback)")


def cb8_shape(n, c, x, y, z):
    return (n, (c + 7) // 8, x, y, z, 8)


def act_dims(a: torch.Tensor):
    n, cb, x, y, z, e = a.shape
    assert e == 8 and a.dtype == BF16
    return n, cb * 8, x, y, z


def _f32(n, device):
    return torch.empty(int(n), dtype=torch.float32, device=device)


def _direct(param):
    """Gradient target for kernels that can accumulate in place: the parameter's view of the flat gradient arena
    (set up by NetRuntime.ensure_grad_arena), or None to return the gradient through autograd as usual."""
    if param is None or not getattr(param, "_bcp_direct", False):
        return None
    g = param.grad
    return g if (g is not None and g.is_contiguous()) else None


_COUNTERS = {}


def _counter(device):
    """Zero-initialised device ints used by the 'last block finalises' reductions (self-resetting, stream-ordered).  One set
    per (device, stream): the step bodies run the teacher's forward on a side stream next to the student's, and two kernels
    in flight must not share an arrival counter."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    c = _COUNTERS.get(key)
    if c is None:
        c = torch.zeros(16, dtype=torch.int32, device=device)
        _COUNTERS[key] = c
    return c


# ----------------------------------------------------------------------------------------------
# layout converts
# ----------------------------------------------------------------------------------------------
class PlanarToCB8(Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "planar_to_cb8")
        x = x.contiguous().float()
        n, c = x.shape[:2]
        sp = tuple(x.shape[2:])
        x3 = (1,) + sp if len(sp) == 2 else sp
        s = x3[0] * x3[1] * x3[2]
        out = torch.empty(cb8_shape(n, c, *x3), dtype=BF16, device=x.device)
        LIB.call("bcp_planar_to_cb8", ptr(x), ptr(out), n, c, s, stream())
        ctx.meta = (n, c, sp, s)
        return out

    @staticmethod
    def backward(ctx, g):
        n, c, sp, s = ctx.meta
        g = g.contiguous()
        out = torch.empty((n, c) + sp, dtype=torch.float32, device=g.device)
        LIB.call("bcp_cb8_to_planar", ptr(g), ptr(out), n, c, s, stream())
        return out


class CB8ToPlanar(Function):
    @staticmethod
    def forward(ctx, a, c, two_d):
        _require_cuda(a, "cb8_to_planar")
        a = a.contiguous()
        n, cc, x, y, z = act_dims(a)
        sp = (y, z) if two_d else (x, y, z)
        out = torch.empty((n, c) + sp, dtype=torch.float32, device=a.device)
        LIB.call("bcp_cb8_to_planar", ptr(a), ptr(out), n, c, x * y * z, stream())
        ctx.meta = (n, c, x, y, z)
        return out

    @staticmethod
    def backward(ctx, g):
        n, c, x, y, z = ctx.meta
        g = g.contiguous().float()
        out = torch.empty(cb8_shape(n, c, x, y, z), dtype=BF16, device=g.device)
        LIB.call("bcp_planar_to_cb8", ptr(g), ptr(out), n, c, x * y * z, stream())
        return out, None, None


# ----------------------------------------------------------------------------------------------
# convolutions
# ----------------------------------------------------------------------------------------------
class ConvPack:
    """bf16 operand packs of one conv layer, keyed by repack kind (include/bcp_b200.h): views into the owning
    network's packed buffer."""
    __slots__ = ("k",)

    def __init__(self, by_kind):
        self.k = dict(by_kind)


def _conv_same(a, wpack, bias, cout, kernel, allow_tc=True):
    """kx*ky*kz stride-1 'same' conv on CB8; picks the tcgen05 kernel when the shape qualifies."""
    n, cin, x, y, z = act_dims(a)
    out = torch.empty(cb8_shape(n, cout, x, y, z), dtype=BF16, device=a.device)
    dims, k = i3(x, y, z), i3(*kernel)
    if allow_tc and _TC_FWD and _TC_FOLD and LIB.query("bcp_conv_tc_fold_supported", cin, cout, dims, k):
        LIB.call("bcp_conv_tc_fold_fwd", ptr(a), ptr(wpack), ptr(bias), ptr(out), n, cin, cout, dims, k, stream())
    elif allow_tc and _TC_FWD and LIB.query("bcp_conv_tc_supported", cin, cout, dims, k):
        LIB.call("bcp_conv_tc_fwd", ptr(a), ptr(wpack), ptr(bias), ptr(out), n, cin, cout, dims, k, stream())
    else:
        LIB.call("bcp_conv_direct_fwd", ptr(a), ptr(wpack), ptr(bias), ptr(out), n, cin, cout, dims, k,
                 i3(1, 1, 1), i3(kernel[0] // 2, kernel[1] // 2, kernel[2] // 2), 0, stream())
    return out


def _conv_same_stats(a, wpack, bias, cout, kernel, req):
    """conv + fused train-mode normalisation statistics (bcp_conv_tc_fwd_stats); None when the layer is not eligible."""
    n, cin, x, y, z = act_dims(a)
    dims, k = i3(x, y, z), i3(*kernel)
    spg = int(req["spg"])
    if not LIB.query("bcp_conv_tc_supported", cin, cout, dims, k):
        return None
    wsb = LIB.query("bcp_conv_tc_stats_workspace_bytes", n, cin, cout, dims, k, spg)
    if wsb <= 0:
        return None
    dev = a.device
    out = torch.empty(cb8_shape(n, cout, x, y, z), dtype=BF16, device=dev)
    groups = n // spg
    stat = torch.empty(groups, cout, 2, dtype=torch.float32, device=dev)
    coef = torch.empty(groups, cout, 2, dtype=torch.float32, device=dev)
    ws = torch.empty(wsb // 8, dtype=torch.float64, device=dev)
    LIB.call("bcp_conv_tc_fwd_stats", ptr(a), ptr(wpack), ptr(bias), ptr(out), n, cin, cout, dims, k, ptr(req["gamma"]),
             ptr(req["beta"]), ptr(req["running_mean"]), ptr(req["running_var"]), ptr(req["nbt"]), ptr(stat), ptr(coef), ptr(ws),
             ptr(_counter(dev)), spg, float(req["eps"]), float(req["momentum"]), stream())
    req["out"] = (stat, coef)
    return out


# ---- weight gradients on their own stream --------------------------------------------------------------------------------
# Within a layer's backward the weight gradient and the data gradient are independent, and nothing downstream in backward
# needs dW: only the optimiser does.  When the gradient accumulates straight into the flat arena (``into`` is given) the
# launch goes to ONE dedicated side stream per device and is joined when the autograd engine finishes the backward pass.
# The tensor-core weight-gradient kernels (tensor-pipe-bound, one CTA per SM) then run next to the normalisation backward
# kernels (HBM-bound, small shared memory, co-resident on the same SMs) and the data-gradient chain.  ONE stream, because
# those kernels meet at an in-kernel grid barrier: two of them in flight at once could each hold SMs the other waits for.
WGRAD_ON_SIDE_STREAM = os.environ.get("BCP_WGRAD_STREAM", "1") != "0"
_WG_STREAMS, _WG_PENDING = {}, set()


def _wgrad_stream(dev):
    key = (dev.type, dev.index)
    st = _WG_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=dev)
        _WG_STREAMS[key] = st
    return st


def join_wgrad_stream(dev):
    """Make the current stream wait for every weight-gradient launch issued so far (no-op when none is pending)."""
    key = (dev.type, dev.index)
    if key in _WG_PENDING:
        torch.cuda.current_stream(dev).wait_stream(_WG_STREAMS[key])
        _WG_PENDING.discard(key)


def _wgrad_async(dev, tensors, fn):
    if not WGRAD_ON_SIDE_STREAM:
        return fn()
    key = (dev.type, dev.index)
    cur, side = torch.cuda.current_stream(dev), _wgrad_stream(dev)
    side.wait_stream(cur)                              # the operands (dy just produced on `cur`) are ready
    with torch.cuda.stream(side):
        r = fn()
    for t in tensors:
        t.record_stream(side)                          # keep the operands alive until the side stream has consumed them
    if key not in _WG_PENDING:
        _WG_PENDING.add(key)
        from torch.autograd import Variable
        Variable._execution_engine.queue_callback(lambda: join_wgrad_stream(dev))      # end of this backward pass
    return r


def _wgrad(inp, outgrad, cin, cout, in_dims, kernel, stride, pad, wshape, allow_tc=True, into=None):
    """Weight gradient [cout][cin][taps].  ``into``: accumulate into this tensor (flat-arena view) and return None."""
    n = inp.shape[0]
    acc = 1 if into is not None else 0
    same = tuple(stride) == (1, 1, 1) and tuple(pad) == tuple(k // 2 for k in kernel)
    if allow_tc and _TC_WGRAD and same and LIB.query("bcp_conv_tc_wgrad_supported", cin, cout, i3(*in_dims), i3(*kernel)):
        ws = _f32(LIB.query("bcp_conv_tc_wgrad_workspace_floats", n, cin, cout, i3(*in_dims), i3(*kernel)), inp.device)
        dw = into if into is not None else torch.empty(wshape, dtype=torch.float32, device=inp.device)
        LIB.call("bcp_conv_tc_wgrad", ptr(inp), ptr(outgrad), ptr(dw), ptr(ws), _counter(inp.device).data_ptr() + 16, n, cin, cout,
                 i3(*in_dims), i3(*kernel), acc, stream())
        return None if into is not None else dw
    od = [(in_dims[i] + 2 * pad[i] - kernel[i]) // stride[i] + 1 for i in range(3)]
    ws = _f32(LIB.query("bcp_conv_wgrad_workspace_floats", n, cin, cout, i3(*od), i3(*kernel)), inp.device)
    dw = into if into is not None else torch.empty(wshape, dtype=torch.float32, device=inp.device)
    LIB.call("bcp_conv_direct_wgrad", ptr(inp), ptr(outgrad), ptr(dw), ptr(ws), n, cin, cout, i3(*in_dims), i3(*kernel),
             i3(*stride), i3(*pad), acc, stream())
    return None if into is not None else dw


def _bias_grad(dy, c, is_zero, into=None):
    """Conv bias gradient = per-channel sum of dy.  When the conv feeds a batch-statistic normalisation, dy is the
    norm's input gradient and sums to zero per channel identically (sum(g - mean(g) - xhat*mean(g*xhat)) = 0 because
    sum(xhat) = 0); the reference's value there is fp32 rounding noise (~1e-7 relative).  Skip the reduction then."""
    if is_zero:
        return None if into is not None else torch.zeros(c, dtype=torch.float32, device=dy.device)
    g = _chan_sum(dy, c)
    if into is not None:
        into.add_(g)
        return None
    return g


def _chan_sum(t, c):
    n, _, x, y, z = act_dims(t)
    s = x * y * z
    ws = _f32(LIB.query("bcp_chan_sum_workspace_floats", n, c, s), t.device)
    out = torch.empty(c, dtype=torch.float32, device=t.device)
    LIB.call("bcp_chan_sum", ptr(t), ptr(out), ptr(ws), n, c, s, stream())
    return out


class ConvSame(Function):
    """nn.Conv3d(k=3,p=1) / nn.Conv2d(k=3,p=1 | k=1) on CB8 (networks/VNet.py:17, networks/unet.py:20,24,49)."""

    @staticmethod
    def forward(ctx, a, weight, bias, pack: ConvPack, kernel, bias_grad_is_zero=False, stats_req=None):
        """stats_req: optional dict(gamma, beta, running_mean, running_var, nbt, spg, eps, momentum) describing the
        train-mode normalisation that follows; when the layer is eligible the statistics are produced by the conv
        epilogue and returned as stats_req["out"] = (stat, coef) for NormAct(precomputed=...)."""
        _require_cuda(a, "conv")
        a = a.contiguous()
        cout = weight.shape[0]
        ctx.save_for_backward(a, weight)
        ctx.pack, ctx.kernel, ctx.has_bias, ctx.bz = pack, tuple(kernel), bias is not None, bias_grad_is_zero
        ctx.bias_ref = bias
        if stats_req is not None and _TC_FWD and _FUSE_STATS:
            y = _conv_same_stats(a, pack.k[0], bias, cout, kernel, stats_req)
            if y is not None:
                return y
        return _conv_same(a, pack.k[0], bias, cout, kernel)

    @staticmethod
    def backward(ctx, dy):
        a, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, cin, x, y, z = act_dims(a)
        cout, k = weight.shape[0], ctx.kernel
        da = dw = db = None
        if ctx.needs_input_grad[0]:
            da = _conv_same(dy, ctx.pack.k[1], None, cin, k)
        if ctx.needs_input_grad[1]:
            into = _direct(weight)
            args = (a, dy, cin, cout, (x, y, z), k, (1, 1, 1), (k[0] // 2, k[1] // 2, k[2] // 2), weight.shape)
            if into is not None:
                _wgrad_async(a.device, (a, dy), lambda: _wgrad(*args, into=into))
            else:
                dw = _wgrad(*args)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy, cout, ctx.bz, _direct(ctx.bias_ref))
        return da, dw, db, None, None, None, None


def _s2_fwd(inp, pack: ConvPack, bias, cin, cout, half_dims, mode):
    """stride-2 conv (mode 1: full->half) / transposed conv (mode 2: half->full) on CB8; tcgen05 when the shape qualifies."""
    n = inp.shape[0]
    hx, hy, hz = half_dims
    od = (hx, hy, hz) if mode == 1 else (2 * hx, 2 * hy, 2 * hz)
    out = torch.empty(cb8_shape(n, cout, *od), dtype=BF16, device=inp.device)
    if _TC_FWD and 3 in pack.k and LIB.query("bcp_conv_tc_s2_supported", cin, cout, i3(*half_dims), mode):
        LIB.call("bcp_conv_tc_s2_fwd", ptr(inp), ptr(pack.k[0] if mode == 1 else pack.k[3]), ptr(bias), ptr(out), n, cin, cout,
                 i3(*half_dims), mode, stream())
    else:
        ind = (2 * hx, 2 * hy, 2 * hz) if mode == 1 else (hx, hy, hz)
        LIB.call("bcp_conv_direct_fwd", ptr(inp), ptr(pack.k[0] if mode == 1 else pack.k[2]), ptr(bias), ptr(out), n, cin, cout,
                 i3(*ind), i3(2, 2, 2), i3(2, 2, 2), i3(0, 0, 0), 0 if mode == 1 else 1, stream())
    return out


def _s2_wgrad(full, half, c_full, c_half, half_dims, wshape, into=None):
    """dW[c_half][c_full][8] = sum_i half[i] (x) full[2i+t]."""
    n = full.shape[0]
    hx, hy, hz = half_dims
    if _TC_WGRAD and LIB.query("bcp_conv_tc_s2_wgrad_supported", c_half, c_full, i3(*half_dims)):
        ws = _f32(LIB.query("bcp_conv_tc_s2_wgrad_workspace_floats", n, c_half, c_full, i3(*half_dims)), full.device)
        dw = into if into is not None else torch.empty(wshape, dtype=torch.float32, device=full.device)
        LIB.call("bcp_conv_tc_s2_wgrad", ptr(full), ptr(half), ptr(dw), ptr(ws), _counter(full.device).data_ptr() + 16, n, c_half,
                 c_full, i3(*half_dims), 1 if into is not None else 0, stream())
        return None if into is not None else dw
    return _wgrad(full, half, c_full, c_half, (2 * hx, 2 * hy, 2 * hz), (2, 2, 2), (2, 2, 2), (0, 0, 0), wshape, allow_tc=False,
                  into=into)


class ConvDown2(Function):
    """nn.Conv3d(k=2, s=2) (networks/VNet.py:74).  weight [Cout][Cin][2,2,2]; packs: kind 0 (fwd), kinds 2/3 (dgrad)."""

    @staticmethod
    def forward(ctx, a, weight, bias, pack: ConvPack, bias_grad_is_zero=False):
        ctx.bz, ctx.bias_ref = bias_grad_is_zero, bias
        _require_cuda(a, "conv_down2")
        a = a.contiguous()
        n, cin, x, y, z = act_dims(a)
        cout = weight.shape[0]
        ctx.save_for_backward(a, weight)
        ctx.pack, ctx.has_bias = pack, bias is not None
        return _s2_fwd(a, pack, bias, cin, cout, (x // 2, y // 2, z // 2), 1)

    @staticmethod
    def backward(ctx, dy):
        a, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, cin, x, y, z = act_dims(a)
        cout = weight.shape[0]
        half = (x // 2, y // 2, z // 2)
        da = dw = db = None
        if ctx.needs_input_grad[0]:
            da = _s2_fwd(dy, ctx.pack, None, cout, cin, half, 2)
        if ctx.needs_input_grad[1]:
            into = _direct(weight)
            if into is not None:
                _wgrad_async(a.device, (a, dy), lambda: _s2_wgrad(a, dy, cin, cout, half, weight.shape, into=into))
            else:
                dw = _s2_wgrad(a, dy, cin, cout, half, weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy, cout, ctx.bz, _direct(ctx.bias_ref))
        return da, dw, db, None, None


class ConvUp2(Function):
    """nn.ConvTranspose3d(k=2, s=2) (networks/VNet.py:101).  weight [Cin][Cout][2,2,2]; packs: kinds 2/3 (fwd), kind 0 (dgrad)."""

    @staticmethod
    def forward(ctx, a, weight, bias, pack: ConvPack, bias_grad_is_zero=False):
        ctx.bz, ctx.bias_ref = bias_grad_is_zero, bias
        _require_cuda(a, "conv_up2")
        a = a.contiguous()
        n, cin, x, y, z = act_dims(a)
        cout = weight.shape[1]
        ctx.save_for_backward(a, weight)
        ctx.pack, ctx.has_bias = pack, bias is not None
        return _s2_fwd(a, pack, bias, cin, cout, (x, y, z), 2)

    @staticmethod
    def backward(ctx, dy):
        a, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, cin, x, y, z = act_dims(a)
        cout = weight.shape[1]
        da = dw = db = None
        if ctx.needs_input_grad[0]:
            da = _s2_fwd(dy, ctx.pack, None, cout, cin, (x, y, z), 1)
        if ctx.needs_input_grad[1]:
            into = _direct(weight)                                                                # half = layer input, full = dy
            if into is not None:
                _wgrad_async(a.device, (a, dy), lambda: _s2_wgrad(dy, a, cout, cin, (x, y, z), weight.shape, into=into))
            else:
                dw = _s2_wgrad(dy, a, cout, cin, (x, y, z), weight.shape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy, cout, ctx.bz, _direct(ctx.bias_ref))
        return da, dw, db, None, None


class ConvFirst(Function):
    """First layer, Cin = 1, planar fp32 input (networks/VNet.py:151 block_one, networks/unet.py:72 in_conv)."""

    @staticmethod
    def forward(ctx, x, weight, bias, bias_grad_is_zero=False):
        ctx.bz, ctx.bias_ref = bias_grad_is_zero, bias
        _require_cuda(x, "conv_first")
        x = x.contiguous().float()
        assert x.shape[1] == 1, "conv_first takes single-channel input"
        n = x.shape[0]
        sp = tuple(x.shape[2:])
        dims = (1,) + sp if len(sp) == 2 else sp
        kernel = (1,) + tuple(weight.shape[2:]) if len(sp) == 2 else tuple(weight.shape[2:])
        cout = weight.shape[0]
        w = weight.contiguous()
        out = torch.empty(cb8_shape(n, cout, *dims), dtype=BF16, device=x.device)
        LIB.call("bcp_conv_first_fwd", ptr(x), ptr(w), ptr(bias), ptr(out), n, cout, i3(*dims), i3(*kernel), stream())
        ctx.save_for_backward(x, weight)
        ctx.meta = (n, cout, dims, kernel, bias is not None)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        n, cout, dims, kernel, has_bias = ctx.meta
        dy = dy.contiguous()
        dw = db = None
        if ctx.needs_input_grad[1]:
            into = _direct(weight)

            def run(dst, acc):
                ws = _f32(LIB.query("bcp_conv_first_wgrad_workspace_floats", n, cout, i3(*dims), i3(*kernel)), x.device)
                LIB.call("bcp_conv_first_wgrad", ptr(x), ptr(dy), ptr(dst), ptr(ws), n, cout, i3(*dims), i3(*kernel), acc, stream())
            if into is not None:
                _wgrad_async(x.device, (x, dy), lambda: run(into, 1))
            else:
                dw = torch.empty(weight.shape, dtype=torch.float32, device=x.device)
                run(dw, 0)
        if has_bias and ctx.needs_input_grad[2]:
            db = _bias_grad(dy, cout, ctx.bz, _direct(ctx.bias_ref))
        return None, dw, db, None


class Head(Function):
    """Classifier conv to planar fp32 logits: nn.Conv3d(16, 2, 1) (networks/VNet.py:210) / nn.Conv2d(16, 4, 3, p=1)
    (networks/unet.py:102)."""

    @staticmethod
    def forward(ctx, a, weight, bias, two_d):
        _require_cuda(a, "head")
        a = a.contiguous()
        n, cin, x, y, z = act_dims(a)
        ncls = weight.shape[0]
        kernel = (1,) + tuple(weight.shape[2:]) if two_d else tuple(weight.shape[2:])
        w = weight.contiguous()
        sp = (y, z) if two_d else (x, y, z)
        out = torch.empty((n, ncls) + sp, dtype=torch.float32, device=a.device)
        LIB.call("bcp_head_fwd", ptr(a), ptr(w), ptr(bias), ptr(out), n, cin, ncls, i3(x, y, z), i3(*kernel), stream())
        ctx.save_for_backward(a, weight)
        ctx.meta = (n, cin, ncls, (x, y, z), kernel, bias is not None)
        ctx.bias_ref = bias
        return out

    @staticmethod
    def backward(ctx, dlog):
        a, weight = ctx.saved_tensors
        n, cin, ncls, dims, kernel, has_bias = ctx.meta
        dlog = dlog.contiguous().float()
        w = weight.contiguous()
        da = dw = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            LIB.call("bcp_head_dgrad", ptr(dlog), ptr(w), ptr(da), n, cin, ncls, i3(*dims), i3(*kernel), stream())
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            iw, ib = _direct(weight), _direct(ctx.bias_ref)
            direct = iw is not None and (ib is not None or not has_bias)
            dw = iw if direct else torch.empty(weight.shape, dtype=torch.float32, device=a.device)
            db = (ib if direct else torch.empty(ncls, dtype=torch.float32, device=a.device)) if has_bias else None

            def run():
                ws = _f32(LIB.query("bcp_head_wgrad_workspace_floats", n, cin, ncls, i3(*dims), i3(*kernel)), a.device)
                LIB.call("bcp_head_wgrad", ptr(a), ptr(dlog), ptr(dw), ptr(db), ptr(ws), n, cin, ncls, i3(*dims), i3(*kernel),
                         1 if direct else 0, stream())
            if direct:
                _wgrad_async(a.device, (a, dlog), run)
                dw = db = None
            else:
                run()
        return da, dw, db, None


class GradBoundary(Function):
    """Identity in forward.  In backward it tells the optimiser that every parameter gradient stored at arena offset >=
    ``lo`` is final (parameters are laid out in forward order and autograd runs later-created nodes first), so the
    data-parallel all-reduce of that slice can start on the side stream while backward continues."""

    @staticmethod
    def forward(ctx, a, rt, lo):
        ctx.rt, ctx.lo = rt, lo
        return a.view_as(a)

    @staticmethod
    def backward(ctx, g):
        cb = getattr(ctx.rt, "grad_ready_cb", None)
        if cb is not None:
            cb(ctx.lo)
        return g, None, None


# ----------------------------------------------------------------------------------------------
# normalisation + activation (+ dropout) (+ residual)
# ----------------------------------------------------------------------------------------------
class NormAct(Function):
    """act(norm(y)) [* dropout] [+ residual] on CB8.

    mode 'batch': statistics over groups of `spg` samples (train-mode BatchNorm, InstanceNorm with spg=1);
    running stats (if given) are updated once per group, in order.  mode 'eval': running statistics.
    mode 'none': no normalisation (normalization='none' of networks/VNet.py:6)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, nbt, mode, spg, eps, momentum, slope,
                chan_scale, elem_keep, elem_scale, residual, precomputed=None):
        """precomputed: (stat, coef) already produced by the convolution's epilogue (ConvSame stats_req) for mode 'batch'."""
        _require_cuda(y, "norm_act")
        y = y.contiguous()
        n, c, x, yy, z = act_dims(y)
        s = x * yy * z
        dev = y.device
        if mode != "batch":
            spg = n
        groups = n // spg
        if mode == "batch" and precomputed is not None:
            stat, coef = precomputed
        else:
            stat = torch.empty(groups, c, 2, dtype=torch.float32, device=dev)
            coef = torch.empty(groups, c, 2, dtype=torch.float32, device=dev)
        if residual is not None:
            residual = residual.contiguous()
        if mode == "batch" and precomputed is None and _NORM_FUSED and LIB.query("bcp_norm_fused_supported", n, c, s, spg):
            # statistics + finalize + normalise/activation in ONE cluster kernel
            out = torch.empty_like(y)
            ws = _f32(groups * c * 2, dev)
            LIB.call("bcp_norm_fused_fwd", ptr(y), ptr(out), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(nbt),
                     ptr(stat), ptr(coef), ptr(ws), _counter(dev).data_ptr() + 32, ptr(chan_scale), ptr(elem_keep),
                     float(elem_scale), ptr(residual), n, c, s, spg, float(eps), float(momentum), float(slope), stream())
            ctx.save_for_backward(y, stat, coef, chan_scale, elem_keep)
            ctx.meta = (n, c, s, spg, float(slope), float(elem_scale), mode, gamma is not None, beta is not None,
                        residual is not None)
            ctx.affine_refs = (gamma, beta)
            return out
        if mode == "batch" and precomputed is not None:
            pass
        elif mode == "batch":
            ws = _f32(LIB.query("bcp_norm_workspace_floats", n, c, s), dev)
            LIB.call("bcp_norm_stats", ptr(y), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(nbt),
                     ptr(stat), ptr(coef), ptr(ws), ptr(_counter(dev)), n, c, s, spg, float(eps), float(momentum), stream())
        elif mode == "eval":
            LIB.call("bcp_norm_eval_coef", ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(stat), ptr(coef),
                     c, groups, float(eps), stream())
        else:
            stat[..., 0] = 0.0
            stat[..., 1] = 1.0
            coef[..., 0] = 1.0
            coef[..., 1] = 0.0
        out = torch.empty_like(y)
        LIB.call("bcp_norm_apply", ptr(y), ptr(out), ptr(coef), ptr(chan_scale), ptr(elem_keep), float(elem_scale),
                 ptr(residual), n, c, s, spg, float(slope), stream())
        ctx.save_for_backward(y, stat, coef, chan_scale, elem_keep)
        ctx.meta = (n, c, s, spg, float(slope), float(elem_scale), mode, gamma is not None, beta is not None,
                    residual is not None)
        ctx.affine_refs = (gamma, beta)
        return out

    @staticmethod
    def backward(ctx, da):
        y, stat, coef, chan_scale, elem_keep = ctx.saved_tensors
        n, c, s, spg, slope, elem_scale, mode, has_g, has_b, has_res = ctx.meta
        da = da.contiguous()
        dev = y.device
        dy = torch.empty_like(y)
        need_affine = mode != "none" and (has_g or has_b)
        ig, ib = _direct(ctx.affine_refs[0]), _direct(ctx.affine_refs[1])
        direct = need_affine and ig is not None and ib is not None
        dgamma = (ig if direct else torch.empty(c, dtype=torch.float32, device=dev)) if need_affine else None
        dbeta = (ib if direct else torch.empty(c, dtype=torch.float32, device=dev)) if need_affine else None
        sums = torch.empty(n // spg, c, 2, dtype=torch.float32, device=dev)
        if _NORM_FUSED and LIB.query("bcp_norm_fused_supported", n, c, s, spg):
            LIB.call("bcp_norm_fused_bwd", ptr(da), ptr(y), ptr(dy), ptr(stat), ptr(coef), ptr(chan_scale), ptr(elem_keep), elem_scale,
                     ptr(dgamma), ptr(dbeta), ptr(sums), _counter(dev).data_ptr() + 36, n, c, s, spg, slope,
                     1 if mode == "batch" else 0, 1 if direct else 0, stream())
            if direct:
                dgamma = dbeta = None
            return (dy, dgamma if has_g else None, dbeta if has_b else None, None, None, None, None, None, None, None, None,
                    None, None, None, da if has_res else None, None)
        ws = _f32(LIB.query("bcp_norm_workspace_floats", n, c, s), dev)
        LIB.call("bcp_norm_bwd", ptr(da), ptr(y), ptr(dy), ptr(stat), ptr(coef), ptr(chan_scale), ptr(elem_keep), elem_scale,
                 ptr(dgamma), ptr(dbeta), ptr(sums), ptr(ws), ptr(_counter(dev)), n, c, s, spg, slope,
                 1 if mode == "batch" else 0, 1 if direct else 0, stream())
        if direct:
            dgamma = dbeta = None
        return (dy, dgamma if has_g else None, dbeta if has_b else None, None, None, None, None, None, None, None, None,
                None, None, None, da if has_res else None, None)


# ----------------------------------------------------------------------------------------------
# resampling
# ----------------------------------------------------------------------------------------------
class MaxPool2(Function):
    @staticmethod
    def forward(ctx, a):
        a = a.contiguous()
        n, c, x, y, z = act_dims(a)
        out = torch.empty(cb8_shape(n, c, x, y // 2, z // 2), dtype=BF16, device=a.device)
        LIB.call("bcp_maxpool2_fwd", ptr(a), ptr(out), n * (c // 8) * x, y, z, stream())
        ctx.save_for_backward(a)
        return out

    @staticmethod
    def backward(ctx, g):
        (a,) = ctx.saved_tensors
        n, c, x, y, z = act_dims(a)
        din = torch.empty_like(a)
        LIB.call("bcp_maxpool2_bwd", ptr(a), ptr(g.contiguous()), ptr(din), n * (c // 8) * x, y, z, stream())
        return din


class Upsample2(Function):
    @staticmethod
    def forward(ctx, a):
        a = a.contiguous()
        n, c, x, y, z = act_dims(a)
        out = torch.empty(cb8_shape(n, c, x, 2 * y, 2 * z), dtype=BF16, device=a.device)
        LIB.call("bcp_upsample2_fwd", ptr(a), ptr(out), n * (c // 8) * x, y, z, stream())
        ctx.meta = (n, c, x, y, z)
        return out

    @staticmethod
    def backward(ctx, g):
        n, c, x, y, z = ctx.meta
        din = torch.empty(cb8_shape(n, c, x, y, z), dtype=BF16, device=g.device)
        LIB.call("bcp_upsample2_bwd", ptr(g.contiguous()), ptr(din), n * (c // 8) * x, y, z, stream())
        return din


def maxpool3d_k3s2(a, c):
    n, _, x, y, z = act_dims(a)
    out = torch.empty((n, c, (x - 3) // 2 + 1, (y - 3) // 2 + 1, (z - 3) // 2 + 1), dtype=torch.float32, device=a.device)
    LIB.call("bcp_maxpool3d_k3s2", ptr(a.contiguous()), ptr(out), n, c, x, y, z, stream())
    return out


# ----------------------------------------------------------------------------------------------
# step primitives
# ----------------------------------------------------------------------------------------------
_CONST_BOXES = {}


def box_tensor(box, device):
    """{x0,y0,z0,px,py,pz} int32 on the device.  Accepts a 6-/4-tuple (2-D boxes get z0=0,pz=1) or an int32 device tensor
    (used as is -- the captured-graph path keeps one static tensor and refreshes it before each replay)."""
    if torch.is_tensor(box):
        assert box.dtype == torch.int32 and box.numel() == 6 and box.is_cuda
        return box
    box = tuple(int(v) for v in box)
    if len(box) == 4:
        box = (box[0], box[1], 0, box[2], box[3], 1)
    if torch.cuda.is_current_stream_capturing():
        # a host constant inside a captured step (e.g. the all-zero box of the pre-training loss): one cached device tensor per
        # value, fed from a pinned buffer that stays alive, so the capture records at most one memcpy node for it
        key = (device.type, device.index, box)
        hit = _CONST_BOXES.get(key)
        if hit is None:
            pinned = torch.tensor(box, dtype=torch.int32).pin_memory()
            hit = (pinned, pinned.to(device, non_blocking=True))
            _CONST_BOXES[key] = hit
        return hit[1]
    return torch.tensor(box, dtype=torch.int32).to(device, non_blocking=True)


def _dims3(t, lead):
    sp = tuple(t.shape[lead:])
    return (sp[0], sp[1], 1) if len(sp) == 2 else sp


def mask_mix(a: torch.Tensor, b: torch.Tensor, box, out: torch.Tensor | None = None) -> torch.Tensor:
    """out = a*M + b*(1-M) with the implicit box mask (LA_BCP_train.py:248-249).  a, b: fp32 [N,C,...]."""
    _require_cuda(a, "mask_mix")
    a, b = a.contiguous().float(), b.contiguous().float()
    X, Y, Z = _dims3(a, 2)
    bt = box_tensor(box, a.device)
    if out is None:
        out = torch.empty_like(a)
    assert out.is_contiguous() and out.shape == a.shape and out.dtype == torch.float32
    LIB.call("bcp_mask_mix", ptr(a), ptr(b), ptr(out), a.shape[0], a.shape[1], X, Y, Z, ptr(bt), stream())
    return out


def label_mix(a: torch.Tensor, b: torch.Tensor, box) -> torch.Tensor:
    """uint8 label maps: a outside the box, b inside (label_batch = lab_a*M + lab_b*(1-M), LA_BCP_train.py:156)."""
    _require_cuda(a, "label_mix")
    a, b = to_u8_labels(a).contiguous(), to_u8_labels(b).contiguous()
    X, Y, Z = _dims3(a, 1)
    bt = box_tensor(box, a.device)
    out = torch.empty_like(a)
    LIB.call("bcp_label_mix", ptr(a), ptr(b), ptr(out), a.shape[0], X, Y, Z, ptr(bt), stream())
    return out


def pseudo_label(logits: torch.Tensor, mode: str = "thresh", thr: float = 0.5) -> torch.Tensor:
    """uint8 pseudo labels from planar fp32 logits (LA_BCP_train.py:57-60 / ACDC_BCP_train.py:112-114)."""
    _require_cuda(logits, "pseudo_label")
    logits = logits.contiguous().float()
    n, c = logits.shape[:2]
    v = logits[0, 0].numel()
    out = torch.empty((n,) + tuple(logits.shape[2:]), dtype=torch.uint8, device=logits.device)
    LIB.call("bcp_pseudo_label", ptr(logits), ptr(out), n, c, v, 0 if mode == "thresh" else 1, float(thr), stream())
    return out


def largest_cc(seg_u8: torch.Tensor, connectivity: int | None = None, out_float: bool = False) -> torch.Tensor:
    """Per-(sample, class) largest connected component on the GPU (LA_BCP_train.py:65-77, ACDC_BCP_train.py:89-109)."""
    _require_cuda(seg_u8, "largest_cc")
    seg = seg_u8.contiguous()
    assert seg.dtype == torch.uint8
    n = seg.shape[0]
    sp = tuple(seg.shape[1:])
    X, Y, Z = (1,) + sp if len(sp) == 2 else sp
    conn = connectivity if connectivity is not None else len(sp)
    ws = torch.empty(LIB.query("bcp_largest_cc_workspace_bytes", n, X * Y * Z), dtype=torch.uint8, device=seg.device)
    out = torch.empty(seg.shape, dtype=torch.float32 if out_float else torch.uint8, device=seg.device)
    LIB.call("bcp_largest_cc", ptr(seg), None if out_float else ptr(out), ptr(out) if out_float else None, ptr(ws),
             n, X, Y, Z, conn, stream())
    return out


class MixLoss(Function):
    """Returns a 3-vector [ (dice+ce)/2, dice, ce ] (device).  form 0 = LA/PAN, form 1 = ACDC."""

    @staticmethod
    def forward(ctx, logits, lab_img, lab_patch, box, mask_u8, form, w_img, w_patch):
        _require_cuda(logits, "mix_loss")
        logits = logits.contiguous().float()
        n, c = logits.shape[:2]
        X, Y, Z = _dims3(logits, 2)
        if box is None:
            box = (0, 0, 0, 0, 0, 0)
        box6 = box_tensor(box, logits.device)
        if mask_u8 is not None:
            mask_u8 = mask_u8.contiguous()
            assert mask_u8.dtype == torch.uint8 and mask_u8.numel() == n * X * Y * Z
        li = lab_img.contiguous()
        lp = lab_patch.contiguous()
        assert li.dtype == torch.uint8 and lp.dtype == torch.uint8
        dev = logits.device
        cbuf = _f32(LIB.query("bcp_mix_loss_ctx_floats", n, c), dev)
        ws = _f32(LIB.query("bcp_mix_loss_workspace_floats", n, c, X * Y * Z), dev)
        LIB.call("bcp_mix_loss_fwd", ptr(logits), ptr(li), ptr(lp), ptr(mask_u8), ptr(cbuf), ptr(ws), n, c, X, Y, Z, ptr(box6),
                 int(form), float(w_img), float(w_patch), stream())
        ctx.save_for_backward(logits, li, lp, cbuf, mask_u8, box6)
        ctx.meta = (n, c, X, Y, Z)
        return cbuf[:3].clone()

    @staticmethod
    def backward(ctx, g3):
        logits, li, lp, cbuf, mask_u8, box6 = ctx.saved_tensors
        n, c, X, Y, Z = ctx.meta
        g3 = g3.contiguous().float()
        dlog = torch.empty_like(logits)
        LIB.call("bcp_mix_loss_bwd", ptr(logits), ptr(li), ptr(lp), ptr(mask_u8), ptr(cbuf), ptr(g3), ptr(dlog), n, c, X, Y, Z,
                 ptr(box6), stream())
        return dlog, None, None, None, None, None, None, None


class DiceProb(Function):
    """DiceLoss on probabilities (utils/losses.py:113-134 with softmax=False, as ACDC_BCP_train.py:170,175 calls it)."""

    @staticmethod
    def forward(ctx, probs, target_u8, mask_u8):
        _require_cuda(probs, "dice_prob")
        probs = probs.contiguous().float()
        n, c = probs.shape[:2]
        v = probs[0, 0].numel()
        t = target_u8.contiguous()
        assert t.dtype == torch.uint8 and t.numel() == n * v
        if mask_u8 is not None:
            mask_u8 = mask_u8.contiguous()
            assert mask_u8.dtype == torch.uint8 and mask_u8.numel() == n * v
        cbuf = _f32(LIB.query("bcp_dice_prob_ctx_floats", c), probs.device)
        ws = _f32(LIB.query("bcp_dice_prob_workspace_floats", n, c, v), probs.device)
        LIB.call("bcp_dice_prob_fwd", ptr(probs), ptr(t), ptr(mask_u8), ptr(cbuf), ptr(ws), n, c, v, stream())
        ctx.save_for_backward(probs, t, mask_u8, cbuf)
        ctx.meta = (n, c, v)
        return cbuf[0].clone()

    @staticmethod
    def backward(ctx, g):
        probs, t, mask_u8, cbuf = ctx.saved_tensors
        n, c, v = ctx.meta
        g = g.contiguous().float().reshape(1)
        d = torch.empty_like(probs)
        LIB.call("bcp_dice_prob_bwd", ptr(probs), ptr(t), ptr(mask_u8), ptr(cbuf), ptr(g), ptr(d), n, c, v, stream())
        return d, None, None


def to_u8_labels(t: torch.Tensor) -> torch.Tensor:
    """int64 / float32 / uint8 label maps -> uint8 (the reference casts with .type(torch.int64), utils/BCP_utils.py:59)."""
    return t if t.dtype == torch.uint8 else t.to(torch.uint8)
