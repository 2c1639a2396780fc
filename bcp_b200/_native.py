"""ctypes binding of libbcp_b200.so (the C ABI in include/bcp_b200.h).

The prototypes are parsed from the header itself so the Python side cannot drift from the ABI.
There is NO fallback: if the library is missing it is built with nvcc (bcp_b200/build.py); if that
fails, or a kernel reports an error, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "bcp_b200.h")
LIB_PATH = os.path.join(HERE, "libbcp_b200.so")

_CT = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "cudaStream_t": ctypes.c_void_p,
}

# kernels launched per C-ABI call (for bench.py's gpu_launches claim)
def _norm_bwd_kernels(args):
    # (dact, y, dy, stat, coef, chan_scale, elem_keep, elem_scale, dgamma, dbeta, sums, workspace, counter, n, c, s, spg, slope,
    #  stats_grad, accumulate, stream): small layers run reduce+apply in one launch (norm.cu SMALL_LIMIT)
    n, s, stats_grad = int(args[13]), int(args[15]), int(args[18])
    if n * s <= 2048:
        return 1
    return 2 if (stats_grad or args[8] or args[9]) else 1


def _tc_wgrad_kernels(args):
    """bcp_conv_tc_wgrad launches one kernel per z-window (z-lines longer than a TMA box: 254 columns)."""
    z = int(args[8][2])
    return 1 if z + 2 <= 256 else (z + 127) // 128


KERNELS_PER_CALL = {
    "bcp_conv_tc_wgrad": _tc_wgrad_kernels,
    "bcp_norm_stats": 1, "bcp_norm_bwd": _norm_bwd_kernels, "bcp_mix_loss_fwd": 2, "bcp_conv_direct_wgrad": 2,
    "bcp_chan_sum": 2, "bcp_dice_prob_fwd": 2, "bcp_conv_first_wgrad": 2, "bcp_head_wgrad": 2, "bcp_largest_cc": 5,
}


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every function declared in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"typedef struct.*?\}\s*\w+;", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|long long|int)\s+(bcp_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = {"const char*": ctypes.c_char_p, "long long": ctypes.c_longlong, "int": ctypes.c_int}[ret]
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(a.split("*")[-1].strip())
                else:
                    ty, nm = a.rsplit(" ", 1)
                    argtypes.append(_CT[ty.replace("const ", "")])
                    argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


class _Lib:
    def __init__(self):
        self._lib = None
        self._lock = threading.Lock()
        self.protos = parse_header()
        self.launches = 0
        self.calls = {}

    def load(self):
        if self._lib is not None:
            return self._lib
        with self._lock:
            if self._lib is not None:
                return self._lib
            if not os.path.exists(LIB_PATH):
                from . import build as _build
                _build.build()
            lib = ctypes.CDLL(LIB_PATH)
            for name, (restype, argtypes, _) in self.protos.items():
                fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
                fn.restype = restype
                fn.argtypes = argtypes
            if lib.bcp_abi_version() != 1:
                raise RuntimeError("libbcp_b200.so ABI version mismatch")
            self._lib = lib
        return self._lib

    def call(self, name: str, *args):
        lib = self.load()
        rc = getattr(lib, name)(*args)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.bcp_last_error().decode()))
        k = KERNELS_PER_CALL.get(name, 1)
        self.launches += k(args) if callable(k) else k

    def query(self, name: str, *args):
        return getattr(self.load(), name)(*args)


LIB = _Lib()


def i3(a, b, c):
    return (ctypes.c_int * 3)(int(a), int(b), int(c))


def i6(vals):
    return (ctypes.c_int * 6)(*[int(v) for v in vals])


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
