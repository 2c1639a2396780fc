"""Import the UNMODIFIED reference modules from /root/reference/code (this container only).

TEST INFRASTRUCTURE.  Used by tests/golden/make_golden.py (fixture generator) and by the CPU
tests that pin oracle/bcp_oracle.py against the real reference when /root/reference exists.
Nothing here is importable on the GPU box (no /root/reference there) and nothing in the product
package imports it.

Shims: dummy ``turtle`` (utils/BCP_utils.py:4 does ``from turtle import pd``),
``skimage`` stub whose ``measure.label`` maps to scipy.ndimage.label with the equivalent
structuring element, dummy ``matplotlib`` for pancreas/Vnet.py:4, and ``.cuda()`` -> identity when
no GPU is present.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("BCP_REFERENCE_ROOT", "/root/reference")
REF_CODE = os.path.join(REF_ROOT, "code")


def available() -> bool:
    return os.path.isdir(REF_CODE)


def _install_stubs():
    import numpy as np
    import torch
    from scipy import ndimage

    if "turtle" not in sys.modules:
        t = types.ModuleType("turtle")
        t.pd = None
        sys.modules["turtle"] = t

    if "skimage" not in sys.modules:
        sk = types.ModuleType("skimage")
        meas = types.ModuleType("skimage.measure")
        seg = types.ModuleType("skimage.segmentation")

        def label(a, connectivity=None, **kw):
            a = np.asarray(a)
            c = a.ndim if connectivity is None else connectivity
            lab, _ = ndimage.label(a != 0, structure=ndimage.generate_binary_structure(a.ndim, c))
            return lab
        meas.label = label
        sk.measure, sk.segmentation = meas, seg
        sys.modules.update({"skimage": sk, "skimage.measure": meas, "skimage.segmentation": seg})

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def load():
    """Returns a namespace with the reference modules: VNet, unet, net_factory, BCP_utils, losses,
    pan_Vnet, pan_losses."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_CODE)
    _install_stubs()
    for p in (REF_CODE, os.path.join(REF_CODE, "pancreas")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    ns = types.SimpleNamespace()
    ns.VNet = importlib.import_module("networks.VNet")
    ns.unet = importlib.import_module("networks.unet")
    ns.net_factory = importlib.import_module("networks.net_factory")
    ns.losses = importlib.import_module("utils.losses")
    ns.BCP_utils = importlib.import_module("utils.BCP_utils")
    # pancreas/ has its own top-level names (Vnet.py, losses.py): load by path to avoid clashes
    import importlib.util

    def by_path(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    ns.pan_Vnet = by_path("ref_pan_Vnet", os.path.join(REF_CODE, "pancreas", "Vnet.py"))
    ns.pan_losses = by_path("ref_pan_losses", os.path.join(REF_CODE, "pancreas", "losses.py"))
    return ns


def extract_defs(path, names):
    """exec the named top-level ``def``s of a reference script that cannot be imported whole
    (LA_BCP_train.py / ACDC_BCP_train.py / pancreas_utils.py parse argv and import py<3.12-only
    modules at import time).  The function sources are taken verbatim from the file."""
    import ast
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    _install_stubs()
    from skimage.measure import label
    src = open(path).read()
    tree = ast.parse(src)
    glb = dict(np=np, torch=torch, nn=nn, F=F, label=label)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = ast.get_source_segment(src, node)
            exec(compile(code, path, "exec"), glb)
    return types.SimpleNamespace(**{n: glb[n] for n in names}), glb
