"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim that lets the UNMODIFIED reference modules under /root/reference
(`AVE/model/Swin_AVE.py`, `AVE/model/CLIP_AVE.py`) be imported in this container, where
`timm`, `clip` and `loratorch` are not installed.  The reference only needs three symbols
from timm (`Swin_AVE.py:1,5`: DropPath, to_2tuple, trunc_normal_); `clip` is touched only
when a pretrained path is given (`CLIP_AVE.py:818`) and `loratorch` is imported but unused
(`CLIP_AVE.py:8`).

/root/reference exists only in the build container.  `oracle/build_ref.py` stages verbatim copies of the files
the tests and the baseline arms execute in the git-ignored `oracle/_ref/`, which travels to the GPU box: when
/root/reference is absent `load_reference()` reads from there.  Used by `oracle/make_goldens.py`,
`oracle/validate_oracle.py` (pin the oracle restatement to the real reference), tests/test_reference_gpu.py and
the `--impl reference` / `--impl eager` arms of bench.py.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("STGCMA_REFERENCE_ROOT", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """The reference checkout when it exists (build container), else the staged copy oracle/_ref (GPU box), else None."""
    for root in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isfile(os.path.join(root, "AVE", "model", "Swin_AVE.py")):
            return root
    return None


class _DropPath(nn.Module):
    """Per-sample stochastic depth (the standard timm definition)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


def _to_2tuple(x):
    if isinstance(x, (tuple, list)):
        return tuple(x)
    return (x, x)


def install_shims():
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = _to_2tuple
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules["timm"] = timm
        sys.modules["timm.models"] = models
        sys.modules["timm.models.layers"] = layers
    for name in ("clip", "loratorch"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)


def reference_available():
    return reference_root() is not None


def load_reference(relpath="AVE/model/Swin_AVE.py", modname="_ref_Swin_AVE"):
    """Import one reference source file (unmodified) as a module."""
    if modname in sys.modules:
        return sys.modules[modname]
    install_shims()
    path = os.path.join(reference_root(), relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_caller():
    """The UNCHANGED caller `AVE/traintest_adapt_ave29.py` (train / validate) with its `from utilities import *`."""
    root = reference_root()
    if root not in sys.path:
        sys.path.insert(0, root)          # `utilities` is a top-level package of the reference checkout
    return load_reference("AVE/traintest_adapt_ave29.py", "_ref_traintest_adapt_ave29")
