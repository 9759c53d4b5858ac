"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference `pytorch/bts.py`.

Where /root/reference exists (the build container) the file is loaded from there; `make -C oracle` also copies it,
unmodified, into the git-ignored oracle/_ref/ so that it travels to the GPU box with the working tree (never into
git history).  Used by oracle/make_golden.py to generate tests/golden/*.npz, by the CPU tests that pin
oracle/bts_oracle.py against the real reference, and by bench.py's reference legs (`--impl reference`,
`cpu_baseline`, `gpu_baseline`) -- never by the product path.

Two shims, source untouched:
  (1) torchvision backbone ctors are called with weights=None
      (reference hard-codes pretrained=True, pytorch/bts.py:274-298; no network here)
  (2) on a CPU-only host Tensor.cuda() is the identity
      (reference calls .cuda() inside LPG forward, pytorch/bts.py:140,143)
"""
import importlib.util
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_dir():
    for d in (os.environ.get("BTS_REFERENCE_DIR"), "/root/reference/pytorch", os.path.join(_HERE, "_ref")):
        if d and os.path.isfile(os.path.join(d, "bts.py")):
            return d
    return None


REFERENCE_DIR = _reference_dir()


def reference_available() -> bool:
    return REFERENCE_DIR is not None


_cached = None


class cuda_is_identity:
    """Context: Tensor.cuda() is the identity -- lets the reference's LPG (which calls .cuda() inside forward,
    pytorch/bts.py:140,143) run on the CPU of a box that does have a GPU (bench.py's CPU reference arm)."""

    def __enter__(self):
        self.saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self.saved
        return False


def load_reference():
    """Returns the reference `bts` module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError("reference not found (neither /root/reference/pytorch nor oracle/_ref)")
    import torchvision.models as tvm

    for name in ("densenet121", "densenet161", "resnet50", "resnet101",
                 "resnext50_32x4d", "resnext101_32x8d", "mobilenet_v2"):
        orig = getattr(tvm, name)
        if getattr(orig, "_bts_shimmed", False):
            continue

        def make(o):
            def ctor(pretrained=False, **kw):
                kw.pop("weights", None)
                return o(weights=None, **kw)
            ctor._bts_shimmed = True
            return ctor
        setattr(tvm, name, make(orig))
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    spec = importlib.util.spec_from_file_location("bts_reference", os.path.join(REFERENCE_DIR, "bts.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bts_reference"] = mod
    spec.loader.exec_module(mod)
    _cached = mod
    return mod
