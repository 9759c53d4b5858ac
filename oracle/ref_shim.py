"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference `pytorch/bts.py`.

Only usable where /root/reference exists (the build container).  It does not
travel to the GPU box; nothing under tests -m gpu, smoke() or bench.py imports it.
It is used by oracle/make_golden.py to generate tests/golden/*.npz and by the
CPU tests that pin oracle/bts_oracle.py against the real reference.

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

REFERENCE_DIR = os.environ.get("BTS_REFERENCE_DIR", "/root/reference/pytorch")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "bts.py"))


_cached = None


def load_reference():
    """Returns the reference `bts` module object (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError("reference not mounted at %s" % REFERENCE_DIR)
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
