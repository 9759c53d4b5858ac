"""Runs an UNMODIFIED reference script (pytorch/bts_main.py, pytorch/bts_test.py) against the drop-in `bts` module.

    python tests/dropin/boot.py <script.py> <script args...>

Nothing in the script is edited.  What this launcher adds is the ENVIRONMENT the 2019 scripts assume and this image lacks:
  * stub `tensorboardX` / `matplotlib` modules (not installed here);
  * Tensor.__array__ falls back to .cpu() for CUDA tensors -- bts_main.py:429 does np.sum over a list of CUDA scalars,
    which torch 1.2 tolerated and torch 2.x refuses;
  * TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1 -- the checkpoint dict holds a numpy array (bts_main.py:366,536), which
    torch >= 2.6's default weights_only=True load rejects.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "stubs"))
os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")

import torch  # noqa: E402

_orig_array = torch.Tensor.__array__


def _array(self, *a, **k):
    if self.is_cuda:
        self = self.detach().cpu()
    return _orig_array(self, *a, **k)


torch.Tensor.__array__ = _array

script = os.path.abspath(sys.argv[1])
sys.argv = sys.argv[1:]
sys.path.insert(0, os.path.dirname(script))      # what `python script.py` does
runpy.run_path(script, run_name="__main__")
