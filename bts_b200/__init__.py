"""bts_b200 -- B200-native (sm_100a) implementation of the BTS dense forward/backward hot path.

Public surface = the reference's `bts` module surface (see bts_b200.model and the root-level bts.py);
native boundary = the C ABI in include/bts_b200.h (libbts_b200.so, loaded by bts_b200._lib).
"""
from .model import (BtsModel, atrous_conv, bn_init_as_tf, bts, encoder, local_planar_guidance,  # noqa: F401
                    reduction_1x1, silog_loss, upconv, weights_init_xavier)

__all__ = ["BtsModel", "atrous_conv", "bn_init_as_tf", "bts", "encoder", "local_planar_guidance",
           "reduction_1x1", "silog_loss", "upconv", "weights_init_xavier"]

import os as _os

import torch as _torch

# Parity is defined in fp32 (north star: <=1e-3 relative on depth; SURVEY Appendix F shows single-pass TF32
# operands miss that bar).  Any convolution that still goes through cuDNN must therefore not silently use
# TF32; our own tcgen05 engine uses the 3xTF32 split.  Opt out with BTS_B200_ALLOW_TF32=1.
if _os.environ.get("BTS_B200_ALLOW_TF32", "0") != "1":
    _torch.backends.cudnn.allow_tf32 = False
    _torch.backends.cuda.matmul.allow_tf32 = False
