"""Host side of the tcgen05 implicit-GEMM convolution engine (csrc/conv_tc.cu, C-ABI bts_conv_*).

Activations are NHWC in memory: torch tensors of logical shape (B,C,H,W) in `channels_last` format, so the same
tensor is understood by the rest of torch (cat, BatchNorm, autograd) without transposes.  Weights keep the
reference's (Cout,Cin,kh,kw) parameter layout (checkpoint wire format) and are re-packed into the engine's
pre-split, pre-swizzled tile stream whenever the parameter's version counter changes (i.e. once per optimizer step).
"""
import ctypes
import weakref

import torch

from . import _lib
from .ops import _need_cuda, _ptr, _stream

ACT = {None: 0, "none": 0, "elu": 1, "sigmoid": 2}
import os as _os
WGRAD_BACKEND = _os.environ.get("BTS_B200_WGRAD", "tc")     # tc: tcgen05 wgrad kernel | aten: library scaffold

_pack_cache = {}   # id(weight) -> (weakref, version, data_ptr, transpose) -> packed tensor


def pack_weights(weight, transpose_flip=False):
    """(Cout,Cin,KH,KW) fp32 parameter -> packed operator.  Cached on the tensor's version counter."""
    _need_cuda(weight)
    key = (id(weight), bool(transpose_flip))
    ent = _pack_cache.get(key)
    w = weight.detach()
    if ent is not None:
        ref, ver, ptr, packed = ent
        if ref() is weight and ver == weight._version and ptr == w.data_ptr():
            return packed
    Cout, Cin, KH, KW = w.shape
    rows, kch = (Cin, Cout) if transpose_flip else (Cout, Cin)
    L = _lib.lib()
    n = L.bts_conv_packed_floats(rows, kch, KH, KW)
    packed = torch.empty(n, device=w.device, dtype=torch.float32)
    s = w.stride()
    with torch.cuda.device(w.device):
        _lib.check(L.bts_conv_pack_weights(_ptr(w), s[0], s[1], s[2], s[3], Cout, Cin, KH, KW, int(transpose_flip),
                                           _ptr(packed), _stream()), "bts_conv_pack_weights")
    _lib.count()
    _pack_cache[key] = (weakref.ref(weight), weight._version, w.data_ptr(), packed)
    if len(_pack_cache) > 4096:
        for k in [k for k, v in _pack_cache.items() if v[0]() is None]:
            del _pack_cache[k]
    return packed


def _nhwc_view(x):
    """(ptr tensor, pixel stride) of a (B,C,H,W) tensor whose memory is NHWC (channels_last, possibly a channel
    slice of a wider slab).  Anything else is converted (one copy)."""
    B, C, H, W = x.shape
    sb, sc, sh, sw = x.stride()
    ok = (sc == 1 or C == 1) and sh == W * sw and sb == H * sh and sw >= C
    if not ok:
        x = x.contiguous(memory_format=torch.channels_last)
        sb, sc, sh, sw = x.stride()
        if C == 1:               # channels_last of a 1-channel tensor keeps NCHW strides; pixel stride is 1
            sw = 1
    return x, sw


def conv2d_tc(x, weight, stride=1, padding=0, dilation=1, pre_scale=None, pre_shift=None, pre_relu=False,
              upsample2=False, act=None, out=None, precision=0, packed=None, cout=None, transpose_flip=False):
    """Runs the engine.  x: (B,Cin,Hs,Ws) NHWC-in-memory fp32 CUDA.  Returns (B,Cout,Hout,Wout) channels_last.
    `out` may be a pre-allocated channels_last tensor or a channel slice of one (concat-free writes)."""
    _need_cuda(x, weight)
    if x.dtype != torch.float32:
        raise TypeError("conv2d_tc computes in fp32 (3xTF32 on tcgen05); got %s" % x.dtype)
    x, xs = _nhwc_view(x)
    B, Cin, Hs, Ws = x.shape
    Co, Ci, KH, KW = weight.shape
    if transpose_flip:
        Co, Ci = Ci, Co
    if Ci != Cin:
        raise ValueError("weight expects %d input channels, got %d" % (Ci, Cin))
    if packed is None:
        packed = pack_weights(weight, transpose_flip)
    Hin, Win = (2 * Hs, 2 * Ws) if upsample2 else (Hs, Ws)
    Hout = (Hin + 2 * padding - dilation * (KH - 1) - 1) // stride + 1
    Wout = (Win + 2 * padding - dilation * (KW - 1) - 1) // stride + 1
    if out is None:
        out = torch.empty((B, Co, Hout, Wout), device=x.device, dtype=torch.float32,
                          memory_format=torch.channels_last)
        os_ = Co
    else:
        if tuple(out.shape) != (B, Co, Hout, Wout):
            raise ValueError("out has shape %s, expected %s" % (tuple(out.shape), (B, Co, Hout, Wout)))
        o2, os_ = _nhwc_view(out)
        if o2 is not out:
            raise ValueError("out must be NHWC in memory")
    if pre_scale is not None:
        pre_scale = pre_scale.contiguous()
        pre_shift = pre_shift.contiguous()
    with torch.cuda.device(x.device):
        rc = _lib.lib().bts_conv_fwd(_ptr(x), xs, B, Hs, Ws, int(upsample2), Cin, KH, KW, stride, padding, dilation,
                                     _ptr(packed), Co, _ptr(pre_scale), _ptr(pre_shift), int(pre_relu), _ptr(out), os_,
                                     ACT[act], int(precision), _stream())
    _lib.check(rc, "bts_conv_fwd")
    _lib.count()
    return out


def wgrad_tc(x, gy, weight_shape, weight_strides, stride=1, padding=0, dilation=1, pre_scale=None, pre_shift=None,
             pre_relu=False, upsample2=False, precision=0):
    """dW (shaped/strided like the weight parameter) on the tcgen05 engine."""
    _need_cuda(x, gy)
    x, xs = _nhwc_view(x)
    gy, gs = _nhwc_view(gy)
    B, Cin, Hs, Ws = x.shape
    Cout, _, KH, KW = weight_shape
    L = _lib.lib()
    split = ctypes.c_int(0)
    wsf = ctypes.c_longlong(0)
    _lib.check(L.bts_conv_wgrad_plan(B, gy.shape[2], gy.shape[3], Cin, Cout, KH, KW, ctypes.byref(split), ctypes.byref(wsf)),
               "bts_conv_wgrad_plan")
    ws = torch.empty(wsf.value, device=x.device, dtype=torch.float32)
    gw = torch.empty_strided(tuple(weight_shape), tuple(weight_strides), device=x.device, dtype=torch.float32)
    s = weight_strides
    if pre_scale is not None:
        pre_scale, pre_shift = pre_scale.contiguous(), pre_shift.contiguous()
    with torch.cuda.device(x.device):
        rc = L.bts_conv_wgrad(_ptr(x), xs, B, Hs, Ws, int(upsample2), Cin, KH, KW, stride, padding, dilation,
                              _ptr(pre_scale), _ptr(pre_shift), int(pre_relu), _ptr(gy), gs, Cout, _ptr(ws), split.value,
                              _ptr(gw), s[0], s[1], s[2], s[3], int(precision), _stream())
    _lib.check(rc, "bts_conv_wgrad")
    _lib.count(2)
    return gw


class _ConvTC(torch.autograd.Function):
    """Plain convolution (no fused pre/post ops) with autograd, all three GEMMs on the tcgen05 engine: forward,
    dgrad (the same kernel over the transposed, tap-flipped packed operator) and wgrad (MN-major operands)."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding, dilation):
        y = conv2d_tc(x, weight, stride, padding, dilation)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding, dilation)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        stride, padding, dilation = ctx.cfg
        gx = gw = None
        KH = weight.shape[2]
        if ctx.needs_input_grad[0]:
            if stride == 1:
                gx = conv2d_tc(gy, weight, 1, dilation * (KH - 1) - padding, dilation, transpose_flip=True)
            else:
                gx = torch.ops.aten.convolution_backward(gy, x, weight, None, [stride] * 2, [padding] * 2, [dilation] * 2,
                                                         False, [0, 0], 1, [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            if WGRAD_BACKEND == "tc":
                gw = wgrad_tc(x, gy, weight.shape, weight.stride(), stride, padding, dilation)
            else:
                gw = torch.ops.aten.convolution_backward(gy.contiguous(memory_format=torch.channels_last), x, weight,
                                                         None, [stride] * 2, [padding] * 2, [dilation] * 2, False,
                                                         [0, 0], 1, [False, True, False])[1]
        return gx, gw, None, None, None


def conv2d(x, weight, stride=1, padding=0, dilation=1):
    return _ConvTC.apply(x, weight, stride, padding, dilation)
