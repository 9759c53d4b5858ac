"""Routing sweep of the narrow-output wgrad (wgrad2_tc.cu) against the tap-in-grid kernel (wgrad_tc.cu) on the K16 shapes:
per layer shape, time conv.wgrad_tc with the production routing and with the shifted-dY kernel forced at several split-K
granularities.  Usage: python tools/w2_sweep.py   (prints one line per (shape, setting), CUDA-event times)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bts_b200 import _lib, conv  # noqa: E402

POINTWISE = [  # B, Cin, Cout, H, W  (dense-layer conv1 + transitions of DenseNet-161 at K16)
    (16, 96, 192, 88, 176), (16, 336, 192, 88, 176), (16, 192, 192, 44, 88), (16, 720, 192, 44, 88), (16, 384, 192, 22, 44),
    (16, 1200, 192, 22, 44), (16, 2064, 192, 22, 44), (16, 1056, 192, 11, 22), (16, 2160, 192, 11, 22), (16, 384, 192, 88, 176),
    (16, 512, 256, 22, 44), (16, 960, 256, 44, 88),
]

SHAPES = [  # B, Cin, Cout, H, W, up, pre
    (16, 192, 48, 11, 22, False, True), (16, 192, 48, 22, 44, False, True), (16, 192, 48, 44, 88, False, True),
    (16, 192, 48, 88, 176, False, True), (16, 128, 64, 44, 88, True, False), (16, 64, 32, 88, 176, True, False),
    (16, 161, 64, 176, 352, False, False), (16, 36, 32, 352, 704, False, False),
]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    for (B, Cin, Cout, H, W, up, pre) in SHAPES:
        x = torch.randn(B, Cin, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        Ho, Wo = (2 * H, 2 * W) if up else (H, W)
        gy = torch.randn(B, Cout, Ho, Wo, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        sc = (torch.rand(Cin, generator=g) + 0.5).to(dev) if pre else None
        sh = (torch.randn(Cin, generator=g) * 0.3).to(dev) if pre else None
        wshape, wstr = (Cout, Cin, 3, 3), (Cin * 9, 9, 3, 1)

        def run():
            return conv.wgrad_tc(x, gy, wshape, wstr, 1, 1, 1, pre_scale=sc, pre_shift=sh, pre_relu=pre, upsample2=up)

        flops = 2.0 * B * Ho * Wo * Cout * Cin * 9
        L.bts_wgrad2_set_min_pixels(1 << 40)
        t = timed(run)
        ref = run().clone()
        print("%-34s tap-in-grid kernel          %8.1f us %7.1f TF/s" % ((B, Cin, Cout, H, W, up), t, flops / t / 1e6), flush=True)
        L.bts_wgrad2_set_min_pixels(0)
        for tma in (1, 0):
            L.bts_wgrad2_set_tma(tma)
            for kb in (4, 8, 16, 32, 64):
                L.bts_wgrad2_set_min_kblocks(kb)
                t = timed(run)
                err = float((run() - ref).abs().max() / ref.abs().max())
                print("%-34s shifted-dY %s min_kb %2d %8.1f us %7.1f TF/s  (max diff to tap-in-grid %.1e)"
                      % ("", "tma" if tma else "ldg", kb, t, flops / t / 1e6, err), flush=True)
        L.bts_wgrad2_set_tma(1)
        L.bts_wgrad2_set_min_kblocks(0)
        L.bts_wgrad2_set_min_pixels(-1)


def main_pointwise():
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    for (B, Cin, Cout, H, W) in POINTWISE:
        x = torch.randn(B, Cin, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        gy = torch.randn(B, Cout, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
        sc = (torch.rand(Cin, generator=g) + 0.5).to(dev)
        sh = (torch.randn(Cin, generator=g) * 0.3).to(dev)

        def run():
            return conv.wgrad_tc(x, gy, (Cout, Cin, 1, 1), (Cin, 1, 1, 1), 1, 0, 1, pre_scale=sc, pre_shift=sh, pre_relu=True)

        flops = 2.0 * B * H * W * Cout * Cin
        L.bts_wgrad2_set_pointwise(0)
        t = timed(run)
        ref = run().clone()
        print("1x1 %-30s wgrad_tc (tap-in-grid)      %8.1f us %7.1f TF/s" % ((B, Cin, Cout, H, W), t, flops / t / 1e6), flush=True)
        L.bts_wgrad2_set_pointwise(1)
        L.bts_wgrad2_set_min_pixels(0)
        for kb in (8, 16, 32):
            L.bts_wgrad2_set_min_kblocks(kb)
            t = timed(run)
            err = float((run() - ref).abs().max() / ref.abs().max())
            print("1x1 %-30s wgrad2 tma ring min_kb %2d     %8.1f us %7.1f TF/s  (max diff %.1e)"
                  % ("", kb, t, flops / t / 1e6, err), flush=True)
        L.bts_wgrad2_set_min_kblocks(0)
        L.bts_wgrad2_set_min_pixels(-1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pointwise":
        main_pointwise()
    else:
        main()
