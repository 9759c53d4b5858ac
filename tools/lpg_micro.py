"""LPG-u microbench driver (BASELINE.json configs[4]) -- used under ncu and for the H x W sweep.
usage: python tools/lpg_micro.py [r] [side] [batch] [iters]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bts_b200 import ops  # noqa: E402

r = int(sys.argv[1]) if len(sys.argv) > 1 else 8
side = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
Bn = int(sys.argv[3]) if len(sys.argv) > 3 else 128
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
h = side // r
z = torch.randn(Bn, 3, h, h, device=dev)
th = torch.sigmoid(z[:, 0]) * math.pi / 3
ph = torch.sigmoid(z[:, 1]) * math.pi * 2
plane = torch.stack([torch.sin(th) * torch.cos(ph), torch.sin(th) * torch.sin(ph), torch.cos(th),
                     torch.sigmoid(z[:, 2]) * 80.0], 1).contiguous().requires_grad_(True)
dy = torch.randn(Bn, side, side, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for it in range(iters + 2):
    ev[0].record()
    d = ops.lpg(plane, r)
    ev[1].record()
    (g,) = torch.autograd.grad(d, plane, dy)
    ev[2].record()
    torch.cuda.synchronize()
    if it >= 2:
        tf += ev[0].elapsed_time(ev[1])
        tb += ev[1].elapsed_time(ev[2])
px = Bn * side * side
bf, bb = 4.0 * px * (1 + 4.0 / r ** 2), 4.0 * px * (1 + 8.0 / r ** 2)
print("r=%d side=%d B=%d  fwd %.3f ms %.0f GB/s | bwd %.3f ms %.0f GB/s | fwd+bwd %.0f GB/s" %
      (r, side, Bn, tf / iters, bf / (tf / iters) / 1e6, tb / iters, bb / (tb / iters) / 1e6,
       (bf + bb) / ((tf + tb) / iters) / 1e6))
