"""Bring-up probe of the TMA im2col convention: prints the error of modes 1 and 2 (bts_conv_set_tma) against fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from bts_b200 import _lib, conv
L = _lib.lib()
for (B, Cin, H, W, Cout, k, pad, dil) in [(1, 32, 16, 16, 16, 1, 0, 1), (1, 32, 16, 16, 16, 3, 1, 1), (2, 64, 9, 11, 48, 3, 2, 2), (2, 192, 20, 24, 48, 3, 1, 1)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    ref = F.conv2d(x.double(), w.double(), None, 1, pad, dil)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    out = []
    for mode in (0, 1, 2):
        L.bts_conv_set_tma(mode)
        try:
            y = conv.conv2d_tc(xc, w.cuda(), 1, pad, dil)
            torch.cuda.synchronize()
            out.append("%.3g" % float((y.cpu().double() - ref).abs().max() / ref.abs().max()))
        except Exception as e:
            out.append("ERR %s" % e)
    print("B%d C%d %dx%d ->%d k%d p%d d%d : mode0 %s  mode1 %s  mode2 %s" % (B, Cin, H, W, Cout, k, pad, dil, *out), flush=True)
L.bts_conv_set_tma(0)
