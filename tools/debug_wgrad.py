import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from bts_b200 import conv
torch.manual_seed(0)
for (Cin, Cout, k, pad, dil, H, W) in [(64, 48, 1, 0, 1, 10, 12), (32, 32, 1, 0, 1, 8, 8), (128, 128, 3, 1, 1, 16, 16)]:
    x = torch.randn(2, Cin, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    gy = torch.randn(2, Cout, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device="cuda")
    ref = torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [pad, pad], [dil, dil], False, [0, 0], 1, [False, True, False])[1]
    for prec in (0, 1):
        gw = conv.wgrad_tc(x, gy, w.shape, w.stride(), 1, pad, dil, precision=prec)
        torch.cuda.synchronize()
        print(Cin, Cout, k, "prec", prec, "max|gw|", float(gw.abs().max()), "max|ref|", float(ref.abs().max()),
              "err", float((gw - ref).abs().max() / ref.abs().max()), "nonzero frac", float((gw != 0).float().mean()))
