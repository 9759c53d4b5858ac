import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bts_b200 import conv
Cin, H, W, Cout, k, dil, up = [int(a) for a in sys.argv[1:8]]
B = 16
x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
Ho, Wo = (2 * H, 2 * W) if up else (H, W)
gy = torch.randn(B, Cout, Ho, Wo, device="cuda").contiguous(memory_format=torch.channels_last)
w = torch.randn(Cout, Cin, k, k, device="cuda")
for _ in range(3):
    gw = conv.wgrad_tc(x, gy, w.shape, w.stride(), 1, dil * (k // 2), dil, upsample2=bool(up))
torch.cuda.synchronize()
print("ok")
