"""run one conv layer shape a few times (for ncu).  usage: conv_one.py Cin H W Cout k dil up [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bts_b200 import conv
Cin, H, W, Cout, k, dil, up = [int(a) for a in sys.argv[1:8]]
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 3
B = 16
x = torch.randn(B, Cin, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
packed = conv.pack_weights(w)
pre = os.environ.get("PRE", "0") == "1"            # PRE=1: folded BatchNorm + ReLU prologue (the dense-layer form)
sc = (torch.rand(Cin, device="cuda") + 0.5) if pre else None
sh = (torch.randn(Cin, device="cuda") * 0.3) if pre else None
for _ in range(iters):
    y = conv.conv2d_tc(x, w, 1, dil * (k // 2), dil, pre_scale=sc, pre_shift=sh, pre_relu=pre, upsample2=bool(up))
torch.cuda.synchronize()
print("ok", tuple(y.shape))
