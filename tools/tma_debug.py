"""Bring-up: one K16-shaped training step (small batch) with TMA staging forced on and a synchronize after every engine call;
prints every distinct (kind, layer) that ran and names the call that fails."""
import os, sys, types
os.environ["BTS_B200_SYNC"] = "1"
os.environ.setdefault("BTS_B200_TMA", "3")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bts
from bts_b200 import conv
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "2"))
cfg = bench.CONFIGS["K16"]
p = types.SimpleNamespace(encoder=cfg["encoder"], max_depth=80.0, dataset="kitti", bts_size=512, pretrained=False)
torch.manual_seed(0)
m = bts.BtsModel(p); m.train(); m.decoder.apply(bts.weights_init_xavier); bench.freeze_like_set_misc(m); m.to(dev)
img, focal, gt = bench.synth_batch(cfg, B, 1, dev)
seen = []
orig = conv._traced
def spy(kind, desc, fn, flops=0.0):
    print("CALL %s %s" % (kind, desc), flush=True)
    return orig(kind, desc, fn, flops)
conv._traced = spy
out = m(img, focal)
loss = bts.silog_loss(0.85)(out[4], gt, gt > 1.0)
loss.backward()
torch.cuda.synchronize()
print("STEP OK loss", float(loss))
