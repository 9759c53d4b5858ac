"""Per-shape time of every conv-engine call in one K16 training step (BTS_B200_TRACE=1)."""
import os, sys, types
os.environ["BTS_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bts
from bts_b200 import conv
dev = torch.device("cuda:0")
torch.manual_seed(0)
p = types.SimpleNamespace(encoder="densenet161_bts", max_depth=80.0, dataset="kitti", bts_size=512, pretrained=False)
model = bts.BtsModel(p); model.train(); model.decoder.apply(bts.weights_init_xavier); bench.freeze_like_set_misc(model); model.to(dev)
opt = bench.make_optimizer(model, torch); crit = bts.silog_loss(0.85)
img, focal, gt = bench.synth_batch(16, 1, dev)
def step():
    opt.zero_grad(); out = model(img, focal); loss = crit(out[4], gt, gt > 1.0); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize(); conv.trace_log.clear()
step()
rep = conv.trace_report()
tot = sum(r[0] for r in rep)
bykind = {}
for t, n, (kind, desc) in rep: bykind[kind] = bykind.get(kind, 0) + t
print("total conv-engine time %.1f ms: %s" % (tot, {k: round(v, 1) for k, v in bykind.items()}))
for t, n, (kind, desc) in rep[:60]:
    print("%7.2f ms %4d  %-6s %s" % (t, n, kind, desc))
