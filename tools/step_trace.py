"""Per-shape conv-engine time inside one real K16 training step (BTS_B200_TRACE=1: CUDA events around every call)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bts
from bts_b200 import conv

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(os.environ.get("B", "16"))
p = types.SimpleNamespace(encoder="densenet161_bts", max_depth=80.0, dataset="kitti", bts_size=512, pretrained=False)
model = bts.BtsModel(p); model.train(); model.decoder.apply(bts.weights_init_xavier); bench.freeze_like_set_misc(model); model.to(dev)
opt = bench.make_optimizer(model, torch); crit = bts.silog_loss(0.85)
img, focal, gt = bench.synth_batch(bench.CONFIGS["K16"], B, 1, dev)
def step():
    opt.zero_grad()
    out = model(img, focal)
    loss = crit(out[4], gt, gt > 1.0)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
conv.set_trace(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
rep = conv.trace_report()
tot = sum(r[0] for r in rep)
print("step %.1f ms; traced native conv calls %.1f ms" % (e0.elapsed_time(e1), tot))
bykind = {}
for t, n, (kind, desc), fl in rep:
    bykind[kind] = bykind.get(kind, 0.0) + t
print("by kind:", {k: round(v, 2) for k, v in bykind.items()})
for t, n, (kind, desc), fl in rep[:int(os.environ.get("TOP", "400"))]:
    print("%8.3f ms %4d x %7.3f  %-7s %-40s %7.1f TF/s" % (t, n, t / n, kind, desc, fl / t / 1e9 if t > 0 else 0.0))
