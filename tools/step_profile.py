"""Kernel-time breakdown of one K16 training step (torch profiler / CUPTI)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bts
from torch.profiler import profile, ProfilerActivity

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
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = prof.key_averages()
rows = sorted(((e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count, e.key) for e in evs), reverse=True)
tot = sum(r[0] for r in rows)
print("total device time %.1f ms over %d kernel names" % (tot / 1e3, len(rows)))
for t, c, k in rows[:40]:
    print("%7.2f ms %5.1f%% %6d  %s" % (t / 1e3, 100 * t / tot, c, k[:110]))
