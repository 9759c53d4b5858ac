"""How reproducible is one training step run twice?  Two identically initialised models, the same batch, gradients compared.
The engine's reductions that are order-dependent: fp32 shared-memory atomics of the conv-epilogue BatchNorm statistics and
fp64 global atomics of the reduce passes.  usage: determinism_probe.py [encoder] [H] [W]"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BTS_B200_PRETRAINED", "0")
import bts  # noqa: E402

enc = sys.argv[1] if len(sys.argv) > 1 else "densenet121_bts"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 160)
dev = torch.device("cuda", 0)
p = types.SimpleNamespace(encoder=enc, max_depth=10.0, dataset="nyu", bts_size=512, pretrained=False)


def build():
    torch.manual_seed(0)
    m = bts.BtsModel(p)
    m.decoder.apply(bts.weights_init_xavier)
    return m.to(dev).train()


g = torch.Generator().manual_seed(100)
x = torch.randn(2, 3, H, W, generator=g).to(dev)
gt = (torch.rand(2, 1, H, W, generator=g) * 10).to(dev)
focal = torch.full((2,), 518.8579, device=dev)
crit = bts.silog_loss(0.85)


def grads(mode):
    from bts_b200 import fused
    fused.EPI_STATS = mode != "no_epilogue_stats"
    m = build()
    out = m(x, focal)
    crit(out[4], gt, gt > 0.1).backward()
    torch.cuda.synchronize()
    return {k: q.grad.clone() for k, q in m.named_parameters() if q.grad is not None}, [o.detach().clone() for o in out]


for mode in ("default", "no_epilogue_stats"):
    (a, oa), (b, ob) = grads(mode), grads(mode)
    num = sum(float((a[k] - b[k]).double().pow(2).sum()) for k in a)
    den = sum(float(a[k].double().pow(2).sum()) for k in a)
    per = sorted(((float((a[k] - b[k]).norm() / a[k].norm().clamp_min(1e-30)), k) for k in a), reverse=True)
    oerr = [float((u - v).abs().max() / u.abs().max()) for u, v in zip(oa, ob)]
    print("%s %dx%d %-18s run-to-run: whole-gradient rel diff %.3g | median per-parameter %.3g | forward outputs %s"
          % (enc, H, W, mode, (num / den) ** 0.5, per[len(per) // 2][0], ["%.2g" % e for e in oerr]))
    for e, k in per[:4]:
        print("      %.3g  %s" % (e, k))
