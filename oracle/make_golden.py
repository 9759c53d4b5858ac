"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/pytorch/bts.py through oracle/ref_shim.py) on seeded inputs, on the CPU, fp32.

Run in the build container (the reference is mounted there, not on the GPU box):
    python oracle/make_golden.py
The committed fixtures are what pins oracle/bts_oracle.py and the CUDA path to the reference.
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ref_shim import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def head_planes(B, h, w, max_depth, gen):
    """Plane coefficients through the real head parametrisation (BASELINE.md LPG-u inputs)."""
    z = torch.randn(B, 3, h, w, generator=gen)
    th = torch.sigmoid(z[:, 0]) * np.pi / 3
    ph = torch.sigmoid(z[:, 1]) * np.pi * 2
    d = torch.sigmoid(z[:, 2]) * max_depth
    return torch.stack([torch.sin(th) * torch.cos(ph), torch.sin(th) * torch.sin(ph), torch.cos(th), d], 1).contiguous()


def main():
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    torch.set_num_threads(4)

    # ---- LPG: forward + autograd gradient, r = 2, 4, 8 (pytorch/bts.py:124-146)
    for r in (2, 4, 8):
        gen = torch.Generator().manual_seed(100 + r)
        plane = head_planes(2, 5, 6, 10.0, gen).requires_grad_(True)
        mod = R.local_planar_guidance(r)
        depth = mod(plane, torch.tensor([518.8579, 518.8579]))
        dy = torch.randn(depth.shape, generator=gen)
        depth.backward(dy)
        np.savez_compressed(os.path.join(OUT, "lpg_r%d.npz" % r), plane=plane.detach().numpy(),
                            depth=depth.detach().numpy(), dy=dy.numpy(), dplane=plane.grad.numpy())

    # ---- reduction_1x1 (non-final and final) (pytorch/bts.py:83-122)
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    for final in (False, True):
        mod = R.reduction_1x1(32, 16, 10.0, is_final=final)
        x = torch.randn(2, 32, 6, 7, generator=gen)
        y = mod(x)
        d = {"x": x.numpy(), "y": y.detach().numpy()}
        d.update({"sd." + k: v.numpy() for k, v in mod.state_dict().items()})
        np.savez_compressed(os.path.join(OUT, "reduc_%s.npz" % ("final" if final else "plane")), **d)

    # ---- silog_loss forward + autograd gradient (pytorch/bts.py:41-48)
    gen = torch.Generator().manual_seed(11)
    est = (torch.rand(2, 1, 16, 24, generator=gen) * 79 + 1).requires_grad_(True)
    gt = torch.rand(2, 1, 16, 24, generator=gen) * 80
    gt = torch.where(torch.rand(2, 1, 16, 24, generator=gen) < 0.3, gt, torch.zeros_like(gt))
    mask = gt > 1.0
    loss = R.silog_loss(0.85).forward(est, gt, mask)
    loss.backward()
    np.savez_compressed(os.path.join(OUT, "silog.npz"), est=est.detach().numpy(), gt=gt.numpy(),
                        mask=mask.numpy(), loss=loss.detach().numpy(), dest=est.grad.numpy())

    # ---- whole decoder `bts` (pytorch/bts.py:148-266), small widths, eval and train mode
    feat = [8, 8, 16, 24, 40]
    for dataset, max_depth in (("kitti", 80.0), ("nyu", 10.0)):
        torch.manual_seed(3)
        params = types.SimpleNamespace(max_depth=max_depth, dataset=dataset)
        dec = R.bts(params, feat, num_features=128)
        dec.apply(R.weights_init_xavier)
        gen = torch.Generator().manual_seed(5)
        with torch.no_grad():      # non-trivial BN affine / running stats
            for m in dec.modules():
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.weight.copy_(torch.rand(m.weight.shape, generator=gen) + 0.5)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)
                    m.running_mean.copy_(torch.randn(m.bias.shape, generator=gen) * 0.1)
                    m.running_var.copy_(torch.rand(m.bias.shape, generator=gen) + 0.5)
        H = W = 64
        B = 2
        feats = [torch.randn(B, feat[i], H >> (i + 1), W >> (i + 1), generator=gen) for i in range(5)]
        focal = torch.tensor([721.5377, 707.0912])
        d = {"focal": focal.numpy(), "feat_channels": np.array(feat), "num_features": np.array(128),
             "max_depth": np.array(max_depth)}
        d.update({"feat%d" % i: f.numpy() for i, f in enumerate(feats)})
        d.update({"sd." + k: v.clone().numpy() for k, v in dec.state_dict().items()})
        dec.eval()
        with torch.no_grad():
            outs = dec(feats, focal)
        d.update({"eval_out%d" % i: o.numpy() for i, o in enumerate(outs)})
        dec.train()
        with torch.no_grad():
            outs = dec(feats, focal)
        d.update({"train_out%d" % i: o.numpy() for i, o in enumerate(outs)})
        d.update({"sd_after." + k: v.numpy() for k, v in dec.state_dict().items() if "running" in k})
        np.savez_compressed(os.path.join(OUT, "decoder_%s.npz" % dataset), **d)
    print("golden fixtures written to", os.path.abspath(OUT))
    for f in sorted(os.listdir(OUT)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
