"""GPU parity tests: fused plane head + LPG, silog -- checker: oracle/ and tests/golden."""
import numpy as np
import pytest
import torch

import bts_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_head(c3, r, md, S):
    eq = O.plane_eq_from_head(O.plane_params(c3, md))
    scaled = O.lpg_forward(eq, r).unsqueeze(1) / md
    ds = scaled[:, :, ::S, ::S] if S else None
    return eq, scaled, ds


@pytest.mark.parametrize("r,S,h,w", [(8, 4, 5, 6), (4, 2, 6, 7), (2, 0, 6, 8), (8, 4, 44, 88), (2, 0, 3, 5)])
def test_plane_head_forward(r, S, h, w):
    from bts_b200 import ops
    md = 80.0
    c3 = torch.randn(2, 3, h, w, generator=torch.Generator().manual_seed(r))
    eq, scaled, ds = _oracle_head(c3.double(), r, md, S)
    out = ops.plane_head_lpg(c3.cuda(), r, md, S)
    s = out[0] if S else out
    # compare where the plane denominator is well conditioned (SURVEY 8c hazard (i))
    g = O.lpg_grid(r, torch.float64)
    e = eq.repeat_interleave(r, 2).repeat_interleave(r, 3)
    den = e[:, 0] * g.repeat(w).view(1, 1, -1) + e[:, 1] * g.repeat(h).view(1, -1, 1) + e[:, 2]
    ok = (den.abs() > 0.05).unsqueeze(1)
    rel = ((s.cpu().double() - scaled).abs() / scaled.abs().clamp_min(1e-6))[ok]
    assert rel.max() < 1e-4
    if S:
        assert torch.equal(out[1], s[:, :, ::S, ::S])          # nearest down-sample rule (Q14), exact


@pytest.mark.parametrize("r,S,h,w", [(8, 4, 5, 6), (4, 2, 6, 7), (2, 0, 6, 8), (2, 0, 3, 5)])
def test_plane_head_backward_vs_autograd_fp64(r, S, h, w):
    from bts_b200 import ops
    md = 10.0
    gen = torch.Generator().manual_seed(10 + r)
    c3 = torch.randn(2, 3, h, w, generator=gen) * 0.7          # keeps theta moderate -> den well away from 0
    c3d = c3.double().requires_grad_(True)
    _, scaled, ds = _oracle_head(c3d, r, md, S)
    gs = torch.randn(scaled.shape, generator=gen)
    loss = (scaled * gs.double()).sum()
    gd = None
    if S:
        gd = torch.randn(ds.shape, generator=gen)
        loss = loss + (ds * gd.double()).sum()
    loss.backward()
    c = c3.cuda().requires_grad_(True)
    out = ops.plane_head_lpg(c, r, md, S)
    if S:
        torch.autograd.backward(out, [gs.cuda(), gd.cuda()])
    else:
        out.backward(gs.cuda())
    ref = c3d.grad
    err = (c.grad.cpu().double() - ref).abs().max() / ref.abs().max()
    assert err < 5e-5


def test_silog_golden(golden):
    from bts_b200 import ops
    g = golden("silog")
    est = torch.from_numpy(g["est"]).cuda().requires_grad_(True)
    gt, mask = torch.from_numpy(g["gt"]).cuda(), torch.from_numpy(g["mask"]).cuda()
    loss = ops.silog(est, gt, mask, 0.85)
    assert abs(float(loss.detach()) - float(g["loss"])) < 2e-6 * abs(float(g["loss"]))
    loss.backward()
    np.testing.assert_allclose(est.grad.cpu().numpy(), g["dest"], rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize("n,keep", [(16 * 352 * 704, 0.2), (1000003, 0.95), (7, 1.0), (5, 0.0)])
def test_silog_vs_oracle_sizes_and_masks(n, keep):
    """K16 full size (sparse KITTI-like mask), ragged length (scalar tail), tiny, and the EMPTY mask
    (reference: mean of empty -> NaN loss, zero gradient)."""
    from bts_b200 import ops
    gen = torch.Generator().manual_seed(n % 1000)
    est = torch.rand(n, generator=gen) * 79 + 1
    gt = torch.rand(n, generator=gen) * 79 + 1.5
    mask = torch.rand(n, generator=gen) < keep
    e = est.cuda().requires_grad_(True)
    loss = ops.silog(e, gt.cuda(), mask.cuda(), 0.85)
    loss.backward(torch.tensor(2.0, device="cuda"))
    if keep == 0.0:
        assert torch.isnan(loss.detach()) and float(e.grad.abs().sum()) == 0.0
        return
    ref = O.silog(est.double(), gt.double(), mask, 0.85)
    assert abs(float(loss.detach()) - float(ref)) < 5e-6 * abs(float(ref))
    gref = 2.0 * O.silog_grad(est, gt, mask, 0.85)
    assert ((e.grad.cpu() - gref).abs().max() / gref.abs().max()) < 2e-5
