"""GPU parity of the data-format kernels (csrc/io.cu) against oracle/io_oracle.py (itself pinned to the
reference's loader / eval functions by tests/test_io_oracle.py).  Index work (crop / flip / uint16 quantisation) is
bit-exact; the float transforms within 2 ulp-ish (powf) -- tolerances stated per assertion."""
import numpy as np
import pytest
import torch

import io_oracle as IO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dataset,div", [("nyu", 1000.0), ("kitti", 256.0)])
def test_input_prep_matches_oracle(dataset, div):
    from bts_b200 import ops
    rng = np.random.RandomState(3)
    B, Hs, Ws, H, W = 4, 48, 72, 32, 64
    img = rng.randint(0, 256, (B, Hs, Ws, 3)).astype(np.uint8)
    dep = rng.randint(0, 65535, (B, Hs, Ws)).astype(np.uint16)
    par = np.zeros((B, 9), dtype=np.float32)
    for b in range(B):
        par[b] = [rng.randint(0, Hs - H + 1), rng.randint(0, Ws - W + 1), b % 2, (b // 2) % 2, rng.uniform(0.9, 1.1),
                  rng.uniform(0.75, 1.25), *rng.uniform(0.9, 1.1, 3)]
    gi, gd = ops.input_prep(torch.from_numpy(img).cuda(), torch.from_numpy(par).cuda(), (H, W),
                            torch.from_numpy(dep.view(np.int16)).cuda().view(torch.uint16), div)
    assert gi.shape == (B, 3, H, W) and gi.is_contiguous(memory_format=torch.channels_last) and gd.shape == (B, 1, H, W)
    for b in range(B):
        wi, wd = IO.input_prep(img[b], dep[b], div, int(par[b, 0]), int(par[b, 1]), H, W, par[b, 2] > 0.5, par[b, 3] > 0.5,
                               par[b, 4], par[b, 5], par[b, 6:9])
        np.testing.assert_allclose(gi[b].cpu().numpy(), wi, rtol=2e-6, atol=2e-6)
        np.testing.assert_array_equal(gd[b].cpu().numpy(), wd)          # integer / exact-division work: bit-exact


@pytest.mark.parametrize("crop", [None, (3, 41, 5, 60)])
def test_eval_errors_match_oracle(crop):
    from bts_b200 import ops
    rng = np.random.RandomState(4)
    H, W = 44, 64
    gt = rng.uniform(0, 90, (H, W)).astype(np.float32)
    gt[rng.uniform(size=(H, W)) < 0.3] = 0                          # sparse ground truth (KITTI-like)
    pred = (np.abs(gt) * rng.uniform(0.5, 1.7, (H, W)) + rng.uniform(0, 2, (H, W))).astype(np.float32)
    pred[0, 0], pred[5, 7], pred[9, 9] = np.inf, np.nan, -3.0
    want, n = IO.eval_errors(pred, gt, 1e-3, 80.0, crop)
    got = ops.eval_errors(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 1e-3, 80.0, crop).cpu().numpy()
    assert int(got[9]) == n
    np.testing.assert_allclose(got[:9], np.array(want, dtype=np.float64), rtol=2e-5)


def test_depth_to_u16_is_bit_exact():
    from bts_b200 import ops
    rng = np.random.RandomState(5)
    d = rng.uniform(0, 65, (3, 50, 70)).astype(np.float32)
    for scale in (256.0, 1000.0):
        got = ops.depth_to_u16(torch.from_numpy(d).cuda(), scale).cpu().numpy()
        np.testing.assert_array_equal(got, IO.depth_to_u16(d, scale))
