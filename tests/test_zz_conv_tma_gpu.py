"""GPU parity of the TMA-staged conv path (bts_conv_set_tma(1): activation tiles loaded by cp.async.bulk.tensor im2col,
zero padding by out-of-bounds fill, A_lo derived in shared memory) against torch fp64 and against the LDG-staged path of the
same engine.  Shapes: the layer classes that are eligible in the K16 step (K channels % 32 == 0, stride 1)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, Cin, H, W, Cout, k, pad, dil, pre
    (2, 64, 12, 20, 48, 1, 0, 1, False),       # 1x1, M tail (480 px)
    (1, 32, 16, 16, 16, 3, 1, 1, False),       # 3x3, exact 2 tiles, one chunk per tap
    (2, 192, 9, 11, 48, 3, 1, 1, True),        # dense-layer 3x3 with the BN+ReLU prologue (padding after the pre-op)
    (1, 256, 11, 22, 128, 3, 3, 3, True),      # atrous d=3
    (1, 64, 11, 22, 32, 3, 12, 12, False),     # atrous d=12 (> map height on one axis)
    (1, 256, 6, 8, 256, 1, 0, 1, True),        # 1x1 with prologue, 256-wide tile
    (2, 96, 10, 14, 192, 1, 0, 1, True),       # dense 1x1
    (3, 128, 7, 9, 64, 3, 1, 1, False),        # tiles straddle rows and images
]


@pytest.fixture
def tma():
    from bts_b200 import _lib, conv
    L = _lib.lib()
    prev, prev_cm = L.bts_conv_get_tma(), conv.CHUNK_MAJOR
    conv.CHUNK_MAJOR = False               # TMA staging enumerates K tap-major; the chunk-major layers keep LDG producers
    yield L
    L.bts_conv_set_tma(prev)
    conv.CHUNK_MAJOR = prev_cm


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,pad,dil,pre", CASES)
def test_tma_staged_conv_matches_fp64_and_ldg_path(tma, B, Cin, H, W, Cout, k, pad, dil, pre):
    from bts_b200 import conv
    g = torch.Generator().manual_seed(Cin + Cout + k + dil)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    sc = sh = None
    xd = x.double()
    if pre:
        sc, sh = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.3
        xd = F.relu(xd * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    ref = F.conv2d(xd, w.double(), None, 1, pad, dil)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    kw = dict(pre_scale=sc.cuda() if pre else None, pre_shift=sh.cuda() if pre else None, pre_relu=pre)
    tma.bts_conv_set_tma(0)
    y0 = conv.conv2d_tc(xc, w.cuda(), 1, pad, dil, **kw)
    tma.bts_conv_set_tma(3)                                    # TMA on, strict: a tensor-map failure is an error, not a fallback
    y1 = conv.conv2d_tc(xc, w.cuda(), 1, pad, dil, **kw)
    torch.cuda.synchronize()
    scale = ref.abs().max()
    e0 = float((y0.cpu().double() - ref).abs().max() / scale)
    e1 = float((y1.cpu().double() - ref).abs().max() / scale)
    assert e0 < 2e-5
    # hi = truncation instead of round-to-nearest: |lo| doubles, its own truncation error with it -- still fp32-grade
    assert e1 < 4e-5, "TMA path: rel-to-scale error %.3g (LDG path %.3g)" % (e1, e0)


def test_tma_path_reads_a_channel_slice_of_a_slab_and_feeds_the_stats_epilogue(tma):
    from bts_b200 import conv
    g = torch.Generator().manual_seed(5)
    slab = torch.randn(2, 160, 9, 13, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    x = slab[:, 32:128]                                        # 96 channels at offset 32 of a 160-channel slab
    w = (torch.randn(48, 96, 3, 3, generator=g) / 30).cuda()
    ref = F.conv2d(x.double().cpu(), w.double().cpu(), None, 1, 1, 1)
    tma.bts_conv_set_tma(3)
    st = torch.zeros(2, 48, device="cuda", dtype=torch.float64)
    y = conv.conv2d_tc(x, w, 1, 1, 1, stats=st)
    torch.cuda.synchronize()
    assert float((y.cpu().double() - ref).abs().max() / ref.abs().max()) < 4e-5
    assert torch.allclose(st[0].cpu(), ref.sum((0, 2, 3)), rtol=1e-4, atol=1e-3)
    assert torch.allclose(st[1].cpu(), (ref * ref).sum((0, 2, 3)), rtol=1e-4, atol=1e-3)
