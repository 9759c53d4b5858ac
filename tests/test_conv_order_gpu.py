"""GPU parity of the two K orders of the conv engine: dense tap-major quads and chunk-major (32-channel chunk outer, taps
inner -- the L1-friendly order used for multi-tap layers whose K channels are whole chunks).  Same checker and tolerance as
tests/test_conv_gpu.py; forward, dgrad and the fused prologue / statistics epilogue."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[False, True], ids=["tap_major", "chunk_major"])
def order(request):
    from bts_b200 import conv
    prev = conv.CHUNK_MAJOR
    conv.CHUNK_MAJOR = request.param
    conv.invalidate_packed()
    yield request.param
    conv.CHUNK_MAJOR = prev
    conv.invalidate_packed()


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,pad,dil,pre", [
    (2, 192, 9, 11, 48, 3, 1, 1, True),        # dense-layer 3x3 with the BN+ReLU prologue
    (1, 256, 11, 22, 128, 3, 3, 3, True),      # atrous d=3
    (1, 64, 11, 22, 32, 3, 12, 12, False),     # atrous d=12
    (3, 128, 7, 9, 64, 3, 1, 1, False),        # tiles straddle rows and images
    (1, 96, 12, 16, 256, 3, 1, 1, False),      # 256-wide tile
    (1, 32, 20, 24, 16, 7, 3, 1, False),       # 49 taps (more than the 32-bit tap mask covers)
])
def test_conv_fwd_dgrad_both_k_orders(order, B, Cin, H, W, Cout, k, pad, dil, pre):
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
    st = torch.zeros(2, Cout, device="cuda", dtype=torch.float64)
    y = conv.conv2d_tc(xc, w.cuda(), 1, pad, dil, pre_scale=sc.cuda() if pre else None, pre_shift=sh.cuda() if pre else None,
                       pre_relu=pre, stats=st)
    torch.cuda.synchronize()
    assert float((y.cpu().double() - ref).abs().max() / ref.abs().max()) < 2e-5
    assert torch.allclose(st[0].cpu(), ref.sum((0, 2, 3)), rtol=1e-4, atol=1e-3)
    # dgrad = the same engine over the transposed, tap-flipped operator (its K channels are the layer's Cout)
    gy = torch.randn(ref.shape, generator=g)
    gref = torch.nn.grad.conv2d_input(x.shape, w.double(), gy.double(), 1, pad, dil)
    gx = conv.conv2d_tc(gy.cuda().contiguous(memory_format=torch.channels_last), w.cuda(), 1, dil * (k - 1) - pad, dil,
                        transpose_flip=True)
    torch.cuda.synchronize()
    assert float((gx.cpu().double() - gref).abs().max() / gref.abs().max()) < 2e-5
