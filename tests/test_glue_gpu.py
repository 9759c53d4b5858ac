"""GPU parity of the fused decoder / transition units (bts_b200/glue.py) against the same arithmetic written with plain
torch ops in fp64 on the CPU (the reference's own op sequence, pytorch/bts.py:51-80,154-182 and torchvision's
_Transition): outputs, input gradients, parameter gradients, BatchNorm running statistics."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.cuda().contiguous(memory_format=torch.channels_last)


def _close(a, b, tol):
    a, b = a.detach().cpu().double(), b.detach().double()
    err = (a - b).abs().max() / b.abs().max().clamp_min(1e-12)
    assert err < tol, "rel err %.3g" % float(err)


@pytest.mark.parametrize("pre_relu,up,act,Cin,Cout,k,dil", [
    (False, False, "elu", 36, 32, 3, 1), (True, True, "elu", 40, 24, 3, 1), (False, True, "elu", 64, 32, 3, 1),
    (False, False, "elu", 32, 16, 1, 1), (True, False, None, 64, 48, 1, 1), (False, False, None, 16, 16, 3, 6)])
def test_conv_act_unit(pre_relu, up, act, Cin, Cout, k, dil):
    from bts_b200 import glue
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, Cin, 10, 14, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    pad = dil * (k // 2)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    h = F.relu(xr) if pre_relu else xr
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    yr = F.conv2d(h, wr, None, 1, pad, dil)
    if act == "elu":
        yr = F.elu(yr)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy.double())
    xo, wo = _cl(x).requires_grad_(True), w.cuda().requires_grad_(True)
    yo = glue.conv_act(xo, wo, pad, dil, pre_relu=pre_relu, up=up, act=act)
    yo.backward(_cl(gy))
    _close(yo, yr, 2e-5)
    _close(xo.grad, xr.grad, 1e-4)
    _close(wo.grad, wr.grad, 1e-4)


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("relu", [False, True])
def test_bn_act_unit(mode, relu):
    from bts_b200 import glue
    g = torch.Generator().manual_seed(1)
    C = 24
    bn = nn.BatchNorm2d(C, momentum=0.01, eps=1.1e-5)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ref = copy.deepcopy(bn).double()
    ours = copy.deepcopy(bn).cuda()
    getattr(ref, mode)()
    getattr(ours, mode)()
    x = torch.randn(3, C, 9, 11, generator=g) * 2 + 0.5
    gy = torch.randn(3, C, 9, 11, generator=g)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    if relu:
        yr = F.relu(yr)
    yr.backward(gy.double())
    xo = _cl(x).requires_grad_(True)
    yo = glue.bn_act(xo, ours, relu=relu)
    yo.backward(_cl(gy))
    _close(yo, yr, 1e-5)
    _close(xo.grad, xr.grad, 1e-4)
    _close(ours.weight.grad, ref.weight.grad, 1e-4)
    _close(ours.bias.grad, ref.bias.grad, 1e-4)
    _close(ours.running_mean, ref.running_mean, 1e-5)
    _close(ours.running_var, ref.running_var, 1e-5)
    assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked)


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("k,dil", [(1, 1), (3, 3)])
def test_bn_relu_conv_unit(mode, k, dil):
    from bts_b200 import glue
    g = torch.Generator().manual_seed(2)
    Cin, Cout = 40, 32
    bn = nn.BatchNorm2d(Cin, momentum=0.01, eps=1.1e-5)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(Cin, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(Cin, generator=g) * 0.3)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    ref = copy.deepcopy(bn).double()
    ours = copy.deepcopy(bn).cuda()
    getattr(ref, mode)()
    getattr(ours, mode)()
    x = torch.randn(2, Cin, 12, 10, generator=g)
    pad = dil * (k // 2)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = F.conv2d(F.relu(ref(xr)), wr, None, 1, pad, dil)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy.double())
    xo, wo = _cl(x).requires_grad_(True), w.cuda().requires_grad_(True)
    yo = glue.bn_relu_conv(xo, ours, wo, pad, dil)
    yo.backward(_cl(gy))
    _close(yo, yr, 2e-5)
    _close(xo.grad, xr.grad, 2e-4)
    _close(wo.grad, wr.grad, 2e-4)
    _close(ours.weight.grad, ref.weight.grad, 2e-4)
    _close(ours.bias.grad, ref.bias.grad, 2e-4)
    _close(ours.running_var, ref.running_var, 1e-5)


def test_cat_nhwc_pads_rows_and_routes_gradients():
    from bts_b200 import glue
    g = torch.Generator().manual_seed(3)
    parts = [torch.randn(2, c, 6, 8, generator=g) for c in (32, 96, 1)]        # 129 channels -> slab of 132
    ins = [(_cl(t) if t.shape[1] > 1 else t.cuda()).requires_grad_(True) for t in parts]
    y = glue.cat_nhwc(ins)
    assert y.shape == (2, 129, 6, 8)
    assert y.stride(3) == 132 and y.stride(1) == 1                            # 16-byte aligned pixel rows
    assert torch.equal(y.detach().cpu(), torch.cat(parts, 1))
    gy = torch.randn(2, 129, 6, 8, generator=g)
    y.backward(_cl(gy))
    off = 0
    for t, p in zip(ins, parts):
        assert torch.equal(t.grad.cpu(), gy[:, off:off + p.shape[1]])
        off += p.shape[1]


def test_avgpool2_and_transition():
    from torchvision.models.densenet import _Transition
    from bts_b200 import model as M
    torch.manual_seed(4)
    ref = _Transition(48, 24).double()
    ours = copy.deepcopy(ref).float()
    holder = nn.Sequential()
    holder.add_module("transition1", ours)
    M.adopt_convs(holder)
    assert type(ours).__name__ == "TransitionTC"
    ours.cuda().train()
    ref.train()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 48, 8, 12, generator=g)
    gy = torch.randn(2, 24, 4, 6, generator=g)
    xr = x.double().requires_grad_(True)
    yr = ref(xr)
    yr.backward(gy.double())
    xo = _cl(x).requires_grad_(True)
    yo = ours(xo)
    yo.backward(_cl(gy))
    _close(yo, yr, 2e-5)
    _close(xo.grad, xr.grad, 2e-4)
    for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        _close(p.grad, q.grad, 2e-4)
    _close(ours.norm.running_mean, ref.norm.running_mean, 1e-5)
