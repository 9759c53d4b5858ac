"""GPU parity of the fused optimizer step: bts_b200.optim.FusedAdamW against torch.optim.AdamW with the
reference's hyper-parameters (bts_main.py:371-373: two groups, wd 1e-2 / 0, lr 1e-4, eps 1e-3; poly LR :456-460), and the
one-launch re-pack of every cached conv operator against per-layer packing."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed, dev):
    g = torch.Generator().manual_seed(seed)
    shapes = [(48, 192, 3, 3), (192, 240, 1, 1), (7,), (1, 32, 3, 3), (513,), (128, 225, 3, 3), (3, 8, 1, 1)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


def test_fused_adamw_matches_torch_adamw_to_an_ulp():
    from bts_b200.optim import FusedAdamW
    dev = torch.device("cuda")
    pa, pb = _params(0, dev), _params(0, dev)
    mk = lambda cls, ps, **kw: cls([{"params": ps[:4], "weight_decay": 1e-2}, {"params": ps[4:], "weight_decay": 0}],
                                   lr=1e-4, eps=1e-3, **kw)
    ref = mk(torch.optim.AdamW, pa)
    opt = mk(FusedAdamW, pb, repack=False)
    g = torch.Generator().manual_seed(1)
    total = 50
    for step in range(6):
        lr = (1e-4 - 1e-5) * (1 - step / total) ** 0.9 + 1e-5            # the reference's poly schedule
        for o in (ref, opt):
            for grp in o.param_groups:
                grp["lr"] = lr
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).to(dev) * (10.0 ** (step - 3))    # wide dynamic range of gradients
            if step == 2 and a.dim() == 1:
                a.grad = b.grad = None                                  # parameters without a gradient are skipped
                continue
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        opt.step()
        for a, b in zip(pa, pb):
            # identical operation order; the only freedom is FMA contraction inside torch's own kernels: <= 2 ulp
            assert torch.allclose(a, b, rtol=3e-7, atol=1e-10), float((a - b).abs().max())
    sa, sb = ref.state_dict(), opt.state_dict()
    assert sa["param_groups"][0]["lr"] == sb["param_groups"][0]["lr"]
    for k in sa["state"]:
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"])
        assert torch.allclose(sa["state"][k]["exp_avg"], sb["state"][k]["exp_avg"], rtol=3e-7, atol=1e-12)
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=3e-7, atol=1e-20)
    # checkpoints interchange: a torch.optim.AdamW state loads into the fused optimizer and vice versa
    opt.load_state_dict(sa)
    ref.load_state_dict(sb)


def test_fused_adamw_bumps_versions_and_repacks_every_cached_operator_in_one_launch():
    from bts_b200 import _lib, conv
    from bts_b200.optim import FusedAdamW
    dev = torch.device("cuda")
    conv.invalidate_packed()
    g = torch.Generator().manual_seed(3)
    ws = [torch.nn.Parameter((torch.randn(*s, generator=g) / 10).to(dev))
          for s in [(48, 192, 3, 3), (192, 240, 1, 1), (32, 36, 3, 3), (512, 64, 1, 1), (256, 8, 3, 3)]]
    packs = [conv.pack_weights(w, False) for w in ws[:4]] + [conv.pack_weights(ws[0], True)]
    packs.append(conv.pack_weights(ws[4], False, groups=32))         # grouped (ResNeXt) operator
    packs.append(conv.pack_weights(ws[4], True, groups=32))
    opt = FusedAdamW(ws, lr=1e-2, eps=1e-3)
    for w in ws:
        w.grad = torch.randn(w.shape, generator=g).to(dev)
    v0 = [w._version for w in ws]
    before = _lib.launches
    opt.step()
    assert _lib.launches - before == 2                                # one AdamW launch + one re-pack launch
    assert all(w._version > v for w, v in zip(ws, v0))
    cached = [conv.pack_weights(w, False) for w in ws[:4]] + [conv.pack_weights(ws[0], True),
                                                               conv.pack_weights(ws[4], False, groups=32),
                                                               conv.pack_weights(ws[4], True, groups=32)]
    assert _lib.launches - before == 2                                # all cache hits: nothing re-packed lazily
    assert all(a.data_ptr() == b.data_ptr() for a, b in zip(cached, packs))
    conv.invalidate_packed()
    fresh = [conv.pack_weights(w, False) for w in ws[:4]] + [conv.pack_weights(ws[0], True),
                                                              conv.pack_weights(ws[4], False, groups=32),
                                                              conv.pack_weights(ws[4], True, groups=32)]
    for a, b in zip(cached, fresh):
        assert torch.equal(a, b)
