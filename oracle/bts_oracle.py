"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's BTS hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product (bts_b200/, bts.py) never does.

Parity status: the reference (cleinc/bts @5e3406b) ships NO golden vectors or tests
 -- "parity unpinned" upstream.  This restatement is therefore
pinned against outputs of the reference itself, generated in the build container by
oracle/make_golden.py (imports the unmodified pytorch/bts.py through oracle/ref_shim.py)
and committed as tests/golden/*.npz; tests/test_oracle.py replays them, and -- when
/root/reference is mounted -- also compares live against the reference modules.

Everything is written functionally over a `state_dict` with the reference's key names
 in torch on the CPU, dtype-generic (fp32 = parity, fp64 = error
budget).  Each function cites the reference lines it restates.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- LPG
def lpg_grid(r, dtype=torch.float32):
    """u (== v) offsets of the r sub-pixels of a patch: (k - (r-1)/2)/r.
    pytorch/bts.py:128-129,141,144; custom_layer/local_planar_guidance.cc:100-101."""
    k = torch.arange(r, dtype=dtype)
    return (k - (r - 1) * 0.5) / r


def lpg_forward(plane, r, layout="nchw"):
    """depth[b,y,x] = n4 / ((n1*u(x) + n2*v(y)) + n3), plane of patch (y//r, x//r).
    pytorch/bts.py:132-146 (NCHW (B,4,h,w)); local_planar_guidance.cc:85-114 (NHWC (B,h,w,4))."""
    if layout == "nhwc":
        plane = plane.permute(0, 3, 1, 2)
    B, four, h, w = plane.shape
    assert four == 4
    g = lpg_grid(r, plane.dtype)
    u = g.repeat(w).view(1, 1, w * r)          # varies along width
    v = g.repeat(h).view(1, h * r, 1)          # varies along height
    e = plane.repeat_interleave(r, 2).repeat_interleave(r, 3)
    n1, n2, n3, n4 = e[:, 0], e[:, 1], e[:, 2], e[:, 3]
    return n4 / (n1 * u + n2 * v + n3)


def lpg_backward(dy, plane, r, layout="nchw", tf_compat=False, acc_dtype=torch.float64):
    """Per-patch sums over the r x r tile:
        g1 = S -dY*n4*u/den^2, g2 = S -dY*n4*v/den^2, g3 = S -dY*n4/den^2, g4 = S dY/den.
    tf_compat drops n4 from g1..g3 exactly like local_planar_guidance.cc:291-293 / .cu:143-145 (Q5).
    Accumulates in acc_dtype (fp64 by default so it can serve as the error-budget reference)."""
    if layout == "nhwc":
        plane_c = plane.permute(0, 3, 1, 2)
    else:
        plane_c = plane
    B, _, h, w = plane_c.shape
    out_dtype = plane.dtype
    p = plane_c.to(acc_dtype)
    dy = dy.to(acc_dtype).view(B, h, r, w, r)
    g = lpg_grid(r, acc_dtype)
    u = g.view(1, 1, 1, 1, r)
    v = g.view(1, 1, r, 1, 1)
    n1, n2, n3, n4 = (p[:, c].view(B, h, 1, w, 1) for c in range(4))
    den = n1 * u + n2 * v + n3
    k = dy / (den * den)
    if not tf_compat:
        k = k * n4
    g1 = (-k * u).sum(dim=(2, 4))
    g2 = (-k * v).sum(dim=(2, 4))
    g3 = (-k).sum(dim=(2, 4))
    g4 = (dy / den).sum(dim=(2, 4))
    out = torch.stack([g1, g2, g3, g4], dim=1).to(out_dtype)
    if layout == "nhwc":
        out = out.permute(0, 2, 3, 1).contiguous()
    return out


# ----------------------------------------------------------------------------- heads
def plane_params(c3, max_depth):
    """Non-final reduction_1x1 tail, pytorch/bts.py:112-120: 3-channel conv output -> (n1,n2,n3,n4)."""
    theta = torch.sigmoid(c3[:, 0]) * math.pi / 3
    phi = torch.sigmoid(c3[:, 1]) * math.pi * 2
    dist = torch.sigmoid(c3[:, 2]) * max_depth
    n1 = torch.sin(theta) * torch.cos(phi)
    n2 = torch.sin(theta) * torch.sin(phi)
    n3 = torch.cos(theta)
    return torch.stack([n1, n2, n3, dist], dim=1)


def plane_eq_from_head(net4):
    """pytorch/bts.py:223-226: L2-normalise the 3 normal channels (F.normalize eps=1e-12), keep dist."""
    n = net4[:, :3]
    nrm = n.pow(2).sum(1, keepdim=True).sqrt().clamp_min(1e-12)
    return torch.cat([n / nrm, net4[:, 3:4]], dim=1)


def reduction_chain(x, sd, prefix, max_depth, is_final):
    """reduction_1x1, pytorch/bts.py:83-122.  1x1 conv + ELU chain named inter_{in}_{out}, then
    plane_params (3ch) + trig tail, or final (1ch) + sigmoid."""
    tag = prefix + "reduc.inter_"
    # module order in the reference = construction order = (cin, cout) descending (bts.py:91-108)
    names = sorted((k for k in sd if k.startswith(tag)),
                   key=lambda k: tuple(-int(t) for t in k[len(tag):].split(".")[0].split("_")))
    for k in names:
        x = F.elu(F.conv2d(x, sd[k]))
    if is_final:
        return torch.sigmoid(F.conv2d(x, sd[prefix + "reduc.final.0.weight"]))
    return plane_params(F.conv2d(x, sd[prefix + "reduc.plane_params.weight"]), max_depth)


# ----------------------------------------------------------------------------- loss
def silog(est, gt, mask, lam):
    """silog_loss.forward, pytorch/bts.py:46-48."""
    d = torch.log(est[mask]) - torch.log(gt[mask])
    return torch.sqrt((d ** 2).mean() - lam * (d.mean() ** 2)) * 10.0


def silog_grad(est, gt, mask, lam, acc_dtype=torch.float64):
    """dL/d est_i = [i in M] * (10/sqrt(S)) * (d_i - lam*m1) / (N*est_i)   (SURVEY.md Appendix B)."""
    e = est.to(acc_dtype)
    g = gt.to(acc_dtype)
    m = mask
    d = torch.where(m, torch.log(e) - torch.log(torch.where(m, g, torch.ones_like(g))), torch.zeros_like(e))
    n = m.sum().to(acc_dtype)
    m1 = d.sum() / n
    m2 = (d * d).sum() / n
    s = m2 - lam * m1 * m1
    grad = torch.where(m, (10.0 / torch.sqrt(s)) * (d - lam * m1) / (n * e), torch.zeros_like(e))
    return grad.to(est.dtype)


# ----------------------------------------------------------------------------- decoder
def _bn(x, sd, name, training, eps, momentum, stats_out):
    """nn.BatchNorm2d semantics.  Running-stat updates are written to stats_out."""
    w, b = sd[name + ".weight"], sd[name + ".bias"]
    rm, rv = sd[name + ".running_mean"].clone(), sd[name + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, w, b, training, momentum, eps)
    if training and stats_out is not None:
        stats_out[name + ".running_mean"] = rm
        stats_out[name + ".running_var"] = rv
    return y


def _upconv(x, w):
    """upconv, pytorch/bts.py:76-80: nearest x2 -> 3x3 conv pad 1 -> ELU."""
    return F.elu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1))


def _atrous(x, sd, p, dil, first_bn, training, stats_out):
    """atrous_conv, pytorch/bts.py:51-66: [BN eps 1.1e-5] ReLU 1x1 BN(eps 1e-5) ReLU 3x3 dilated."""
    if first_bn:
        x = _bn(x, sd, p + "atrous_conv.first_bn", training, 1.1e-5, 0.01, stats_out)
    x = F.conv2d(F.relu(x), sd[p + "atrous_conv.aconv_sequence.1.weight"])
    x = _bn(x, sd, p + "atrous_conv.aconv_sequence.2", training, 1e-5, 0.01, stats_out)
    return F.conv2d(F.relu(x), sd[p + "atrous_conv.aconv_sequence.4.weight"], padding=dil, dilation=dil)


def decoder_forward(sd, feats, focal, max_depth, dataset, training=False, prefix="", stats_out=None):
    """bts.forward, pytorch/bts.py:196-266, over a state_dict whose keys start with `prefix`."""
    sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    bn = lambda x, n: _bn(x, sd, n, training, 1.1e-5, 0.01, stats_out)
    skip0, skip1, skip2, skip3 = feats[0], feats[1], feats[2], feats[3]
    dense = F.relu(feats[4])
    up5 = bn(_upconv(dense, sd["upconv5.conv.weight"]), "bn5")
    i5 = F.elu(F.conv2d(torch.cat([up5, skip3], 1), sd["conv5.0.weight"], padding=1))
    up4 = bn(_upconv(i5, sd["upconv4.conv.weight"]), "bn4")
    cat4 = torch.cat([up4, skip2], 1)
    i4 = bn(F.elu(F.conv2d(cat4, sd["conv4.0.weight"], padding=1)), "bn4_2")
    d3 = _atrous(i4, sd, "daspp_3.", 3, False, training, stats_out)
    c = torch.cat([cat4, d3], 1)
    d6 = _atrous(c, sd, "daspp_6.", 6, True, training, stats_out)
    c = torch.cat([c, d6], 1)
    d12 = _atrous(c, sd, "daspp_12.", 12, True, training, stats_out)
    c = torch.cat([c, d12], 1)
    d18 = _atrous(c, sd, "daspp_18.", 18, True, training, stats_out)
    c = torch.cat([c, d18], 1)
    d24 = _atrous(c, sd, "daspp_24.", 24, True, training, stats_out)
    feat = F.elu(F.conv2d(torch.cat([i4, d3, d6, d12, d18, d24], 1), sd["daspp_conv.0.weight"], padding=1))

    def lpg_scale(x, pfx, r):
        eq = plane_eq_from_head(reduction_chain(x, sd, pfx, max_depth, False))
        return lpg_forward(eq, r).unsqueeze(1) / max_depth

    d8 = lpg_scale(feat, "reduc8x8.", 8)
    d8_ds = d8[:, :, ::4, ::4]                                       # nearest x0.25 (Q14)
    up3 = bn(_upconv(feat, sd["upconv3.conv.weight"]), "bn3")
    i3 = F.elu(F.conv2d(torch.cat([up3, skip1, d8_ds], 1), sd["conv3.0.weight"], padding=1))
    d4 = lpg_scale(i3, "reduc4x4.", 4)
    d4_ds = d4[:, :, ::2, ::2]                                       # nearest x0.5
    up2 = bn(_upconv(i3, sd["upconv2.conv.weight"]), "bn2")
    i2 = F.elu(F.conv2d(torch.cat([up2, skip0, d4_ds], 1), sd["conv2.0.weight"], padding=1))
    d2 = lpg_scale(i2, "reduc2x2.", 2)
    up1 = _upconv(i2, sd["upconv1.conv.weight"])
    r1 = reduction_chain(up1, sd, "reduc1x1.", max_depth, True)
    i1 = F.elu(F.conv2d(torch.cat([up1, r1, d2, d4, d8], 1), sd["conv1.0.weight"], padding=1))
    final = max_depth * torch.sigmoid(F.conv2d(i1, sd["get_depth.0.weight"], padding=1))
    if dataset == "kitti":
        final = final * focal.view(-1, 1, 1, 1).to(final.dtype) / 715.0873
    return d8, d4, d2, r1, final


# ----------------------------------------------------------------------------- encoder
ENCODERS = {
    # name: (torchvision ctor, takes .features, skip names, channels)     pytorch/bts.py:273-300
    "densenet121_bts": ("densenet121", True, ["relu0", "pool0", "transition1", "transition2", "norm5"], [64, 64, 128, 256, 1024]),
    "densenet161_bts": ("densenet161", True, ["relu0", "pool0", "transition1", "transition2", "norm5"], [96, 96, 192, 384, 2208]),
    "resnet50_bts": ("resnet50", False, ["relu", "layer1", "layer2", "layer3", "layer4"], [64, 256, 512, 1024, 2048]),
    "resnet101_bts": ("resnet101", False, ["relu", "layer1", "layer2", "layer3", "layer4"], [64, 256, 512, 1024, 2048]),
    "resnext50_bts": ("resnext50_32x4d", False, ["relu", "layer1", "layer2", "layer3", "layer4"], [64, 256, 512, 1024, 2048]),
    "resnext101_bts": ("resnext101_32x8d", False, ["relu", "layer1", "layer2", "layer3", "layer4"], [64, 256, 512, 1024, 2048]),
    # pytorch/bts.py:297-301: no tap names, the skips are the outputs of the 2nd/4th/7th/11th/19th module of `.features`
    "mobilenetv2_bts": ("mobilenet_v2", True, [], [16, 24, 32, 64, 1280]),
}
MOBILENET_TAPS = (2, 4, 7, 11, 19)


def build_encoder(name):
    """The arithmetic of the encoder is torchvision's (un-vendored third-party dependency of the
    reference, no version pin) -- the oracle uses the torchvision installed in the image,
    random init (weights=None)."""
    import torchvision.models as tvm
    ctor, feats, names, ch = ENCODERS[name]
    m = getattr(tvm, ctor)(weights=None)
    return (m.features if feats else m), names, ch


def encoder_forward(base_model, names, x):
    """encoder.forward, pytorch/bts.py:305-320 (names == [] selects the MobileNetV2 index taps, :312-314)."""
    skips = []
    i = 1
    for k, v in base_model._modules.items():
        if "fc" in k or "avgpool" in k:
            continue
        x = v(x)
        if not names:
            if i in MOBILENET_TAPS:
                skips.append(x)
        elif any(n in k for n in names):
            skips.append(x)
        i += 1
    return skips


class OracleModel(torch.nn.Module):
    """Restatement of BtsModel (pytorch/bts.py:323-331) with the same state_dict keys, built from the
    functional pieces above; used as the CPU checker for the whole path and as bench.py's CPU arm."""

    def __init__(self, encoder="densenet161_bts", max_depth=80.0, dataset="kitti", bts_size=512):
        super().__init__()
        self.max_depth, self.dataset = max_depth, dataset
        enc = torch.nn.Module()
        enc.base_model, self._names, self.feat = build_encoder(encoder)
        self.encoder = enc
        self.decoder = make_decoder_params(self.feat, bts_size)

    def forward(self, x, focal):
        feats = encoder_forward(self.encoder.base_model, self._names, x)
        sd = dict(self.decoder.named_parameters())
        sd.update(dict(self.decoder.named_buffers()))
        stats = {} if self.training else None
        out = decoder_forward(sd, feats, focal, self.max_depth, self.dataset, self.training, "", stats)
        if stats:
            with torch.no_grad():
                bufs = dict(self.decoder.named_buffers())
                for k, v in stats.items():
                    bufs[k].copy_(v)
        return out


class _ParamTree(torch.nn.Module):
    pass


def decoder_shapes(feat, nf=512):
    """(key, shape) of every decoder parameter -- SURVEY.md Appendix C / pytorch/bts.py:153-194."""
    out = []
    conv = lambda k, co, ci, ks: out.append((k, (co, ci, ks, ks)))

    def bn(k, c):
        out.append((k, ("bn", c)))

    conv("upconv5.conv.weight", nf, feat[4], 3); bn("bn5", nf)
    conv("conv5.0.weight", nf, nf + feat[3], 3)
    conv("upconv4.conv.weight", nf // 2, nf, 3); bn("bn4", nf // 2)
    conv("conv4.0.weight", nf // 2, nf // 2 + feat[2], 3); bn("bn4_2", nf // 2)
    cins = [nf // 2, nf // 2 + nf // 4 + feat[2], nf + feat[2], nf + nf // 4 + feat[2], nf + nf // 2 + feat[2]]
    for d, cin in zip((3, 6, 12, 18, 24), cins):
        p = "daspp_%d.atrous_conv." % d
        if d != 3:
            bn(p + "first_bn", cin)
        conv(p + "aconv_sequence.1.weight", nf // 2, cin, 1)
        bn(p + "aconv_sequence.2", nf // 2)
        conv(p + "aconv_sequence.4.weight", nf // 4, nf // 2, 3)
    conv("daspp_conv.0.weight", nf // 4, nf + nf // 2 + nf // 4, 3)

    def reduc(pfx, cin, cout, final):
        while cout >= 4:
            if cout < 8:
                if final:
                    conv(pfx + "reduc.final.0.weight", 1, cin, 1)
                else:
                    conv(pfx + "reduc.plane_params.weight", 3, cin, 1)
                break
            conv(pfx + "reduc.inter_%d_%d.0.weight" % (cin, cout), cout, cin, 1)
            cin, cout = cout, cout // 2

    reduc("reduc8x8.", nf // 4, nf // 4, False)
    conv("upconv3.conv.weight", nf // 4, nf // 4, 3); bn("bn3", nf // 4)
    conv("conv3.0.weight", nf // 4, nf // 4 + feat[1] + 1, 3)
    reduc("reduc4x4.", nf // 4, nf // 8, False)
    conv("upconv2.conv.weight", nf // 8, nf // 4, 3); bn("bn2", nf // 8)
    conv("conv2.0.weight", nf // 8, nf // 8 + feat[0] + 1, 3)
    reduc("reduc2x2.", nf // 8, nf // 16, False)
    conv("upconv1.conv.weight", nf // 16, nf // 8, 3)
    reduc("reduc1x1.", nf // 16, nf // 32, True)
    conv("conv1.0.weight", nf // 16, nf // 16 + 4, 3)
    conv("get_depth.0.weight", 1, nf // 16, 3)
    return out


def make_decoder_params(feat, nf=512):
    """A bare parameter container whose state_dict has the reference decoder's keys; conv weights
    xavier-uniform (bts_main.py:338), BN default init."""
    root = _ParamTree()

    def put(path, tensor, buffer=False):
        mod = root
        parts = path.split(".")
        for p in parts[:-1]:
            if not hasattr(mod, p):
                setattr(mod, p, _ParamTree())
            mod = getattr(mod, p)
        if buffer:
            mod.register_buffer(parts[-1], tensor)
        else:
            setattr(mod, parts[-1], torch.nn.Parameter(tensor))

    for k, shp in decoder_shapes(feat, nf):
        if shp[0] == "bn":
            c = shp[1]
            put(k + ".weight", torch.ones(c)); put(k + ".bias", torch.zeros(c))
            put(k + ".running_mean", torch.zeros(c), True); put(k + ".running_var", torch.ones(c), True)
            put(k + ".num_batches_tracked", torch.zeros((), dtype=torch.long), True)
        else:
            w = torch.empty(*shp)
            torch.nn.init.xavier_uniform_(w)
            put(k, w)
    return root
