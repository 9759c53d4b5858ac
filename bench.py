#!/usr/bin/env python
"""bench.py -- training images/sec of the BTS hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun launches N>1, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference algorithm on host cores

Workload (config K16, BASELINE.json configs[1]): DenseNet-161 encoder, 352x704 synthetic RGB->depth, batch 16
PER GPU (weak scaling), train mode (batch-stat BN), silog loss (lambda .85), backward, AdamW step with the two
parameter groups and poly LR of bts_main.py:371-373,456-460.  A "step" = zero_grad + forward + loss + backward +
optimizer step over one batch.  `value` is timed with inputs resident in HBM; `e2e` repeats the measurement with
the batch starting in pinned host memory (H2D inside the timed region) and the loss read back (D2H) every step.
`roofline` is the LPG plane-to-depth kernel pair (the kernel BASELINE.json's metric names), timed live with CUDA
events on the launching stream; `roofline_conv` is one layer of the step's dominant kernel (the tcgen05 conv engine)
timed the same way against the 3xTF32 tensor roof; `roofline_step` relates the whole step to the conv-FLOP roof.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, B_PER_GPU = 352, 704, 16
FOCAL = 721.5377
MAX_DEPTH = 80.0
GFLOP_PER_IMG_TRAIN = 588.0      # nominal conv FLOPs, fwd+dgrad+wgrad, DN-161 @352x704


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def synth_batch(B, seed, dev="cpu", pin=False):
    """SURVEY 8d synthetic inputs: image randn, focal 721.5377 (float64 like the DataLoader collate), KITTI-like
    sparse depth (20% valid, else 0)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(B, 3, H, W, generator=g)
    gt = torch.rand(B, 1, H, W, generator=g) * MAX_DEPTH
    gt = torch.where(torch.rand(B, 1, H, W, generator=g) < 0.2, gt, torch.zeros_like(gt))
    focal = torch.full((B,), FOCAL, dtype=torch.float64)
    if pin:
        return img.pin_memory(), focal.pin_memory(), gt.pin_memory()
    return img.to(dev), focal.to(dev), gt.to(dev)


def make_optimizer(model, torch):
    """bts_main.py:371-373 (lr 1e-4, eps 1e-3, wd 1e-2 encoder / 0 decoder: arguments_train_eigen.txt)."""
    m = model.module if hasattr(model, "module") else model
    return torch.optim.AdamW([{"params": m.encoder.parameters(), "weight_decay": 1e-2},
                              {"params": m.decoder.parameters(), "weight_decay": 0}], lr=1e-4, eps=1e-3)


def freeze_like_set_misc(model):
    """bts_main.py:217-247 default: 'Fixing first conv layer' -> conv0 + every encoder BN affine param frozen."""
    for name, p in model.encoder.named_parameters():
        if any(x in name for x in ("conv0", "norm")):
            p.requires_grad = False


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.p = [], None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def usable_cores():
    """host cores this process may actually use: affinity mask, capped by the cgroup CPU quota"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def cpu_reference_steps(steps, warmup, B=2, budget_s=150.0):
    """The reference algorithm (oracle port of pytorch/bts.py; the Python reference itself does not travel to the
    GPU box) on all host cores: DN-161, 352x704, one train step per sample batch of B images."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bts_oracle as O
    cores = usable_cores()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = O.OracleModel("densenet161_bts", MAX_DEPTH, "kitti", 512)
    m.train()
    opt = torch.optim.AdamW([{"params": m.encoder.parameters(), "weight_decay": 1e-2},
                             {"params": m.decoder.parameters(), "weight_decay": 0}], lr=1e-4, eps=1e-3)
    img, focal, gt = synth_batch(B, 1)
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        if times and time.perf_counter() - t_start > budget_s:
            break                                  # bounded sample: keep the bench within minutes on slow hosts
        t0 = time.perf_counter()
        opt.zero_grad()
        out = m(img, focal)
        loss = O.silog(out[4], gt, gt > 1.0, 0.85)
        loss.backward()
        opt.step()
        float(loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": B * len(times) / total, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "oracle/bts_oracle.py OracleModel (torch CPU fp32, %d threads), DN-161 352x704, B=%d train step "
                      "(fwd+silog+bwd+AdamW), %d timed steps after %d warm-up" % (cores, B, len(times), warmup),
            "ms_per_step": 1e3 * total / len(times)}


def lpg_roofline(torch, dev, pk):
    """LPG-u microbench (BASELINE.json configs[4]), HBM regime: r=8/4/2 at 1024^2, batch 128 so the output is
    512 MB (>> 126 MB L2; consecutive launches cannot hit in L2).  The kernels are launched through the C ABI
    (bts_lpg_fwd / bts_lpg_bwd) back to back -- 20 launches between two CUDA events on the launching stream, so
    the figure is the kernels' average duration, not Python launch latency.
    Algorithmic bytes: fwd 4*px*(1+4/r^2), bwd 4*px*(1+8/r^2)  (BASELINE.md section 3)."""
    import ctypes
    import math
    from bts_b200 import _lib
    L = _lib.lib()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    st = torch.cuda.current_stream()
    sp = ctypes.c_void_p(st.cuda_stream)
    out = {}
    side, Bn, reps = 1024, 128, 20
    depth = torch.empty(Bn, side, side, device=dev)
    dy = torch.randn(Bn, side, side, device=dev)
    for r in (8, 4, 2):
        h = side // r
        g = torch.Generator(device=dev).manual_seed(r)
        z = torch.randn(Bn, 3, h, h, device=dev, generator=g)
        th = torch.sigmoid(z[:, 0]) * math.pi / 3
        ph = torch.sigmoid(z[:, 1]) * math.pi * 2
        plane = torch.stack([torch.sin(th) * torch.cos(ph), torch.sin(th) * torch.sin(ph), torch.cos(th),
                             torch.sigmoid(z[:, 2]) * MAX_DEPTH], 1).contiguous()
        dplane = torch.empty_like(plane)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(2):
            if it == 1:
                ev[0].record(st)
            for _ in range(reps):
                assert L.bts_lpg_fwd(vp(plane), vp(depth), Bn, h, h, r, 0, sp) == 0
            if it == 1:
                ev[1].record(st)
            for _ in range(reps):
                assert L.bts_lpg_bwd(vp(dy), vp(plane), vp(dplane), Bn, h, h, r, 0, 0, sp) == 0
            if it == 1:
                ev[2].record(st)
        torch.cuda.synchronize()
        _lib.count(4 * reps)
        mf, mb = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        px = Bn * side * side
        bf, bb = 4.0 * px * (1 + 4.0 / r ** 2), 4.0 * px * (1 + 8.0 / r ** 2)
        out["r%d" % r] = {"fwd_gbs": bf / mf / 1e6, "bwd_gbs": bb / mb / 1e6, "fwd_bwd_gbs": (bf + bb) / (mf + mb) / 1e6,
                          "fwd_ms": mf, "bwd_ms": mb, "shape": [Bn, side, side]}
        del plane, dplane, z
    del depth, dy
    torch.cuda.empty_cache()
    a = out["r8"]["fwd_bwd_gbs"]
    return {"bound": "hbm", "kernel": "lpg_fwd_vec<8> + lpg_bwd_vec<8> (fwd+bwd, r=8, 128x1024x1024, output 512 MB >> L2, "
                                      "20 back-to-back launches each)",
            "achieved": a, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": a / pk["hbm_gbs"], "peak_source": pk["source"],
            # dram__bytes_read.sum + dram__bytes_write.sum of one (fwd, bwd) launch pair, `ncu --set full` capture
            # profiles/r01_lpg_r8_v2.ncu-rep of this same microbench: fwd 33.6 + 479.3 MB, bwd 570.4 + 32.3 MB
            # (algorithmic 570.4 + 604.0 MB: no re-reads; part of the forward's output is still dirty in L2 at kernel end)
            "traffic": 1115.6e6, "traffic_unit": "bytes per fwd+bwd launch pair (ncu, profiles/r01_lpg_r8_v2.ncu-rep)",
            "algorithmic_bytes": (4.0 * 128 * 1024 * 1024) * ((1 + 4.0 / 64) + (1 + 8.0 / 64)), "sweep": out}


def conv_roofline(torch, dev, pk):
    """The dominant kernel of the step by time is the conv engine (conv_tc_kernel: every forward and dgrad).  One of its
    layers, timed live: daspp_conv 896->128 3x3 at 44x88, batch 16 (input 222 MB > 126 MB L2, so consecutive launches
    cannot be served from L2), 20 back-to-back launches between two CUDA events on the launching stream.
    achieved = nominal dense FLOPs (2*B*H*W*Cout*Cin*9) / average launch time; peak = the parity-mode (3xTF32) tensor
    roof derived from the measured bf16 throughput: bf16 / 2 (tf32) / 3 (three products)."""
    from bts_b200 import conv
    B, Cin, H, W, Cout, k = 16, 896, 44, 88, 128, 3
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(B, Cin, H, W, device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    w = torch.randn(Cout, Cin, k, k, device=dev, generator=g) / (Cin * k * k) ** 0.5
    packed = conv.pack_weights(w)
    run = lambda: conv.conv2d_tc(x, w, 1, 1, 1, packed=packed)
    reps = 20
    run()
    run()
    torch.cuda.synchronize()
    st = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        run()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * H * W * Cout * Cin * k * k
    a = flops / ms / 1e9
    peak = pk["bf16_tflops"] / 6.0
    del x, w, packed
    torch.cuda.empty_cache()
    return {"bound": "tensor", "kernel": "conv_tc_kernel, daspp_conv 896->128 3x3 @16x44x88 (3xTF32, N-stacked hi/lo MMAs), "
                                         "%d back-to-back launches" % reps,
            "achieved": a, "peak": peak, "unit": "TFLOP/s", "frac": a / peak, "ms_per_launch": ms,
            "peak_source": "bf16 sustained / 2 / 3; " + pk["source"], "traffic": None,
            "algorithmic_flops": flops}


class StdoutToStderr:
    """While active, file descriptor 1 points at stderr: library chatter (e.g. the "NCCL version ..." banner NCCL prints on
    stdout at communicator creation) cannot precede the ONE JSON line this script owes its caller on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_ours(args):
    with StdoutToStderr():
        res, rank, dist = _run_ours(args)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if dist is not None:
        with StdoutToStderr():
            dist.destroy_process_group()


def _run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import bts
    from bts_b200 import _lib
    _lib.lib()                                   # fail loudly if the CUDA library is missing
    torch.backends.cudnn.benchmark = True        # bts_main.py:402
    torch.manual_seed(0)
    p = types.SimpleNamespace(encoder="densenet161_bts", max_depth=MAX_DEPTH, dataset="kitti", bts_size=512,
                              pretrained=False)
    model = bts.BtsModel(p)
    model.train()
    model.decoder.apply(bts.weights_init_xavier)
    freeze_like_set_misc(model)
    model.to(dev)
    if world > 1:
        # bts_main.py:352 passes find_unused_parameters=True because the ResNet/ResNeXt encoders keep a never-used `fc`
        # (SURVEY Q11); DenseNet-161 has no unused parameter, so the extra autograd traversal is switched off here.
        unused = any(k.startswith("encoder.base_model.fc") for k, _ in model.named_parameters())
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=unused,
                                                          gradient_as_bucket_view=True)
    opt = make_optimizer(model, torch)
    crit = bts.silog_loss(0.85)
    B = B_PER_GPU
    img, focal, gt = synth_batch(B, 1 + rank, dev)
    himg, hfocal, hgt = synth_batch(B, 1 + rank, pin=True)
    total_steps = 10000

    def step(i, x, f, g):
        opt.zero_grad()
        out = model(x, f)
        loss = crit(out[4], g, g > 1.0)
        loss.backward()
        lr = (1e-4 - 1e-5) * (1 - i / total_steps) ** 0.9 + 1e-5
        for grp in opt.param_groups:
            grp["lr"] = lr
        opt.step()
        return loss

    def timed(fn, K, Wm):
        for i in range(Wm):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if rank == 0 else None
        l0 = _lib.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = torch.cuda.current_stream()
        e0.record(st)
        for i in range(K):
            fn(Wm + i)
        e1.record(st)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        clocks = sampler.stop() if sampler else None
        return float(ms), _lib.launches - l0, clocks

    K, Wm = args.steps, max(args.warmup, 3)
    ms, launches, clocks = timed(lambda i: step(i, img, focal, gt), K, Wm)

    def e2e_step(i):
        x = himg.to(dev, non_blocking=True)
        f = hfocal.to(dev, non_blocking=True)
        g = hgt.to(dev, non_blocking=True)
        return float(step(i, x, f, g).detach())  # D2H read of the loss every step (bts_main.py:463)

    ms_e, _, _ = timed(e2e_step, K, 1)
    h2d = himg.numel() * 4 + hfocal.numel() * 8 + hgt.numel() * 4

    pk = peaks()
    res = {
        "metric": "training images/sec (352x704, DenseNet-161)", "value": world * B * K / (ms / 1e3), "unit": "images/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "K16: DenseNet-161 encoder, 352x704 KITTI-shape synthetic, batch 16/GPU, train step "
                               "(fwd + silog + bwd + AdamW), random-init weights",
                   "global_batch": world * B, "parallelism": "dp%d" % world,
                   "l2": "no explicit flush: per-step working set (~31 GB saved activations) >> 126 MB L2",
                   "conv_path": os.environ.get("BTS_B200_CONV", "auto")},
        "e2e": {"value": world * B * K / (ms_e / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e / K},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        res["roofline_step"] = {"bound": "tensor", "achieved": res["value"] / world * GFLOP_PER_IMG_TRAIN / 1e3,
                                "unit": "TFLOP/s per GPU (nominal conv FLOPs 588 GFLOP/img)",
                                "peak": pk["bf16_tflops"] / 2.0 / 3.0,
                                "peak_note": "parity-grade 3xTF32 = (bf16 sustained / 2) / 3; " + pk["source"],
                                "frac": res["value"] / world * GFLOP_PER_IMG_TRAIN / 1e3 / (pk["bf16_tflops"] / 6.0)}
        if world == 1:
            if not args.no_lpg:
                del model, opt
                torch.cuda.empty_cache()
                res["roofline"] = lpg_roofline(torch, dev, pk)
                try:
                    res["roofline_conv"] = conv_roofline(torch, dev, pk)
                except Exception as e:             # an evidence leg must never cost the bench line
                    res["roofline_conv"] = {"error": "%s: %s" % (type(e).__name__, e)}
            if not args.no_cpu:
                res["cpu_baseline"] = cpu_reference_steps(3, 1)
    return res, rank, (dist if world > 1 else None)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with StdoutToStderr():
        r = cpu_reference_steps(args.steps, args.warmup)
    res = {"impl": "reference", "metric": "training images/sec (352x704, DenseNet-161)", "value": r["value"],
           "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "K16: DenseNet-161 encoder, 352x704 KITTI-shape synthetic, train step "
                                  "(fwd + silog + bwd + AdamW); CPU sample batch = 2 images/step"},
           "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
           "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-lpg", action="store_true", help="skip the LPG roofline microbench")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
