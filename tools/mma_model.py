"""MMA-issue model of the conv engine vs the measured per-layer times of one K16 step.

Round-1 finding: a tcgen05.mma with M=128, K=8 (tf32) occupies the tensor pipe's operand fetch for
~CYC cycles whatever N is, so a layer's floor on this engine is  (#MMA instructions on the busiest SM) x CYC / clock.
This script counts the instructions each launch issues (same tiling rules as csrc/conv_tc.cu, wgrad_tc.cu, wgrad2_tc.cu),
and prints measured / modelled time per layer shape from a `tools/step_trace.py` log -- layers with a ratio near 1 sit
on the MMA-issue bound, layers far above it are limited by something else (producers, epilogue, split-K tail, launch).

usage: python tools/mma_model.py [profiles/r01_step_trace_s2_final.txt] [cycles_per_mma=120]
"""
import math
import re
import sys

SMS, CLOCK_GHZ = 148, 1.965


def ceil16(n):
    return (n + 15) // 16 * 16


def conv_n_tile(cout, cap=256):
    n = ceil16(cout)
    if n > cap:
        tiles = -(-n // cap)
        n = ceil16(-(-cout // tiles))
    return n


def mma_fwd(B, H, W, cin, cout, k, up):
    """conv_tc_kernel (forward and dgrad): persistent, tiles round-robin over the SMs"""
    Ho, Wo = (2 * H, 2 * W) if up else (H, W)
    m_tiles = -(-(B * Ho * Wo) // 128)
    nt = conv_n_tile(cout)
    n_tiles = -(-cout // nt)
    kb = -(-(k * k * -(-cin // 4)) // 8)
    per_kb = 4 * (2 if nt <= 128 else 3)
    tiles_per_sm = -(-(m_tiles * n_tiles) // SMS)
    return tiles_per_sm * kb * per_kb


def wgrad2_cg(cout, taps):
    cap = min((512 // taps) // 16 * 16, 48)
    groups = -(-cout // cap)
    return min(ceil16(-(-cout // groups)), cap)


def mma_wgrad(B, H, W, cin, cout, k, up):
    Ho, Wo = (2 * H, 2 * W) if up else (H, W)
    px = B * Ho * Wo
    taps = k * k
    if 1 < taps <= 9 and cout <= 64 and px >= 60000:                       # wgrad2_tc: all taps per CTA, 16-pixel k-blocks
        cg = wgrad2_cg(cout, taps)
        ctas = -(-cin // 128) * -(-cout // cg)
        kbs = -(-px // 16)
        pieces = -(-(taps * cg) // 256)
        total = ctas * kbs * 2 * 3 * pieces                                # 2 k-groups x 3 products x pieces
        return -(-total // SMS)                                            # split-K spreads the k-blocks over the SMs
    nt = conv_n_tile(cout)                                                 # wgrad_tc: 32-pixel k-blocks, taps in the grid
    ctas = -(-cin // 128) * -(-cout // nt) * taps
    kbs = -(-px // 32)
    total = ctas * kbs * 4 * (2 if nt <= 128 else 3)
    return -(-total // SMS)


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_step_trace_s2_final.txt"
    cyc = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
    pat = re.compile(r"\s*([\d.]+) ms\s+(\d+) x\s+([\d.]+)\s+(\w+)\s+(\d+)x(\d+)x(\d+) (\d+)->(\d+) k(\d+) d(\d+) s(\d+)( up)?")
    rows, tot_meas, tot_model = [], {}, {}
    for line in open(path):
        m = pat.match(line)
        if not m:
            continue
        total, n, each, kind = float(m.group(1)), int(m.group(2)), float(m.group(3)), m.group(4)
        B, H, W, cin, cout, k = (int(m.group(i)) for i in range(5, 11))
        stride, up = int(m.group(12)), bool(m.group(13))
        if kind in ("fwd", "dgrad"):
            if stride != 1:
                H, W = H // stride, W // stride                            # the stem: output pixels
            instr = mma_fwd(B, H, W, cin, cout, k, up)
        elif kind == "wgrad":
            instr = mma_wgrad(B, H, W, cin, cout, k, up)
        else:
            continue
        model_ms = instr * cyc / (CLOCK_GHZ * 1e6)
        rows.append((total, n, each, kind, "%dx%dx%d %d->%d k%d%s" % (B, H, W, cin, cout, k, " up" if up else ""), instr, model_ms))
        tot_meas[kind] = tot_meas.get(kind, 0.0) + total
        tot_model[kind] = tot_model.get(kind, 0.0) + model_ms * n
    print("%-8s %-34s %5s %9s %9s %7s" % ("kind", "layer (B x H x W  Cin->Cout)", "calls", "meas ms", "model ms", "ratio"))
    for total, n, each, kind, desc, instr, model_ms in rows:
        print("%-8s %-34s %5d %9.3f %9.3f %7.2f" % (kind, desc, n, each, model_ms, each / model_ms))
    print()
    for kind in tot_meas:
        print("%-6s listed layers: measured %.1f ms, MMA-issue model %.1f ms (%.0f %%)"
              % (kind, tot_meas[kind], tot_model[kind], 100 * tot_model[kind] / tot_meas[kind]))


if __name__ == "__main__":
    main()
