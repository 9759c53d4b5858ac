"""SASS mnemonic counts of bts_b200/libbts_b200.so per kernel family (cuobjdump -sass; no GPU needed):
the Blackwell-native instructions B200_PROFILING.md names (UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR =
tcgen05.commit, UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier ops ...).
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bts_b200", "libbts_b200.so")
MNEM = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "HMMA", "ELECT",
        "BRA.U.ANY", "FENCE.VIEW.ASYNC", "LDG.E.128", "STG.E.128", "SHFL"]
FAMILIES = ["conv_tc_kernel", "wgrad_tc_kernel", "wgrad2_tc_kernel", "lpg_bwd", "lpg_fwd", "thin_", "pw_wgrad", "pw_fwd",
            "bn_", "wgrad_reduce", "silog", "pack_weights", "eval_", "maxpool", "avgpool2", "input_prep", "adamw", "plane_head",
            "elu_bwd", "upsample2", "copy_channels"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fam = None
    counts = collections.defaultdict(lambda: collections.Counter())
    inst = collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            fam = next((f for f in FAMILIES if f in name), "other")
            inst[fam] += 1
            continue
        if fam is None or "/*" not in line:
            continue
        m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        op = m.group(1)
        counts[fam]["_ins"] += 1
        for k in MNEM:
            if op == k or op.startswith(k + ".") or (k == "BRA.U.ANY" and op.startswith("BRA.U") and "ANY" in op):
                counts[fam][k] += 1
    total = collections.Counter()
    for f in counts:
        total.update(counts[f])
    print("# SASS evidence of bts_b200/libbts_b200.so (cuobjdump -sass, sm_100a), final build of round 2 (tools/sass_summary.py)")
    print("# whole library: " + ", ".join("%s %d" % (k, total[k]) for k in MNEM))
    print()
    print("%-18s %5s %8s " % ("kernel family", "inst.", "SASS ins") + " ".join("%s" % k for k in MNEM))
    for f in sorted(counts, key=lambda f: -counts[f]["_ins"]):
        print("%-18s %5d %8d " % (f, inst[f], counts[f]["_ins"]) + " ".join("%*d" % (len(k), counts[f][k]) for k in MNEM))


if __name__ == "__main__":
    sys.exit(main())
