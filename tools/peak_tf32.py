"""Measures the TF32 dense GEMM peak of this GPU the way MEASURED_PEAKS.json was made for bf16 (cuBLAS via torch.matmul,
8192^3: best of 10 = burst; back to back for `secs` seconds = sustained).  Prints one JSON object."""
import json, sys, time
import torch


def measure(secs=4.0, n=8192):
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    c = torch.empty(n, n, device="cuda")
    fl = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); k = 0
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            torch.matmul(a, b, out=c)
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / k
    out = {"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl / sus / 1e9,
           "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32) %d^3: best of 10 (burst), back to back %.0f s (sustained)" % (n, secs)}
    # same for bf16 as a cross-check against MEASURED_PEAKS.json
    a16, b16 = a.bfloat16(), b.bfloat16()
    c16 = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        torch.matmul(a16, b16, out=c16)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a16, b16, out=c16); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["bf16_tflops_check"] = fl / best / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    return out


if __name__ == "__main__":
    print(json.dumps(measure(float(sys.argv[1]) if len(sys.argv) > 1 else 4.0)))
