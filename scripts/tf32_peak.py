""""TF32 tensor peak: not measured -- builder measures it the same way (`allow_tf32`) and records it".
The same recipe as MEASURED_PEAKS.json's bf16 figure: torch.matmul 8192^3 (2 N^3 FLOPs), best of 10 (burst) and back to back
for a few seconds (sustained), CUDA events; bf16 re-measured beside it on the same box so that the ratio is same-box.
Library GEMMs (cuBLAS) as a yardstick only -- nothing here is on the product path.  One JSON line."""
import json
import sys
import time

import torch


def peak(dtype, tf32, seconds):
    n = 8192
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    flop = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    per = flop / (best * 1e12)
    iters = max(10, int(seconds / per))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        torch.matmul(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    return {"burst_tflops": best, "sustained_tflops": iters * flop / (e0.elapsed_time(e1) * 1e-3) / 1e12, "sustained_iters": iters}


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
    out = {"how": f"torch.matmul 8192^3, best of 10 (burst) and back to back for ~{seconds:g} s (sustained), CUDA events",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "tf32": peak(torch.float32, True, seconds), "bf16": peak(torch.bfloat16, False, seconds),
           "fp32_no_tf32": peak(torch.float32, False, min(seconds, 1.0))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
