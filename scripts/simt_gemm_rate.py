"""fp32 SIMT GEMM (usf_linear) rate at the stack's shapes vs torch (cuBLAS SGEMM, TF32 off) on the same box."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nf4ad_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
for (B, N, K) in [(65536, 800, 784), (65536, 256, 392), (65536, 256, 256), (65536, 896, 256), (16384, 800, 784)]:
    x = torch.randn(B, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.zeros(N, device="cuda")
    ours = t(lambda: ops.linear(x, W, b, False))
    ref = t(lambda: torch.addmm(b, x, W.t()))
    fl = 2.0 * B * N * K
    print(f"{B}x{N}x{K}: ours {ours*1e6:8.1f} us = {fl/ours/1e12:5.1f} TFLOP/s | cuBLAS sgemm {ref*1e6:8.1f} us = {fl/ref/1e12:5.1f} TFLOP/s", flush=True)
