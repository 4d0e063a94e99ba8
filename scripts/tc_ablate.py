"""Debug: which resource bounds the tcgen05 G GEMM?  Times usf_linear_bf16 (65536 x 800 x 784 and the conditioner
shapes) with the kernel's ablation switches (bits 8+ of usf_debug_tc_trace's `on`): 1 = no global stores, 2 = no
epilogue at all, 4 = no A loads, 8 = no W loads, 16 = no MMAs.  Results are wrong by construction; only times matter."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nf4ad_b200 import _lib
from nf4ad_b200._lib import lib, ptr, stream

out = {}
for (B, N, K, relu) in [(65536, 800, 784, 0), (65536, 256, 392, 1), (65536, 832, 256, 0)]:
    x = torch.randn(B, K + (-K) % 8, device="cuda").bfloat16()
    ldx = x.shape[1]
    W = (torch.randn(N, ldx, device="cuda") / K ** 0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    y = torch.empty(B, N, device="cuda", dtype=torch.bfloat16)
    res = {}
    for dbg in (0, 0x200 | 4, 0x200 | 8, 0x200 | 12, 128, 12):
        _lib.check(lib().usf_debug_tc_trace(dbg << 8, None, 0))
        for _ in range(3):
            _lib.check(lib().usf_linear_bf16(ptr(x), ldx, ptr(W), ldx, ptr(b), relu, ptr(y), N, 1, B, N, K, stream()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(lib().usf_linear_bf16(ptr(x), ldx, ptr(W), ldx, ptr(b), relu, ptr(y), N, 1, B, N, K, stream()))
        e1.record()
        torch.cuda.synchronize()
        res[dbg] = e0.elapsed_time(e1) / 20 * 1e3
    _lib.check(lib().usf_debug_tc_trace(0, None, 0))
    flag = _lib.C.c_int(0) if hasattr(_lib, "C") else None
    print(f"{B}x{N}x{K}: " + "  ".join(f"dbg{d}={t:.1f}us" for d, t in res.items()), flush=True)
    out[f"{B}x{N}x{K}"] = res
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tc_ablate.json", "w"), indent=1)
