"""Debug: run the tcgen05 GEMM standalone with the in-kernel pipeline trace on and dump per-role timelines
of CTA 0 (records: tile, event, SM clock) to gpurun_out/tc_trace_<tag>.json."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nf4ad_b200 import _lib
from nf4ad_b200._lib import lib, ptr, stream

CAP = 2048
tag = os.environ.get("USF_TC_CTA_GROUP", "2")
out = {}
for (B, N, K, relu) in [(65536, 256, 256, 1), (65536, 256, 392, 1), (65536, 800, 784, 0)]:
    x = torch.randn(B, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    y = torch.empty(B, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        _lib.check(lib().usf_linear_bf16(ptr(x), K, ptr(W), K, ptr(b), relu, ptr(y), N, 1, B, N, K, stream()))
    torch.cuda.synchronize()
    _lib.check(lib().usf_debug_tc_trace(1, None, 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib().usf_linear_bf16(ptr(x), K, ptr(W), K, ptr(b), relu, ptr(y), N, 1, B, N, K, stream()))
    e1.record()
    torch.cuda.synchronize()
    buf = (C.c_uint64 * (2 * 3 * CAP * 2))()
    _lib.check(lib().usf_debug_tc_trace(0, buf, 2 * 3 * CAP))
    roles = {}
    for cta in range(2):
        for role in range(3):
            base = (cta * 3 + role) * CAP * 2
            recs = []
            for i in range(CAP):
                tagv, clk = buf[base + 2 * i], buf[base + 2 * i + 1]
                if clk == 0:
                    break
                recs.append((int(tagv >> 8), int(tagv & 255), int(clk)))
            roles[f"cta{cta}_role{role}"] = recs
    out[f"{B}x{N}x{K}"] = {"ms": e0.elapsed_time(e1), "roles": roles}
    # quick text summary for CTA 0
    r = roles["cta0_role2"]
    t0 = min(v[2] for vs in roles.values() for v in vs) if any(roles.values()) else 0
    print(f"== {B}x{N}x{K} cg={tag}: {e0.elapsed_time(e1)*1e3:.1f} us")
    for name in ("cta0_role0", "cta0_role1", "cta0_role2"):
        recs = roles[name]
        print(name, " ".join(f"t{t}e{e}@{c - t0}" for t, e, c in recs[:40]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/tc_trace_cg{tag}.json", "w"))
