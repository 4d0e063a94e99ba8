"""Debug: per-k-block timeline of the G GEMM's producer / MMA threads (CTA 0) under the ablation switches."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nf4ad_b200 import _lib
from nf4ad_b200._lib import lib, ptr, stream

CAP = 2048
B, N, K = 65536, 800, 784
x = torch.randn(B, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
b = torch.zeros(N, device="cuda")
y = torch.empty(B, N, device="cuda", dtype=torch.bfloat16)
for dbg in (0, 14, 30):
    _lib.check(lib().usf_debug_tc_trace(dbg << 8, None, 0))
    for _ in range(3):
        _lib.check(lib().usf_linear_bf16(ptr(x), K, ptr(W), K, ptr(b), 0, ptr(y), N, 1, B, N, K, stream()))
    torch.cuda.synchronize()
    _lib.check(lib().usf_debug_tc_trace((dbg << 8) | 1, None, 0))
    _lib.check(lib().usf_linear_bf16(ptr(x), K, ptr(W), K, ptr(b), 0, ptr(y), N, 1, B, N, K, stream()))
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
            roles[(cta, role)] = recs
    t0 = min(v[2] for vs in roles.values() for v in vs)
    print(f"==== dbg={dbg}")
    for key in ((0, 0), (1, 0), (0, 1), (0, 2), (1, 2)):
        recs = roles[key]
        tiles = sorted(set(r[0] for r in recs))
        if len(tiles) < 6:
            continue
        for tl in tiles[4:6]:
            print(f"cta{key[0]} role{key[1]} tile{tl}: " + " ".join(f"e{e}@{c - t0}" for t, e, c in recs if t == tl))
