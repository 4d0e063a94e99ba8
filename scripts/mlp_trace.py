"""Debug: timeline of the fused conditioner kernel (last traced launch = last block of the bench stack)."""
import ctypes as C, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["USF_GRAPHS"] = "0"
import bench, nf4ad_b200
from nf4ad_b200 import _lib
from nf4ad_b200._lib import lib
P = nf4ad_b200.namespace()
flow = bench.build_flow(P, "cuda"); flow.precision = "bf16"
x = torch.randn(65536, bench.D, device="cuda")
CAP = 2048
with torch.no_grad():
    for _ in range(3): flow.log_prob(x)
    torch.cuda.synchronize()
    # trace only a 1-block prefix: build a K=1 flow so that the LAST traced launch is the fused kernel
    _lib.check(lib().usf_debug_tc_trace(2 | (int(os.environ.get("CPL_DBG", "0")) << 16), None, 0))
    flow.log_prob(x)
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
            if clk == 0: break
            recs.append((int(tagv >> 8), int(tagv & 255), int(clk)))
        roles[f"cta{cta}_role{role}"] = recs
os.makedirs("gpurun_out", exist_ok=True)
json.dump(roles, open("gpurun_out/mlp_trace.json", "w"))
print({k: len(v) for k, v in roles.items()})
t0 = min(v[2] for vs in roles.values() for v in vs)
for name in ("cta0_role1", "cta0_role2", "cta1_role2"):
    recs = roles[name]
    tiles = sorted(set(r[0] >> 4 for r in recs))
    for tl in tiles[1:3]:
        print(name, "rowtile", tl, " ".join(f"g{t & 15}e{e}@{c - t0}" for t, e, c in recs if (t >> 4) == tl))
