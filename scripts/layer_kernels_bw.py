"""Achieved HBM bandwidth of the layer-wise (element-wise / row-reduction) kernels at the benchmark shape
(65536 x 784 fp32), CUDA events, against MEASURED_PEAKS.json's copy bandwidth."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nf4ad_b200 import ops, _lib
from nf4ad_b200._lib import lib, ptr, stream, check

peak = 6546.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
B, D = 65536, 784
g = torch.Generator().manual_seed(0)
x = torch.randn(B, D, generator=g).cuda()
s = (torch.randn(B, D, generator=g) * 0.3).cuda()
t = torch.randn(B, D, generator=g).cuda()
mask = (torch.arange(D) % 2).float().cuda()
V = torch.randn(1, D, generator=g).cuda()
sc = (torch.rand(D, generator=g) + 0.5).cuda()
loc = torch.zeros(D).cuda()

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

rows = []
f4 = 4 * B * D
cases = [
    ("usf_coupling (affine, inverse; x,s,t in, y + ladj out)", lambda: ops.coupling(x, s, t, mask, 5.0, True), 4 * f4 + 4 * B),
    ("usf_coupling (additive, inverse; x,t in, y out)", lambda: ops.coupling(x, None, t, mask, 5.0, True), 3 * f4 + 4 * B),
    ("usf_householder (1 vector; x in, y out)", lambda: ops.householder(x, V), 2 * f4),
    ("usf_scale (x in, y out)", lambda: ops.scale(x, sc), 2 * f4),
    ("usf_base_logprob (Normal; z in, (B,) out)", lambda: ops.base_logprob(0, x, loc, sc), f4 + 4 * B),
    ("usf_to_bf16 (rows + transposed + colsum)", lambda: ops.to_bf16(x, want_rows=True, want_transposed=True, want_colsum=True), f4 + 2 * (2 * B * D)),
]
out = {}
for name, fn, nbytes in cases:
    sec = timeit(fn)
    gbs = nbytes / sec / 1e9
    out[name] = {"us": sec * 1e6, "algorithmic_bytes": nbytes, "GBps": gbs, "frac_of_measured_copy_peak": gbs / peak}
    print(f"{name:62s} {sec*1e6:8.1f} us  {gbs:8.1f} GB/s  {gbs/peak:5.2f} of {peak:.0f}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/layer_kernels_bw.json", "w"), indent=1)
