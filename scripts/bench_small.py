"""Small event shapes: the one-kernel path (csrc/usf_small.cu) against the launch chain (3xTF32 kernels), resident rows,
device time per call at several batch sizes.  Decides the row threshold in Flow._small_ok."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200
from _cases import build_flow, tame
P = nf4ad_b200.namespace()


def timed(flow, x):
    with torch.no_grad():
        for _ in range(5):
            flow.log_prob(x)
        torch.cuda.synchronize()
        reps = 30 if x.shape[0] <= 65536 else 8
        g = torch.cuda.CUDAGraph()          # graph of `reps` calls: device time without the Python host cost per call
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            flow.log_prob(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(reps):
                flow.log_prob(x)
            e1.record(s)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


cases = [("C4-adbench-D6", None), ("fixture-D20-K3-h20", ("NonUSFlow", 20, 3, ("mlp", [20]))),
         ("fixture-D32-K3-h128", ("NonUSFlow", 32, 3, ("mlp", [128]))), ("C4-adbench-D64", None), ("C1-gmm-D2", None)]
for name, spec in cases:
    if spec is None:
        d = bench.CONFIGS[name][1]
        mk = lambda: bench.build_config_flow(P, name, "cuda")
    else:
        d = spec[1]
        def mk():
            torch.manual_seed(0)
            f = build_flow(P, spec[0], spec[1], spec[2], spec[3], affine_conjugation=True, prior_scale=1.0)
            tame(f, 0.25)
            return f.to("cuda").eval()
    for rows in (32, 1024, 8192, 65536, 524288):
        x = torch.randn(rows, d, device="cuda")
        out = []
        for small in (True, False):
            flow = mk()
            flow.precision = "bf16"
            flow.SMALL_MAX_DIM = 64 if small else 0
            flow.SMALL_ALWAYS = True
            ms = timed(flow, x)
            out.append((ms, flow.last_launches, flow.effective_precision))
        (a, la, ta), (b, lb, tb) = out
        print(f"{name:22s} rows {rows:7d}  one-kernel {a * 1e3:9.1f} us ({la} launch, {ta})   chain {b * 1e3:9.1f} us ({lb} launches, {tb})   "
              f"ratio {b / a:5.2f}  one-kernel HBM frac {rows * (4 * d + 4) / (a * 1e-3) / 6546.9e9:.4f}")
