"""Per-step device time of the replayed scoring step (events after every step), to see whether slow runs are uniformly
slow or have slow stretches."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200
from nf4ad_b200.parallel import ShardedScorer
P = nf4ad_b200.namespace()
flow = bench.build_flow(P, "cuda"); flow.precision = "bf16"
x = torch.randn(65536, bench.D, device="cuda")
scorer = ShardedScorer(flow)
N = int(os.environ.get("STEPS", "300"))
with torch.no_grad():
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0:
        for _ in range(10): lp = scorer.score_local(x)
        torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    evs[0].record()
    th = time.perf_counter()
    for i in range(N):
        lp = scorer.score_local(x)
        evs[i + 1].record()
    th = (time.perf_counter() - th) / N * 1e3
    torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(N)]
chunks = [sum(ms[i:i + 25]) / 25 for i in range(0, N, 25)]
print(f"host {th:.3f} ms/step | mean {sum(ms)/N:.3f} min {min(ms):.3f} max {max(ms):.3f} | per-25-step means: " + " ".join(f"{c:.2f}" for c in chunks))
