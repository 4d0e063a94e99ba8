"""End-to-end scoring call (pinned fp32 rows -> scores on the host) by pipeline chunk size and fp32-head size."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200
from nf4ad_b200.parallel import ShardedScorer
P = nf4ad_b200.namespace()
flow = bench.build_flow(P, "cuda"); flow.precision = "bf16"
x_host = torch.randn(65536, 784, generator=torch.Generator().manual_seed(42)).pin_memory()
with torch.no_grad():
    flow.log_prob(x_host[:512].cuda())
for chunk in (16384, 8192, 12288, 4096):
    for raw in (16384, 8192, 24576, 12288):
        sc = ShardedScorer(flow)
        sc.chunk_rows, sc.raw_rows, sc.autotune = chunk, raw, False
        for _ in range(4):
            sc.predict_score_host(x_host)
        torch.cuda.synchronize()
        ts = []
        for _ in range(12):
            t0 = time.perf_counter()
            sc.predict_score_host(x_host)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        print(f"chunk {chunk:6d} head {raw:6d}: median {ts[len(ts)//2]*1e3:.3f} ms  best {ts[0]*1e3:.3f} ms  h2d {sc.last_h2d_bytes/1e6:.0f} MB  narrow {sc.host_bf16} threads {sc.host_threads}")
