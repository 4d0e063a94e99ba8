"""Where a small-batch log_prob call spends its time: host enqueue vs device, eager vs graph replay."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200
P = nf4ad_b200.namespace()
flow = bench.build_flow(P, "cuda")
for prec in ("bf16", "fp32"):
    flow.precision = prec
    for rows in (16, 64, 1024, 16384):
        x = torch.randn(rows, bench.D, device="cuda")
        with torch.no_grad():
            for _ in range(5):
                flow.log_prob(x)
            torch.cuda.synchronize()
            reps = 50
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(reps):
                flow.log_prob(x)
            e1.record()
            t_host = (time.perf_counter() - t0) / reps * 1e3
            torch.cuda.synchronize()
            t_wall = (time.perf_counter() - t0) / reps * 1e3
            # single call latency (sync each time)
            t1 = time.perf_counter()
            for _ in range(20):
                flow.log_prob(x); torch.cuda.synchronize()
            t_lat = (time.perf_counter() - t1) / 20 * 1e3
        print(f"{prec} rows={rows}: host enqueue {t_host:.3f} ms/call, device {e0.elapsed_time(e1)/reps:.3f} ms/call, wall {t_wall:.3f}, sync latency {t_lat:.3f} ms, launches {flow.last_launches}", flush=True)
import cProfile, pstats
flow.precision = "bf16"
x = torch.randn(64, bench.D, device="cuda")
with torch.no_grad():
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): flow.log_prob(x)
    pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
