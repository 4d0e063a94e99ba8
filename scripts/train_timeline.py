"""Where one graph-replayed training step (C2, B = 4096) spends its time: per-kernel totals, per-stream busy time and
the idle gaps of the union of all streams (torch.profiler over graph replays)."""
import json, os, sys, tempfile, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import nf4ad_b200
from nf4ad_b200.parallel import DataParallelTrainer
from nf4ad_b200.optim import FusedAdam
from torch.profiler import profile, ProfilerActivity
P = nf4ad_b200.namespace()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = os.environ.get("TRAIN_PREC", "bf16")
torch.manual_seed(0)
flow = bench.build_flow(P, "cuda").train()
flow.precision = prec
opt = FusedAdam(flow.parameters(), lr=1e-4)
tr = DataParallelTrainer(flow, opt)
x = torch.randn(B, bench.D, device="cuda")
for _ in range(8): tr.step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): tr.step(x)
e1.record(); torch.cuda.synchronize()
print(f"precision {prec} B {B}: {e0.elapsed_time(e1) / 20:.3f} ms/step; graph_error={tr.graph_error}")
eager = os.environ.get("TIMELINE_EAGER") == "1"      # eager: the profiler shows the real torch streams
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    if eager:
        tr._eager_step(x)
    else:
        for _ in range(int(os.environ.get("TIMELINE_REPLAYS", "4"))):      # back to back: the last one is steady state
            tr.step(x)
    torch.cuda.synchronize()
if eager:
    print("side streams:", [s.cuda_stream for s in flow.__dict__.get("_side_streams", [])])
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
def short(n):
    n = n.replace("(anonymous namespace)::", "").replace("void ", "").replace("usf::", "").replace("at::native::", "")
    return n.split("(")[0][:70]
for e in ev: e["name"] = short(e["name"])
ev.sort(key=lambda e: e["ts"])
if not eager:                       # keep the last replay: from its input copy (the step's first device op) on
    starts = [i for i, e in enumerate(ev) if e["name"].startswith("Memcpy DtoD") and e["args"].get("bytes", 0) >= x.numel() * 4]
    if len(starts) > 1:
        print(f"replay starts at {[round(ev[i]['ts'] - ev[0]['ts']) for i in starts]} us")
        ev = ev[starts[-1]:]
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"one replay: {len(ev)} device ops, span {(t1 - t0) / 1e3:.3f} ms")
tot = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"]
    tot[n][0] += 1; tot[n][1] += e["dur"]
print(f"{'kernel':72s} {'n':>5s} {'total us':>10s} {'avg us':>8s}")
for n, (c, d) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{n:72s} {c:5d} {d:10.1f} {d / c:8.2f}")
print(f"{'sum of all device ops':72s} {len(ev):5d} {sum(e['dur'] for e in ev):10.1f}")
streams = collections.defaultdict(float)
for e in ev: streams[e["args"].get("stream")] += e["dur"]
print("busy us per stream:", {k: round(v) for k, v in streams.items()})
# union coverage + gaps
cover, gaps, end = 0.0, [], t0
for e in ev:
    s, f = e["ts"], e["ts"] + e["dur"]
    if s > end: gaps.append((s - end, e["name"][:50])); cover += f - s; end = f
    elif f > end: cover += f - end; end = f
print(f"union busy {cover / 1e3:.3f} ms, idle {sum(g for g, _ in gaps) / 1e3:.3f} ms in {len(gaps)} gaps")
# timeline in 20 slices: which kernels dominate each slice
N = 20
sl = [collections.defaultdict(float) for _ in range(N)]
w = (t1 - t0) / N
for e in ev:
    i = min(N - 1, int((e["ts"] - t0) / w))
    sl[i][e["name"][:34]] += e["dur"]
for i, d in enumerate(sl):
    top = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    print(f"  slice {i:2d} ({i * w / 1e3:5.2f} ms): " + ", ".join(f"{k} {v:.0f}us" for k, v in top))
if os.environ.get("TIMELINE_DUMP"):
    with open(os.environ["TIMELINE_DUMP"], "w") as f:
        for e in ev:
            f.write(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} {e['args'].get('stream'):4d} {e['name']}\n")
