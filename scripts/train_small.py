"""Training step time (fwd + bwd + Adam) at the reference's own batch sizes."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200
from _cases import build_flow
from nf4ad_b200.parallel import DataParallelTrainer
from nf4ad_b200.optim import FusedAdam
P = nf4ad_b200.namespace()
for name, D, K, hid in (("C2 D=784 K=8 [256,256]", 784, 8, [256, 256]), ("test D=32 K=3 [128]", 32, 3, [128]), ("C5 D=128 K=10 [512,256]", 128, 10, [512, 256])):
    torch.manual_seed(0)
    flow = build_flow(P, "NonUSFlow", D, K, ("mlp", hid), base="normal", affine_conjugation=True, prior_scale=1.0).to("cuda").train()
    flow.precision = os.environ.get("TRAIN_PREC", "fp32")
    opt = (torch.optim.Adam(flow.parameters(), lr=1e-4, capturable=True, fused=True) if os.environ.get("TRAIN_OPT") == "torch"
           else FusedAdam(flow.parameters(), lr=1e-4))
    tr = DataParallelTrainer(flow, opt)
    for B in (32, 64, 4096):
        x = torch.randn(B, D, device="cuda")
        for _ in range(5): tr.step(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(10): tr.step(x)
        e1.record(); th = (time.perf_counter() - t0) / 10 * 1e3
        torch.cuda.synchronize()
        print(f"{name} B={B}: {e0.elapsed_time(e1)/10:.2f} ms/step device, host enqueue {th:.2f} ms -> {B/(e0.elapsed_time(e1)/10*1e-3):.0f} samples/s", flush=True)
from torch.profiler import profile, ProfilerActivity
torch.manual_seed(0)
flow = build_flow(P, "NonUSFlow", 784, 8, ("mlp", [256, 256]), base="normal", affine_conjugation=True, prior_scale=1.0).to("cuda").train()
flow.precision = os.environ.get("TRAIN_PREC", "fp32")
opt = torch.optim.Adam(flow.parameters(), lr=1e-4); tr = DataParallelTrainer(flow, opt)
x = torch.randn(int(os.environ.get("PROF_B", "64")), 784, device="cuda")
for _ in range(5): tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): tr.step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
