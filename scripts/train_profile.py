"""Kernel-level breakdown of one training step (torch.profiler, CUDA activity)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
import nf4ad_b200
P = nf4ad_b200.namespace()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flow = bench.build_flow(P, "cuda").train()
opt = torch.optim.Adam(flow.parameters(), lr=1e-4)
x = torch.randn(B, bench.D, device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    loss = -flow.log_prob(x).mean()
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
import time
t=time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter()-t)/5*1e3, "B", B)
