"""Kernel-time breakdown of one graph-replayed training step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200
from _cases import build_flow
from nf4ad_b200.parallel import DataParallelTrainer
from torch.profiler import profile, ProfilerActivity
P = nf4ad_b200.namespace()
for name, D, K, hid, B in (("C5 D=128 K=10 [512,256]", 128, 10, [512, 256], 64), ("C2 D=784 K=8 [256,256]", 784, 8, [256, 256], 64), ("test D=32 K=3 [128]", 32, 3, [128], 64)):
    torch.manual_seed(0)
    flow = build_flow(P, "NonUSFlow", D, K, ("mlp", hid), base="normal", affine_conjugation=True, prior_scale=1.0).to("cuda").train()
    flow.precision = os.environ.get("TRAIN_PREC", "fp32")
    opt = torch.optim.Adam(flow.parameters(), lr=1e-4, capturable=True, fused=True)
    tr = DataParallelTrainer(flow, opt)
    x = torch.randn(B, D, device="cuda")
    for _ in range(6): tr.step(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): tr.step(x)
        torch.cuda.synchronize()
    print("=====", name, "B", B)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70).split("Self CPU time")[0][-4200:])
