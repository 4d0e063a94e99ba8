"""Extra measurements for the record: every BASELINE.json / SURVEY section 8(d) config shape, resident-input
log_prob samples/s on one B200 at both precisions + max relative log_prob error vs the fp64 CPU oracle
(256 rows).  The headline number stays bench.py's C2 line; this table shows the path is general."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200
import oracle
from _cases import build_flow, tame

O, P = oracle.load(), nf4ad_b200.namespace()
CONFIGS = [
    ("C1 GMM D=2 USFlow K=10 DenseNN[128,128]", "USFlow", 2, 10, ("densenn1", [128, 128]), "usnormal", dict(affine_conjugation=True, householder=0, prior_scale=1.0)),
    ("C1 GMM D=128 USFlow K=10 DenseNN[128,128]", "USFlow", 128, 10, ("densenn1", [128, 128]), "usnormal", dict(affine_conjugation=True, householder=0, prior_scale=1.0)),
    ("C2 MNIST D=784 NonUSFlow K=8 MLP[256,256]", "NonUSFlow", 784, 8, ("mlp", [256, 256]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
    ("C2' MNIST D=784 USFlow (additive) K=8 MLP[256,256]", "USFlow", 784, 8, ("mlp_add", [256, 256]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
    ("C3 Fashion D=784 NonUSFlow K=11 MLP[200,200,200] Laplace", "NonUSFlow", 784, 11, ("mlp", [200, 200, 200]), "laplace", dict(affine_conjugation=True, householder=0)),
    ("C4 ADBench D=6 K=3 MLP[6]", "NonUSFlow", 6, 3, ("mlp", [6]), "normal", dict(affine_conjugation=True)),
    ("C4 ADBench D=64 K=3 MLP[64]", "NonUSFlow", 64, 3, ("mlp", [64]), "normal", dict(affine_conjugation=True)),
    ("C4 ADBench D=500 K=3 MLP[128]", "NonUSFlow", 500, 3, ("mlp", [128]), "normal", dict(affine_conjugation=True)),
    ("C4 ADBench D=500 K=8 MLP[256,256]", "NonUSFlow", 500, 8, ("mlp", [256, 256]), "normal", dict(affine_conjugation=True)),
    ("C5 MVTec D=128 USFlow K=10 DenseNN[512,256]", "USFlow", 128, 10, ("densenn1", [512, 256]), "normal", dict(affine_conjugation=True, householder=0)),
    ("C5 MVTec D=256 NonUSFlow K=10 MLP[512,256]", "NonUSFlow", 256, 10, ("mlp", [512, 256]), "normal", dict(affine_conjugation=True, householder=0)),
]
rows = 65536
out = []
for name, kind, D, K, cond, base, kw in CONFIGS:
    torch.manual_seed(0)
    fo = build_flow(O, kind, D, K, cond, base=base, **kw)
    tame(fo, 0.25)
    fp = build_flow(P, kind, D, K, cond, base=base, **kw)
    fp.load_state_dict(fo.state_dict())
    fp = fp.to("cuda").eval()
    x = torch.randn(rows, D, generator=torch.Generator().manual_seed(42))
    xc = x.cuda()
    rec = {"config": name, "rows": rows}
    with torch.no_grad():
        ref = fo.double().log_prob(x[:256].double())
        for prec in ("bf16", "fp32"):
            fp.precision = prec
            for _ in range(3):
                lp = fp.log_prob(xc)
            torch.cuda.synchronize()
            reps = 10 if prec == "bf16" else 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                lp = fp.log_prob(xc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            err = float(((lp[:256].double().cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max())
            med = float(((lp[:256].double().cpu() - ref).abs() / ref.abs().clamp_min(1.0)).median())
            rec[prec] = {"ms_per_call": ms, "samples_per_s": rows / (ms * 1e-3), "launches": fp.last_launches,
                         "max_rel_err_vs_fp64_oracle": err, "median_rel_err": med}
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
