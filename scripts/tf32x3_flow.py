"""3xTF32 tier of the fused stack: log_prob error vs the fp64 oracle and throughput vs the fp32 SIMT / bf16 tiers."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200, oracle
from _cases import build_flow, tame
from nf4ad_b200 import _lib
O, P = oracle.load(), nf4ad_b200.namespace()
CONFIGS = [
    ("C2 D=784 K=8 [256,256]", "NonUSFlow", 784, 8, ("mlp", [256, 256]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
    ("C2' additive", "USFlow", 784, 8, ("mlp_add", [256, 256]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
    ("C3 K=11 [200x3] Laplace", "NonUSFlow", 784, 11, ("mlp", [200, 200, 200]), "laplace", dict(affine_conjugation=True, householder=0)),
    ("C4 D=500 K=8", "NonUSFlow", 500, 8, ("mlp", [256, 256]), "normal", dict(affine_conjugation=True)),
    ("C5 D=128 USFlow [512,256]", "USFlow", 128, 10, ("densenn1", [512, 256]), "normal", dict(affine_conjugation=True, householder=0)),
    ("test D=32 K=3 [128]", "NonUSFlow", 32, 3, ("mlp", [128]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
]
rows = int(os.environ.get("ROWS", "65536"))
for name, kind, D, K, cond, base, kw in CONFIGS:
    torch.manual_seed(0)
    fo = build_flow(O, kind, D, K, cond, base=base, **kw); tame(fo, 0.25)
    fp = build_flow(P, kind, D, K, cond, base=base, **kw); fp.load_state_dict(fo.state_dict()); fp = fp.to("cuda").eval()
    x = torch.randn(rows, D, generator=torch.Generator().manual_seed(42)); xc = x.cuda()
    out = [name]
    with torch.no_grad():
        ref = fo.double().log_prob(x[:256].double())
        for prec in ("tf32x3", "fp32", "bf16"):
            fp.precision = prec
            for _ in range(3): lp = fp.log_prob(xc)
            torch.cuda.synchronize()
            flag = _lib.C.c_int(0); _lib.check(_lib.lib().usf_debug_tc_timeout(_lib.C.byref(flag), 1))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps): lp = fp.log_prob(xc)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            err = float(((lp[:256].double().cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max())
            out.append(f"{prec}: {ms:.2f} ms {rows/ms/1e3:.2f} M/s err {err:.1e} L{fp.last_launches}" + (" TIMEOUT" if flag.value else ""))
    print(" | ".join(out), flush=True)
