"""Soak: thousands of replays of the tensor-core chains on fixed inputs; every result must equal the first one (up to
the atomics' last-ulp order effects) and no bounded mbarrier wait may expire.  Ragged row counts included."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200
from _cases import build_flow, tame
from nf4ad_b200 import _lib
P = nf4ad_b200.namespace()
CONFIGS = [
    ("C2", "NonUSFlow", 784, 8, ("mlp", [256, 256]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
    ("C3", "NonUSFlow", 784, 11, ("mlp", [200, 200, 200]), "laplace", dict(affine_conjugation=True, householder=0)),
    ("C2add", "USFlow", 784, 8, ("mlp_add", [256, 256]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
    ("C5", "USFlow", 128, 10, ("densenn1", [512, 256]), "normal", dict(affine_conjugation=True, householder=0)),
    ("D32", "NonUSFlow", 32, 3, ("mlp", [128]), "normal", dict(affine_conjugation=True, prior_scale=1.0)),
]
N = int(os.environ.get("SOAK_ITERS", "600"))
bad = 0
for name, kind, D, K, cond, base, kw in CONFIGS:
    torch.manual_seed(0)
    f = build_flow(P, kind, D, K, cond, base=base, **kw); tame(f, 0.25); f = f.to("cuda").eval()
    for prec in ("bf16", "tf32x3"):
        f.precision = prec
        for rows in (65536, 40001, 257, 5):
            x = torch.randn(rows, D, device="cuda")
            with torch.no_grad():
                first = f.log_prob(x).clone()
                worst = 0.0
                for i in range(N if rows >= 40001 else N // 4):
                    lp = f.log_prob(x)
                    if i % 50 == 49:
                        worst = max(worst, float(((lp - first).abs() / first.abs().clamp_min(1.0)).max()))
                torch.cuda.synchronize()
            flag = _lib.C.c_int(0); _lib.check(_lib.lib().usf_debug_tc_timeout(_lib.C.byref(flag), 1))
            ok = flag.value == 0 and worst < 1e-5 and bool(torch.isfinite(first).all())
            bad += 0 if ok else 1
            print(f"{name} {prec} rows={rows}: worst drift {worst:.1e} timeout={flag.value} {'ok' if ok else 'FAIL'}", flush=True)
print("SOAK", "PASSED" if bad == 0 else f"FAILED ({bad})")
