"""The ADBench wrapper's scoring call on a pageable numpy test set (what the reference's callers pass,
adbench_wrapper.py:406-433): host narrowing on / off, by precision tier."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200
from nf4ad_b200.parallel import ShardedScorer
from _cases import build_flow
P = nf4ad_b200.namespace()
rows, D = 65536, 784
torch.manual_seed(0)
flow = build_flow(P, "NonUSFlow", D, 8, ("mlp", [256, 256]), base="normal", affine_conjugation=True, prior_scale=1.0).to("cuda").eval()
X = np.random.default_rng(0).standard_normal((rows, D), dtype=np.float32)
for prec in ("bf16", "fp32"):
    flow.precision = prec
    for narrow in (False, True):
        sc = ShardedScorer(flow)
        sc.host_bf16 = sc.host_staging = narrow
        xt = torch.as_tensor(X)
        for _ in range(2):
            s = sc.predict_score_host(xt)
        t = time.perf_counter()
        for _ in range(5):
            s = sc.predict_score_host(xt)
        dt = (time.perf_counter() - t) / 5
        print(f"pageable numpy rows, {prec}, host staging/narrowing={narrow}: {dt*1e3:.2f} ms per {rows} rows = {rows/dt/1e6:.2f} M samples/s", flush=True)

# the wrapper call itself (numpy in, numpy out), bf16 tier
from nf4ad_b200.adbench import ADBenchFlow
flow.precision = "bf16"
w = ADBenchFlow(flow, verbose=False)
for _ in range(2):
    w.predict_score(X)
t = time.perf_counter()
for _ in range(5):
    sc_np = w.predict_score(X)
dt = (time.perf_counter() - t) / 5
print(f"ADBenchFlow.predict_score(numpy {rows}x{D}), bf16 tier: {dt*1e3:.2f} ms = {rows/dt/1e6:.2f} M samples/s", flush=True)
# the previous form of that call: pin the whole array, then the plain chunked copy
sc = ShardedScorer(flow); sc.host_bf16 = sc.host_staging = False
t = time.perf_counter()
for _ in range(3):
    s_old = sc.predict_score_host(torch.as_tensor(X).pin_memory()).numpy()
dt = (time.perf_counter() - t) / 3
print(f"previous form (pin_memory of the whole array + plain chunked copy): {dt*1e3:.2f} ms = {rows/dt/1e6:.2f} M samples/s", flush=True)
print("same scores:", bool(np.allclose(sc_np, s_old, rtol=1e-6, atol=1e-3)))
