"""Host narrowing (fp32 -> bf16 on the host cores before the PCIe copy): converter throughput by thread count and the
end-to-end scoring call (pinned host rows -> scores on the host) with and without it, C2 shape."""
import ctypes as C, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import nf4ad_b200
from nf4ad_b200 import _lib
from nf4ad_b200.parallel import ShardedScorer
from _cases import build_flow
L = _lib.lib()
print("hardware threads", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
rows, D = 65536, 784
x = torch.randn(rows, D).pin_memory()
out = torch.empty(rows, D, dtype=torch.bfloat16).pin_memory()
for th in (1, 2, 4, 8, 16, 32, 0):
    for _ in range(2):
        L.usf_host_f32_to_bf16(C.c_void_p(x.data_ptr()), D, C.c_void_p(out.data_ptr()), D, rows, D, th)
    t = time.perf_counter()
    for _ in range(10):
        L.usf_host_f32_to_bf16(C.c_void_p(x.data_ptr()), D, C.c_void_p(out.data_ptr()), D, rows, D, th)
    dt = (time.perf_counter() - t) / 10
    print(f"threads {th:2d}: {dt*1e3:7.2f} ms per {rows} rows = {rows*D*4/dt/1e9:6.1f} GB/s of fp32 read", flush=True)
for th in (4, 16, 0):       # chunk-sized calls (what the pipeline issues)
    t = time.perf_counter()
    for _ in range(40):
        L.usf_host_f32_to_bf16(C.c_void_p(x.data_ptr()), D, C.c_void_p(out.data_ptr()), D, 16384, D, th)
    print(f"threads {th:2d}: {(time.perf_counter()-t)/40*1e3:.3f} ms per 16384-row chunk", flush=True)
P = nf4ad_b200.namespace()
torch.manual_seed(0)
flow = build_flow(P, "NonUSFlow", D, 8, ("mlp", [256, 256]), base="normal", affine_conjugation=True, prior_scale=1.0).to("cuda").eval()
flow.precision = "bf16"
sc = ShardedScorer(flow)
for narrow, chunk, raw, th in ((False, 16384, 0, 0), (True, 16384, 0, 0), (True, 16384, 8192, 0), (True, 16384, 16384, 0),
                               (True, 16384, 24576, 0), (True, 8192, 16384, 0), (True, 16384, 16384, 12), (True, 16384, 16384, 8)):
    sc.host_bf16, sc.chunk_rows, sc.raw_rows, sc._ring, sc.autotune = narrow, chunk, raw, None, False
    sc.host_threads = th or os.cpu_count()
    for _ in range(3):
        s = sc.predict_score_host(x)
    t = time.perf_counter()
    for _ in range(20):
        s = sc.predict_score_host(x)
    dt = (time.perf_counter() - t) / 20
    print(f"predict_score_host narrow={narrow} chunk={chunk} fp32-head={raw} threads={sc.host_threads}: {dt*1e3:.2f} ms per "
          f"{rows} rows = {rows/dt/1e6:.1f} M samples/s, {sc.last_h2d_bytes/1e6:.0f} MB over the link", flush=True)
