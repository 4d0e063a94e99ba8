"""Host cost of one replayed scoring call, split into the torch allocations and the usf_stack_run (cudaGraphLaunch) call."""
import ctypes as C, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200
from nf4ad_b200 import _lib, ops
from nf4ad_b200._lib import lib, ptr, stream, check
P = nf4ad_b200.namespace()
flow = bench.build_flow(P, "cuda"); flow.precision = "bf16"
x = torch.randn(65536, bench.D, device="cuda")
with torch.no_grad():
    for _ in range(20): flow.log_prob(x)
    torch.cuda.synchronize()
    cs = flow._stack(True, x.device)
    xr, ldx = ops._rows(x)
    B = x.shape[0]
    nbytes = lib().usf_stack_workspace_bytes(C.byref(cs.desc), B, cs.precision)
    n = C.c_int(0)
    t_alloc, t_call, t_total = [], [], []
    for it in range(300):
        t0 = time.perf_counter()
        lp = torch.empty(B, device="cuda"); ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
        t1 = time.perf_counter()
        check(lib().usf_stack_run(C.byref(cs.desc), ptr(xr), ldx, B, ptr(lp), None, cs.D, None, ptr(ws), nbytes, cs.precision, C.byref(n), stream()))
        t2 = time.perf_counter()
        t_alloc.append(t1 - t0); t_call.append(t2 - t1)
        if it % 50 == 49: torch.cuda.synchronize()
    torch.cuda.synchronize()
    import statistics as st
    print(f"alloc mean {1e3*st.mean(t_alloc):.3f} ms max {1e3*max(t_alloc):.3f} | stack_run mean {1e3*st.mean(t_call):.3f} ms median {1e3*st.median(t_call):.3f} max {1e3*max(t_call):.3f}")
    g = (C.c_longlong * 4)(); lib().usf_debug_graph_stats(g, None, 0); print("graph stats", list(g))
    # same with a persistent workspace / output (no allocation, one key)
    lp = torch.empty(B, device="cuda"); ws = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
    t_call = []
    for it in range(300):
        t1 = time.perf_counter()
        check(lib().usf_stack_run(C.byref(cs.desc), ptr(xr), ldx, B, ptr(lp), None, cs.D, None, ptr(ws), nbytes, cs.precision, C.byref(n), stream()))
        t_call.append(time.perf_counter() - t1)
        if it % 50 == 49: torch.cuda.synchronize()
    torch.cuda.synchronize()
    print(f"persistent buffers: stack_run mean {1e3*st.mean(t_call):.3f} ms median {1e3*st.median(t_call):.3f} max {1e3*max(t_call):.3f}")
