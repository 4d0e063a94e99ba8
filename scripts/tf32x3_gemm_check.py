"""3xTF32 GEMM: accuracy vs fp64 and speed vs the fp32 SIMT GEMM / bf16 tcgen05 GEMM."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nf4ad_b200 import ops, _lib
from nf4ad_b200._lib import lib, ptr, stream, check

def lo_of(x):
    out = torch.empty_like(x)
    check(lib().usf_split_lo(ptr(x), x.stride(0), ptr(out), out.stride(0), x.shape[0], x.shape[1], stream()))
    return out

def run(B, N, K, relu):
    g = torch.Generator().manual_seed(B + N + K)
    ld = (K + 3) // 4 * 4
    x = torch.zeros(B, ld); x[:, :K] = torch.randn(B, K, generator=g); x = x.cuda()
    W = torch.zeros(N, ld); W[:, :K] = torch.randn(N, K, generator=g) / K ** 0.5; W = W.cuda()
    b = torch.randn(N, generator=g).cuda()
    xlo, Wlo = lo_of(x), lo_of(W)
    y = torch.empty(B, N, device="cuda"); ylo = torch.empty(B, N, device="cuda")
    def call():
        check(lib().usf_linear_tf32x3(ptr(x), ptr(xlo), ld, ptr(W), ptr(Wlo), ld, ptr(b), int(relu), ptr(y), ptr(ylo), N, B, N, K, stream()))
    call(); torch.cuda.synchronize()
    flag = _lib.C.c_int(0)
    check(lib().usf_debug_tc_timeout(_lib.C.byref(flag), 1))
    ref = x[:, :K].double().cpu() @ W[:, :K].double().cpu().t() + b.double().cpu()
    if relu: ref = ref.clamp_min(0)
    yy = y.double().cpu() + ylo.double().cpu()          # the output travels as (hi, lo)
    err = float((yy - ref).abs().max() / ref.abs().max())
    y32 = ops.linear(x[:, :K].contiguous(), W[:, :K].contiguous(), b, relu)
    err32 = float((y32.double().cpu() - ref).abs().max() / ref.abs().max())
    lo_ok = float(((y.view(torch.int32) & 8191) != 0).sum())   # hi must be tf32-exact (low 13 bits zero)
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"{B}x{N}x{K} relu={relu}: timeout_flag={flag.value} max rel err 3xTF32 {err:.2e} (fp32 SIMT {err32:.2e}) hi not tf32-exact: {lo_ok:.0f} | {us:.1f} us = {2.0*B*N*K/us/1e6:.1f} TFLOP/s", flush=True)

for cfg in [(256, 208, 64, False), (300, 256, 392, True), (4096, 800, 784, False), (65536, 800, 784, False), (65536, 256, 392, True), (65536, 896, 256, False)]:
    run(*cfg)
