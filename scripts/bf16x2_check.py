"""bf16x2 tier (3 bf16 MMAs per K step on (hi, lo) operand pairs) against the fp64 oracle and the 3xTF32 tier:
log_prob max-row error and time per 65536 rows on the BASELINE config shapes with D >= 128."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, nf4ad_b200, oracle
P, O = nf4ad_b200.namespace(), oracle.load()
for name in ("C2-mnist-D784", "C3-fashion-D784", "C4-adbench-D500-K3", "C4-adbench-D500-K8", "C1-gmm-D128", "C5-mvtec-D128", "C5-mvtec-D256"):
    d = bench.CONFIGS[name][1]
    flow = bench.build_config_flow(P, name, "cuda")
    fo = bench.build_config_flow(O, name, "cpu").double()
    x = torch.randn(65536, d, generator=torch.Generator().manual_seed(42))
    xc = x.cuda()
    with torch.no_grad():
        ref = fo.log_prob(x[:256].double())
        z_ref = fo.backward(x[:256].double())
        line = f"{name:20s}"
        for prec in ("bf16x2", "tf32x3", "bf16"):
            flow.precision = prec
            flow.bf16_trust = True
            lp = flow.log_prob(xc)
            err = float(((lp[:256].double().cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max())
            z = flow.backward(xc[:256].contiguous()).double().cpu()
            zerr = float(((z - z_ref).abs().amax(1) / z_ref.abs().amax(1).clamp_min(1.0)).max())
            for _ in range(5):
                flow.log_prob(xc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                flow.log_prob(xc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line += f" | {prec}: {ms:7.3f} ms {65536 / ms / 1e3:6.1f} M/s err {err:.1e} z {zerr:.1e} L{flow.last_launches}"
        print(line)
