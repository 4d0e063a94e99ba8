import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import nf4ad_b200, oracle
from nf4ad_b200 import ops
from _cases import build_flow, randomize_constants, tame
O = oracle.load(); P = nf4ad_b200.namespace()
D, K, hidden, B = 784, 2, [256], 1024
torch.manual_seed(0)
fo = build_flow(O, "NonUSFlow", D, K, ("mlp", hidden), affine_conjugation=True, prior_scale=1.0)
tame(fo, 0.25); randomize_constants(fo, 3)
flow = build_flow(P, "NonUSFlow", D, K, ("mlp", hidden), affine_conjugation=True, prior_scale=1.0)
flow.load_state_dict(fo.state_dict()); flow = flow.to("cuda").train()
xc = torch.randn(B, D, generator=torch.Generator().manual_seed(5)); x = xc.cuda()
fo = fo.double(); l64 = -fo.log_prob(xc.double()).mean(); l64.backward()
g64 = {n: p.grad.detach() for n, p in fo.named_parameters() if p.grad is not None}
def run(tag, **attrs):
    for k, v in attrs.items(): setattr(flow, k, v)
    flow.zero_grad(set_to_none=True)
    (-flow.log_prob(x).mean()).backward(); torch.cuda.synchronize()
    worst = sorted(((float((p.grad.double().cpu() - g64[n]).norm() / g64[n].norm()), n) for n, p in flow.named_parameters() if p.grad is not None and float(g64[n].norm()) > 1e-9), reverse=True)[:4]
    print(tag, [(f"{e:.2e}", n.replace("trainable_layers.", "L").replace("block_transform.transforms.", "t")) for e, n in worst])
run("fp32", precision="fp32")
run("t3 layerwise", precision="tf32x3", compose_affine=False)
run("t3 composed", precision="tf32x3", compose_affine=True)
ops._LU_T3_PRODUCT = False
run("t3 composed, fp32 LU product", precision="tf32x3", compose_affine=True)
run("t3 layerwise, fp32 LU product", precision="tf32x3", compose_affine=False)
