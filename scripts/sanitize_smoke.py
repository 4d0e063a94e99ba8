"""Small end-to-end exercise for compute-sanitizer: fused scoring (bf16 + fp32), latent/sample directions, one
fp32 and one mixed-precision training step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["USF_GRAPHS"] = "0"
import nf4ad_b200
from _cases import build_flow, tame
P = nf4ad_b200.namespace()
torch.manual_seed(0)
for D, K, hid, rows in ((64, 2, [128, 128], 600), (80, 2, [48], 300)):
    flow = build_flow(P, "NonUSFlow", D, K, ("mlp", hid), affine_conjugation=True, prior_scale=1.0)
    tame(flow, 0.25)
    flow = flow.to("cuda").eval()
    x = torch.randn(rows, D, device="cuda")
    with torch.no_grad():
        for prec in ("bf16", "fp32"):
            flow.precision = prec
            lp = flow.log_prob(x); z = flow.backward(x); xs = flow.sample([33]); lp16 = flow.log_prob(x[:16])
            torch.cuda.synchronize()
            assert torch.isfinite(lp).all() and torch.isfinite(z).all() and torch.isfinite(xs).all()
    flow.train()
    for prec in ("fp32", "bf16"):
        flow.precision = prec
        flow.zero_grad(set_to_none=True)
        loss = -flow.log_prob(x[:64]).mean(); loss.backward()
        torch.cuda.synchronize()
        assert torch.isfinite(loss)
    print("ok", D, K, hid)
