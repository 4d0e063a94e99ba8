"""Round-2 golden fixtures under tests/golden/r2/ (run in the BUILD container only; needs /root/reference).

Same method as make_golden.py -- the reference's OWN unmodified `nf4ad.flows.NonUSFlow` /
`nf4ad.transforms.MaskedAffineCoupling` on the oracle's `src.usflows` / `pyro` shim, fp64, seeded -- for the parts of
the path the first set does not reach:

  * image-shaped events `in_dims=[C, H, W]`: N-D checkerboard and channel masks (`flows.py:127-145`), per-pixel C x C
    LU blocks, a conv conditioner whose 2C output channels are split along dim 1 (`transforms.py:52-55`), the
    `view(B, -1).sum` log-det (`transforms.py:137-140`);
  * `scale_activation="softplus"` incl. its +1e-6 / +1e-12 inconsistency (`transforms.py:83-85,107,132`);
  * a conditioner that returns ONE tensor of x's shape = additive shift through the affine class (`transforms.py:46-49`);
  * a conditional flow: `conditioner(x_masked, context)` (`transforms.py:71-74`) with a per-sample context.

    python tests/golden/make_golden_r2.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from _cases import flow_from_r2_case, randomize_constants, tame  # noqa: E402

CASES = {
    "nonus_img_c4_8x8_k2_conv": dict(kind="NonUSFlow", D=[4, 8, 8], K=2, cond=("conv", [8]), base="normal", B=3, gain=0.25,
                                     kw=dict(affine_conjugation=True, prior_scale=1.0)),
    "nonus_img_c3_4x4_k3_channel_laplace": dict(kind="NonUSFlow", D=[3, 4, 4], K=3, cond=("conv", [6, 6]), base="laplace",
                                                B=4, gain=0.5, kw=dict(affine_conjugation=False, masktype="channel",
                                                                       householder=2)),
    "nonus_d8_k2_softplus": dict(kind="NonUSFlow", D=8, K=2, cond=("mlp", [16]), base="normal", B=6, gain=0.25,
                                 scale_activation="softplus", kw=dict(affine_conjugation=True, prior_scale=1.0)),
    "nonus_d7_k2_same_shape_params": dict(kind="NonUSFlow", D=7, K=2, cond=("mlp_add", [12]), base="normal", B=5, gain=1.0,
                                          kw=dict(affine_conjugation=True)),
    "nonus_d8_k3_context": dict(kind="NonUSFlow", D=8, K=3, cond=("cond2", [16, 16]), base="normal", B=6, gain=0.5,
                                context=True, kw=dict(affine_conjugation=True, prior_scale=1.0)),
}


def main():
    assert oracle.ref_available() and os.path.isdir("/root/reference"), "needs the reference checkout (oracle/make_ref.py)"
    R = oracle.load_ref()             # the reference's own NonUSFlow / MaskedAffineCoupling over the oracle shim
    os.makedirs(os.path.join(HERE, "r2"), exist_ok=True)
    for idx, (name, case) in enumerate(CASES.items()):
        torch.manual_seed(2000 + idx)
        flow = flow_from_r2_case(R, case)
        assert type(flow).__module__ == "nf4ad.flows" and type(flow.layers[1]).__module__ == "nf4ad.transforms"
        tame(flow, case["gain"])
        randomize_constants(flow, 2000 + idx)
        with torch.no_grad():             # image-shaped Scale: non-trivial too
            for n, p in flow.named_parameters():
                if n.endswith("scale") and p.dim() > 1:
                    p.copy_(0.5 + torch.rand_like(p))
        flow = flow.double()
        shape = [case["D"]] if isinstance(case["D"], int) else list(case["D"])
        x = torch.randn(case["B"], *shape, dtype=torch.float64) * 1.3 + 0.2
        x.requires_grad_(True)
        ctx = (0.05 + 0.3 * torch.rand(case["B"], 1, dtype=torch.float64)) if case.get("context") else None
        lp = flow.log_prob(x, ctx) if ctx is not None else flow.log_prob(x)
        grads = torch.autograd.grad(-lp.mean(), [x] + list(flow.parameters()), allow_unused=True)
        with torch.no_grad():
            z = flow.backward(x, ctx) if ctx is not None else flow.backward(x)
            zs = torch.randn(case["B"], *shape, dtype=torch.float64)
            xs = flow.latent_to_data(zs, ctx) if ctx is not None else flow.latent_to_data(zs)
            rt = flow.latent_to_data(z, ctx) if ctx is not None else flow.latent_to_data(z)
            quirk = flow.log_abs_det_jacobian(x) if ctx is None else None       # flows.py:160-169 (x never advanced)
        out = dict(case=case, seed=2000 + idx, dtype="float64",
                   state_dict={k: v.detach().clone() for k, v in flow.state_dict().items()},
                   x=x.detach().clone(), context=ctx, log_prob=lp.detach().clone(), latent=z.clone(), z_sample=zs,
                   x_from_z=xs.clone(), grad_x=grads[0].clone(),
                   ladj_quirk=None if quirk is None else torch.as_tensor(quirk).detach().clone(),
                   grad_params={n: (g.clone() if g is not None else None)
                                for (n, _), g in zip(flow.named_parameters(), grads[1:])},
                   generator="reference nf4ad.flows.NonUSFlow + nf4ad.transforms.MaskedAffineCoupling (oracle/_ref)")
        path = os.path.join(HERE, "r2", name + ".pt")
        torch.save(out, path)
        print(f"{name}: log_prob[:3]={lp[:3].tolist()} roundtrip_err={float((rt - x).abs().max()):.2e} "
              f"-> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
