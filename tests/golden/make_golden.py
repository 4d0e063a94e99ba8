"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

Runs the reference's OWN, unmodified in-tree classes
(`/root/reference/src/nf4ad/flows.py:NonUSFlow`,
`/root/reference/src/nf4ad/transforms.py:MaskedAffineCoupling`) on top of the
oracle's `src.usflows` / `pyro` shim, in fp64, on seeded weights and inputs, and
freezes state_dict + inputs + outputs.  The USFlows layers underneath are the
oracle's restatement (upstream is absent -> "parity unpinned" for those), so
these vectors pin (a) the reference's in-tree arithmetic and stack order and
(b) the oracle's own semantics against regressions.

    python tests/golden/make_golden.py          # needs /root/reference

`/root/reference` is never read by the tests themselves.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

REF_SRC = "/root/reference/src"


class MLP(torch.nn.Module):
    """Same shape of conditioner as tests/conftest.py:111-121 of the reference."""

    def __init__(self, in_dim, hidden, out_dim):
        super().__init__()
        dims = [in_dim] + list(hidden)
        mods = []
        for i in range(len(hidden)):
            mods += [torch.nn.Linear(dims[i], dims[i + 1]), torch.nn.ReLU()]
        mods.append(torch.nn.Linear(dims[-1], out_dim))
        self.net = torch.nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)


CASES = {
    # name: (flow kind, D, K, conditioner spec, kwargs, base, B, last-layer gain)
    "nonus_d8_k2_conj_hh1": dict(kind="NonUSFlow", D=8, K=2, cond=("mlp", [16]), base="normal", B=7, gain=0.5,
                                 kw=dict(affine_conjugation=True, prior_scale=1.0)),
    "nonus_d6_k3_noconj_hh0_densenn_laplace": dict(kind="NonUSFlow", D=6, K=3, cond=("densenn2", [16, 8]),
                                                   base="laplace", B=5, gain=0.5,
                                                   kw=dict(affine_conjugation=False, householder=0)),
    "nonus_d5_k2_lu2_hh2": dict(kind="NonUSFlow", D=5, K=2, cond=("mlp", [12, 12]), base="normal", B=4, gain=1.0,
                                kw=dict(affine_conjugation=True, lu_transform=2, householder=2)),
    "nonus_d32_k3_fixture": dict(kind="NonUSFlow", D=32, K=3, cond=("mlp", [128]), base="normal", B=5, gain=0.25,
                                 kw=dict(affine_conjugation=True, prior_scale=1.0)),
    "us_d8_k2_densenn_trainable_normal": dict(kind="USFlow", D=8, K=2, cond=("densenn1", [16]), base="usnormal",
                                              B=6, gain=1.0,
                                              kw=dict(affine_conjugation=True, householder=0, prior_scale=1.0)),
}


def build(case, NonUSFlow, USFlow, dist, DenseNN, USNormal):
    D = case["D"]
    kind, hidden = case["cond"]
    if kind == "mlp":
        ccls, cargs = MLP, dict(in_dim=D, hidden=hidden, out_dim=2 * D)
    elif kind == "densenn2":
        ccls, cargs = DenseNN, dict(input_dim=D, hidden_dims=hidden, param_dims=[D, D])
    else:
        ccls, cargs = DenseNN, dict(input_dim=D, hidden_dims=hidden, param_dims=[D])
    if case["base"] == "normal":
        base = dist.Normal(torch.zeros(D), torch.ones(D))
    elif case["base"] == "laplace":
        base = dist.Laplace(torch.zeros(D), torch.ones(D))
    else:
        base = USNormal(torch.zeros(D), torch.tensor(1.5))
    cls = NonUSFlow if case["kind"] == "NonUSFlow" else USFlow
    return cls(base_distribution=base, in_dims=[D], coupling_blocks=case["K"],
               conditioner_cls=ccls, conditioner_args=cargs, **case["kw"])


def tame(flow, gain):
    """Scale the last conditioner layer so activations stay O(1) (as in a trained flow)."""
    with torch.no_grad():
        for layer in flow.layers:
            cond = getattr(layer, "conditioner", None)
            if cond is None:
                continue
            last = [m for m in cond.modules() if isinstance(m, torch.nn.Linear)][-1]
            last.weight.mul_(gain)
            last.bias.mul_(gain)


def main():
    assert os.path.isdir(REF_SRC), "golden vectors are generated against /root/reference"
    with oracle.activated():
        sys.path.insert(1, REF_SRC)
        try:
            from nf4ad.flows import NonUSFlow          # the reference's own class
            import nf4ad.transforms as ref_t
            assert ref_t.__file__.startswith(REF_SRC)
            from src.usflows.flows import USFlow
            from src.usflows.distributions import Normal as USNormal
            import pyro.distributions as dist
            from pyro.nn import DenseNN
            for idx, (name, case) in enumerate(CASES.items()):
                torch.manual_seed(1000 + idx)
                flow = build(case, NonUSFlow, USFlow, dist, DenseNN, USNormal)
                tame(flow, case["gain"])
                # make the data-independent terms non-trivial
                with torch.no_grad():
                    for p_name, p in flow.named_parameters():
                        if p_name.endswith("scale") and p.dim() == 1:
                            p.copy_(0.5 + torch.rand_like(p))
                        if p_name.endswith("U_raw"):
                            p.diagonal().copy_((0.6 + 0.8 * torch.rand(p.shape[0])) *
                                               torch.where(torch.rand(p.shape[0]) < 0.3, -1.0, 1.0))
                flow = flow.double()
                x = torch.randn(case["B"], case["D"], dtype=torch.float64) * 1.3 + 0.2
                x.requires_grad_(True)
                lp = flow.log_prob(x)
                loss = -lp.mean()
                grads = torch.autograd.grad(loss, [x] + list(flow.parameters()), allow_unused=True)
                with torch.no_grad():
                    z = flow.backward(x)
                    x_rt = z
                    for layer in flow.layers:
                        x_rt = layer.forward(x_rt)
                    zs = torch.randn(case["B"], case["D"], dtype=torch.float64)
                    xs = zs
                    for layer in flow.layers:
                        xs = layer.forward(xs)
                out = dict(
                    case={k: v for k, v in case.items()}, seed=1000 + idx, dtype="float64",
                    state_dict={k: v.detach().clone() for k, v in flow.state_dict().items()},
                    x=x.detach().clone(), log_prob=lp.detach().clone(), latent=z.clone(),
                    roundtrip=x_rt.clone(), z_sample=zs, x_from_z=xs.clone(),
                    grad_x=grads[0].clone(),
                    grad_params={n: (g.clone() if g is not None else None)
                                 for (n, _), g in zip(flow.named_parameters(), grads[1:])},
                    generator="reference nf4ad.flows.NonUSFlow + nf4ad.transforms.MaskedAffineCoupling"
                              if case["kind"] == "NonUSFlow" else "oracle USFlow (unpinned)",
                )
                path = os.path.join(HERE, name + ".pt")
                torch.save(out, path)
                print(f"{name}: log_prob[:3]={lp[:3].tolist()} roundtrip_err={float((x_rt - x).abs().max()):.2e} "
                      f"-> {os.path.getsize(path)} bytes")
        finally:
            sys.path.remove(REF_SRC)
            for k in [k for k in sys.modules if k.split('.')[0] == "nf4ad"]:
                del sys.modules[k]


if __name__ == "__main__":
    main()
