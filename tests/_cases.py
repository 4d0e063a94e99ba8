"""Helpers shared by the parity tests: build the same flow from either the
oracle namespace or the product namespace, load golden fixtures."""
import glob
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class MLP(torch.nn.Module):
    """Linear/ReLU chain wrapped in `.net`, as the reference's test conditioners
    (`tests/conftest.py:111-121`, `examples/adbench_flow_example.py:20-32`)."""

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


def golden_names():
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=True)


def golden_r2_names():
    """Round-2 fixtures (tests/golden/r2, made by tests/golden/make_golden_r2.py): image-shaped events, the softplus scale
    activation, the additive-through-affine conditioner form, a conditional (context) flow."""
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, "r2", "*.pt")))


def load_golden_r2(name):
    return torch.load(os.path.join(GOLDEN_DIR, "r2", name + ".pt"), weights_only=True)


def flow_from_r2_case(ns, case, **over):
    """Rebuilds the flow of a round-2 golden case from either namespace (incl. its post-construction tweaks)."""
    flow = build_flow(ns, case["kind"], case["D"], case["K"], tuple(case["cond"]), base=case["base"],
                      **{**case["kw"], **over})
    act = case.get("scale_activation")
    if act:
        for layer in flow.layers:
            if hasattr(layer, "scale_activation"):
                layer.scale_activation = act
    return flow


def build_flow(ns, kind, D, K, cond, base="normal", device="cpu", dtype=torch.float32, **kw):
    """`ns` exposes NonUSFlow, USFlow, dist, DenseNN, Normal (oracle.load() or the product's namespace).
    `D` is an int (flat event, in_dims=[D]) or an event shape such as [4, 8, 8] (image-shaped)."""
    ckind, hidden = cond
    in_dims = [D] if isinstance(D, int) else list(D)
    if not isinstance(D, int):
        C = in_dims[0]
        if ckind == "conv":          # 2C output channels: `[s | t]` split along dim 1 (transforms.py:52-55)
            ccls, cargs = ns.ConvNet, dict(in_dims=in_dims, c_hidden=hidden, c_out=2 * C)
        elif ckind == "conv_add":
            ccls, cargs = ns.ConvNet, dict(in_dims=in_dims, c_hidden=hidden, c_out=C)
        else:
            raise ValueError(ckind)
        shape = tuple(in_dims)
        bd = {"normal": ns.dist.Normal, "laplace": ns.dist.Laplace}[base](torch.zeros(shape, dtype=dtype),
                                                                        torch.ones(shape, dtype=dtype))
        cls = ns.NonUSFlow if kind == "NonUSFlow" else ns.USFlow
        return cls(base_distribution=bd, in_dims=in_dims, coupling_blocks=K, conditioner_cls=ccls,
                   conditioner_args=cargs, device=device, **kw)
    if ckind == "cond2":             # conditional conditioner (context), affine (s, t) tuple
        ccls, cargs = ns.ConditionalDenseNN, dict(input_dim=D, context_dim=1, hidden_dims=hidden, param_dims=[D, D])
    elif ckind == "cond1":
        ccls, cargs = ns.ConditionalDenseNN, dict(input_dim=D, context_dim=1, hidden_dims=hidden, param_dims=[D])
    elif ckind == "mlp":
        ccls, cargs = MLP, dict(in_dim=D, hidden=hidden, out_dim=2 * D)
    elif ckind == "mlp_add":
        ccls, cargs = MLP, dict(in_dim=D, hidden=hidden, out_dim=D)
    elif ckind == "densenn2":
        ccls, cargs = ns.DenseNN, dict(input_dim=D, hidden_dims=hidden, param_dims=[D, D])
    elif ckind == "densenn1":
        ccls, cargs = ns.DenseNN, dict(input_dim=D, hidden_dims=hidden, param_dims=[D])
    elif ckind in ("cond2", "cond1"):
        pass
    else:
        raise ValueError(ckind)
    if base == "normal":
        bd = ns.dist.Normal(torch.zeros(D, dtype=dtype), torch.ones(D, dtype=dtype))
    elif base == "laplace":
        bd = ns.dist.Laplace(torch.zeros(D, dtype=dtype), torch.ones(D, dtype=dtype))
    elif base == "usnormal":
        bd = ns.Normal(torch.zeros(D, dtype=dtype), torch.tensor(1.5, dtype=dtype))
    else:
        raise ValueError(base)
    cls = ns.NonUSFlow if kind == "NonUSFlow" else ns.USFlow
    flow = cls(base_distribution=bd, in_dims=[D], coupling_blocks=K,
               conditioner_cls=ccls, conditioner_args=cargs, device=device, **kw)
    return flow


def flow_from_case(ns, case, **over):
    return build_flow(ns, case["kind"], case["D"], case["K"], tuple(case["cond"]),
                      base=case["base"], **{**case["kw"], **over})


def tame(flow, gain):
    """Scale each conditioner's last Linear (trained-flow-like activations)."""
    with torch.no_grad():
        for layer in flow.layers:
            cond = getattr(layer, "conditioner", None)
            if cond is None:
                continue
            last = [m for m in cond.modules() if isinstance(m, (torch.nn.Linear, torch.nn.Conv2d))][-1]
            last.weight.mul_(gain)
            last.bias.mul_(gain)


def randomize_constants(flow, seed=0):
    """Make the data-independent log-det terms non-trivial (scale, diag U with mixed signs)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in flow.named_parameters():
            if name.endswith("scale") and p.dim() == 1:
                p.copy_((0.5 + torch.rand(p.shape, generator=g)).to(p))
            if name.endswith("U_raw"):
                n = p.shape[0]
                d = (0.6 + 0.8 * torch.rand(n, generator=g)) * torch.where(
                    torch.rand(n, generator=g) < 0.3, -1.0, 1.0)
                p.diagonal().copy_(d.to(p))
