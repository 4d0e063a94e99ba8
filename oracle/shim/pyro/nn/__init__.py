"""ORACLE (test infrastructure): `pyro.nn.DenseNN` restated.

Published pyro semantics: `DenseNN(input_dim, hidden_dims, param_dims=[1,1],
nonlinearity=ReLU())` is Linear(input_dim,h0) - act - ... - Linear(h_last,
sum(param_dims)); the output is returned whole when `len(param_dims)==1`
and otherwise split along the last dim into a tuple.  Reference call sites:
`experiments/gmm/gaussian_mixture_standart_base.yaml:66-72` (`param_dims:[D]`,
additive) and `experiments/synthetic/gaussian_mixture_non_usflow.yaml:66-73`
(`param_dims:[D,D]`, affine (s,t) tuple -> `nf4ad/transforms.py:42-43`).
"""
import torch


class DenseNN(torch.nn.Module):
    def __init__(self, input_dim, hidden_dims, param_dims=(1, 1), nonlinearity=None):
        super().__init__()
        self.input_dim = int(input_dim)
        self.hidden_dims = [int(h) for h in hidden_dims]
        self.param_dims = [int(p) for p in param_dims]
        self.count_params = len(self.param_dims)
        self.output_multiplier = sum(self.param_dims)
        ends = torch.cumsum(torch.tensor(self.param_dims), dim=0)
        starts = torch.cat((torch.zeros(1).type_as(ends), ends[:-1]))
        self.param_slices = [slice(int(s), int(e)) for s, e in zip(starts, ends)]
        dims = [self.input_dim] + self.hidden_dims
        layers = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
        layers.append(torch.nn.Linear(dims[-1], self.output_multiplier))
        self.layers = torch.nn.ModuleList(layers)
        self.f = nonlinearity if nonlinearity is not None else torch.nn.ReLU()

    def forward(self, x):
        h = x
        for layer in self.layers[:-1]:
            h = self.f(layer(h))
        h = self.layers[-1](h)
        if self.count_params == 1:
            return h
        return tuple(h[..., s] for s in self.param_slices)


class ConditionalDenseNN(DenseNN):
    """Published pyro semantics: `ConditionalDenseNN(input_dim, context_dim, hidden_dims, param_dims)` is a DenseNN over
    `torch.cat([context, x], dim=-1)` with the context broadcast over x's batch dims.  The conditioner form of USFlows'
    soft training [RECALL]; reached through `conditioner(x_masked, context)` (`nf4ad/transforms.py:71-74`)."""

    def __init__(self, input_dim, context_dim, hidden_dims, param_dims=(1, 1), nonlinearity=None):
        super().__init__(int(input_dim) + int(context_dim), hidden_dims, param_dims, nonlinearity)
        self.input_dim = int(input_dim)
        self.context_dim = int(context_dim)

    def forward(self, x, context):
        context = context.expand(x.shape[:-1] + (context.shape[-1],))
        return super().forward(torch.cat([context, x], dim=-1))
