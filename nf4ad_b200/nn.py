"""`pyro.nn.DenseNN` stand-in: Linear(input_dim,h0)-ReLU-...-Linear(h_last, sum(param_dims)); tensor
output for one param group, tuple otherwise
(`/root/reference/experiments/gmm/gaussian_mixture_standart_base.yaml:66-72`).  `Flow` recognises it as a
Linear/ReLU chain and runs it on the library's GEMM kernels."""
import torch


class DenseNN(torch.nn.Module):
    def __init__(self, input_dim, hidden_dims, param_dims=(1, 1), nonlinearity=None):
        super().__init__()
        self.input_dim = int(input_dim)
        self.hidden_dims = [int(h) for h in hidden_dims]
        self.param_dims = [int(p) for p in param_dims]
        self.count_params = len(self.param_dims)
        self.output_multiplier = sum(self.param_dims)
        dims = [self.input_dim] + self.hidden_dims
        mods = [torch.nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
        mods.append(torch.nn.Linear(dims[-1], self.output_multiplier))
        self.layers = torch.nn.ModuleList(mods)
        self.f = nonlinearity if nonlinearity is not None else torch.nn.ReLU()

    def forward(self, x):
        h = x
        for layer in self.layers[:-1]:
            h = self.f(layer(h))
        h = self.layers[-1](h)
        if self.count_params == 1:
            return h
        outs, o = [], 0
        for p in self.param_dims:
            outs.append(h[..., o:o + p])
            o += p
        return tuple(outs)
