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


class ConditionalDenseNN(DenseNN):
    """`pyro.nn.ConditionalDenseNN(input_dim, context_dim, hidden_dims, param_dims)`: a DenseNN over
    `cat([context, x], -1)` -- the conditioner form of USFlows' soft training, where the context is the per-sample noise
    level (`Flow.fit(soft_training=True)`; `MaskedAffineCoupling` passes it on as `conditioner(x_masked, context)`,
    `/root/reference/src/nf4ad/transforms.py:71-74`).  `Flow` recognises it as a Linear/ReLU chain whose first layer has
    `context_dim` extra input columns and carries the context through the fused launch chain."""

    def __init__(self, input_dim, context_dim, hidden_dims, param_dims=(1, 1), nonlinearity=None):
        super().__init__(int(input_dim) + int(context_dim), hidden_dims, param_dims, nonlinearity)
        self.input_dim = int(input_dim)
        self.context_dim = int(context_dim)

    def forward(self, x, context):
        context = context.expand(x.shape[:-1] + (context.shape[-1],))
        return super().forward(torch.cat([context, x], dim=-1))


class ConvNet(torch.nn.Module):
    """`src.usflows.networks.ConvNet(in_dims, c_hidden, c_out, nonlinearity)` stand-in for image-shaped events
    (`/root/reference/experiments/MVTec/mvtec_trainable_encoder_us.yaml:67-74`): 3x3 convolutions that keep the spatial
    size, `in_dims[0] -> c_hidden... -> c_out` channels (c_out = C for USFlow, 2C for `NonUSFlow`'s `[s | t]` split along
    dim 1).  Upstream's exact layer list cannot be pinned from the reference tree; to `Flow` this is an opaque conditioner
    (cuDNN convolutions, evaluated as given) feeding the coupling kernel."""

    def __init__(self, in_dims, c_hidden, c_out=None, nonlinearity=None, kernel_size=3):
        super().__init__()
        c_in = int(in_dims[0])
        hidden = [int(c) for c in (c_hidden if isinstance(c_hidden, (list, tuple)) else [c_hidden])]
        c_out = c_in if c_out is None else int(c_out)
        chans = [c_in] + hidden
        mods = []
        for i in range(len(hidden)):
            mods += [torch.nn.Conv2d(chans[i], chans[i + 1], kernel_size, padding=kernel_size // 2),
                     nonlinearity if nonlinearity is not None else torch.nn.ReLU()]
        mods.append(torch.nn.Conv2d(chans[-1], c_out, kernel_size, padding=kernel_size // 2))
        self.net = torch.nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)
