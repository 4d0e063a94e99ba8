"""`src.usflows.distributions.Normal` on the B200 path: isotropic normal with a trainable scalar
`scale_unconstrained` (softplus), event shape = `loc.shape`
(`/root/reference/experiments/gmm/gaussian_mixture_standart_base.yaml:77-83`,
`/root/reference/scripts/gmm_eval_usflows.py:61,105`).  Its log-density is evaluated by the fused
base-density epilogue / `usf_base_logprob` through `Flow`; the methods below are the generic API."""
import math

import torch
from torch.distributions import constraints


class Normal(torch.nn.Module, torch.distributions.Distribution):
    arg_constraints = {}
    support = constraints.real_vector
    has_rsample = True

    def __init__(self, loc, scale, device="cpu", *args, **kwargs):
        torch.nn.Module.__init__(self)
        loc = torch.as_tensor(loc).to(device)
        scale = torch.as_tensor(scale, dtype=loc.dtype).to(device)
        torch.distributions.Distribution.__init__(
            self, batch_shape=torch.Size(), event_shape=loc.shape, validate_args=False)
        self.register_buffer("loc", loc)
        self.scale_unconstrained = torch.nn.Parameter(scale + torch.log(-torch.expm1(-scale)))

    def __hash__(self):
        return torch.nn.Module.__hash__(self)

    @property
    def scale(self):
        return torch.nn.functional.softplus(self.scale_unconstrained)

    def log_prob(self, value):
        if value.is_cuda and value.dim() == 2:
            from . import ops
            return ops.BaseLogProbFn.apply(value, self.loc, self.scale, 0)
        s = self.scale
        z = (value - self.loc) / s
        lp = -0.5 * z * z - torch.log(s) - 0.5 * math.log(2 * math.pi)
        n = len(self.event_shape)
        return lp.sum(dim=tuple(range(-n, 0))) if n > 0 else lp

    def rsample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self.loc.shape
        eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        return self.loc + eps * self.scale

    def sample(self, sample_shape=torch.Size()):
        with torch.no_grad():
            return self.rsample(sample_shape)
