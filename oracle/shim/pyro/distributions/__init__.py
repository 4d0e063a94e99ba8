"""ORACLE (test infrastructure): `pyro.distributions` surface used by nf4ad.

pyro.distributions re-exports torch.distributions and adds `TransformModule`
(a `Transform` that is also an `nn.Module`).  Call sites in the reference:
`flows.py:3,37`, `transforms.py:6,37-38` (`dist.constraints.real_vector`),
`tests/conftest.py:105-108` (`dist.Normal`), YAMLs (`Laplace`, `Uniform`).
"""
import torch
from torch.distributions import *  # noqa: F401,F403
from torch.distributions import constraints, transforms  # noqa: F401
from torch.distributions import (  # noqa: F401
    Distribution, Independent, Laplace, Normal, TransformedDistribution, Uniform,
)
from torch.distributions.transforms import Transform


class TransformModule(Transform, torch.nn.Module):
    """A bijector with learnable parameters (pyro's TransformModule, restated)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def __hash__(self):
        return torch.nn.Module.__hash__(self)

    def __eq__(self, other):
        return self is other
