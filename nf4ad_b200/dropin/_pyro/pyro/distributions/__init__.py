from torch.distributions import *  # noqa: F401,F403
from torch.distributions import constraints, transforms  # noqa: F401
from torch.distributions import (  # noqa: F401
    Distribution, Independent, Laplace, Normal, TransformedDistribution, Uniform,
)
from nf4ad_b200.transforms import TransformModule  # noqa: F401
