"""Minimal `pyro` stand-in used only when pyro-ppl is not installed: the surface nf4ad touches is
`pyro.distributions` (= torch.distributions + TransformModule) and `pyro.nn.DenseNN`."""
from . import distributions, nn  # noqa: F401
