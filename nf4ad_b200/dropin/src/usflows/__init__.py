"""`src.usflows` drop-in: the B200-native classes under the names nf4ad imports
(`/root/reference/src/nf4ad/flows.py:6-17`, `transforms.py:5`)."""
from . import distributions, flows, networks, transforms  # noqa: F401
