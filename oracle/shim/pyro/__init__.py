"""ORACLE (test infrastructure, not product code).

Minimal stand-in for the `pyro` import root that the reference imports
(`/root/reference/src/nf4ad/flows.py:3`, `transforms.py:6`,
`tests/conftest.py:9`).  pyro-ppl is an un-pinned, un-vendored dependency of
the reference and is not installable here (no network); the surface nf4ad
touches is a thin re-export of `torch.distributions`, which is what this
package provides.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline legs may import anything under `oracle/`.
"""
from . import distributions, nn  # noqa: F401
