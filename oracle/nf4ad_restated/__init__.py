"""ORACLE (test infrastructure): restatement of the reference's IN-TREE hot-path code.

`MaskedAffineCoupling` follows `/root/reference/src/nf4ad/transforms.py:8-149`,
`NonUSFlow` follows `/root/reference/src/nf4ad/flows.py:27-169`.  Unlike the
`src.usflows` shim these two ARE pinned: `tests/golden/make_golden.py` runs the
reference's own unmodified classes (imported from `/root/reference/src` in the
build container) and `tests/test_oracle_golden.py` checks this restatement
reproduces the frozen outputs bit-for-bit in fp64.
Requires `oracle/shim` on `sys.path` (see `oracle/__init__.py:activate`).
"""
from .transforms import MaskedAffineCoupling  # noqa: F401
from .flows import NonUSFlow, Flow  # noqa: F401
