"""ORACLE -- test infrastructure only.

CPU restatement of the density path nf4ad drives through USFlows
(`Flow.log_prob / sample / forward / backward`).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this
package; the product (`nf4ad_b200/`) never does.

Parity status: the `src.usflows` / `pyro` shim is PARITY UNPINNED (upstream
sources absent, see `oracle/shim/src/usflows/__init__.py`); the restatement of
the reference's in-tree classes (`oracle/nf4ad_restated`) is pinned against the
reference's own code through `tests/golden/`.
"""
import contextlib
import os
import sys

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
_SHIM_ROOTS = ("pyro", "src")


@contextlib.contextmanager
def activated():
    """Temporarily make `pyro` / `src.usflows` resolve to the oracle shim.

    The product ships modules under the same import names (its drop-in), so the
    two are kept apart: inside the context the shim owns the names, and on exit
    the previous `sys.modules` entries are restored.
    """
    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _SHIM_ROOTS or k.startswith("oracle.nf4ad_restated")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, SHIM)
    try:
        yield
    finally:
        sys.path.remove(SHIM)
        for k in [k for k in sys.modules if k.split(".")[0] in _SHIM_ROOTS]:
            del sys.modules[k]
        sys.modules.update(saved)


def load():
    """Import the oracle once and return a namespace of its classes.

    The classes stay usable after the shim is deactivated (they hold references
    to their own modules), so tests can build oracle and product models side by
    side even though both define `src.usflows.*`.
    """
    global _NS
    try:
        return _NS
    except NameError:
        pass
    import types
    with activated():
        import pyro.distributions as pdist
        import pyro.nn as pnn
        import src.usflows.flows as uflows
        import src.usflows.transforms as utransforms
        import src.usflows.distributions as udist
        from oracle import nf4ad_restated
        _NS = types.SimpleNamespace(
            dist=pdist, DenseNN=pnn.DenseNN, Flow=uflows.Flow, USFlow=uflows.USFlow,
            transforms=utransforms, Normal=udist.Normal,
            NonUSFlow=nf4ad_restated.NonUSFlow,
            MaskedAffineCoupling=nf4ad_restated.MaskedAffineCoupling,
            MaskedCoupling=utransforms.MaskedCoupling,
        )
    return _NS
