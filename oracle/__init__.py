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
_SHIM_ROOTS = ("pyro", "src", "nf4ad")
_OWN = {}


@contextlib.contextmanager
def activated(extra_path=None):
    """Temporarily make `pyro` / `src.usflows` resolve to the oracle shim.

    The product ships modules under the same import names (its drop-in), so the
    two are kept apart: inside the context the shim owns the names, and on exit
    the previous `sys.modules` entries are restored.
    """
    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _SHIM_ROOTS or k.startswith("oracle.nf4ad_restated")}
    for k in saved:
        del sys.modules[k]
    sys.modules.update(_OWN)          # the shim's modules from earlier activations: ONE copy of every oracle class
    sys.path.insert(0, SHIM)
    if extra_path:
        sys.path.insert(0, extra_path)
    try:
        yield
    finally:
        sys.path.remove(SHIM)
        if extra_path:
            sys.path.remove(extra_path)
        for k in [k for k in sys.modules if k.split(".")[0] in _SHIM_ROOTS or k.startswith("oracle.nf4ad_restated")]:
            _OWN[k] = sys.modules.pop(k)
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
        import src.usflows.networks as unets
        from oracle import nf4ad_restated
        _NS = types.SimpleNamespace(
            dist=pdist, DenseNN=pnn.DenseNN, ConditionalDenseNN=pnn.ConditionalDenseNN, ConvNet=unets.ConvNet,
            Flow=uflows.Flow, USFlow=uflows.USFlow,
            transforms=utransforms, Normal=udist.Normal,
            NonUSFlow=nf4ad_restated.NonUSFlow,
            MaskedAffineCoupling=nf4ad_restated.MaskedAffineCoupling,
            MaskedCoupling=utransforms.MaskedCoupling,
        )
    return _NS


def ref_available():
    from oracle import make_ref
    return make_ref.available()


def load_ref():
    """The reference's OWN classes -- `nf4ad.flows.NonUSFlow`, `nf4ad.transforms.MaskedAffineCoupling`,
    `nf4ad.adbench_wrapper.ADBenchFlow`, unmodified, from the `oracle/_ref` snapshot -- on top of the oracle's
    `src.usflows` / `pyro` shim; same namespace layout as `load()`.  This is the CPU baseline of kind "reference"."""
    global _REF
    try:
        return _REF
    except NameError:
        pass
    import types
    from oracle import make_ref
    base = load()
    root = make_ref.unpack()
    with activated(os.path.join(root, "src")):
        import nf4ad.flows as rflows
        import nf4ad.transforms as rtransforms
        import nf4ad.adbench_wrapper as rwrap
        _REF = types.SimpleNamespace(
            dist=base.dist, DenseNN=base.DenseNN, ConditionalDenseNN=base.ConditionalDenseNN, ConvNet=base.ConvNet,
            Flow=rflows.Flow, USFlow=base.USFlow, transforms=base.transforms,
            Normal=base.Normal, NonUSFlow=rflows.NonUSFlow, MaskedAffineCoupling=rtransforms.MaskedAffineCoupling,
            MaskedCoupling=base.MaskedCoupling, ADBenchFlow=rwrap.ADBenchFlow, root=root,
        )
    return _REF
