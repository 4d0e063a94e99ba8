"""nf4ad_b200 -- B200-native (sm_100a) implementation of the uniformly-scaling-flow density path that
nf4ad drives through USFlows: `Flow.log_prob / sample / forward / backward`.

Layout
  csrc/     hand-written CUDA (tcgen05 bf16 GEMM with fused layer epilogues, fp32 SIMT kernels) + C ABI
  _lib.py   ctypes binding of libusflow_b200.so (include/usflow_b200.h)
  ops.py    tensor-level wrappers and the autograd Functions of the training path
  transforms.py / flows.py / distributions.py / nn.py   the USFlows module API (same names)
  stack.py  stack compiler (weights -> packed descriptor array) for the fused path
  parallel.py  batch-sharded scoring and data-parallel training (torch.distributed / NCCL)
  dropin/   import roots `src.usflows.*` and `pyro.*` for unmodified nf4ad code
"""
import os
import sys

from ._lib import LIB_PATH, USFError  # noqa: F401

DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropin(with_pyro=None):
    """Puts the drop-in import roots on `sys.path` so unmodified nf4ad code (`from src.usflows...`,
    `from pyro import distributions`) resolves to this package.  `pyro` is only shimmed when the real
    pyro-ppl is not importable (or when `with_pyro=True`)."""
    if DROPIN not in sys.path:
        sys.path.insert(0, DROPIN)
    if with_pyro is None:
        import importlib.util
        try:
            with_pyro = importlib.util.find_spec("pyro") is None or \
                os.path.dirname(importlib.util.find_spec("pyro").origin or "").startswith(DROPIN)
        except (ImportError, ValueError):
            with_pyro = True
    pyro_dir = os.path.join(DROPIN, "_pyro")
    if with_pyro and pyro_dir not in sys.path:
        sys.path.insert(0, pyro_dir)
    return DROPIN


def namespace():
    """Classes under the names the parity tests use (mirrors `oracle.load()`)."""
    import types
    import torch.distributions as tdist
    from . import distributions, flows, nn, transforms
    return types.SimpleNamespace(
        dist=tdist, DenseNN=nn.DenseNN, ConditionalDenseNN=nn.ConditionalDenseNN, ConvNet=nn.ConvNet, Flow=flows.Flow, USFlow=flows.USFlow, NonUSFlow=flows.NonUSFlow,
        transforms=transforms, Normal=distributions.Normal,
        MaskedAffineCoupling=transforms.MaskedAffineCoupling, MaskedCoupling=transforms.MaskedCoupling)
