"""ORACLE (test infrastructure) -- `src.usflows.transforms` restated on CPU.  PARITY UNPINNED.

The bijector layers nf4ad imports (`/root/reference/src/nf4ad/flows.py:9-17`,
`/root/reference/src/nf4ad/transforms.py:5`).  Upstream source is absent, so
each class states the contract inferred from its nf4ad call site [INFER] and
the recalled upstream semantics [RECALL] that this oracle fixes.  All
arithmetic is plain PyTorch and dtype-generic: the parity tests run it in fp64.

Protocol every layer follows (read off `nf4ad/transforms.py:8-149`):
`forward(x, context=None) -> y`, `backward(y, context=None) -> x`,
`log_abs_det_jacobian(x, y, context=None) -> (B,) or scalar`,
`is_feasible()`, `jitter()/add_jitter()`, `log_prior()`; usable as a
`torch.distributions.Transform` (`_call -> forward`, `_inverse -> backward`).
"""
import math
from typing import Iterable, List, Optional

import torch
from torch.nn import init

from pyro import distributions as dist


class BaseTransform(dist.TransformModule):
    """[INFER transforms.py:8,31-38] nn.Module + Transform base of every layer."""

    bijective = True
    domain = dist.constraints.real_vector
    codomain = dist.constraints.real_vector

    def __init__(self, *args, **kwargs):
        super().__init__(cache_size=0)

    # torch.distributions.Transform plumbing
    def _call(self, x):
        return self.forward(x)

    def _inverse(self, y):
        return self.backward(y)

    def forward(self, x, context=None):
        raise NotImplementedError

    def backward(self, y, context=None):
        raise NotImplementedError

    def log_abs_det_jacobian(self, x, y, context=None):
        raise NotImplementedError

    def is_feasible(self) -> bool:
        return True

    def jitter(self, jitter: float = 1e-6) -> None:
        return None

    def add_jitter(self, jitter: float = 1e-6) -> None:
        return self.jitter(jitter)

    def log_prior(self):
        return 0.0

    def with_cache(self, cache_size=1):
        return self


class LUTransform(BaseTransform):
    """[INFER flows.py:85,110] `LUTransform(dim, prior_scale)`.

    [RECALL] parameters `L_raw`, `U_raw` (D,D) and `bias` (D);
    `L = tril(L_raw,-1) + I`, `U = triu(U_raw)`; forward `y = (L U) x + bias`;
    backward = two triangular solves of `y - bias`; log|det J| = sum log|U_ii|
    (data independent); feasible iff every U_ii != 0.  Init as torch's
    `nn.Linear` (Kaiming-uniform, a=sqrt 5: off-diagonals U(-1/sqrt D, 1/sqrt D)),
    unit diagonals, bias U(-1/sqrt D, 1/sqrt D).
    """

    def __init__(self, dim: int, prior_scale: Optional[float] = 1.0, *args, **kwargs):
        super().__init__()
        self.dim = int(dim)
        self.prior_scale = prior_scale
        self.L_raw = torch.nn.Parameter(torch.empty(self.dim, self.dim))
        self.U_raw = torch.nn.Parameter(torch.empty(self.dim, self.dim))
        self.bias = torch.nn.Parameter(torch.empty(self.dim))
        self.init_params()

    def init_params(self):
        init.kaiming_uniform_(self.L_raw, a=math.sqrt(5))
        init.kaiming_uniform_(self.U_raw, a=math.sqrt(5))
        with torch.no_grad():
            self.L_raw.copy_(self.L_raw.tril(-1) + torch.eye(self.dim))
            self.U_raw.copy_(self.U_raw.triu(1) + torch.eye(self.dim))
        bound = 1.0 / math.sqrt(self.dim) if self.dim > 0 else 0.0
        init.uniform_(self.bias, -bound, bound)

    @property
    def L(self):
        eye = torch.eye(self.dim, dtype=self.L_raw.dtype, device=self.L_raw.device)
        return self.L_raw.tril(-1) + eye

    @property
    def U(self):
        return self.U_raw.triu(0)

    @property
    def weight(self):
        return self.L @ self.U

    def forward(self, x, context=None):
        return torch.nn.functional.linear(x, self.weight, self.bias)

    def backward(self, y, context=None):
        rhs = (y - self.bias).reshape(-1, self.dim).t()
        z = torch.linalg.solve_triangular(self.L, rhs, upper=False, unitriangular=True)
        x = torch.linalg.solve_triangular(self.U, z, upper=True)
        return x.t().reshape(y.shape)

    def log_abs_det_jacobian(self, x, y, context=None):
        return self.U_raw.diagonal().abs().log().sum()

    def is_feasible(self) -> bool:
        return bool((self.U_raw.diagonal() != 0).all())

    def jitter(self, jitter: float = 1e-6) -> None:
        with torch.no_grad():
            d = self.U_raw.diagonal()
            small = d.abs() < jitter
            d[small] = torch.where(d[small] < 0, -jitter, jitter).to(d.dtype)

    def log_prior(self):
        """[RECALL] zero-mean Gaussian prior of std `prior_scale` on the active entries."""
        if self.prior_scale is None:
            return 0.0
        s = float(self.prior_scale)
        active = torch.cat([
            self.L_raw[torch.tril(torch.ones_like(self.L_raw), -1) > 0],
            self.U_raw[torch.triu(torch.ones_like(self.U_raw), 0) > 0],
            self.bias,
        ])
        return (-0.5 * (active / s) ** 2 - math.log(s) - 0.5 * math.log(2 * math.pi)).sum()


class HouseholderTransform(BaseTransform):
    """[INFER flows.py:90] `HouseholderTransform(dim, nvs, device)`.

    [RECALL] product of `nvs` reflections `H_v = I - 2 v v^T / |v|^2` applied in
    storage order; orthogonal, log|det J| = 0; parameter `vk_householder`
    (nvs, D), init N(0,1).
    """

    def __init__(self, dim: int, nvs: int = 1, device="cpu", *args, **kwargs):
        super().__init__()
        self.dim = int(dim)
        self.nvs = int(nvs)
        self.vk_householder = torch.nn.Parameter(torch.randn(self.nvs, self.dim, device=device))

    @staticmethod
    def _reflect(x, v):
        return x - (2.0 / (v * v).sum()) * (x * v).sum(-1, keepdim=True) * v

    def forward(self, x, context=None):
        for k in range(self.nvs):
            x = self._reflect(x, self.vk_householder[k])
        return x

    def backward(self, y, context=None):
        for k in reversed(range(self.nvs)):
            y = self._reflect(y, self.vk_householder[k])
        return y

    def log_abs_det_jacobian(self, x, y, context=None):
        return torch.zeros((), dtype=x.dtype, device=x.device)

    def is_feasible(self) -> bool:
        return bool(((self.vk_householder ** 2).sum(-1) > 0).all())


class ScaleTransform(BaseTransform):
    """[INFER flows.py:113] `ScaleTransform(in_dims)`.

    [RECALL] learnable elementwise `scale` of shape `in_dims` (init ones);
    `y = x * scale`; log|det J| = sum log|scale|; feasible iff no zero entry.
    """

    def __init__(self, dim: Iterable[int], *args, **kwargs):
        super().__init__()
        self.dim = tuple(int(d) for d in dim)
        self.scale = torch.nn.Parameter(torch.ones(self.dim))

    def forward(self, x, context=None):
        return x * self.scale

    def backward(self, y, context=None):
        return y / self.scale

    def log_abs_det_jacobian(self, x, y, context=None):
        return self.scale.abs().log().sum()

    def is_feasible(self) -> bool:
        return bool((self.scale != 0).all())

    def jitter(self, jitter: float = 1e-6) -> None:
        with torch.no_grad():
            small = self.scale.abs() < jitter
            self.scale[small] = jitter


class SequentialAffineTransform(BaseTransform):
    """[INFER flows.py:95] composition of affine layers: forward in list order,
    backward in reverse, log-dets add."""

    def __init__(self, transforms: List[BaseTransform], *args, **kwargs):
        super().__init__()
        self.transforms = torch.nn.ModuleList(transforms)

    def forward(self, x, context=None):
        for t in self.transforms:
            x = t.forward(x)
        return x

    def backward(self, y, context=None):
        for t in reversed(self.transforms):
            y = t.backward(y)
        return y

    def log_abs_det_jacobian(self, x, y, context=None):
        total = 0.0
        for t in self.transforms:
            total = total + t.log_abs_det_jacobian(x, y)
        return total

    def is_feasible(self) -> bool:
        return all(t.is_feasible() for t in self.transforms)

    def jitter(self, jitter: float = 1e-6) -> None:
        for t in self.transforms:
            t.jitter(jitter)

    def log_prior(self):
        total = 0.0
        for t in self.transforms:
            total = total + t.log_prior()
        return total


class BlockAffineTransform(BaseTransform):
    """[INFER flows.py:95,111] `BlockAffineTransform(in_dims, block_transform)`.

    [RECALL] applies the `dim = in_dims[0]` affine layer along the LEADING event
    dimension: for `in_dims=[D]` (every nf4ad config) it is the wrapped
    transform itself; for an image-shaped event `[C, H, W]` it is the same C x C
    map at every pixel (a 1x1 convolution), so the log-det counts H*W times.
    """

    def __init__(self, in_dims, block_transform: BaseTransform, *args, **kwargs):
        super().__init__()
        self.in_dims = tuple(int(d) for d in in_dims)
        self.block_transform = block_transform

    def _per_position(self, v, fn):
        n = len(self.in_dims)
        if n == 1:
            return fn(v)
        vm = v.movedim(v.dim() - n, -1)                  # (*batch, *spatial, C)
        return fn(vm.reshape(-1, self.in_dims[0])).reshape(vm.shape).movedim(-1, v.dim() - n)

    def forward(self, x, context=None):
        return self._per_position(x, self.block_transform.forward)

    def backward(self, y, context=None):
        return self._per_position(y, self.block_transform.backward)

    def log_abs_det_jacobian(self, x, y, context=None):
        positions = 1
        for d in self.in_dims[1:]:
            positions *= d
        return self.block_transform.log_abs_det_jacobian(x, y) * positions

    def is_feasible(self) -> bool:
        return self.block_transform.is_feasible()

    def jitter(self, jitter: float = 1e-6) -> None:
        self.block_transform.jitter(jitter)

    def log_prior(self):
        return self.block_transform.log_prior()


class InverseTransform(BaseTransform):
    """[INFER flows.py:104] swaps forward/backward of the wrapped layer (sharing
    its parameters) and negates the log-det."""

    def __init__(self, transform: BaseTransform, *args, **kwargs):
        super().__init__()
        self.transform = transform

    def forward(self, x, context=None):
        return self.transform.backward(x)

    def backward(self, y, context=None):
        return self.transform.forward(y)

    def log_abs_det_jacobian(self, x, y, context=None):
        return -self.transform.log_abs_det_jacobian(y, x)

    def is_feasible(self) -> bool:
        return self.transform.is_feasible()

    def jitter(self, jitter: float = 1e-6) -> None:
        self.transform.jitter(jitter)


class MaskedCoupling(BaseTransform):
    """[INFER flows.py:28-30 docstring, YAML `param_dims:[D]`] additive coupling of
    USFlow: `y = x + (1-m) * cond(m*x)`, log|det J| = 0."""

    def __init__(self, mask: torch.Tensor, conditioner: torch.nn.Module, *args, **kwargs):
        super().__init__()
        self.register_buffer("mask", mask.float())
        self.conditioner = conditioner

    def _shift(self, xm, context=None):
        t = self.conditioner(xm) if context is None else self.conditioner(xm, context)
        if isinstance(t, (tuple, list)):
            t = t[-1]
        return t.to(xm.dtype)

    def forward(self, x, context=None):
        return x + (1.0 - self.mask) * self._shift(x * self.mask, context)

    def backward(self, y, context=None):
        return y - (1.0 - self.mask) * self._shift(y * self.mask, context)

    def log_abs_det_jacobian(self, x, y, context=None):
        batched = x.dim() > self.mask.dim() - 1
        return torch.zeros(x.shape[0] if batched else (), dtype=x.dtype, device=x.device)

    def is_feasible(self) -> bool:
        m = self.mask
        return bool(((m == 0) | (m == 1)).all())
