"""B200 bijector layers behind the USFlows `src.usflows.transforms` names.

Same classes, constructor arguments, parameter names and `BaseTransform` protocol as the layers nf4ad
imports (`/root/reference/src/nf4ad/flows.py:9-17`, `/root/reference/src/nf4ad/transforms.py:5`), but
every `forward / backward / log_abs_det_jacobian` enqueues hand-written sm_100a kernels through the C
ABI (`nf4ad_b200.ops`).  CUDA tensors only: there is no CPU path (the CPU oracle lives in `oracle/`).

`MaskedAffineCoupling` mirrors the reference's own class (`nf4ad/transforms.py:8-149`) so flows can be
built without the reference checkout; the reference's class itself also works on top of these layers.
"""
import math
from typing import Iterable, List, Optional

import torch
from torch.distributions import Transform, constraints
from torch.nn import init

from . import ops


class TransformModule(Transform, torch.nn.Module):
    """`pyro.distributions.TransformModule`: a Transform that owns parameters."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    def __hash__(self):
        return torch.nn.Module.__hash__(self)

    def __eq__(self, other):
        return self is other


def _as2d(x):
    if x.dim() == 1:
        return x.unsqueeze(0), True
    if x.dim() != 2:
        raise ValueError(f"expected (B, D) or (D,) input, got shape {tuple(x.shape)}")
    return x, False


class BaseTransform(TransformModule):
    """Protocol base (`nf4ad/transforms.py:8,31-38`): forward / backward / log_abs_det_jacobian,
    is_feasible, jitter / add_jitter, log_prior."""

    bijective = True
    domain = constraints.real_vector
    codomain = constraints.real_vector

    def __init__(self, *args, **kwargs):
        super().__init__(cache_size=0)

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

    # affine layers override: number of dense (D x D) maps, used by the stack compiler
    is_affine = False


class LUTransform(BaseTransform):
    """`LUTransform(dim, prior_scale)` (call sites `nf4ad/flows.py:85,110`).

    Parameters `L_raw`, `U_raw` (D,D), `bias` (D); `L = tril(L_raw,-1)+I`, `U = triu(U_raw)`.
    forward  y = (L U) x + b      -> usf_lu_pack (factor product) + usf_linear (GEMM)
    backward x = U^-1 L^-1 (y-b)  -> usf_lu_solve (blocked triangular solve)
    log|det| = sum log|U_ii| (data independent).
    """

    is_affine = True

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
            eye = torch.eye(self.dim, device=self.L_raw.device, dtype=self.L_raw.dtype)
            self.L_raw.copy_(self.L_raw.tril(-1) + eye)
            self.U_raw.copy_(self.U_raw.triu(1) + eye)
        bound = 1.0 / math.sqrt(self.dim) if self.dim > 0 else 0.0
        init.uniform_(self.bias, -bound, bound)

    @property
    def weight(self):
        return ops.LUPackFn.apply(self.L_raw, self.U_raw)

    def forward(self, x, context=None):
        x2, squeeze = _as2d(x)
        y = ops.linear_fn(x2, self.weight, self.bias, False)
        return y[0] if squeeze else y

    def backward(self, y, context=None):
        y2, squeeze = _as2d(y)
        if ops.tc_train_enabled() and y2.is_cuda and self.dim % 16 == 0 and y2.shape[0] % 8 == 0 and y2.shape[0] >= 8:
            # mixed-precision training: x = A (y - b) with the dense inverse A = (LU)^{-1} (one weight-space solve per
            # step) applied by a tensor-core GEMM; the shift -A b is a weight-space matrix-vector product
            pre = self.__dict__.pop("_A_pre", None)      # issued up front by Flow._prefetch_lu_inverses
            if pre is not None:
                A, side = pre
                torch.cuda.current_stream(y2.device).wait_stream(side)
            else:
                A = ops.LUInverseFn.apply(self.L_raw, self.U_raw)
            x = ops.linear_fn(y2, A, -(A @ self.bias), False)
        else:
            x = ops.LUSolveFn.apply(y2, self.L_raw, self.U_raw, self.bias)
        return x[0] if squeeze else x

    def log_abs_det_jacobian(self, x, y, context=None):
        # O(D) parameter-only constant: plain tensor ops (not on the data path)
        return self.U_raw.diagonal().abs().log().sum()

    def is_feasible(self) -> bool:
        return bool((self.U_raw.diagonal() != 0).all())

    def jitter(self, jitter: float = 1e-6) -> None:
        with torch.no_grad():
            d = self.U_raw.diagonal()
            small = d.abs() < jitter
            d[small] = torch.where(d[small] < 0, -jitter, jitter).to(d.dtype)

    def log_prior(self):
        if self.prior_scale is None:
            return 0.0
        # Gaussian prior N(0, s) over the active entries (strict lower triangle of L, upper triangle of U, bias): fixed-shape
        # tensor ops only (no boolean indexing), so that a training step with the MAP term replays as a CUDA graph
        s = float(self.prior_scale)
        D = self.L_raw.shape[0]
        sq = (torch.tril(self.L_raw, -1) ** 2).sum() + (torch.triu(self.U_raw) ** 2).sum() + (self.bias ** 2).sum()
        return -0.5 * sq / (s * s) - (D * D + D) * (math.log(s) + 0.5 * math.log(2 * math.pi))


class HouseholderTransform(BaseTransform):
    """`HouseholderTransform(dim, nvs, device)` (`nf4ad/flows.py:90`): product of `nvs` reflections
    `I - 2 v v^T/|v|^2`, parameter `vk_householder` (nvs, D); log|det| = 0.  usf_householder."""

    is_affine = True

    def __init__(self, dim: int, nvs: int = 1, device="cpu", *args, **kwargs):
        super().__init__()
        self.dim = int(dim)
        self.nvs = int(nvs)
        self.vk_householder = torch.nn.Parameter(torch.randn(self.nvs, self.dim, device=device))

    def forward(self, x, context=None):
        x2, squeeze = _as2d(x)
        y = ops.HouseholderFn.apply(x2, self.vk_householder, False)
        return y[0] if squeeze else y

    def backward(self, y, context=None):
        y2, squeeze = _as2d(y)
        x = ops.HouseholderFn.apply(y2, self.vk_householder, True)
        return x[0] if squeeze else x

    def log_abs_det_jacobian(self, x, y, context=None):
        return torch.zeros((), dtype=x.dtype, device=x.device)

    def is_feasible(self) -> bool:
        return bool(((self.vk_householder ** 2).sum(-1) > 0).all())


class ScaleTransform(BaseTransform):
    """`ScaleTransform(in_dims)` (`nf4ad/flows.py:113`): y = x * scale, log|det| = sum log|scale|.  usf_scale."""

    is_affine = True

    def __init__(self, dim: Iterable[int], *args, **kwargs):
        super().__init__()
        self.dim = tuple(int(d) for d in dim)
        if len(self.dim) != 1:
            raise NotImplementedError("only 1-D event shapes in_dims=[D] are supported")
        self.scale = torch.nn.Parameter(torch.ones(self.dim))

    def forward(self, x, context=None):
        x2, squeeze = _as2d(x)
        y = ops.ScaleFn.apply(x2, self.scale, False)
        return y[0] if squeeze else y

    def backward(self, y, context=None):
        y2, squeeze = _as2d(y)
        x = ops.ScaleFn.apply(y2, self.scale, True)
        return x[0] if squeeze else x

    def log_abs_det_jacobian(self, x, y, context=None):
        return self.scale.abs().log().sum()

    def is_feasible(self) -> bool:
        return bool((self.scale != 0).all())

    def jitter(self, jitter: float = 1e-6) -> None:
        with torch.no_grad():
            small = self.scale.abs() < jitter
            self.scale[small] = jitter


class SequentialAffineTransform(BaseTransform):
    """`SequentialAffineTransform([..])` (`nf4ad/flows.py:95`): composition, log-dets add."""

    is_affine = True

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
    """`BlockAffineTransform(in_dims, block_transform)` (`nf4ad/flows.py:95,111`); for the 1-D event
    shapes nf4ad uses it is the wrapped transform itself (image-shaped in_dims: not supported)."""

    is_affine = True

    def __init__(self, in_dims, block_transform: BaseTransform, *args, **kwargs):
        super().__init__()
        self.in_dims = tuple(int(d) for d in in_dims)
        if len(self.in_dims) != 1:
            raise NotImplementedError("only 1-D event shapes in_dims=[D] are supported")
        self.block_transform = block_transform

    def forward(self, x, context=None):
        return self.block_transform.forward(x)

    def backward(self, y, context=None):
        return self.block_transform.backward(y)

    def log_abs_det_jacobian(self, x, y, context=None):
        return self.block_transform.log_abs_det_jacobian(x, y)

    def is_feasible(self) -> bool:
        return self.block_transform.is_feasible()

    def jitter(self, jitter: float = 1e-6) -> None:
        self.block_transform.jitter(jitter)

    def log_prior(self):
        return self.block_transform.log_prior()


class InverseTransform(BaseTransform):
    """`InverseTransform(t)` (`nf4ad/flows.py:104`): swaps the directions of `t` (sharing its
    parameters) and negates the log-det."""

    def __init__(self, transform: BaseTransform, *args, **kwargs):
        super().__init__()
        self.transform = transform

    @property
    def is_affine(self):
        return bool(getattr(self.transform, "is_affine", False))

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


# ------------------------------------------------------------------------------------------------
# couplings
# ------------------------------------------------------------------------------------------------
def mlp_layers(module):
    """If `module` is a plain Linear/ReLU chain (the conditioners nf4ad configures: `nn.Sequential`
    MLPs wrapped in `.net` -- `tests/conftest.py:111-121` -- or `pyro.nn.DenseNN`), return its Linear
    layers in call order, else None (opaque conditioner)."""
    if isinstance(module, torch.nn.Linear):
        return [module]
    layers = getattr(module, "layers", None)
    if isinstance(layers, torch.nn.ModuleList) and isinstance(getattr(module, "f", None), torch.nn.ReLU) \
            and len(layers) > 0 and all(isinstance(l, torch.nn.Linear) for l in layers):
        linears = list(layers)                                   # DenseNN duck type
    else:
        leaves = [m for m in module.modules()
                  if len(list(m.children())) == 0 and not isinstance(m, torch.nn.Identity)]
        if len(leaves) % 2 == 0:
            return None
        for i, m in enumerate(leaves):
            if not isinstance(m, torch.nn.Linear if i % 2 == 0 else torch.nn.ReLU):
                return None
        linears = leaves[0::2]
    for a, b in zip(linears[:-1], linears[1:]):
        if a.out_features != b.in_features:
            return None
    return linears


def run_conditioner(cond, xm, context=None):
    """Evaluates the conditioner.  Recognised Linear/ReLU chains run on usf_linear (our GEMM +
    bias/ReLU epilogue); anything else is an opaque user module evaluated as given."""
    if context is not None:
        return cond(xm, context)
    linears = mlp_layers(cond)
    if linears is None or not xm.is_cuda:
        return cond(xm)
    h = xm
    for i, lin in enumerate(linears):
        h = ops.linear_fn(h, lin.weight, lin.bias, i + 1 < len(linears))
    pd = getattr(cond, "param_dims", None)
    if pd is not None and len(pd) > 1:           # DenseNN tuple output
        outs, o = [], 0
        for p in pd:
            outs.append(h[..., o:o + p])
            o += p
        return tuple(outs)
    return h


def split_params(params, x):
    """`MaskedAffineCoupling._parse_params` contract (`nf4ad/transforms.py:40-64`)."""
    if isinstance(params, (list, tuple)) and len(params) == 2:
        s, t = params
    elif params.shape == x.shape:
        s, t = None, params
    elif params.dim() >= 2 and x.dim() >= 2 and params.shape[1] == 2 * x.shape[1]:
        c = x.shape[1]
        s, t = params[:, :c], params[:, c:]
    else:
        raise ValueError(
            "Conditioner output shape not compatible. "
            "Expected (s,t) tuple, tensor same shape as x, or tensor with 2*C channels.")
    return (None if s is None else s.to(x.dtype)), t.to(x.dtype)


class MaskedAffineCoupling(BaseTransform):
    """Affine masked coupling (`nf4ad/transforms.py:8-149`):
    `y = m x + (1-m)(x exp(clamp tanh s(m x)) + t(m x))`, inverse, per-row log-det.
    The elementwise part is one fused kernel (usf_coupling); the conditioner is evaluated ONCE per call
    pair (the reference evaluates it again inside log_abs_det_jacobian, `:121-125`)."""

    bijective = True

    def __init__(self, mask, conditioner, scale_activation: str = "exp", clamp: float = 5.0):
        super().__init__()
        self.register_buffer("mask", mask.float())
        self.conditioner = conditioner
        self.scale_activation = scale_activation
        self.clamp = float(clamp)
        self.domain = constraints.real_vector
        self.codomain = constraints.real_vector
        self._memo = None

    additive = False

    def _apply_dir(self, v, inverse, context=None):
        if self.scale_activation != "exp":
            raise NotImplementedError("only scale_activation='exp' is implemented on the B200 path")
        v2, squeeze = _as2d(v)
        mask = self.mask.reshape(-1)
        vm = ops.ScaleFn.apply(v2, mask, False)
        s, t = split_params(run_conditioner(self.conditioner, vm, context), v2)
        if self.additive:
            s = None
        out, ladj = ops.CouplingFn.apply(v2, s, t, mask, self.clamp, inverse)
        return (out[0] if squeeze else out), ladj

    def forward(self, x, context=None):
        y, ladj = self._apply_dir(x, False, context)
        self._memo = (x, y, ladj)
        return y

    def backward(self, y, context=None):
        x, ladj = self._apply_dir(y, True, context)
        self._memo = (x, y, ladj)
        return x

    def inverse_and_ladj(self, y, context=None):
        x, ladj = self._apply_dir(y, True, context)
        return x, ladj

    def forward_and_ladj(self, x, context=None):
        return self._apply_dir(x, False, context)

    def log_abs_det_jacobian(self, x, y, context=None):
        memo, self._memo = self._memo, None
        if memo is not None and (memo[0] is x or memo[1] is y):
            return memo[2]
        _, ladj = self._apply_dir(x, False, context)
        return ladj

    def is_feasible(self):
        m = self.mask
        return bool(((m == 0) | (m == 1)).all())

    def jitter(self, jitter: float = 1e-6) -> None:
        return None


class MaskedCoupling(MaskedAffineCoupling):
    """Additive coupling of USFlow (`nf4ad/flows.py:28-30` docstring; YAML `param_dims:[D]`):
    `y = x + (1-m) cond(m x)`, log|det| = 0."""

    additive = True

    def __init__(self, mask, conditioner, *args, **kwargs):
        super().__init__(mask, conditioner)

    def _apply_dir(self, v, inverse, context=None):
        v2, squeeze = _as2d(v)
        mask = self.mask.reshape(-1)
        vm = ops.ScaleFn.apply(v2, mask, False)
        t = run_conditioner(self.conditioner, vm, context)
        if isinstance(t, (tuple, list)):
            t = t[-1]
        out, ladj = ops.CouplingFn.apply(v2, None, t.to(v2.dtype), mask, self.clamp, inverse)
        return (out[0] if squeeze else out), ladj
