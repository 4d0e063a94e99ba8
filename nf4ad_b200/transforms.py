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
    """(..., D) -> ((rows, D), restore): the 1-D layers act on the last dimension, whatever precedes it."""
    if x.dim() == 2:
        return x, False
    if x.dim() == 1:
        return x.unsqueeze(0), True
    if x.dim() == 0:
        raise ValueError("expected at least a (D,) input")
    return x.reshape(-1, x.shape[-1]), tuple(x.shape)


def _as_event_rows(x, event_ndim):
    """(*batch, *event) -> ((rows, prod(event)), restore) for the element-wise layers of an N-D event shape."""
    if event_ndim == 1:
        return _as2d(x)
    if x.dim() < event_ndim:
        raise ValueError(f"input of shape {tuple(x.shape)} has fewer dimensions than the event shape")
    n = 1
    for d in x.shape[x.dim() - event_ndim:]:
        n *= int(d)
    if x.dim() == event_ndim:
        return x.reshape(1, n), tuple(x.shape)
    return x.reshape(-1, n), (False if x.dim() == 2 and event_ndim == 1 else tuple(x.shape))


def _restore(y, how):
    if how is False:
        return y
    if how is True:
        return y[0]
    return y.reshape(how)


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


# Set while `Flow._compose_affine_runs` evaluates an affine run on probe rows: "linear_only" drops the shifts (the probe
# rows are the identity: the run's matrix), "A" / "W" share a layer's dense inverse / factor product between the
# matrix pass and the shift pass of one composition.
_COMPOSE = None


def _mark(t):
    """(stream, event) after the work enqueued so far on the current stream of `t`'s device (None on the CPU)."""
    if not t.is_cuda:
        return None
    ev = torch.cuda.Event()
    st = torch.cuda.current_stream(t.device)
    ev.record(st)
    return st, ev


def _shared(hit):
    """A tensor cached by a composition, made safe to read on the current stream (the matrix pass and the shift pass
    of a run are enqueued on two streams)."""
    t, mark = hit
    if mark is not None:
        cur = torch.cuda.current_stream(t.device)
        if cur != mark[0]:
            cur.wait_event(mark[1])
            t.record_stream(cur)
    return t


class composing:
    def __enter__(self):
        global _COMPOSE
        self.prev = _COMPOSE
        _COMPOSE = {"linear_only": False, "A": {}, "W": {}}
        return _COMPOSE

    def __exit__(self, *exc):
        global _COMPOSE
        _COMPOSE = self.prev
        return False


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
        comp = _COMPOSE
        if comp is not None:                                  # one product per layer per composition
            hit = comp["W"].get(id(self))
            if hit is None:
                # on the composition's auxiliary stream: the product (and, in the backward pass, its factor gradients --
                # three D^3 products) runs beside the chain that consumes it instead of inside it
                aux = comp.get("aux")
                if aux is not None and self.L_raw.is_cuda:
                    aux.wait_stream(torch.cuda.current_stream(self.L_raw.device))
                    with torch.cuda.stream(aux):
                        hit = (ops.LUPackFn.apply(self.L_raw, self.U_raw), _mark(self.L_raw))
                else:
                    hit = (ops.LUPackFn.apply(self.L_raw, self.U_raw), _mark(self.L_raw))
                comp["W"][id(self)] = hit
            return _shared(hit)
        return ops.LUPackFn.apply(self.L_raw, self.U_raw)

    def forward(self, x, context=None):
        x2, squeeze = _as2d(x)
        comp = _COMPOSE
        if comp is not None and comp["linear_only"]:
            if x2 is comp.get("eye"):                    # the identity probe: I W^T, no product needed
                return _restore(self.weight.t().contiguous(), squeeze)
            return _restore(ops.linear_fn(x2, self.weight, None, False), squeeze)
        y = ops.linear_fn(x2, self.weight, self.bias, False)
        return _restore(y, squeeze)

    def backward(self, y, context=None):
        y2, squeeze = _as2d(y)
        if ops.tc_train_enabled() and y2.is_cuda and self.dim % 16 == 0 and y2.shape[0] % 8 == 0 and y2.shape[0] >= 8:
            # mixed-precision training: x = A (y - b) with the dense inverse A = (LU)^{-1} (one weight-space solve per
            # step) applied by a tensor-core GEMM; the shift -A b is a weight-space matrix-vector product
            comp = _COMPOSE
            pre = self.__dict__.pop("_A_pre", None)      # issued up front by Flow._prefetch_lu_inverses
            if comp is not None and id(self) in comp["A"]:
                A = _shared(comp["A"][id(self)])
            elif pre is not None:
                A, side = pre
                torch.cuda.current_stream(y2.device).wait_stream(side)
            else:
                A = ops.LUInverseFn.apply(self.L_raw, self.U_raw)
            if comp is not None:
                if id(self) not in comp["A"]:
                    comp["A"][id(self)] = (A, _mark(A))
                if comp["linear_only"]:
                    if y2 is comp.get("eye"):            # the identity probe: I A^T, no product needed
                        return _restore(A.t().contiguous(), squeeze)
                    return _restore(ops.linear_fn(y2, A, None, False), squeeze)
            # the shift -A b through the library's own GEMM (a 1 x D x D product; no cuBLAS call on the training path)
            shift = ops.LinearFn.apply(self.bias.unsqueeze(0), A, None, False).squeeze(0)
            x = ops.linear_fn(y2, A, -shift, False)
        else:
            x = ops.LUSolveFn.apply(y2, self.L_raw, self.U_raw, self.bias)
        return _restore(x, squeeze)

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
        return _restore(y, squeeze)

    def backward(self, y, context=None):
        y2, squeeze = _as2d(y)
        x = ops.HouseholderFn.apply(y2, self.vk_householder, True)
        return _restore(x, squeeze)

    def log_abs_det_jacobian(self, x, y, context=None):
        return torch.zeros((), dtype=x.dtype, device=x.device)

    def is_feasible(self) -> bool:
        return bool(((self.vk_householder ** 2).sum(-1) > 0).all())


class ScaleTransform(BaseTransform):
    """`ScaleTransform(in_dims)` (`nf4ad/flows.py:113`): y = x * scale, log|det| = sum log|scale|.  usf_scale."""

    is_affine = True

    def __init__(self, dim: Iterable[int], *args, **kwargs):
        super().__init__()
        self.dim = tuple(int(d) for d in dim)            # the event shape: [D] or image-shaped [C, H, W]
        self.scale = torch.nn.Parameter(torch.ones(self.dim))

    def _run(self, v, inverse):
        v2, how = _as_event_rows(v, len(self.dim))
        out = ops.ScaleFn.apply(v2, self.scale.reshape(-1), inverse)
        return _restore(out, how)

    def forward(self, x, context=None):
        return self._run(x, False)

    def backward(self, y, context=None):
        return self._run(y, True)

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
    """`BlockAffineTransform(in_dims, block_transform)` (`nf4ad/flows.py:95,111`): applies a `dim = in_dims[0]` affine
    layer along the LEADING event dimension.  For `in_dims=[D]` that is the wrapped transform itself; for an image-shaped
    event `[C, H, W]` it is the same C x C map at every pixel (a 1x1 convolution): the input is viewed as
    `(B*H*W, C)` rows -- one GEMM / triangular solve over all pixels -- and the log-det counts once per pixel."""

    is_affine = True

    def __init__(self, in_dims, block_transform: BaseTransform, *args, **kwargs):
        super().__init__()
        self.in_dims = tuple(int(d) for d in in_dims)
        self.block_transform = block_transform
        self.n_positions = 1
        for d in self.in_dims[1:]:
            self.n_positions *= d

    def _run(self, v, fn):
        n = len(self.in_dims)
        if n == 1:
            return fn(v)
        if v.dim() < n:
            raise ValueError(f"input of shape {tuple(v.shape)} does not end in the event shape {self.in_dims}")
        vm = v.movedim(v.dim() - n, -1)                   # (*batch, *spatial, C)
        out = fn(vm.reshape(-1, self.in_dims[0]))
        return out.reshape(vm.shape).movedim(-1, v.dim() - n)

    def forward(self, x, context=None):
        return self._run(x, self.block_transform.forward)

    def backward(self, y, context=None):
        return self._run(y, self.block_transform.backward)

    def log_abs_det_jacobian(self, x, y, context=None):
        return self.block_transform.log_abs_det_jacobian(x, y) * self.n_positions

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
def context_dim(module):
    """Width of the context a conditional conditioner takes (`pyro.nn.ConditionalDenseNN`: `forward(x, context)` =
    DenseNN on `cat([context, x])`); 0 for an unconditional one."""
    return int(getattr(module, "context_dim", 0) or 0)


def mlp_layers(module):
    """If `module` is a plain Linear/ReLU chain (the conditioners nf4ad configures: `nn.Sequential`
    MLPs wrapped in `.net` -- `tests/conftest.py:111-121` -- or `pyro.nn.DenseNN` / `ConditionalDenseNN`), return its
    Linear layers in call order, else None (opaque conditioner)."""
    if isinstance(module, torch.nn.Linear):
        return [module]
    layers = getattr(module, "layers", None)
    if isinstance(layers, torch.nn.ModuleList) and isinstance(getattr(module, "f", None), torch.nn.ReLU) \
            and len(layers) > 0 and all(isinstance(l, torch.nn.Linear) for l in layers):
        linears = list(layers)                                   # DenseNN duck type
    else:
        if context_dim(module):
            return None
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


def conditioner_weights(cond, mask=None):
    """The (weight, bias) pairs of a recognised Linear/ReLU conditioner as the training GEMMs take them, or None.  With
    `mask` (the coupling's 0/1 mask over the event) the masking of the conditioner's input is folded into the first
    layer -- `(x m) W^T = x (W m)^T`, a weight-sized product whose autograd node also masks the weight gradient -- so
    that no batch-sized masking pass runs."""
    linears = mlp_layers(cond)
    if linears is None:
        return None
    out = []
    for i, lin in enumerate(linears):
        W = lin.weight
        if i == 0 and mask is not None:
            cd = context_dim(cond)
            m = mask.reshape(-1).to(W.dtype)
            if cd:
                m = torch.cat([torch.ones(cd, device=m.device, dtype=m.dtype), m])
            if m.numel() != W.shape[1]:
                return None
            W = W * m
        out.append((W, lin.bias))
    return out


def run_conditioner(cond, xm, context=None, mask=None):
    """Evaluates the conditioner on flat rows `xm` (B, D).  Recognised Linear/ReLU chains run on usf_linear (our GEMM +
    bias/ReLU epilogue) -- a conditional one on `cat([context, xm])`; anything else is an opaque user module evaluated
    as given (`conditioner(x_masked)` / `conditioner(x_masked, context)`, `nf4ad/transforms.py:71-74`).  `mask`: `xm` is
    the UNMASKED input and the mask is to be applied here (tensor-core training: folded into the first layer's weight,
    `conditioner_weights`; the layers' bf16 operands may have been prepared ahead by `Flow._prefetch_conditioners`)."""
    linears = mlp_layers(cond) if xm.is_cuda and xm.dim() == 2 else None
    cd = context_dim(cond)
    fast = not (linears is None or (context is not None) != (cd > 0) or linears[0].in_features != xm.shape[1] + cd)
    pre = cond.__dict__.pop("_usf_pre", None) if isinstance(cond, torch.nn.Module) else None
    weights = None
    if fast and mask is not None:
        if pre is not None:
            weights, side = pre
            torch.cuda.current_stream(xm.device).wait_stream(side)
        else:
            weights = conditioner_weights(cond, mask)
            weights = None if weights is None else [(W, b, None) for W, b in weights]
    if mask is not None and weights is None:
        xm = ops.MaskFn.apply(xm, mask.reshape(-1))
    if not fast:
        return cond(xm) if context is None else cond(xm, context)
    h = xm
    if cd:
        ctx = context.to(xm.dtype)
        ctx = ctx.reshape(-1, cd) if ctx.dim() != 2 else ctx
        h = torch.cat([ctx.expand(xm.shape[0], cd), xm], dim=-1)
    if weights is None:
        weights = [(lin.weight, lin.bias, None) for lin in linears]
    for i, (W, b, operands) in enumerate(weights):
        h = ops.linear_fn(h, W, b, i + 1 < len(weights), operands=operands)
    pd = getattr(cond, "param_dims", None)
    if pd is not None and len(pd) > 1:           # DenseNN tuple output
        outs, o = [], 0
        for p in pd:
            outs.append(h[..., o:o + p])
            o += p
        return tuple(outs)
    return h


def split_params(params, x):
    """`MaskedAffineCoupling._parse_params` contract (`nf4ad/transforms.py:40-64`): an `(s, t)` pair; ONE tensor of x's
    shape = additive shift (s = 0, returned as None: the kernels skip the scale arithmetic); ONE tensor with twice x's
    size along dim 1 = `[s | t]` split there (channels for an image-shaped event); anything else is a ValueError."""
    if isinstance(params, (list, tuple)) and len(params) == 2:
        s, t = params
    elif params.shape == x.shape:
        s, t = None, params
    elif params.dim() >= 2 and x.dim() >= 2 and params.shape[1] == 2 * x.shape[1]:
        c = x.shape[1]
        s, t = params[:, :c, ...], params[:, c:, ...]
    else:
        raise ValueError(
            "Conditioner output shape not compatible. "
            "Expected (s,t) tuple, tensor same shape as x, or tensor with 2*C channels.")
    return (None if s is None else s.to(x.dtype)), t.to(x.dtype)


def coupling_apply(layer, v, inverse, context=None, additive=False):
    """One masked coupling in either direction -> (output, per-row log-det of the FORWARD map), for this package's
    classes and for any layer with the same attributes (`mask`, `conditioner`, `clamp`, `scale_activation`: the
    reference's own `MaskedAffineCoupling` running unmodified on the drop-in).  The event shape is the mask's
    (`(1, D)`, or image-shaped `(1, C, H, W)`): the conditioner sees the masked input in that shape
    (`nf4ad/transforms.py:70-74`), its output is split per `_parse_params` (`:40-64`), and the element-wise part
    -- mask select, scale activation, scale-shift or its inverse, per-row log-det -- is ONE kernel over the flattened
    rows (usf_coupling)."""
    act = ops.scale_activation_id(getattr(layer, "scale_activation", "exp"))
    mask = layer.mask
    ev = tuple(mask.shape[1:]) if mask.dim() > 1 else tuple(mask.shape)
    n = len(ev)
    if v.dim() == n:
        vb, how = v.unsqueeze(0), True
    elif v.dim() == n + 1:
        vb, how = v, False
    elif v.dim() > n + 1:
        vb, how = v.reshape(-1, *v.shape[v.dim() - n:]), tuple(v.shape)
    else:
        raise ValueError(f"input of shape {tuple(v.shape)} does not end in the event shape {ev}")
    B = vb.shape[0]
    v2 = vb.reshape(B, -1)
    m = mask.reshape(-1)
    if vb.dim() == 2 and ops.tc_train_enabled() and v2.is_cuda:
        # tensor-core training: the mask goes into the conditioner's first layer (no batch-sized masking pass)
        params = run_conditioner(layer.conditioner, v2, context, mask=m)
    elif vb.dim() == 2:
        params = run_conditioner(layer.conditioner, ops.ScaleFn.apply(v2, m, False), context)
    else:                                   # image-shaped event: the conditioner sees (B, C, H, W)
        vm = ops.ScaleFn.apply(v2, m, False)
        vmb = vm.reshape(vb.shape)
        params = layer.conditioner(vmb) if context is None else layer.conditioner(vmb, context)
    if not additive and vb.dim() == 2 and torch.is_tensor(params) and params.dim() == 2 and v2.is_cuda \
            and params.shape == (B, 2 * v2.shape[1]) and params.dtype == v2.dtype == torch.float32:
        # ONE [s | t] tensor (`_parse_params`' 2C form on flat rows): read in place, gradient written as one tensor
        out, ladj = ops.CouplingPackedFn.apply(v2, params, m, float(getattr(layer, "clamp", 5.0)), inverse, act)
        return _restore(out.reshape(vb.shape), how), ladj
    if additive:
        t = params[-1] if isinstance(params, (tuple, list)) else params
        s, t = None, t.to(vb.dtype)
    else:
        s, t = split_params(params, vb)
    s = None if s is None else s.reshape(B, -1)
    out, ladj = ops.CouplingFn.apply(v2, s, t.reshape(B, -1), m, float(getattr(layer, "clamp", 5.0)), inverse, act)
    return _restore(out.reshape(vb.shape), how), ladj


class MaskedAffineCoupling(BaseTransform):
    """Affine masked coupling (`nf4ad/transforms.py:8-149`):
    `y = m x + (1-m)(x scale(s(m x)) + t(m x))`, `scale = exp(clamp tanh s)` (or the `softplus` variant), inverse,
    per-row log-det.  Event shape = the mask's (`(1, D)` or image-shaped `(1, C, H, W)`): the input keeps its shape for
    the conditioner, the element-wise part runs on the flattened rows as one fused kernel (usf_coupling); the
    conditioner is evaluated ONCE per call pair (the reference evaluates it again inside log_abs_det_jacobian,
    `:121-125`)."""

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
        return coupling_apply(self, v, inverse, context, additive=self.additive)

    # `TransformedDistribution.log_prob` asks for `x = T.inv(y)` and then `T.log_abs_det_jacobian(x, y)`; the reference
    # evaluates the conditioner again for the second call (`nf4ad/transforms.py:121-125`).  Here the pair of calls shares
    # ONE evaluation through a single-use memo: it is replaced by every forward / backward, dropped by train() / eval(),
    # consumed by the first log_abs_det_jacobian and only honoured for the very (x, y, context) objects it was made from
    # -- so at most one batch is ever referenced, and never a log-det computed under another context.
    def forward(self, x, context=None):
        self._memo = None
        y, ladj = self._apply_dir(x, False, context)
        self._memo = (x, y, context, ladj)
        return y

    def backward(self, y, context=None):
        self._memo = None
        x, ladj = self._apply_dir(y, True, context)
        self._memo = (x, y, context, ladj)
        return x

    def train(self, mode: bool = True):
        self._memo = None
        return super().train(mode)

    def inverse_and_ladj(self, y, context=None):
        x, ladj = self._apply_dir(y, True, context)
        return x, ladj

    def forward_and_ladj(self, x, context=None):
        return self._apply_dir(x, False, context)

    def log_abs_det_jacobian(self, x, y, context=None):
        memo, self._memo = self._memo, None
        if memo is not None and memo[0] is x and memo[1] is y and memo[2] is context:
            return memo[3]
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
