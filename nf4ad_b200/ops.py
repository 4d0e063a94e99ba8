"""Tensor-level wrappers of the C ABI (forward kernels) and the autograd Functions of the training path.

Every function here enqueues hand-written sm_100a kernels through libusflow_b200.so on torch's current
stream; torch is used for device memory and autograd bookkeeping only.
"""
import os

import torch

from . import _lib
from ._lib import check, f32c, lib, ptr, require_cuda, stream


def _rows(t):
    """(tensor2d, leading dimension) with a unit inner stride."""
    t = f32c(t)
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    ld = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)
    return t, ld


# ------------------------------------------------------------------------------------------------
# forward kernels
# ------------------------------------------------------------------------------------------------
def linear(x, W, bias, relu=False):
    """y = x W^T + bias (optionally ReLU).  usf_linear."""
    require_cuda(x, W)
    x, ldx = _rows(x)
    W, ldw = _rows(W)
    B, K = x.shape
    N = W.shape[0]
    y = torch.empty(B, N, device=x.device, dtype=torch.float32)
    b = f32c(bias) if bias is not None else None
    check(lib().usf_linear(ptr(x), ldx, ptr(W), ldw, ptr(b), int(relu), ptr(y), N, B, N, K, stream()), "usf_linear")
    return y


def lu_pack(L_raw, U_raw):
    """W = (tril(L,-1)+I) triu(U).  usf_lu_pack."""
    require_cuda(L_raw, U_raw)
    L_raw, U_raw = f32c(L_raw).contiguous(), f32c(U_raw).contiguous()
    D = L_raw.shape[0]
    W = torch.empty(D, D, device=L_raw.device, dtype=torch.float32)
    scratch = torch.empty(2 * D * D, device=L_raw.device, dtype=torch.float32)
    check(lib().usf_lu_pack(ptr(L_raw), ptr(U_raw), D, ptr(W), None, ptr(scratch), stream()), "usf_lu_pack")
    return W


def lu_solve(y, L_raw, U_raw, bias=None, transpose=False):
    """x = (L U)^{-1} (y - bias) by triangular solves (transpose: (L U)^{-T}).  usf_lu_solve."""
    require_cuda(y, L_raw, U_raw)
    y, ldy = _rows(y)
    L_raw, U_raw = f32c(L_raw).contiguous(), f32c(U_raw).contiguous()
    B, D = y.shape
    x = torch.empty(B, D, device=y.device, dtype=torch.float32)
    b = f32c(bias) if bias is not None else None
    scratch = torch.empty(lib().usf_lu_solve_scratch_floats(D), device=y.device, dtype=torch.float32)
    check(lib().usf_lu_solve(ptr(y), ldy, ptr(L_raw), ptr(U_raw), ptr(b), int(transpose), ptr(x), D, B, D, ptr(scratch),
                             stream()), "usf_lu_solve")
    return x


_LU_T3_PRODUCT = os.environ.get("USF_LU_T3", "1") != "0"


def lu_inverse(L_raw, U_raw):
    """A = (L U)^{-1} as a dense matrix (usf_lu_inverse: both triangular inverses in one launch + one fp32 GEMM)."""
    require_cuda(L_raw, U_raw)
    D = L_raw.shape[0]
    Lc, Uc = f32c(L_raw).contiguous(), f32c(U_raw).contiguous()
    n = int(lib().usf_lu_inverse_scratch_floats(D))
    if n == 0:          # too large for the resident triangular solve: the blocked solve on the identity
        return lu_solve(torch.eye(D, device=Lc.device, dtype=torch.float32), Lc, Uc, None, transpose=True)
    scratch = torch.empty(n, device=Lc.device, dtype=torch.float32)
    if D % 16 == 0 and D >= 256 and _LU_T3_PRODUCT:
        # the product U^{-1} L^{-1} on the 3xTF32 tensor-core GEMM (fp32-grade; the FFMA GEMM takes 4x as long)
        check(lib().usf_lu_inverse(ptr(Lc), ptr(Uc), D, None, ptr(scratch), stream()), "usf_lu_inverse")
        Z, W = scratch[:D * D].view(D, D), scratch[D * D:2 * D * D].view(D, D)
        Zr, _, _ = to_t3(Z, want_rows=True)
        _, Wt, _ = to_t3(W, want_rows=False, want_transposed=True)
        return gemm_t3(Zr, Wt, D, D, D)
    A = torch.empty(D, D, device=Lc.device, dtype=torch.float32)
    check(lib().usf_lu_inverse(ptr(Lc), ptr(Uc), D, ptr(A), ptr(scratch), stream()), "usf_lu_inverse")
    return A


def householder(x, V, reverse=False):
    require_cuda(x, V)
    x, ldx = _rows(x)
    V = f32c(V).contiguous()
    B, D = x.shape
    y = torch.empty(B, D, device=x.device, dtype=torch.float32)
    check(lib().usf_householder(ptr(x), ldx, ptr(V), V.shape[0], int(reverse), ptr(y), D, B, D, stream()),
          "usf_householder")
    return y


def scale(x, s, inverse=False):
    require_cuda(x, s)
    x, ldx = _rows(x)
    s = f32c(s).contiguous()
    B, D = x.shape
    y = torch.empty(B, D, device=x.device, dtype=torch.float32)
    check(lib().usf_scale(ptr(x), ldx, ptr(s), int(inverse), ptr(y), D, B, D, stream()), "usf_scale")
    return y


SCALE_ACTIVATIONS = {"exp": 0, "softplus": 1}


def scale_activation_id(name):
    """`MaskedAffineCoupling.scale_activation` -> kernel flag; anything else is the reference's ValueError
    (`nf4ad/transforms.py:86-87`)."""
    try:
        return SCALE_ACTIVATIONS[name]
    except KeyError:
        raise ValueError("Unsupported scale_activation") from None


def coupling(x, s, t, mask, clamp, inverse, act=0):
    """Returns (y, ladj) with ladj[b] = sum_d (1-m_d) log_scale[b,d] (zeros if s is None); act: 0 exp, 1 softplus."""
    require_cuda(x, t, mask)
    x, ldx = _rows(x)
    t, ldt = _rows(t)
    if s is not None:
        s, lds = _rows(s)
    else:
        lds = 0
    mask = f32c(mask).reshape(-1).contiguous()
    B, D = x.shape
    y = torch.empty(B, D, device=x.device, dtype=torch.float32)
    ladj = torch.zeros(B, device=x.device, dtype=torch.float32)
    check(lib().usf_coupling(ptr(x), ldx, ptr(s), lds, ptr(t), ldt, ptr(mask), float(clamp), int(inverse), int(act), ptr(y), D,
                             ptr(ladj), 1.0, B, D, stream()), "usf_coupling")
    return y, ladj


def base_logprob(kind, z, loc, scale_t):
    require_cuda(z, loc, scale_t)
    z, ldz = _rows(z)
    loc = f32c(loc).reshape(-1).contiguous()
    sc = f32c(scale_t).reshape(-1).contiguous()
    B, D = z.shape
    out = torch.empty(B, device=z.device, dtype=torch.float32)
    check(lib().usf_base_logprob(int(kind), ptr(z), ldz, ptr(loc), ptr(sc), sc.numel(), None, 0.0, ptr(out), B, D,
                                 stream()), "usf_base_logprob")
    return out


def pack_matrix(src, row_idx, col_idx, n_rows, n_cols, ldo, sub_row0=False, transpose_src=False, want_f32=True,
                want_bf16=False):
    """out[r,c] = src[row_idx[r], col_idx[c]] (- src[0, col_idx[c]]); idx < 0 -> 0.  usf_pack_matrix."""
    require_cuda(src)
    src, lds = _rows(src)
    out = torch.empty(n_rows, ldo, device=src.device, dtype=torch.float32) if want_f32 else None
    outb = torch.empty(n_rows, ldo, device=src.device, dtype=torch.bfloat16) if want_bf16 else None
    check(lib().usf_pack_matrix(ptr(src), lds, ptr(row_idx), ptr(col_idx), int(sub_row0), int(transpose_src), n_rows, n_cols, ptr(out),
                                ptr(outb), ldo, stream()), "usf_pack_matrix")
    return out, outb


def _pad8(n):
    return (n + 7) // 8 * 8


def to_bf16(x, relu_mask=None, want_rows=True, want_transposed=False, want_colsum=False):
    """One pass over fp32 x (B,N): bf16 copy (B, pad8(N)), bf16 transposed copy (N, pad8(B)), column sums (N,).
    usf_to_bf16 (pads are zero so the buffers are valid TMA sources)."""
    require_cuda(x)
    x, ldx = _rows(x)
    B, N = x.shape
    dev = x.device
    rows = torch.empty(B, _pad8(N), device=dev, dtype=torch.bfloat16) if want_rows else None
    tr = torch.empty(N, _pad8(B), device=dev, dtype=torch.bfloat16) if want_transposed else None
    cs = torch.zeros(N, device=dev, dtype=torch.float32) if want_colsum else None
    m, ldm = (_rows(relu_mask) if relu_mask is not None else (None, 0))
    check(lib().usf_to_bf16(ptr(x), ldx, ptr(m), ldm, ptr(rows), _pad8(N), ptr(tr), _pad8(B), ptr(cs), B, N, stream()),
          "usf_to_bf16")
    return rows, tr, cs


def gemm_bf16(a, w, M, N, K, bias=None, relu=False):
    """fp32 (M,N) = act(a w^T + bias) on the tcgen05 GEMM; a: (M, ld>=K) bf16, w: (N, ld>=K) bf16, N % 16 == 0."""
    y = torch.empty(M, N, device=a.device, dtype=torch.float32)
    b = f32c(bias) if bias is not None else None
    check(lib().usf_linear_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(b), int(relu), ptr(y), N, 0, M, N, K,
                                stream()), "usf_linear_bf16")
    return y


def _pad4(n):
    return (n + 3) // 4 * 4


def to_t3(x, relu_mask=None, want_rows=True, want_transposed=False, want_colsum=False):
    """fp32 x (B,N) -> (hi, lo) pairs for the 3xTF32 GEMM: rows (B, pad4(N)) and / or transposed (N, pad4(B)), + column
    sums.  usf_to_tf32x3."""
    require_cuda(x)
    x, ldx = _rows(x)
    B, N = x.shape
    dev = x.device
    rh = torch.empty(B, _pad4(N), device=dev, dtype=torch.float32) if want_rows else None
    rl = torch.empty_like(rh) if want_rows else None
    th = torch.empty(N, _pad4(B), device=dev, dtype=torch.float32) if want_transposed else None
    tl = torch.empty_like(th) if want_transposed else None
    cs = torch.zeros(N, device=dev, dtype=torch.float32) if want_colsum else None
    m, ldm = (_rows(relu_mask) if relu_mask is not None else (None, 0))
    check(lib().usf_to_tf32x3(ptr(x), ldx, ptr(m), ldm, ptr(rh), ptr(rl), _pad4(N), ptr(th), ptr(tl), _pad4(B), ptr(cs), B, N,
                              stream()), "usf_to_tf32x3")
    return (rh, rl), (th, tl), cs


def gemm_t3(a, w, M, N, K, bias=None, relu=False):
    """fp32 (M,N) = act(a w^T + bias) on the 3xTF32 tensor-core GEMM; a = (hi, lo) of (M, ld>=K), w = (hi, lo) of
    (N, ld>=K); N % 16 == 0.  Column blocks of <= 1024 (the kernel keeps its per-column vectors resident)."""
    y = torch.empty(M, N, device=a[0].device, dtype=torch.float32)
    b = f32c(bias) if bias is not None else None
    for n0 in range(0, N, 1024):
        n1 = min(N, n0 + 1024)
        wh, wl = w[0][n0:n1], w[1][n0:n1]
        check(lib().usf_linear_tf32x3(ptr(a[0]), ptr(a[1]), a[0].stride(0), ptr(wh), ptr(wl), w[0].stride(0),
                                      ptr(b[n0:n1]) if b is not None else None, int(relu), ptr(y[:, n0:n1]), None, N, M,
                                      n1 - n0, K, stream()), "usf_linear_tf32x3")
    return y


# Mixed-precision training (flow.precision == "bf16" under autograd): the batch-sized GEMMs of the step run on the
# tcgen05 kernel with bf16 operands and fp32 accumulation, parameters / gradients / optimizer state stay fp32, and an
# LU layer is inverted once per step (weight space) so that no triangular solve ever sees the batch.
_TC_TRAIN = 0                 # 0: fp32 kernels, 1: bf16 tensor-core GEMMs, 2: 3xTF32 tensor-core GEMMs (fp32-grade)
_TC_WEIGHT_SPACE = True       # bf16 mode: the D^3 weight-space products run on the tensor cores as well


class tc_training:
    """Context manager set by `Flow` around the autograd-recording forward pass."""

    def __init__(self, on):
        self.on = int(on)          # False / 0, True / 1 (bf16), 2 (3xTF32)

    def __enter__(self):
        global _TC_TRAIN
        self.prev, _TC_TRAIN = _TC_TRAIN, self.on
        return self

    def __exit__(self, *exc):
        global _TC_TRAIN
        _TC_TRAIN = self.prev
        return False


def tc_train_enabled():
    return _TC_TRAIN != 0


def linear_fn(x, W, bias, relu=False, w_transposed=False, operands=None):
    """Autograd linear layer: tensor-core (bf16) GEMMs in mixed-precision training when the shapes allow TMA operands
    (N, K multiples of 16, batch multiple of 8), else the fp32 kernels.  `w_transposed`: `W` holds the map's transpose
    (K, N) -- y = x W + bias -- which is how a composed affine run arrives (`Flow._compose_affine_runs`); its gradient
    comes back in the same layout.  `operands`: the bf16 forms `(W rows, W^T rows)` of `W` if the caller converted them
    ahead of time (`weight_operands`, off the batch-sized chain)."""
    if _TC_TRAIN and x.is_cuda and x.dim() == 2 and x.shape[0] % 8 == 0 and x.shape[0] >= 8 \
            and W.shape[0] % 16 == 0 and W.shape[1] % 16 == 0:
        if _TC_TRAIN == 2:
            if x.shape[0] >= 256:        # below that the fp32 split-K kernels are as fast
                return LinearT3Fn.apply(x, W, bias, relu, w_transposed)
        else:
            W, defer = wgrad_deferred(x, W)
            return LinearTCFn.apply(x, W, bias, relu, w_transposed, operands, defer)
    return LinearFn.apply(x, W.t() if w_transposed else W, bias, relu)


# ------------------------------------------------------------------------------------------------
# autograd Functions (training path; fp32)
# ------------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = relu?(x W^T + b); backward = usf_linear_bwd (dgrad + wgrad GEMMs, bias column-sum)."""

    @staticmethod
    def forward(ctx, x, W, bias, relu):
        y = linear(x, W, bias, relu)
        ctx.relu = bool(relu)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, W, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy, lddy = _rows(dy)
        x, ldx = _rows(x)
        W, ldw = _rows(W)
        B, N = dy.shape
        K = W.shape[1]
        dev = dy.device
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty(B, K, device=dev, dtype=torch.float32) if need_x else None
        dW = torch.empty(N, K, device=dev, dtype=torch.float32) if need_w else None
        db = torch.empty(N, device=dev, dtype=torch.float32) if need_b else None
        scratch = torch.empty(B * N, device=dev, dtype=torch.float32) if ctx.relu else None
        yr, ldyr = (_rows(y) if ctx.relu else (None, 0))
        check(lib().usf_linear_bwd(ptr(dy), lddy, ptr(x), ldx, ptr(W), ldw, ptr(yr), ldyr, ptr(dx), K, ptr(dW), K,
                                   ptr(db), 0, ptr(scratch), B, N, K, stream()), "usf_linear_bwd")
        return dx, dW, db, None


_WGRAD_STREAMS = {}
_WGRAD_OVERLAP = os.environ.get("USF_WGRAD_OVERLAP", "1") != "0"
_WGRAD_DEFER = os.environ.get("USF_WGRAD_DEFER", "1") != "0"
_WGRAD_OVERLAP_MIN_ROWS = 1024     # weight-space products (D or 8 rows) are too short to be worth a fork / join


def _wgrad_stream(dev):
    """The partner stream of the current stream of `dev` for weight-gradient GEMMs.  One per calling stream: the
    weight-space chains of different affine runs (each on its own side stream) must not meet on a shared one -- every
    fork / join there would order them after each other."""
    cur = torch.cuda.current_stream(dev)
    key = (dev.index, cur.cuda_stream)
    side = _WGRAD_STREAMS.get(key)
    if side is None:
        side = _WGRAD_STREAMS[key] = _lib.new_stream(dev)
    return cur, side


def _overlapped_wgrad(fn, operands):
    """The weight-gradient GEMM of a linear layer reduces over the batch into a small (N, K) output -- 16-64 tiles, a
    fraction of the GPU -- and nothing downstream in the backward chain needs it.  It is issued on a side stream while
    the input-gradient GEMM runs on the current one (fork after the operands are ready); both GEMMs then share the SMs.
    -> (dW, side stream): the caller joins the stream before returning -- or, if the weight went through
    `_WgradSide`, not at all.  Survives CUDA-graph capture."""
    dev = operands[0].device
    cur, side = _wgrad_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = fn()
    for t in operands:
        if t is not None:
            t.record_stream(side)
    out.record_stream(cur)
    return out, side


class _WgradSide(torch.autograd.Function):
    """Identity on a weight, applied with the weight-gradient partner stream current.  Autograd runs a node's backward
    on the stream its forward ran on and synchronises the CONSUMERS of its result with that stream: with this node
    between a weight and `LinearTCFn`, the weight gradient that `LinearTCFn.backward` leaves running on the partner
    stream is handed on from that very stream -- ordered behind the GEMM -- and the batch-sized backward chain never
    waits for a weight-gradient GEMM (28 us against the 10-17 us of the input-gradient GEMM it ran beside)."""

    @staticmethod
    def forward(ctx, W):
        return W.view_as(W)

    @staticmethod
    def backward(ctx, g):
        return g


def wgrad_deferred(x, W):
    """-> (W routed through `_WgradSide`, True) if this layer's weight-gradient GEMM may trail behind the backward chain,
    else (W, False)."""
    if not (_WGRAD_OVERLAP and _WGRAD_DEFER and x.is_cuda and x.shape[0] >= _WGRAD_OVERLAP_MIN_ROWS
            and torch.is_grad_enabled() and W.requires_grad and x.requires_grad):
        return W, False
    _, side = _wgrad_stream(x.device)
    with torch.cuda.stream(side):
        return _WgradSide.apply(W), True


class LinearTCFn(torch.autograd.Function):
    """y = relu?(x W^T + b) with bf16 tensor-core GEMMs (fp32 accumulate) forward and backward:
        y  = x  W^T        A = x   (B,K),  W-operand = W   (N,K)
        dx = dy W          A = dy  (B,N),  W-operand = W^T (K,N)
        dW = dy^T x        A = dy^T (N,B), W-operand = x^T (K,B)     (reduces over the batch)
    One `usf_to_bf16` pass per fp32 matrix produces the bf16 row-major / transposed operands (and, for dy, the ReLU
    gating and the bias gradient)."""

    @staticmethod
    def forward(ctx, x, W, bias, relu, wt=False, operands=None, defer=False):
        B, K = x.shape
        N = W.shape[1] if wt else W.shape[0]
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ctx.defer = bool(defer)          # W came through _WgradSide: its gradient is handed on from the partner stream
        xb, xT, _ = to_bf16(x, want_rows=True, want_transposed=need_w)
        if operands is not None:
            Wb, WT = operands
        elif wt:     # W^T (K, N) given: its rows are the dgrad operand, its transpose the forward operand
            WT, Wb, _ = to_bf16(W, want_rows=need_x, want_transposed=True)
        else:
            Wb, WT, _ = to_bf16(W, want_rows=True, want_transposed=need_x)
        y = gemm_bf16(xb, Wb, B, N, K, bias, relu)
        ctx.relu, ctx.has_bias, ctx.shape, ctx.wt = bool(relu), bias is not None, (B, N, K), bool(wt)
        ctx.overlap = _WGRAD_OVERLAP and B >= _WGRAD_OVERLAP_MIN_ROWS
        ctx.save_for_backward(xT, WT, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xT, WT, y = ctx.saved_tensors
        B, N, K = ctx.shape
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dyb, dyT, db = to_bf16(dy, relu_mask=y if ctx.relu else None, want_rows=need_x, want_transposed=need_w,
                               want_colsum=need_b)
        # dW = dy^T x (N, K), or in the transposed layout d(W^T) = x^T dy (K, N)
        wgrad = (lambda: gemm_bf16(xT, dyT, K, N, B)) if ctx.wt else (lambda: gemm_bf16(dyT, xT, N, K, B))
        join = None
        if need_w and need_x and ctx.overlap:
            dW, join = _overlapped_wgrad(wgrad, (dyT, xT))
        else:
            dW = wgrad() if need_w else None
        dx = gemm_bf16(dyb, WT, B, K, N) if need_x else None
        if join is not None and not ctx.defer:
            torch.cuda.current_stream(dy.device).wait_stream(join)
        return dx, dW, db, None, None, None, None


def weight_operands(W, w_transposed=False):
    """The bf16 operand pair `LinearTCFn` builds from a weight: (W rows (N, K), W^T rows (K, N))."""
    if w_transposed:
        WT, Wb, _ = to_bf16(W, want_rows=True, want_transposed=True)
    else:
        Wb, WT, _ = to_bf16(W, want_rows=True, want_transposed=True)
    return Wb, WT


class MaskFn(torch.autograd.Function):
    """x * mask for a constant 0/1 mask: the same element-wise kernel forward and backward (no parameter gradient)."""

    @staticmethod
    def forward(ctx, x, mask):
        ctx.save_for_backward(mask)
        return scale(x, mask, False)

    @staticmethod
    def backward(ctx, dy):
        mask, = ctx.saved_tensors
        return scale(dy, mask, False), None


class LinearT3Fn(torch.autograd.Function):
    """LinearTCFn with 3xTF32 GEMMs: fp32-grade forward / dgrad / wgrad on the tensor cores."""

    @staticmethod
    def forward(ctx, x, W, bias, relu, wt=False):
        B, K = x.shape
        N = W.shape[1] if wt else W.shape[0]
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        xr, xT, _ = to_t3(x, want_rows=True, want_transposed=need_w)
        if wt:
            WT, Wr, _ = to_t3(W, want_rows=need_x, want_transposed=True)
        else:
            Wr, WT, _ = to_t3(W, want_rows=True, want_transposed=need_x)
        y = gemm_t3(xr, Wr, B, N, K, bias, relu)
        ctx.relu, ctx.has_bias, ctx.shape, ctx.wt = bool(relu), bias is not None, (B, N, K), bool(wt)
        ctx.save_for_backward(xT[0], xT[1], WT[0], WT[1], y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xTh, xTl, WTh, WTl, y = ctx.saved_tensors
        B, N, K = ctx.shape
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dyr, dyT, db = to_t3(dy, relu_mask=y if ctx.relu else None, want_rows=need_x, want_transposed=need_w,
                             want_colsum=need_b)
        wgrad = (lambda: gemm_t3((xTh, xTl), dyT, K, N, B)) if ctx.wt else (lambda: gemm_t3(dyT, (xTh, xTl), N, K, B))
        join = None
        if need_w and need_x and _WGRAD_OVERLAP and B >= _WGRAD_OVERLAP_MIN_ROWS:
            dW, join = _overlapped_wgrad(wgrad, (dyT[0], dyT[1], xTh, xTl))
        else:
            dW = wgrad() if need_w else None
        dx = gemm_t3(dyr, (WTh, WTl), B, K, N) if need_x else None
        if join is not None:
            torch.cuda.current_stream(dy.device).wait_stream(join)
        return dx, dW, db, None, None


class LUInverseFn(torch.autograd.Function):
    """A = (L U)^{-1} as a dense matrix, once per step (weight space): the triangular-solve kernel applied to the
    identity.  Backward: dW = -A^T dA A^T, then the factor gradients through usf_lu_pack_bwd."""

    @staticmethod
    def forward(ctx, L_raw, U_raw):
        D = L_raw.shape[0]
        A = lu_inverse(L_raw, U_raw)
        ctx.tc = bool(_TC_TRAIN == 1 and _TC_WEIGHT_SPACE)
        ctx.save_for_backward(A, L_raw, U_raw)
        return A

    @staticmethod
    def backward(ctx, dA):
        A, L_raw, U_raw = ctx.saved_tensors
        D = A.shape[0]
        dev = A.device
        dA = f32c(dA).contiguous()
        P = torch.empty(D, D, device=dev, dtype=torch.float32)
        dW = torch.empty(D, D, device=dev, dtype=torch.float32)
        # P = A^T dA ;  dW = -(P A^T)
        if ctx.tc and D % 16 == 0:
            # mixed precision: the two D^3 products on the tensor cores as well (bf16 operands, fp32 accumulate)
            Ab, At, _ = to_bf16(A, want_rows=True, want_transposed=True)      # A rows, A^T rows: one pass
            _, dAt, _ = to_bf16(dA, want_rows=False, want_transposed=True)    # dA^T rows: operand [n, k] = dA[k, n]
            P = gemm_bf16(At, dAt, D, D, D)
            Pb, _, _ = to_bf16(P, want_rows=True)
            dW = gemm_bf16(Pb, Ab, D, D, D)
        else:
            check(lib().usf_gemm(ptr(A), D, 1, ptr(dA), D, 1, ptr(P), D, 0, D, D, D, stream()), "usf_gemm")
            check(lib().usf_gemm(ptr(P), D, 0, ptr(A), D, 0, ptr(dW), D, 0, D, D, D, stream()), "usf_gemm")
        if ctx.tc and D % 16 == 0:
            Lt = torch.tril(L_raw.detach(), -1)
            Lt.diagonal().fill_(1.0)
            Ut = torch.triu(U_raw.detach())
            dWb, dWT, _ = to_bf16(dW, want_rows=True, want_transposed=True)
            Ub, _, _ = to_bf16(Ut, want_rows=True)
            _, LtT, _ = to_bf16(Lt, want_rows=False, want_transposed=True)
            dL = torch.tril(gemm_bf16(dWb, Ub, D, D, D), -1)
            dU = torch.triu(gemm_bf16(LtT, dWT, D, D, D))
            return dL.neg_(), dU.neg_()
        Lc, Uc = f32c(L_raw).contiguous(), f32c(U_raw).contiguous()
        dL = torch.zeros(D, D, device=dev, dtype=torch.float32)
        dU = torch.zeros(D, D, device=dev, dtype=torch.float32)
        scratch = torch.empty(3 * D * D, device=dev, dtype=torch.float32)
        check(lib().usf_lu_pack_bwd(ptr(dW), ptr(Lc), ptr(Uc), 0.0, D, ptr(dL), ptr(dU), ptr(scratch), stream()),
              "usf_lu_pack_bwd")
        return dL.neg_(), dU.neg_()


class LUPackFn(torch.autograd.Function):
    """W = L U from the raw factors; backward masks the gradient onto the two triangles.  In mixed-precision training
    the three D^3 products (W = Lt Ut, dLt = dW Ut^T, dUt = Lt^T dW) run on the tensor cores like every other GEMM of
    the step (bf16 operands, fp32 accumulate); the triangle masks are weight-space tensor ops."""

    @staticmethod
    def forward(ctx, L_raw, U_raw):
        D = L_raw.shape[0]
        ctx.tc = bool(_TC_TRAIN == 1 and _TC_WEIGHT_SPACE and L_raw.is_cuda and D % 16 == 0)
        ctx.save_for_backward(L_raw, U_raw)
        if not ctx.tc:
            return lu_pack(L_raw, U_raw)
        Lt = torch.tril(L_raw.detach(), -1)
        Lt.diagonal().fill_(1.0)
        Ut = torch.triu(U_raw.detach())
        Lb, _, _ = to_bf16(Lt, want_rows=True)
        _, UtT, _ = to_bf16(Ut, want_rows=False, want_transposed=True)      # operand [n, k] = Ut[k, n]
        return gemm_bf16(Lb, UtT, D, D, D)

    @staticmethod
    def backward(ctx, dW):
        L_raw, U_raw = ctx.saved_tensors
        D = L_raw.shape[0]
        dW = f32c(dW).contiguous()
        if ctx.tc:
            Lt = torch.tril(L_raw.detach(), -1)
            Lt.diagonal().fill_(1.0)
            Ut = torch.triu(U_raw.detach())
            dWb, dWT, _ = to_bf16(dW, want_rows=True, want_transposed=True)
            Ub, _, _ = to_bf16(Ut, want_rows=True)
            _, LtT, _ = to_bf16(Lt, want_rows=False, want_transposed=True)
            dL = torch.tril(gemm_bf16(dWb, Ub, D, D, D), -1)               # dW Ut^T, strictly lower part
            dU = torch.triu(gemm_bf16(LtT, dWT, D, D, D))                  # Lt^T dW, upper part
            return dL, dU
        Lc, Uc = f32c(L_raw).contiguous(), f32c(U_raw).contiguous()
        dL = torch.zeros(D, D, device=dW.device, dtype=torch.float32)
        dU = torch.zeros(D, D, device=dW.device, dtype=torch.float32)
        scratch = torch.empty(3 * D * D, device=dW.device, dtype=torch.float32)
        check(lib().usf_lu_pack_bwd(ptr(dW), ptr(Lc), ptr(Uc), 0.0, D, ptr(dL), ptr(dU), ptr(scratch), stream()),
              "usf_lu_pack_bwd")
        return dL, dU


class LUSolveFn(torch.autograd.Function):
    """x = (L U)^{-1}(y - b) by triangular solves; backward solves with the transposed factors."""

    @staticmethod
    def forward(ctx, y, L_raw, U_raw, bias):
        x = lu_solve(y, L_raw, U_raw, bias)
        ctx.save_for_backward(x, L_raw, U_raw)
        return x

    @staticmethod
    def backward(ctx, dx):
        x, L_raw, U_raw = ctx.saved_tensors
        D = L_raw.shape[0]
        g = lu_solve(dx, L_raw, U_raw, None, transpose=True)     # dL/dy = (LU)^{-T} dx
        B = g.shape[0]
        dev = g.device
        dL = dU = db = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            # dW_eff[i,j] = -sum_b g[b,i] x[b,j]
            dW = torch.empty(D, D, device=dev, dtype=torch.float32)
            xn = x.neg()
            check(lib().usf_gemm(ptr(g), D, 1, ptr(xn), D, 1, ptr(dW), D, 0, D, D, B, stream()), "usf_gemm")
            Lc, Uc = f32c(L_raw).contiguous(), f32c(U_raw).contiguous()
            dL = torch.zeros(D, D, device=dev, dtype=torch.float32)
            dU = torch.zeros(D, D, device=dev, dtype=torch.float32)
            scratch = torch.empty(3 * D * D, device=dev, dtype=torch.float32)
            check(lib().usf_lu_pack_bwd(ptr(dW), ptr(Lc), ptr(Uc), 0.0, D, ptr(dL), ptr(dU), ptr(scratch), stream()),
                  "usf_lu_pack_bwd")
        if ctx.needs_input_grad[3]:
            db = torch.empty(D, device=dev, dtype=torch.float32)
            check(lib().usf_colsum(ptr(g), D, -1.0, 0, ptr(db), B, D, stream()), "usf_colsum")
        return g, dL, dU, db


class HouseholderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, V, reverse):
        ctx.reverse = bool(reverse)
        ctx.save_for_backward(x, V)
        return householder(x, V, reverse)

    @staticmethod
    def backward(ctx, dy):
        x, V = ctx.saved_tensors
        dy, lddy = _rows(dy)
        x, ldx = _rows(x)
        Vc = f32c(V).contiguous()
        B, D = dy.shape
        nvs = Vc.shape[0]
        dx = torch.empty(B, D, device=dy.device, dtype=torch.float32)
        dV = torch.zeros(nvs, D, device=dy.device, dtype=torch.float32)
        scratch = torch.empty(max(nvs, 1) * B * D, device=dy.device, dtype=torch.float32) if nvs > 1 else None
        check(lib().usf_householder_bwd(ptr(dy), lddy, ptr(x), ldx, ptr(Vc), nvs, int(ctx.reverse), ptr(dx), D, ptr(dV),
                                        ptr(scratch), B, D, stream()), "usf_householder_bwd")
        return dx, dV, None


class ScaleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, s, inverse):
        y = scale(x, s, inverse)
        ctx.inverse = bool(inverse)
        ctx.save_for_backward(y if inverse else x, s)
        return y

    @staticmethod
    def backward(ctx, dy):
        xy, s = ctx.saved_tensors
        dy, lddy = _rows(dy)
        xy, ldxy = _rows(xy)
        sc = f32c(s).reshape(-1).contiguous()
        B, D = dy.shape
        dx = torch.empty(B, D, device=dy.device, dtype=torch.float32)
        ds = torch.zeros(D, device=dy.device, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        check(lib().usf_scale_bwd(ptr(dy), lddy, ptr(xy), ldxy, ptr(sc), int(ctx.inverse), ptr(dx), D, ptr(ds), B, D,
                                  stream()), "usf_scale_bwd")
        return dx, (ds.reshape(s.shape) if ds is not None else None), None


class CouplingFn(torch.autograd.Function):
    """(y, ladj) = coupling(x, s, t); s may be None (additive)."""

    @staticmethod
    def forward(ctx, x, s, t, mask, clamp, inverse, act=0):
        y, ladj = coupling(x, s, t, mask, clamp, inverse, act)
        ctx.clamp, ctx.inverse, ctx.has_s, ctx.act = float(clamp), bool(inverse), s is not None, int(act)
        ctx.save_for_backward(x, s, t, mask)
        return y, ladj

    @staticmethod
    def backward(ctx, dy, dladj):
        x, s, t, mask = ctx.saved_tensors
        B, D = x.shape
        dev = x.device
        dy = torch.zeros(B, D, device=dev, dtype=torch.float32) if dy is None else dy
        dy, lddy = _rows(dy)
        x, ldx = _rows(x)
        t, ldt = _rows(t)
        if ctx.has_s:
            s, lds = _rows(s)
        else:
            lds = 0
        m = f32c(mask).reshape(-1).contiguous()
        dl = f32c(dladj).contiguous() if (dladj is not None and ctx.has_s) else None
        dx = torch.empty(B, D, device=dev, dtype=torch.float32)
        dt = torch.empty(B, D, device=dev, dtype=torch.float32)
        ds = torch.empty(B, D, device=dev, dtype=torch.float32) if ctx.has_s else None
        check(lib().usf_coupling_bwd(ptr(dy), lddy, ptr(dl), 1.0, ptr(x), ldx, ptr(s), lds, ptr(t), ldt, ptr(m),
                                     ctx.clamp, int(ctx.inverse), ctx.act, ptr(dx), D, ptr(ds), D, ptr(dt), D, B, D, stream()),
              "usf_coupling_bwd")
        return dx, ds, dt, None, None, None, None


class CouplingPackedFn(torch.autograd.Function):
    """`CouplingFn` for a conditioner output that is ONE `[s | t]` tensor (B, 2 D): the kernels read the two halves in
    place and the backward pass writes `d[s | t]` as one tensor -- no slice / zero-fill / copy / add nodes around it."""

    @staticmethod
    def forward(ctx, x, st, mask, clamp, inverse, act=0):
        D = x.shape[1]
        st, _ = _rows(st)
        y, ladj = coupling(x, st[:, :D], st[:, D:], mask, clamp, inverse, act)
        ctx.clamp, ctx.inverse, ctx.act = float(clamp), bool(inverse), int(act)
        ctx.save_for_backward(x, st, mask)
        return y, ladj

    @staticmethod
    def backward(ctx, dy, dladj):
        x, st, mask = ctx.saved_tensors
        B, D = x.shape
        dev = x.device
        dy = torch.zeros(B, D, device=dev, dtype=torch.float32) if dy is None else dy
        dy, lddy = _rows(dy)
        x, ldx = _rows(x)
        ldst = st.stride(0) if B > 1 else 2 * D
        s, t = st[:, :D], st[:, D:]
        m = f32c(mask).reshape(-1).contiguous()
        dl = f32c(dladj).contiguous() if dladj is not None else None
        dx = torch.empty(B, D, device=dev, dtype=torch.float32)
        dst = torch.empty(B, 2 * D, device=dev, dtype=torch.float32)
        check(lib().usf_coupling_bwd(ptr(dy), lddy, ptr(dl), 1.0, ptr(x), ldx, ptr(s), ldst, ptr(t), ldst, ptr(m),
                                     ctx.clamp, int(ctx.inverse), ctx.act, ptr(dx), D, ptr(dst[:, :D]), 2 * D,
                                     ptr(dst[:, D:]), 2 * D, B, D, stream()), "usf_coupling_bwd")
        return dx, dst, None, None, None, None


class BaseLogProbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, loc, scale_t, kind):
        ctx.kind = int(kind)
        ctx.save_for_backward(z, loc, scale_t)
        return base_logprob(kind, z, loc, scale_t)

    @staticmethod
    def backward(ctx, dout):
        z, loc, scale_t = ctx.saved_tensors
        z, ldz = _rows(z)
        B, D = z.shape
        dev = z.device
        locc = f32c(loc).reshape(-1).contiguous()
        sc = f32c(scale_t).reshape(-1).contiguous()
        dout = f32c(dout).contiguous()
        dz = torch.empty(B, D, device=dev, dtype=torch.float32)
        dloc = torch.zeros(D, device=dev, dtype=torch.float32) if ctx.needs_input_grad[1] else None
        dsc = torch.zeros(sc.numel(), device=dev, dtype=torch.float32) if ctx.needs_input_grad[2] else None
        check(lib().usf_base_logprob_bwd(ctx.kind, ptr(dout), ptr(z), ldz, ptr(locc), ptr(sc), sc.numel(), ptr(dz), D,
                                         ptr(dloc), ptr(dsc), B, D, stream()), "usf_base_logprob_bwd")
        return (dz, dloc.reshape(loc.shape) if dloc is not None else None,
                dsc.reshape(scale_t.shape) if dsc is not None else None, None)
