"""VAE-flow latent tail: the element-wise steps either side of `flow_prior.log_prob` in
`/root/reference/src/nf4ad/vaeflow.py` -- `reparameterize` (:176-179), the posterior log-density `log q(z|x)` (:214-219),
the reconstruction NLL (:208-211, :255-259) -- as single launches with the encoder / decoder facing gradients, and the
two reductions the reference builds from them (`loss_function` :198-235, `anomaly_score` :252-269).

The encoder / decoder networks themselves are out of scope (cuDNN conv nets); these functions take their outputs.
"""
import math

import torch

from . import _lib
from ._lib import check, f32c, lib, ptr, require_cuda, stream


def _rows2(t):
    t = f32c(t)
    t2 = t.reshape(t.shape[0], -1)
    if t2.stride(-1) != 1 or (t2.shape[0] > 1 and t2.stride(0) < t2.shape[1]):
        t2 = t2.contiguous()
    return t2, (t2.stride(0) if t2.shape[0] > 1 else max(t2.shape[1], 1))


class ReparamFn(torch.autograd.Function):
    """(z, log_q) = usf_vae_reparam(mu, logvar, eps); dmu = dz, dlogvar = usf_vae_reparam_bwd."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        require_cuda(mu, logvar, eps)
        m, ldm = _rows2(mu)
        v, ldv = _rows2(logvar)
        e, lde = _rows2(eps)
        B, L = m.shape
        z = torch.empty(B, L, device=m.device, dtype=torch.float32)
        log_q = torch.empty(B, device=m.device, dtype=torch.float32)
        check(lib().usf_vae_reparam(ptr(m), ldm, ptr(v), ldv, ptr(e), lde, ptr(z), L, ptr(log_q), B, L, stream()),
              "usf_vae_reparam")
        ctx.save_for_backward(v, e)
        ctx.shape = mu.shape
        return z.view(mu.shape), log_q

    @staticmethod
    def backward(ctx, dz, dlog_q):
        v, e = ctx.saved_tensors
        B, L = v.shape
        dmu = dz if ctx.needs_input_grad[0] else None
        dlv = None
        if ctx.needs_input_grad[1]:
            dz2, lddz = (_rows2(dz) if dz is not None else (None, 0))
            dq = f32c(dlog_q).contiguous() if dlog_q is not None else None
            dlv = torch.empty(B, L, device=v.device, dtype=torch.float32)
            check(lib().usf_vae_reparam_bwd(ptr(dz2), lddz, ptr(dq), ptr(e), e.stride(0) if B > 1 else L, ptr(v),
                                            v.stride(0) if B > 1 else L, ptr(dlv), L, B, L, stream()), "usf_vae_reparam_bwd")
            dlv = dlv.view(ctx.shape)
        return dmu, dlv, None


class ReconNLLFn(torch.autograd.Function):
    """Per-sample Gaussian reconstruction NLL with fixed variance sigma2 (usf_recon_nll / usf_recon_nll_bwd)."""

    @staticmethod
    def forward(ctx, x, x_recon, sigma2):
        require_cuda(x, x_recon)
        if x.shape != x_recon.shape:
            raise ValueError(f"x {tuple(x.shape)} and its reconstruction {tuple(x_recon.shape)} differ in shape")
        a = f32c(x).reshape(x.shape[0], -1).contiguous()
        b = f32c(x_recon).reshape(x.shape[0], -1).contiguous()
        B, D = a.shape
        out = torch.empty(B, device=a.device, dtype=torch.float32)
        check(lib().usf_recon_nll(ptr(a), ptr(b), B, D, float(sigma2), ptr(out), stream()), "usf_recon_nll")
        ctx.save_for_backward(a, b)
        ctx.sigma2, ctx.shape = float(sigma2), x.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        B, D = a.shape
        g = f32c(dout).contiguous()
        dx = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        dxr = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        if dx is not None or dxr is not None:
            check(lib().usf_recon_nll_bwd(ptr(a), ptr(b), ptr(g), B, D, ctx.sigma2, ptr(dxr), ptr(dx), stream()),
                  "usf_recon_nll_bwd")
        return (dx.view(ctx.shape) if dx is not None else None, dxr.view(ctx.shape) if dxr is not None else None, None)


def reparameterize(mu, logvar, eps=None):
    """`VAEFlow.reparameterize` (vaeflow.py:176-179) plus, from the same pass, log q(z|x) per sample: returns (z, log_q)."""
    if eps is None:
        eps = torch.randn_like(mu)
    return ReparamFn.apply(mu, logvar, eps)


def recon_nll(x, x_recon, sigma_min):
    """0.5 |x - x_recon|^2 / sigma^2 + 0.5 D log(2 pi sigma^2), per sample (vaeflow.py:203-211)."""
    return ReconNLLFn.apply(x, x_recon, float(sigma_min) ** 2)


def loss_function(flow_prior, x, x_recon, mu, logvar, z, log_q, sigma_min, beta=1.0, prior_shape=None):
    """`VAEFlow.loss_function` (vaeflow.py:198-235): sum over the batch of recon NLL + beta (log q(z|x) - log p(z));
    `z, log_q` from `reparameterize` (the reference recomputes log q from mu, std and z; same value)."""
    zf = z if prior_shape is None else z.reshape(z.shape[0], *prior_shape)
    log_p = flow_prior.log_prob(zf)
    if log_p.dim() == 0:
        log_p = log_p.repeat(x.shape[0])
    return recon_nll(x, x_recon, sigma_min).sum() + beta * (log_q - log_p).sum()


def anomaly_score(flow_prior, x, x_recon, z, sigma_min, prior_shape=None):
    """`VAEFlow.anomaly_score` (vaeflow.py:252-269): recon NLL - log p(z), per sample."""
    zf = z if prior_shape is None else z.reshape(z.shape[0], *prior_shape)
    return recon_nll(x, x_recon, sigma_min) - flow_prior.log_prob(zf)
