"""`FusedAdam` / `SophiaG`.  `FusedAdam`: torch.optim.Adam's update (`adbench_wrapper.py:369,391`) through `usf_adam_step` -- one launch per 32
parameter tensors, step count on the device, so the whole training step replays as one CUDA graph
(`DataParallelTrainer` treats it as `capturable`)."""
import ctypes as C

import torch

from . import _lib


class _AdamTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64)]


class FusedAdam(torch.optim.Optimizer):
    """Adam / AdamW (`decoupled=True`) for fp32 CUDA parameters.  Same arithmetic as torch's (no amsgrad / maximize).

    `clip_coef` (optional device scalar tensor) multiplies every gradient inside the update -- the clipping coefficient
    of `clip_grad_norm_` without its extra pass over the gradients (`DataParallelTrainer` sets it)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        defaults = dict(lr=float(lr), betas=tuple(betas), eps=float(eps), weight_decay=float(weight_decay),
                        decoupled=bool(decoupled), capturable=True)
        super().__init__(params, defaults)
        self.clip_coef = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            _lib.require_cuda(*live)
            step = group.get("step")
            if step is None:
                step = group["step"] = torch.zeros((), device=live[0].device, dtype=torch.float32)
            step.add_(1.0)
            arr = (_AdamTensor * len(live))()
            for i, p in enumerate(live):
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad
                if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous():
                    raise _lib.USFError("FusedAdam needs contiguous fp32 parameters and gradients")
                arr[i] = _AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                                     p.numel())
            b1, b2 = group["betas"]
            _lib.check(_lib.lib().usf_adam_step(arr, len(live), _lib.ptr(step), group["lr"], b1, b2, group["eps"],
                                                group["weight_decay"], int(group["decoupled"]),
                                                _lib.ptr(self.clip_coef), _lib.stream()), "usf_adam_step")
        _lib.bump_weights_epoch()      # the kernel wrote the parameters through raw pointers: packed copies are stale
        return loss


class SophiaG(torch.optim.Optimizer):
    """`src.usflows.sophia.SophiaG` (the optimizer `experiments/gmm/gaussian_mixture_standart_base.yaml:45` names) as one
    fused, capturable kernel per 32 tensors (`usf_sophia_step`): Sophia-G of Liu et al. 2023 --
    `m = b1 m + (1-b1) g`, `p *= 1 - lr wd`, `p -= lr sign(m) min(|m| / (rho bs h + 1e-15), 1)` -- with the diagonal
    Hessian estimate `h = b2 h + (1-b2) g^2` refreshed every `hessian_interval` steps from the mini-batch gradient (the
    Gauss-Newton-Bartlett estimator evaluated on the batch itself; upstream's `update_hessian()` call pattern cannot be
    pinned from the reference tree).  `bs` is the batch size of the step (`note_batch`, default the paper's 5120)."""

    def __init__(self, params, lr=1e-4, betas=(0.965, 0.99), rho=0.04, weight_decay=1e-1, hessian_interval=10, bs=5120):
        if lr < 0 or rho < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid SophiaG hyper-parameters")
        defaults = dict(lr=float(lr), betas=tuple(betas), rho=float(rho), weight_decay=float(weight_decay),
                        hessian_interval=int(hessian_interval), capturable=True)
        super().__init__(params, defaults)
        self.bs = float(bs)
        self.clip_coef = None

    def note_batch(self, n):
        self.bs = float(n)

    @torch.no_grad()
    def step(self, closure=None, bs=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if bs is not None:
            self.bs = float(bs)
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                continue
            _lib.require_cuda(*live)
            step = group.get("step")
            if step is None:
                step = group["step"] = torch.zeros((), device=live[0].device, dtype=torch.float32)
            step.add_(1.0)
            arr = (_AdamTensor * len(live))()
            for i, p in enumerate(live):
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["hessian"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                g = p.grad
                if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous() or not g.is_contiguous():
                    raise _lib.USFError("SophiaG needs contiguous fp32 parameters and gradients")
                arr[i] = _AdamTensor(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["hessian"].data_ptr(), p.numel())
            b1, b2 = group["betas"]
            _lib.check(_lib.lib().usf_sophia_step(arr, len(live), _lib.ptr(step), group["lr"], b1, b2, group["rho"], self.bs,
                                                  group["weight_decay"], group["hessian_interval"],
                                                  _lib.ptr(self.clip_coef), _lib.stream()), "usf_sophia_step")
        _lib.bump_weights_epoch()
        return loss
