"""`FusedAdam`: torch.optim.Adam's update (`adbench_wrapper.py:369,391`) through `usf_adam_step` -- one launch per 32
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
        return loss
