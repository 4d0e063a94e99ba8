"""ORACLE: affine masked coupling, restating `/root/reference/src/nf4ad/transforms.py`."""
import torch
import torch.nn.functional as F
from pyro import distributions as dist
from src.usflows.transforms import BaseTransform


class MaskedAffineCoupling(BaseTransform):
    bijective = True

    def __init__(self, mask, conditioner, scale_activation="exp", clamp=5.0):
        # transforms.py:24-38
        super().__init__()
        self.register_buffer("mask", mask.float())
        self.conditioner = conditioner
        self.scale_activation = scale_activation
        self.clamp = float(clamp)
        self.domain = dist.constraints.real_vector
        self.codomain = dist.constraints.real_vector

    # transforms.py:40-64 -- three accepted conditioner output forms
    def _split(self, out, ref):
        if isinstance(out, (list, tuple)) and len(out) == 2:
            s, t = out
        elif out.shape == ref.shape:
            s, t = torch.zeros_like(out), out
        elif out.dim() >= 2 and ref.dim() >= 2 and out.shape[1] == 2 * ref.shape[1]:
            c = ref.shape[1]
            s, t = out[:, :c, ...], out[:, c:, ...]
        else:
            raise ValueError(
                "Conditioner output shape not compatible. "
                "Expected (s,t) tuple, tensor same shape as x, or tensor with 2*C channels.")
        return s.to(ref.dtype), t.to(ref.dtype)

    def _params(self, v, context):
        vm = v * self.mask
        out = self.conditioner(vm) if context is None else self.conditioner(vm, context)
        s, t = self._split(out, v)
        return vm, torch.tanh(s) * self.clamp, t      # transforms.py:79,103,127

    def _scale(self, log_s):
        if self.scale_activation == "exp":
            return torch.exp(log_s)
        if self.scale_activation == "softplus":
            return F.softplus(log_s) + 1e-6            # transforms.py:84,107
        raise ValueError("Unsupported scale_activation")

    def forward(self, x, context=None):
        # transforms.py:66-90
        xm, log_s, t = self._params(x, context)
        return xm + (1.0 - self.mask) * (x * self._scale(log_s) + t)

    def backward(self, y, context=None):
        # transforms.py:92-113 (the +1e-12 is kept literally)
        ym, log_s, t = self._params(y, context)
        return ym + (1.0 - self.mask) * ((y - t) / (self._scale(log_s) + 1e-12))

    def log_abs_det_jacobian(self, x, y, context=None):
        # transforms.py:115-140 -- the conditioner is evaluated again on x*m
        _, log_s, _ = self._params(x, context)
        if self.scale_activation == "exp":
            log_scale = log_s
        elif self.scale_activation == "softplus":
            log_scale = torch.log(F.softplus(log_s) + 1e-12)   # transforms.py:132
        else:
            raise ValueError("Unsupported scale_activation")
        c = (1.0 - self.mask) * log_scale
        return c.view(c.shape[0], -1).sum(dim=1)

    def is_feasible(self):
        return ((self.mask == 0) | (self.mask == 1)).all()      # transforms.py:142-145

    def jitter(self, jitter=1e-6):
        return None                                              # transforms.py:147-149
