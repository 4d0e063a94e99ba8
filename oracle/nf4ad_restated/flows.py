"""ORACLE: `NonUSFlow` stack builder, restating `/root/reference/src/nf4ad/flows.py:27-169`."""
import torch
from src.usflows.flows import Flow
from src.usflows.transforms import (
    BlockAffineTransform, HouseholderTransform, InverseTransform, LUTransform,
    ScaleTransform, SequentialAffineTransform,
)
from .transforms import MaskedAffineCoupling


def _parity_mask(in_dims, channel_only, invert=False):
    # flows.py:127-145: (sum of indices) mod 2, or (leading index) mod 2; shape (1,*in_dims)
    grids = torch.meshgrid(*[torch.arange(d, dtype=torch.int32) for d in in_dims], indexing="ij")
    idx = torch.stack(grids)
    m = torch.fmod(idx[0] if channel_only else idx.sum(dim=0), 2).to(torch.float32).view(1, *in_dims)
    return 1 - m if invert else m


class NonUSFlow(Flow):
    MASKTYPE = ("checkerboard", "channel")

    def __init__(self, base_distribution, in_dims, coupling_blocks, conditioner_cls, conditioner_args,
                 soft_training=False, prior_scale=None, training_noise_prior=None,
                 affine_conjugation=False, nonlinearity=None, lu_transform=1, householder=1,
                 masktype="checkerboard", device="cpu", *args, **kwargs):
        # flows.py:54-76
        self.coupling_blocks, self.in_dims = coupling_blocks, in_dims
        self.conditioner_cls, self.conditioner_args = conditioner_cls, conditioner_args
        self.prior_scale, self.device = prior_scale, device
        if masktype not in self.MASKTYPE:
            raise ValueError(f"Unknown mask type {masktype}")
        if lu_transform < 0:
            raise ValueError("Number of LU transforms must be non-negative")
        if householder < 0:
            raise ValueError("Number of Householder vectors transforms must be non-negative")
        self.lu_transform, self.householder = lu_transform, householder
        D = in_dims[0]
        mask = _parity_mask(in_dims, masktype == "channel")
        stack = []
        for _ in range(coupling_blocks):                       # flows.py:81-107
            affine = [LUTransform(D, prior_scale) for _ in range(lu_transform)]
            if householder > 0:
                affine.append(HouseholderTransform(dim=D, nvs=householder, device=device))
            conj = BlockAffineTransform(in_dims, SequentialAffineTransform(affine)) if affine else None
            if conj is not None:
                stack.append(conj)
            stack.append(MaskedAffineCoupling(mask, conditioner_cls(**conditioner_args)))
            if affine_conjugation and conj is not None:
                stack.append(InverseTransform(conj))
            mask = 1 - mask
        stack.append(BlockAffineTransform(in_dims, LUTransform(D, prior_scale)))   # flows.py:110-112
        stack.append(ScaleTransform(in_dims))                                      # flows.py:113-114
        super().__init__(base_distribution, stack, soft_training=soft_training,
                         training_noise_prior=training_noise_prior, device=device)

    create_checkerboard_mask = staticmethod(lambda in_dims, invert=False: _parity_mask(in_dims, False, invert))
    create_channel_mask = staticmethod(lambda in_dims, invert=False: _parity_mask(in_dims, True, invert))

    def log_prior(self):
        # flows.py:147-158
        if self.prior_scale is None:
            return 0
        total = 0
        for layer in self.layers:
            try:
                total = total + layer.log_prior()
            except Exception:
                continue
        return total

    def log_abs_det_jacobian(self, x):
        # flows.py:160-169 -- quirk kept: x is never advanced through the layers
        total = 0
        for layer in reversed(self.layers):
            try:
                total = total - layer.log_abs_det_jacobian(layer.backward(x), x)
            except Exception:
                continue
        return total
