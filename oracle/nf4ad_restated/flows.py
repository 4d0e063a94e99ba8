"""ORACLE: `NonUSFlow` stack builder, restating `/root/reference/src/nf4ad/flows.py:27-169`.

Layer order produced (generative direction), for K coupling blocks:
    [ Block_k , Coupling_k , Block_k^-1 (only with affine_conjugation) ] * K , Block_final , Scale
with Block_k = BlockAffine(Sequential(LU * lu_transform, Householder(nvs=householder) if householder > 0)),
Block_final = BlockAffine(LU), and the 0/1 mask flipped after every block (flows.py:81-114).
"""
import torch
from src.usflows.flows import Flow
from src.usflows import transforms as T
from .transforms import MaskedAffineCoupling


def parity_mask(in_dims, along_first_axis_only=False, invert=False):
    """flows.py:127-145.  Checkerboard: parity of the index sum; channel: parity of the leading index."""
    index = torch.stack(torch.meshgrid(*[torch.arange(n, dtype=torch.int32) for n in in_dims], indexing="ij"))
    parity = index[0] if along_first_axis_only else index.sum(dim=0)
    m = torch.fmod(parity, 2).to(torch.float32).view(1, *in_dims)
    return (1 - m) if invert else m


def assemble_layers(in_dims, n_blocks, make_conditioner, lu_count, hh_count, conjugate, prior_scale, mask, device):
    dim = in_dims[0]
    out = []
    for _ in range(n_blocks):
        parts = [T.LUTransform(dim, prior_scale) for _ in range(lu_count)]          # flows.py:84-86
        if hh_count > 0:                                                              # flows.py:89-91
            parts.append(T.HouseholderTransform(dim=dim, nvs=hh_count, device=device))
        block = T.BlockAffineTransform(in_dims, T.SequentialAffineTransform(parts)) if parts else None
        if block is not None:
            out.append(block)                                                         # flows.py:93-96
        out.append(MaskedAffineCoupling(mask, make_conditioner()))                   # flows.py:99-100
        if conjugate and block is not None:
            out.append(T.InverseTransform(block))                                     # flows.py:103-104
        mask = 1 - mask                                                               # flows.py:107
    out.append(T.BlockAffineTransform(in_dims, T.LUTransform(dim, prior_scale)))      # flows.py:110-112
    out.append(T.ScaleTransform(in_dims))                                             # flows.py:113-114
    return out


class NonUSFlow(Flow):
    MASKTYPE = ("checkerboard", "channel")

    def __init__(self, base_distribution, in_dims, coupling_blocks, conditioner_cls, conditioner_args,
                 soft_training=False, prior_scale=None, training_noise_prior=None,
                 affine_conjugation=False, nonlinearity=None, lu_transform=1, householder=1,
                 masktype="checkerboard", device="cpu", *args, **kwargs):
        # validation, flows.py:63-76
        if masktype not in self.MASKTYPE:
            raise ValueError(f"Unknown mask type {masktype}")
        for count, what in ((lu_transform, "LU transforms"), (householder, "Householder vectors transforms")):
            if count < 0:
                raise ValueError(f"Number of {what} must be non-negative")
        # plain attributes set before the base constructor runs, flows.py:54-61
        self.coupling_blocks, self.in_dims, self.prior_scale, self.device = coupling_blocks, in_dims, prior_scale, device
        self.conditioner_cls, self.conditioner_args = conditioner_cls, conditioner_args
        self.lu_transform, self.householder = lu_transform, householder
        layers = assemble_layers(in_dims, coupling_blocks, lambda: conditioner_cls(**conditioner_args), lu_transform,
                                 householder, affine_conjugation, prior_scale,
                                 parity_mask(in_dims, masktype == "channel"), device)
        super().__init__(base_distribution, layers, soft_training=soft_training,
                         training_noise_prior=training_noise_prior, device=device)

    @staticmethod
    def create_checkerboard_mask(in_dims, invert=False):
        return parity_mask(in_dims, False, invert)

    @staticmethod
    def create_channel_mask(in_dims, invert=False):
        return parity_mask(in_dims, True, invert)

    def log_prior(self):
        """flows.py:147-158: sum of the layers' log-priors when a prior scale is set; layers without one are skipped."""
        if self.prior_scale is None:
            return 0
        acc = 0
        for layer in self.layers:
            try:
                acc = acc + layer.log_prior()
            except Exception:
                pass
        return acc

    def log_abs_det_jacobian(self, x):
        """flows.py:160-169, quirk kept: every layer is evaluated at the DATA point x (x is never advanced), so
        the value is the true log-det only when every layer's log-det is data independent (USFlow)."""
        acc = 0
        for layer in reversed(self.layers):
            try:
                acc = acc - layer.log_abs_det_jacobian(layer.backward(x), x)
            except Exception:
                pass
        return acc
