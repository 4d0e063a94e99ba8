"""ORACLE (test infrastructure) -- `src.usflows.flows.{Flow,USFlow}` restated.  PARITY UNPINNED.

`Flow` [INFER `/root/reference/src/nf4ad/flows.py:117-125` (ctor),
`adbench_wrapper.py:345,369,383,424` (`to/parameters/log_prob`),
`vaeflow.py:201,240` (`log_prob`, `sample`), `visualization.py:85-86`
(`export`), `scripts/gmm_eval_usflows.py:105` (`backward`)] is
[RECALL] `TransformedDistribution(Independent(base, len(batch_shape)), layers)`;
its log_prob/sample recursion is the container's
`torch/distributions/transformed_distribution.py:143-190`.

`USFlow` [INFER YAMLs `experiments/gmm/...yaml:52`, docstring `flows.py:28-30`]
builds the same layer list as `NonUSFlow.__init__` (`flows.py:78-114`) with the
additive `MaskedCoupling`.
"""
from typing import Any, Dict, List, Optional, Type

import torch
from pyro import distributions as dist

from .transforms import (
    BlockAffineTransform, HouseholderTransform, InverseTransform, LUTransform,
    MaskedCoupling, ScaleTransform, SequentialAffineTransform,
)


class Flow(torch.nn.Module):
    export = "log_prob"

    def __init__(self, base_distribution, layers, soft_training: bool = False,
                 training_noise_prior=None, device="cpu", *args, **kwargs):
        # subclasses assign plain attributes before calling this (flows.py:54-61)
        preset = dict(self.__dict__)
        super().__init__()
        for k, v in preset.items():
            self.__dict__.setdefault(k, v)
        self.soft_training = soft_training
        self.training_noise_prior = training_noise_prior
        self.layers = list(layers)
        self.trainable_layers = torch.nn.ModuleList(
            [l for l in self.layers if isinstance(l, torch.nn.Module)])
        self.base_distribution = base_distribution
        self.device = device
        self._rebuild()

    def _rebuild(self):
        base = self.base_distribution
        nb = len(base.batch_shape)
        # plain attributes (not registered sub-modules): keep state_dict keys = parameters of the layers + base
        self.__dict__["_event_base"] = dist.Independent(base, nb) if nb > 0 else base
        self.__dict__["transform"] = dist.TransformedDistribution(self._event_base, self.layers)

    # ---- density path (transformed_distribution.py:143-190) ----
    @property
    def event_ndim(self):
        return max(1, len(self._event_base.event_shape))

    def _soft_context(self, x, context):
        """[RECALL] soft training: the conditioners take the per-sample noise level as context; scoring without
        one means noise level 0."""
        if context is None and self.soft_training:
            lead = x.shape[:x.dim() - self.event_ndim]
            context = torch.zeros(*lead, 1, dtype=x.dtype, device=x.device)
        return context

    def log_prob(self, x, context=None):
        context = self._soft_context(x, context)
        if context is None and self.event_ndim == 1:
            return self.transform.log_prob(x)
        # the same recursion (transformed_distribution.py:168-190) written out, so that a context reaches every layer
        # and a per-sample (B,) log-det is subtracted as such for an N-D event shape
        lp = 0.0
        y = x
        for layer in reversed(self.layers):
            xin = layer.backward(y) if context is None else layer.backward(y, context)
            ladj = layer.log_abs_det_jacobian(xin, y) if context is None else layer.log_abs_det_jacobian(xin, y, context)
            lp = lp - ladj
            y = xin
        return lp + self._event_base.log_prob(y)

    def sample(self, sample_shape=None, context=None):
        shape = torch.Size() if sample_shape is None else torch.Size(sample_shape)
        if context is None and not self.soft_training:
            return self.transform.sample(shape)
        with torch.no_grad():
            z = self._event_base.sample(shape)
            return self.latent_to_data(z, self._soft_context(z, context))

    def rsample(self, sample_shape=None):
        shape = torch.Size() if sample_shape is None else torch.Size(sample_shape)
        return self.transform.rsample(shape)

    def backward(self, x, context=None):
        """data -> latent: every layer's inverse, last layer first."""
        context = self._soft_context(x, context)
        for layer in reversed(self.layers):
            x = layer.backward(x) if context is None else layer.backward(x, context)
        return x

    def forward(self, x=None, context=None):
        """`export` switch (visualization.py:85-86); default generative direction latent -> data."""
        if self.export == "log_prob":
            return self.log_prob(x)
        if self.export == "sample":
            return self.sample()
        if self.export == "backward":
            return self.backward(x)
        for layer in self.layers:
            x = layer.forward(x)
        return x

    def latent_to_data(self, z, context=None):
        context = self._soft_context(z, context)
        for layer in self.layers:
            z = layer.forward(z) if context is None else layer.forward(z, context)
        return z

    # ---- housekeeping ----
    def is_feasible(self) -> bool:
        return all(l.is_feasible() for l in self.layers if hasattr(l, "is_feasible"))

    def add_jitter(self, jitter: float = 1e-6) -> None:
        for l in self.layers:
            if hasattr(l, "jitter"):
                l.jitter(jitter)

    def log_prior(self):
        total = 0.0
        for l in self.layers:
            if hasattr(l, "log_prior"):
                total = total + l.log_prior()
        return total

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        base = self.base_distribution
        if not isinstance(base, torch.nn.Module):
            for k, v in list(base.__dict__.items()):
                if isinstance(v, torch.Tensor):
                    base.__dict__[k] = fn(v)
        self._rebuild()
        return out

    def to(self, *args, **kwargs):
        out = super().to(*args, **kwargs)
        if args and isinstance(args[0], (str, torch.device)):
            self.device = args[0]
        elif "device" in kwargs:
            self.device = kwargs["device"]
        return out

    def soft_noise(self, batch, generator=None):
        """[RECALL] soft training (SoftFlow-style): one noise level per sample from `training_noise_prior`, the sample is
        perturbed by N(0, level^2) noise and the level is handed to the conditioners as context.
        -> (noisy batch, context of shape (B, 1))."""
        B = batch.shape[0]
        level = self.training_noise_prior.sample([B]).reshape(B).to(batch)
        eps = torch.randn(batch.shape, dtype=batch.dtype, device=batch.device, generator=generator)
        return batch + level.reshape(B, *([1] * (batch.dim() - 1))) * eps, level.reshape(B, 1).detach()

    def fit(self, data_train, optim=torch.optim.Adam, optim_params=None, batch_size=32,
            shuffle=True, gradient_clip=None, device=None, jitter=1e-6, epochs=1):
        """[RECALL] feasibility check + jitter, then minimise -mean log_prob (- log_prior)."""
        if device is not None:
            self.to(device)
        opt = optim(self.parameters(), **(optim_params or {}))
        n = len(data_train)
        losses = []
        for _ in range(epochs):
            perm = torch.randperm(n) if shuffle else torch.arange(n)
            total = 0.0
            for i in range(0, n, batch_size):
                idx = perm[i:i + batch_size]
                batch = torch.stack([data_train[int(j)][0] if isinstance(data_train[int(j)], (tuple, list))
                                     else data_train[int(j)] for j in idx]).to(self.device)
                while not self.is_feasible():
                    self.add_jitter(jitter)
                opt.zero_grad()
                ctx = None
                if self.soft_training:
                    batch, ctx = self.soft_noise(batch)
                loss = -self.log_prob(batch, ctx).mean()
                if getattr(self, "prior_scale", None) is not None:
                    loss = loss - self.log_prior() / n
                loss.backward()
                if gradient_clip is not None:
                    torch.nn.utils.clip_grad_norm_(self.parameters(), gradient_clip)
                opt.step()
                total += float(loss.detach()) * len(idx)
            losses.append(total / n)
        return losses


class USFlow(Flow):
    """Additive-coupling stack; same layer list as `nf4ad/flows.py:78-114`."""

    MASKTYPE = ("checkerboard", "channel")

    def __init__(self, base_distribution, in_dims: List[int], coupling_blocks: int,
                 conditioner_cls: Type[torch.nn.Module], conditioner_args: Dict[str, Any],
                 soft_training: bool = False, prior_scale: Optional[float] = None,
                 training_noise_prior=None, affine_conjugation: bool = False,
                 nonlinearity=None, lu_transform: int = 1, householder: int = 1,
                 masktype: str = "checkerboard", device="cpu", *args, **kwargs):
        self.coupling_blocks = coupling_blocks
        self.in_dims = in_dims
        self.prior_scale = prior_scale
        if masktype not in self.MASKTYPE:
            raise ValueError(f"Unknown mask type {masktype}")
        if lu_transform < 0:
            raise ValueError("Number of LU transforms must be non-negative")
        if householder < 0:
            raise ValueError("Number of Householder vectors transforms must be non-negative")
        mask = self.create_mask(in_dims, masktype)
        layers = []
        for _ in range(coupling_blocks):
            affine = [LUTransform(in_dims[0], prior_scale) for _ in range(lu_transform)]
            if householder > 0:
                affine.append(HouseholderTransform(dim=in_dims[0], nvs=householder, device=device))
            block = None
            if affine:
                block = BlockAffineTransform(in_dims, SequentialAffineTransform(affine))
                layers.append(block)
            layers.append(MaskedCoupling(mask, conditioner_cls(**conditioner_args)))
            if affine_conjugation and block is not None:
                layers.append(InverseTransform(block))
            mask = 1 - mask
        layers.append(BlockAffineTransform(in_dims, LUTransform(in_dims[0], prior_scale)))
        layers.append(ScaleTransform(in_dims))
        super().__init__(base_distribution, layers, soft_training=soft_training,
                         training_noise_prior=training_noise_prior, device=device)

    @staticmethod
    def create_mask(in_dims, masktype="checkerboard"):
        axes = [torch.arange(d, dtype=torch.int32) for d in in_dims]
        idx = torch.stack(torch.meshgrid(*axes, indexing="ij"))
        m = torch.fmod(idx.sum(dim=0) if masktype == "checkerboard" else idx[0], 2)
        return m.to(torch.float32).view(1, *in_dims)
