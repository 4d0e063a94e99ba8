"""ORACLE (test infrastructure) -- `src.usflows.networks.ConvNet` stand-in.  PARITY UNPINNED.

[INFER `/root/reference/experiments/MVTec/mvtec_trainable_encoder_us.yaml:67-74`] `ConvNet(in_dims, c_hidden, c_out,
nonlinearity)` is the conditioner of image-shaped flows; upstream's layer list is not recoverable from the reference
tree, so this fixes it as size-preserving 3x3 convolutions `in_dims[0] -> c_hidden... -> c_out`.
"""
import torch


class ConvNet(torch.nn.Module):
    def __init__(self, in_dims, c_hidden, c_out=None, nonlinearity=None, kernel_size=3):
        super().__init__()
        c_in = int(in_dims[0])
        hidden = [int(c) for c in (c_hidden if isinstance(c_hidden, (list, tuple)) else [c_hidden])]
        c_out = c_in if c_out is None else int(c_out)
        chans = [c_in] + hidden
        mods = []
        for i in range(len(hidden)):
            mods += [torch.nn.Conv2d(chans[i], chans[i + 1], kernel_size, padding=kernel_size // 2),
                     nonlinearity if nonlinearity is not None else torch.nn.ReLU()]
        mods.append(torch.nn.Conv2d(chans[-1], c_out, kernel_size, padding=kernel_size // 2))
        self.net = torch.nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x)
