"""
Encoders -- mirror of src/models/EncodersDecoders/encoders.py.
"""
import torch.nn as nn

from .._common import print_, require_cuda, no_grad_only, f32c
from ..BlocksUtils.model_blocks import ConvBlock

ENCODERS = ["ConvEncoder"]


def get_encoder(in_channels, encoder, **kwargs):
    """ Factory (encoders.py:19-42).  Note: pops 'num_channels'/'kernel_size' from the params dict like upstream. """
    name, params = encoder["encoder_name"], encoder["encoder_params"]
    if name not in ENCODERS:
        raise ValueError(f"Unknwon encoder_name {name}. Use one of {ENCODERS}")
    enc = SimpleConvEncoder(in_channels=in_channels, hidden_dims=params.pop("num_channels"),
                            kernel_size=params.pop("kernel_size"), **params)
    print_("Encoder:")
    print_(f"  --> Encoder_type={name}")
    print_(f"  --> in_channels={in_channels}")
    return enc


class SimpleConvEncoder(nn.Module):
    """ Stack of ConvBlocks, ReLU on all but the last (encoders.py:46-104); `self.encoder` is the Sequential. """

    def __init__(self, in_channels=3, hidden_dims=(64, 64, 64, 64), kernel_size=5, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.hidden_dims = hidden_dims
        self.kernel_size = kernel_size
        self.stride = kwargs.get("stride", 1)
        self.batch_norm = kwargs.get("batch_norm", None)
        self.use_downsample = kwargs.get("downsample_encoder", False)
        self.downsample = kwargs.get("downsample", 2)
        self.out_features = hidden_dims[-1]
        self.encoder = self._build_encoder()

    def _build_encoder(self):
        blocks, cin = [], self.in_channels
        last = len(self.hidden_dims) - 1
        for i, h in enumerate(self.hidden_dims):
            pool = self.downsample if (self.use_downsample and i < last) else False
            blocks.append(ConvBlock(in_channels=cin, out_channels=h, kernel_size=self.kernel_size, stride=self.stride,
                                    padding=self.kernel_size // 2, batch_norm=self.batch_norm, max_pool=pool,
                                    activation=i < last))
            cin = h
        return nn.Sequential(*blocks)

    def run(self, x, last_nhwc_addend=None):
        """ kernel path; with `last_nhwc_addend` the final conv writes NHWC tokens + position embedding """
        n = len(self.encoder)
        for i, blk in enumerate(self.encoder):
            if i == n - 1 and last_nhwc_addend is not None:
                x = blk.run(x, out_nhwc=True, addend=last_nhwc_addend)
            else:
                x = blk.run(x)
        return x

    def forward(self, x):
        require_cuda(x, "SimpleConvEncoder")
        no_grad_only("SimpleConvEncoder", x, self.encoder[0].conv.weight)
        return self.run(f32c(x))
