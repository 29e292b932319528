"""
Decoders -- mirror of src/models/EncodersDecoders/decoders.py.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .._common import print_
from ..BlocksUtils.model_blocks import ConvBlock

DECODERS = ["ConvDecoder"]


def get_decoder(in_channels, decoder, **kwargs):
    """ Factory (decoders.py:18-41); pops num_channels / kernel_size / upsample from the params dict like upstream. """
    name, params = decoder["decoder_name"], decoder["decoder_params"]
    if name not in DECODERS:
        raise ValueError(f"Unknwon decoder_name {name}. Use one of {DECODERS}")
    dec = Decoder(in_channels=in_channels, hidden_dims=params.pop("num_channels"), kernel_size=params.pop("kernel_size"),
                  upsample=params.pop("upsample"), out_channels=kwargs.get("out_channels", 4), **params)
    print_("Decoder:")
    print_(f"  --> Decoder={name}")
    print_(f"  --> in_channels={in_channels}")
    return dec


class Upsample(nn.Module):
    """ nearest-neighbour upsampling module slot of the decoder Sequential (decoders.py:118-140) """

    def __init__(self, scale_factor):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        if self.scale_factor == 1:
            return x
        return F.interpolate(x.contiguous(), scale_factor=self.scale_factor, mode='nearest')

    def __repr__(self):
        return f"Upsample(scale_factor={self.scale_factor})"


class Decoder(nn.Module):
    """ ConvBlocks over reversed(hidden_dims) with Upsample modules in between and a final 3x3 conv to
    RGB + alpha (decoders.py:45-113).  `self.decoder` reproduces the reference's Sequential indices. """

    def __init__(self, in_channels, hidden_dims, kernel_size=5, upsample=None, out_channels=4, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.in_features = in_channels
        self.hidden_dims = hidden_dims
        self.out_features = hidden_dims[0]
        self.kernel_size = kernel_size
        self.stride = kwargs.get("stride", 1)
        self.batch_norm = kwargs.get("batch_norm", None)
        self.upsample = upsample
        self.out_channels = out_channels
        self.decoder = self._build_decoder()

    def _build_decoder(self):
        mods, cin = [], self.in_channels
        for i in range(len(self.hidden_dims) - 1, -1, -1):
            mods.append(ConvBlock(in_channels=cin, out_channels=self.hidden_dims[i], kernel_size=self.kernel_size,
                                  stride=self.stride, padding=self.kernel_size // 2, batch_norm=self.batch_norm))
            cin = self.hidden_dims[i]
            if isinstance(self.upsample, int) and i > 0:
                mods.append(Upsample(scale_factor=self.upsample))
        mods.append(nn.Conv2d(in_channels=self.out_features, out_channels=self.out_channels, kernel_size=3, stride=1, padding=1))
        return nn.Sequential(*mods)

    def conv_blocks(self):
        return [m for m in self.decoder if isinstance(m, ConvBlock)]

    def final_conv(self):
        return self.decoder[-1]

    def forward(self, x):
        """ Stand-alone module forward (decoders.py:103-111): (n, C, H, W) -> (n, out_channels, H, W) through the exact fp32
        convolution kernels.  Not the hot path -- SAVi.decode collapses the first convolution with the broadcast and fuses
        the last one with the compositing -- but callable like the reference's. """
        from ... import ops
        from .._common import require_cuda, no_grad_only, f32c
        require_cuda(x, "Decoder")
        fc = self.final_conv()
        no_grad_only("Decoder", x, fc.weight)
        h = f32c(x)
        for m in self.decoder[:-1]:
            if isinstance(m, ConvBlock):
                h = m.run(h)
            elif isinstance(m, Upsample):
                h = m(h)
                h = h if h.is_contiguous() else h.contiguous()
        co, ci, k, _ = fc.weight.shape
        if k not in (3, 5) or co > 32:
            raise NotImplementedError("Decoder.forward: final convolution must be 3x3 / 5x5 with <= 32 output channels")
        # the direct-convolution kernel computes 32 output channels per pass: pad the 4-channel head with zero filters
        wp = torch.zeros(32, ci, k, k, device=h.device, dtype=torch.float32)
        wp[:co] = fc.weight.detach()
        bp = torch.zeros(32, device=h.device, dtype=torch.float32)
        if fc.bias is not None:
            bp[:co] = fc.bias.detach()
        y = ops.conv2d(h, ops.pack_conv_weight(wp), bp, k, False)
        return y[:, :co].contiguous()
