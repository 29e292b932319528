"""
Basic building blocks -- mirror of src/models/BlocksUtils/model_blocks.py (same class names, constructor
arguments and parameter tree).  Forward passes launch playslot_b200 kernels.
"""
import math

import torch
import torch.nn as nn

from ... import ops
from ... import autograd as AG
from .._common import DerivedWeights, require_cuda, no_grad_only, f32c
from .model_utils import build_grid

__all__ = [
        "ConvBlock",
        "SoftPositionEmbed",
        "LearnedPositionalEncoding",
        "SlotPositionalEncoding",
        "SinusoidalPositionalEncoding",
        "MLP"
    ]


class ConvBlock(nn.Module):
    """
    Conv (+BN) (+pool | +upsample) (+ReLU) block; parameters live in `self.block` exactly as in the
    reference (model_blocks.py:24-79) so that checkpoints map 1:1.  The kernel path supports what the
    shipped configs use: stride 1, kernel 3|5, 32|64 output channels, no BN / pooling / upsampling.
    """

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=None,
                 batch_norm=False, max_pool=None, upsample=None, activation=True):
        super().__init__()
        padding = padding if padding is not None else kernel_size // 2
        layers = [nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding)]
        if batch_norm:
            layers.append(nn.BatchNorm2d(num_features=out_channels))
        if max_pool:
            assert isinstance(max_pool, (int, tuple, list))
            layers.append(nn.MaxPool2d(kernel_size=max_pool, stride=max_pool))
        if upsample is not None:
            assert isinstance(upsample, int)
            assert max_pool is None, "'max_pool' and 'upsample' cannot both be provided"
            layers.append(nn.Upsample(scale_factor=upsample, mode="bilinear"))
        if activation:
            layers.append(nn.ReLU())
        self.block = nn.Sequential(*layers)
        self._plain = (not batch_norm) and (not max_pool) and upsample is None and stride == 1 \
            and padding == kernel_size // 2
        self._relu = bool(activation)
        self._cache = DerivedWeights()

    @property
    def conv(self):
        return self.block[0]

    def packed_weight(self):
        w = self.conv.weight
        return self._cache.get([w], lambda: ops.pack_conv_weight(w))

    def run(self, x, out_nhwc=False, addend=None):
        """ kernel path: x (n, Cin, H, W) fp32 contiguous """
        if not self._plain:
            raise NotImplementedError("ConvBlock: only stride-1 'same' convolutions without BatchNorm / pooling / "
                                      "upsampling are implemented (that is what every shipped config uses)")
        c = self.conv
        return ops.conv2d(x, self.packed_weight(), c.bias, c.kernel_size[0], self._relu, out_nhwc, addend)

    def forward(self, x):
        require_cuda(x, "ConvBlock")
        no_grad_only("ConvBlock", x, self.conv.weight)
        return self.run(f32c(x))


class SoftPositionEmbed(nn.Module):
    """
    Soft positional embedding: a learnable 1x1 projection of the 4-channel coordinate grid, added to the
    inputs (model_blocks.py:85-119).  The projected grid depends on weights only, so it is computed once per
    weight version (`table`) instead of once per call on a batch-times repeated grid.
    """

    def __init__(self, hidden_size, resolution, vmin=-1., vmax=1.):
        super().__init__()
        self.resolution = resolution
        self.projection = nn.Conv2d(4, hidden_size, kernel_size=1)
        self.grid = build_grid(resolution, vmin=vmin, vmax=vmax).permute(0, 3, 1, 2)   # plain attribute, as upstream
        self._cache = DerivedWeights()

    def grid_on(self, device, dtype=torch.float32):
        """ the (1, 4, H, W) coordinate grid on `device` -- `grid` itself stays a plain CPU attribute as upstream; the device copy is
        made once (a host-to-device copy per call would also be illegal inside a CUDA-graph capture of a training step) """
        key = (str(device), dtype)
        c = getattr(self, "_grid_dev", None)
        if c is None or c[0] != key:
            c = (key, self.grid.to(device=device, dtype=dtype))
            self._grid_dev = c
        return c[1]

    def table(self, channels_last=True):
        """ (H*W, C) if channels_last else (C, H, W): projection(grid), weight-only """
        w, b = self.projection.weight, self.projection.bias

        def build():
            g = self.grid_on(w.device, w.dtype)[0]                                    # (4, H, W)
            emb = torch.einsum("ck,khw->chw", w[:, :, 0, 0], g) + b[:, None, None]   # (C, H, W)
            return emb.contiguous(), emb.permute(1, 2, 0).reshape(-1, emb.shape[0]).contiguous()
        chw, hwc = self._cache.get([w, b], build)
        return hwc if channels_last else chw

    def forward(self, inputs, channels_last=True):
        """ inputs (B, H, W, C) [channels_last] or (B, C, H, W) -> inputs + embedding """
        require_cuda(inputs, "SoftPositionEmbed")
        no_grad_only("SoftPositionEmbed", inputs, self.projection.weight)
        if channels_last:
            H, W = self.resolution
            return inputs + self.table(True).view(1, H, W, -1)
        return inputs + self.table(False)[None]


class LearnedPositionalEncoding(nn.Module):
    """ 1-D learned positional encoding (model_blocks.py:123-149) """

    def __init__(self, max_len, token_dim, dropout=0.1):
        super().__init__()
        self.dropout = nn.Dropout(dropout)
        self.pe = nn.Parameter(torch.zeros(1, max_len, token_dim), requires_grad=True)
        nn.init.trunc_normal_(self.pe)

    def forward(self, x, offset=0):
        T = x.shape[1]
        return self.dropout(x + self.pe[:, offset:offset + T])


def _sinusoid_table(max_len, d_model):
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class SlotPositionalEncoding(nn.Module):
    """
    Sinusoidal time encoding shared by all slots of a frame (model_blocks.py:153-216); buffer `pe` has shape
    (1, max_len, 1, d_model) and is part of the state_dict.
    """

    def __init__(self, d_model, dropout=0.1, max_len=50):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.max_len = max_len
        self.register_buffer('pe', _sinusoid_table(max_len, d_model).view(1, max_len, 1, d_model))

    def table(self):
        """ (max_len, d_model) contiguous view used as the broadcast residual of the embedding GEMM """
        return self.pe.view(self.max_len, -1)

    def forward(self, x, batch_size, num_slots, offset=0):
        if len(x.shape) != 4:
            raise ValueError(f"{x.shape = } must have 4 dimensions: (B, SeqLen, Objs, Dim)...")
        if x.device != self.pe.device:
            self.pe = self.pe.to(x.device)
        T = x.shape[1]
        return self.dropout(x + self.pe[:, offset:T + offset])

    def get_pe(self, idx):
        return self.pe[:, idx, :, :]


class SinusoidalPositionalEncoding(nn.Module):
    """ Standard sinusoidal positional encoding (model_blocks.py:220-274) """

    def __init__(self, d_model, dropout=0.1, max_len=5000):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        self.max_len = max_len
        self.register_buffer('pe', _sinusoid_table(max_len, d_model).unsqueeze(0))

    def forward(self, x):
        if x.device != self.pe.device:
            self.pe = self.pe.to(x.device)
        n = x.shape[1]
        if x.dim() == 3:
            pe = self.pe[:, :n]
        elif x.dim() == 4:
            pe = self.pe[:, :n].unsqueeze(2)
        else:
            raise ValueError(f"Wrong shape {x.shape = }. It must have 3 or 4 dims...")
        return self.dropout(x + pe)


class MLP(nn.Module):
    """ 2-layer perceptron of the transformer blocks (model_blocks.py:278-309): Linear, GELU(erf) | ReLU, Linear """

    def __init__(self, in_dim, hidden_dim, out_dim=None, use_gelu=True):
        super().__init__()
        out_dim = out_dim if out_dim is not None else in_dim
        self.mlp = nn.Sequential(
            nn.Linear(in_dim, hidden_dim),
            nn.ReLU() if not use_gelu else nn.GELU(),
            nn.Linear(hidden_dim, out_dim),
        )
        self._act = ops.ACT_GELU if use_gelu else ops.ACT_RELU

    def run(self, x, ln=None, res=None):
        """ kernel path: optional fused pre-LayerNorm and residual """
        fc1, fc2 = self.mlp[0], self.mlp[2]
        rows = x.numel() // x.shape[-1]
        if ops.planes_ok(rows, fc1.out_features) and not AG.needs_grad(x, res, fc1.weight, fc1.bias, fc2.weight, fc2.bias,
                                                                        *(ln[:2] if ln is not None else ())):
            # inference on the tensor cores: fc1's epilogue hands fc2 its operand as fp16 hi/lo planes (loaded by TMA)
            h = ops.linear(x, fc1.weight, fc1.bias, ln=ln, act=self._act, out_planes=True)
            return ops.linear(h, fc2.weight, fc2.bias, res=res).view(*x.shape[:-1], fc2.out_features)
        h = AG.linear(x, fc1.weight, fc1.bias, ln=ln, act=self._act)
        return AG.linear(h, fc2.weight, fc2.bias, res=res)

    def forward(self, x):
        require_cuda(x, "MLP")
        return self.run(f32c(x))
