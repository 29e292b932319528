"""
Unconditional slot predictors -- mirror of src/models/Predictors/transformers.py:
VanillaTransformerPredictor (joint attention over time x slots) and OCVPSeq (alternating object- and
time-attention).  Same constructor arguments and parameter tree; forwards run on the playslot_b200 kernels
(ps_linear with the time encoding fused as a broadcast-row residual, ps_mha).
"""
import torch
import torch.nn as nn

from ... import ops
from .._common import print_, require_cuda, no_grad_only, f32c
from ..BlocksUtils.attention import TransformerEncoderBlock
from ..BlocksUtils.model_blocks import SlotPositionalEncoding
from ..BlocksUtils.model_utils import init_kaiming_

__all__ = ["VanillaTransformerPredictor", "OCVPSeq", "OCVPPar"]


class _SlotTokenPredictor(nn.Module):
    """ shared scaffolding: slot -> token projection + time encoding, token -> slot projection """

    def __init__(self, num_slots, slot_dim, token_dim, hidden_dim, num_layers, n_heads, residual,
                 input_buffer_size, pos_enc_dropout, title):
        if token_dim % n_heads != 0:
            raise ValueError(f"{token_dim = } must be divisible by {n_heads = }...")
        super().__init__()
        self.num_slots = num_slots
        self.slot_dim = slot_dim
        self.token_dim = token_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.num_heads = n_heads
        self.residual = residual
        self.pos_enc_dropout = pos_enc_dropout
        self.input_buffer_size = input_buffer_size
        print_(f"Instanciating {title}:")
        for k in ("num_layers", "slot_dim", "token_dim", "hidden_dim", "num_heads", "pos_enc_dropout", "residual", "input_buffer_size"):
            print_(f"  --> {k}: {getattr(self, k)}")
        self.mlp_in = nn.Linear(slot_dim, token_dim)
        self.mlp_out = nn.Linear(token_dim, slot_dim)

    def _tokens(self, inputs):
        """ (B, F, S, D) -> (B, F, S, E) tokens with the per-frame sinusoidal encoding added in the GEMM epilogue """
        B, Fr, S, _ = inputs.shape
        t = ops.linear(inputs, self.mlp_in.weight, self.mlp_in.bias, res=f32c(self.pe.table()), res_div=S, res_mod=Fr)
        if self.training and self.pe.dropout.p > 0:
            t = self.pe.dropout(t)          # train-mode dropout after the positional encoding (model_blocks.py:211)
        return t

    def _slots(self, tokens, inputs):
        return ops.linear(tokens, self.mlp_out.weight, self.mlp_out.bias, res=inputs if self.residual else None)

    def _prep(self, inputs):
        require_cuda(inputs, type(self).__name__)
        no_grad_only(type(self).__name__, inputs, self.mlp_in.weight)
        x = f32c(inputs)
        if x.shape[1] > self.pe.max_len:
            raise ValueError(f"num_imgs = {x.shape[1]} exceeds input_buffer_size = {self.pe.max_len}")
        return x


class VanillaTransformerPredictor(_SlotTokenPredictor):
    """ joint self-attention over all (time, slot) tokens of the input buffer (transformers.py:21-151) """

    def __init__(self, num_slots, slot_dim, token_dim=128, hidden_dim=256, num_layers=2, n_heads=4,
                 residual=False, input_buffer_size=5, pos_enc_dropout=0.1):
        super().__init__(num_slots, slot_dim, token_dim, hidden_dim, num_layers, n_heads, residual,
                         input_buffer_size, pos_enc_dropout, "Vanilla Transformer Predictor")
        self.transformer_encoders = nn.Sequential(*[
            TransformerEncoderBlock(embed_dim=token_dim, head_dim=token_dim // n_heads, num_heads=n_heads,
                                    mlp_size=hidden_dim, self_attn=True) for _ in range(num_layers)])
        self.pe = SlotPositionalEncoding(d_model=token_dim, max_len=input_buffer_size, dropout=pos_enc_dropout)
        self.attention_masks = None
        self._init_model()

    def forward(self, inputs):
        x = self._prep(inputs)
        B, Fr, S, _ = x.shape
        t = self._tokens(x).view(B, Fr * S, self.token_dim)
        for enc in self.transformer_encoders:
            t = enc.run(t)
        return self._slots(t.view(B, Fr, S, self.token_dim), x)

    def get_attention_masks(self):
        return self.transformer_encoders[-1].get_attention_masks()

    @torch.no_grad()
    def _init_model(self):
        init_kaiming_(self)


class OCVPSeqLayer(nn.Module):
    """ object-attention over the slots of each frame, then time-attention over the frames of each slot
    (transformers.py:263-340) """

    def __init__(self, token_dim=128, hidden_dim=256, n_heads=4):
        if token_dim % n_heads != 0:
            raise ValueError(f"{token_dim = } must be divisible by {n_heads = }...")
        super().__init__()
        self.token_dim = token_dim
        self.hidden_dim = hidden_dim
        self.num_heads = n_heads
        self.object_encoder_block = TransformerEncoderBlock(embed_dim=token_dim, head_dim=token_dim // n_heads,
                                                            num_heads=n_heads, mlp_size=hidden_dim, self_attn=True)
        self.time_encoder_block = TransformerEncoderBlock(embed_dim=token_dim, head_dim=token_dim // n_heads,
                                                          num_heads=n_heads, mlp_size=hidden_dim, self_attn=True)

    def run(self, x):
        B, Fr, S, E = x.shape
        o = self.object_encoder_block.run(x.view(B * Fr, S, E)).view(B, Fr, S, E)
        t = o.transpose(1, 2).contiguous().view(B * S, Fr, E)
        t = self.time_encoder_block.run(t).view(B, S, Fr, E)
        return t.transpose(1, 2).contiguous()

    def forward(self, inputs, time_mask=None):
        require_cuda(inputs, "OCVPSeqLayer")
        no_grad_only("OCVPSeqLayer", inputs, self.object_encoder_block.attn.q.weight)
        return self.run(f32c(inputs))


class OCVPSeq(_SlotTokenPredictor):
    """ OCVP-Seq predictor (transformers.py:155-259) """

    def __init__(self, num_slots, slot_dim, token_dim=128, hidden_dim=256, num_layers=2, n_heads=4,
                 residual=False, input_buffer_size=5, pos_enc_dropout=0.1):
        super().__init__(num_slots, slot_dim, token_dim, hidden_dim, num_layers, n_heads, residual,
                         input_buffer_size, pos_enc_dropout, "OCVP-Seq Predictor Module")
        self.transformer_encoders = nn.Sequential(*[
            OCVPSeqLayer(token_dim=token_dim, hidden_dim=hidden_dim, n_heads=n_heads) for _ in range(num_layers)])
        self.pe = SlotPositionalEncoding(d_model=token_dim, max_len=input_buffer_size, dropout=pos_enc_dropout)

    def forward(self, inputs):
        x = self._prep(inputs)
        t = self._tokens(x)
        for layer in self.transformer_encoders:
            t = layer.run(t)
        return self._slots(t, x)


class OCVPPar(nn.Module):
    """ Upstream's OCVP-Par cannot be constructed (transformers.py:407-413 calls OCVPParLayer with d_model=/nhead=
    while it is defined with token_dim/hidden_dim/n_heads at :477) and no config uses it. """

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("OCVPPar is broken in the reference (constructor argument mismatch) and unused")
