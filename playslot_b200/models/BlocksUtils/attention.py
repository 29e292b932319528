"""
Attention modules -- mirror of src/models/BlocksUtils/attention.py (same class names, constructor arguments,
parameter tree and forward signatures); every forward launches playslot_b200 kernels.
"""
import os

import torch
import torch.nn as nn

from ... import ops
from ... import autograd as AG
from .._common import DerivedWeights, require_cuda, no_grad_only, f32c
from .model_blocks import MLP

__all__ = [
    "SlotAttention",
    "MultiHeadSelfAttention", "MultiHeadCrossAttention",
    "TransformerEncoderBlock", "TransformerDecoder"
]


##################
# SLOT ATTENTION #
##################

class SlotAttention(nn.Module):
    """
    Slot Attention (Locatello et al. 2020) as used by SAVi; reference attention.py:23-135.

    forward(inputs (B, N, dim_feats), slots (B, S, dim_slots), step) -> slots (B, S, dim_slots); the attention
    masks of the last iteration (softmax over slots + eps, before the renormalisation over locations) are kept
    for `get_attention_masks`.

    Kernel path (csrc/slot_attention.cu): per iteration one streaming pass over the feature tokens
    (`ps_sa_main`) and one per-slot update (`ps_sa_update`); k and v are never materialised (re-association
    q.k^T = (q W_k).LN(x)^T and attn.v = (attn.LN(x)) W_v^T, exact up to fp32 rounding).
    """

    TOKENS_PER_CTA = int(os.environ.get("PLAYSLOT_SA_TOKENS", "256"))
    # "tc":   pre-normalised fp16 hi/lo feature planes (one pass per frame) streamed by TMA into tcgen05 contractions
    # "tc1":  the same contractions with the normalisation / split done inside every streaming pass
    # "fp32": exact FFMA kernel
    MODES = ("tc", "tc1", "fp32")

    def __init__(self, dim_feats, dim_slots, num_slots, num_iters_first=2, num_iters=2,
                 mlp_hidden=128, epsilon=1e-8):
        super().__init__()
        self.dim_slots = dim_slots
        self.num_iters_first = num_iters_first
        self.num_iters = num_iters
        self.num_slots = num_slots
        self.epsilon = epsilon
        self.scale = dim_slots ** -0.5

        self.norm_input = nn.LayerNorm(dim_feats)
        self.norm_slot = nn.LayerNorm(dim_slots)
        self.norm_mlp = nn.LayerNorm(dim_slots)

        self.to_q = nn.Linear(dim_slots, dim_slots, bias=False)
        self.to_k = nn.Linear(dim_feats, dim_slots, bias=False)
        self.to_v = nn.Linear(dim_feats, dim_slots, bias=False)

        self.gru = nn.GRUCell(dim_slots, dim_slots)
        self.mlp = nn.Sequential(
            nn.Linear(dim_slots, mlp_hidden),
            nn.ReLU(),
            nn.Linear(mlp_hidden, dim_slots),
        )
        self.attention_masks = None
        self._cache = DerivedWeights()
        self.mode = os.environ.get("PLAYSLOT_SA_MODE", "tc")
        if self.mode not in self.MODES:
            raise ValueError(f"PLAYSLOT_SA_MODE={self.mode!r}; expected one of {self.MODES}")

    def _weights(self):
        """ transposed ([in][out]) copies of the small per-slot matrices, rebuilt when a parameter changes """
        src = [self.norm_input.weight, self.norm_input.bias, self.to_v.weight, self.gru.weight_ih, self.gru.weight_hh,
               self.gru.bias_ih, self.gru.bias_hh, self.norm_mlp.weight, self.norm_mlp.bias,
               self.mlp[0].weight, self.mlp[0].bias, self.mlp[2].weight, self.mlp[2].bias,
               self.norm_slot.weight, self.norm_slot.bias, self.to_q.weight, self.to_k.weight]

        def build():
            t = lambda w: w.detach().t().contiguous()
            c = lambda w: w.detach().contiguous()
            return {
                "ln_in_g": c(self.norm_input.weight), "ln_in_b": c(self.norm_input.bias),
                "WvT": t(self.to_v.weight),
                "WihT": t(self.gru.weight_ih), "WhhT": t(self.gru.weight_hh),
                "b_ih": c(self.gru.bias_ih), "b_hh": c(self.gru.bias_hh),
                "ln_mlp_g": c(self.norm_mlp.weight), "ln_mlp_b": c(self.norm_mlp.bias),
                "W1T": t(self.mlp[0].weight), "b1": c(self.mlp[0].bias),
                "W2T": t(self.mlp[2].weight), "b2": c(self.mlp[2].bias),
                "ln_slot_g": c(self.norm_slot.weight), "ln_slot_b": c(self.norm_slot.bias),
                "WqT": t(self.to_q.weight), "Wk": c(self.to_k.weight),
                # weight-only re-associations of ps_sa_update2 (products formed in fp64, rounded once)
                "WivT": (self.gru.weight_ih.detach().double() @ self.to_v.weight.detach().double()).t().float().contiguous(),
                "MAT": (self.scale * self.norm_input.weight.detach().double()[:, None]
                        * (self.to_k.weight.detach().double().t() @ self.to_q.weight.detach().double())).t().float().contiguous(),
                "v0": (self.scale * (self.norm_input.bias.detach().double()
                                     @ (self.to_k.weight.detach().double().t() @ self.to_q.weight.detach().double()))).float().contiguous(),
            }
        return self._cache.get(src, build)

    def prepare(self, feats):
        """ fp32 features (..., N, dim) -> pre-normalised planes for `forward(inputs=("planes", planes, first_token, B, N))`;
        SAVi.forward prepares all frames of a clip in one launch """
        return ops.sa_prepare(f32c(feats), self.norm_input.eps)

    def forward(self, inputs, slots, step=0, **kwargs):
        planes = None
        if isinstance(inputs, tuple) and inputs[0] == "planes":
            _, planes, first_token, B, N = inputs
            require_cuda(planes, "SlotAttention")
            x = None
        else:
            require_cuda(inputs, "SlotAttention")
            if AG.needs_grad(inputs, slots, self.to_q.weight, self.norm_input.weight):
                # SAVi training (02_train_savi.py): differentiable iterations (savi_autograd.slot_attention_train)
                from ... import savi_autograd as SG
                n_it = self.num_iters_first if step == 0 else self.num_iters
                if n_it <= 0:
                    return slots
                out, masks = SG.slot_attention_train(self, f32c(inputs), f32c(slots), n_it)
                self.attention_masks = masks
                return out
            x = f32c(inputs)
            B, N, _ = x.shape
        no_grad_only("SlotAttention", slots, self.to_q.weight, *([] if x is None else [x]))
        slots = f32c(slots)
        S = slots.shape[1]
        self.attention_masks = None
        num_iters = self.num_iters_first if step == 0 else self.num_iters
        if num_iters <= 0:
            return slots
        mode = self.mode
        if planes is not None and mode != "tc":
            raise ValueError("pre-normalised feature planes are only consumed by SlotAttention.mode == 'tc'")
        w = self._weights()
        dev = slots.device
        if mode == "tc" and planes is None:
            planes, first_token = ops.sa_prepare(x, self.norm_input.eps), 0
        if mode == "tc":
            # persistent kernel with a balanced tile partition: a frame's tokens are shared by a few CTAs, one partial slot each
            n_chunks = ops.sa_tc2_chunks(B, N)
        else:
            tokens = self.TOKENS_PER_CTA
            n_chunks = max(1, (N + tokens - 1) // tokens)
        partial = torch.empty(B, n_chunks, ops.sa_partial_floats(S), device=dev, dtype=torch.float32)
        A = torch.empty(B, S, self.dim_slots, device=dev, dtype=torch.float32)
        a0 = torch.empty(B, S, device=dev, dtype=torch.float32)
        a1 = torch.empty(B, S, device=dev, dtype=torch.float32)
        masks = torch.empty(B, S, N, device=dev, dtype=torch.float32)
        cur = slots
        out = torch.empty_like(slots)
        hidden = self.mlp[0].weight.shape[0]
        eps_ln = self.norm_input.eps
        fast_update = hidden in (128, 256, 384)

        def update(partial_, src, dst, do_update, do_prep):
            if fast_update:
                ops.sa_update2(partial_, n_chunks, src, dst, w, A, a0, hidden, eps_ln, do_update, do_prep)
            else:
                ops.sa_update(partial_, n_chunks, src, dst, w, A, a0, a1, hidden, self.scale, eps_ln, do_update, do_prep)
        # folded queries of the first iteration
        update(None, cur, None, False, True)
        for it in range(num_iters):
            last = it == num_iters - 1
            mk = masks if last else None
            if mode == "tc":
                ops.sa_main_tc2(planes, first_token, B, N, A, a0, partial, mk, n_chunks, self.epsilon)
            elif mode == "tc1":
                ops.sa_main_tc(x, A, a0, partial, mk, n_chunks, eps_ln, self.epsilon)
            else:
                ops.sa_main(x, A, a0, a1, partial, mk, n_chunks, eps_ln, self.epsilon)
            update(partial, cur, out, True, not last)
            cur = out
        self.attention_masks = masks
        return out

    def get_attention_masks(self, shape=None):
        """ last attention masks, (B, N_slots, N_locs) or reshaped to (B, N_slots, *shape) """
        B, N_slots, _ = self.attention_masks.shape
        masks = self.attention_masks
        if shape is not None:
            masks = masks.reshape(B, N_slots, *shape)
        return masks


########################
# ATTENTION MECHANISMS #
########################

def _mask_to_slots_per_frame(mask, L):
    """
    The kernels implement `mask=None` and the block-causal staircase of build_slot_causal_mask.  An explicit
    mask tensor (reference API) is mapped to its slots-per-frame; anything else is rejected loudly.
    """
    if mask is None:
        return 0
    m = (mask[:L, :L] != 0).cpu()
    for spf in range(1, L + 1):
        if L % spf:
            continue
        f = torch.arange(L) // spf
        if torch.equal(m, f[None, :] <= f[:, None]):
            return spf
    raise NotImplementedError("attention mask is not a block-causal (time x slot) staircase; only mask=None and "
                              "build_slot_causal_mask-style masks are implemented")


class MetaAttention(nn.Module):
    """ Base of the key-value attention modules (attention.py:144-263): q/k/v projections without bias and an
    output projection iff num_heads*head_dim != emb_dim or project_out. """

    def __init__(self, emb_dim, head_dim=None, num_heads=1, dropout=0., self_attn=True,
                 kv_dim=None, project_out=False, **kwargs):
        assert num_heads >= 1
        super().__init__()
        head_dim = head_dim if head_dim is not None else emb_dim
        inner_dim = num_heads * head_dim
        project_out = not inner_dim == emb_dim or project_out
        self.emb_dim = emb_dim
        self.head_dim = head_dim
        self.num_heads = num_heads
        self.inner_dim = inner_dim
        if self_attn:
            kv_dim = emb_dim
        elif kv_dim is None:
            raise ValueError(f"{kv_dim = } cannot be None in cross-attention mode")
        self.q = nn.Linear(emb_dim, inner_dim, bias=False)
        self.k = nn.Linear(kv_dim, inner_dim, bias=False)
        self.v = nn.Linear(kv_dim, inner_dim, bias=False)
        self.out_proj = nn.Sequential(
                nn.Linear(inner_dim, emb_dim),
                nn.Dropout(dropout)
            ) if project_out else nn.Identity()
        self.attention_masks = None
        self._mask_cache = {}
        self._cache = DerivedWeights()

    def forward(self, x):
        raise NotImplementedError("Base-Class does not implement a 'forward' method...")

    def attention(self, query, key, value, mask=None, **kwargs):
        """ reference API (attention.py:211-240): softmax(q k^T / sqrt(d_h), mask==0 -> -1e9) v on head-split tensors
        (B * heads, L, d_h), through the attention kernel (self-attention geometry: as many keys as queries; mask=None or the
        block-causal staircase).  The (L, L) attention maps are not materialised. """
        require_cuda(query, type(self).__name__)
        no_grad_only(type(self).__name__, query, key, value)
        if query.shape != key.shape or key.shape != value.shape:
            raise NotImplementedError("MetaAttention.attention: the kernel covers self-attention shapes (num_queries == num_keys)")
        n, L, dh = query.shape
        qkv = torch.cat([f32c(query), f32c(key), f32c(value)], dim=-1).reshape(n * L, 3 * dh)
        scale_fix = (dh ** -0.5) / (self.head_dim ** -0.5)             # the reference scales by self.head_dim, the kernel by d_h
        if scale_fix != 1.0:
            qkv[:, :dh] *= scale_fix
        return ops.mha(qkv, n, L, 1, dh, self.spf_of(mask, L)).view(n, L, dh)

    def get_attention_masks(self, reshape=None):
        """ The fused kernel never writes the (B*heads, L, L) attention maps to memory. """
        if self.attention_masks is None:
            raise ValueError("Attention masks have not yet been computed...")
        return self.attention_masks

    def split_into_heads(self, x):
        b, n, _ = x.shape
        return x.view(b, n, self.num_heads, self.head_dim).transpose(1, 2).reshape(b * self.num_heads, n, self.head_dim)

    def merge_heads(self, x):
        _, n, dh = x.shape
        return x.reshape(-1, self.num_heads, n, dh).transpose(1, 2).reshape(-1, n, self.num_heads * dh)

    # -- kernel path ---------------------------------------------------------------------------------------
    def qkv_weight(self):
        srcs = [self.q.weight, self.k.weight, self.v.weight]
        if torch.is_grad_enabled() and any(w.requires_grad for w in srcs):
            return torch.cat(srcs, dim=0)              # training: keep the autograd link to q / k / v
        return self._cache.get(srcs, lambda: torch.cat([w.detach() for w in srcs], dim=0).contiguous())

    def has_out_proj(self):
        return not isinstance(self.out_proj, nn.Identity)

    def spf_of(self, mask, L):
        if mask is None:
            return 0
        key = (mask.data_ptr(), tuple(mask.shape), L)
        if key not in self._mask_cache:
            self._mask_cache[key] = _mask_to_slots_per_frame(mask, L)
        return self._mask_cache[key]

    def run(self, x, ln=None, res=None, slots_per_frame=0):
        """ x (B, L, E) contiguous -> attention output (B, L, E) (+ res); `ln` fuses the pre-LayerNorm """
        B, L, E = x.shape
        qkv = AG.linear(x, self.qkv_weight(), None, ln=ln)
        if self.has_out_proj():
            lin = self.out_proj[0]
            if (ops.planes_ok(B * L, self.num_heads * self.head_dim) and 12 <= L <= 96 and self.head_dim % 4 == 0
                    and not AG.needs_grad(qkv, res, lin.weight, lin.bias)):
                # inference on the tensor cores: the attention kernel writes the projection's operand as fp16 hi/lo planes
                att = ops.mha(qkv, B, L, self.num_heads, self.head_dim, slots_per_frame, out_planes=True)
                return ops.linear(att, lin.weight, lin.bias, res=res).view(B, L, -1)
            att = AG.mha(qkv, B, L, self.num_heads, self.head_dim, slots_per_frame)
            return AG.linear(att.view(B, L, -1), lin.weight, lin.bias, res=res)
        att = AG.mha(qkv, B, L, self.num_heads, self.head_dim, slots_per_frame, res=res)
        return att.view(B, L, E)


class MultiHeadSelfAttention(MetaAttention):
    """ Multi-head dot-product self-attention (attention.py:267-316) """

    def __init__(self, emb_dim, head_dim, num_heads=8, dropout=0., project_out=False):
        super().__init__(emb_dim=emb_dim, head_dim=head_dim, num_heads=num_heads, self_attn=True,
                         dropout=dropout, project_out=project_out)

    def forward(self, x, mask=None, **kwargs):
        require_cuda(x, "MultiHeadSelfAttention")
        x = f32c(x)
        return self.run(x, slots_per_frame=self.spf_of(mask, x.shape[1]))


class MultiHeadCrossAttention(MetaAttention):
    """ Cross-attention variant (attention.py:320-371).  No shipped configuration instantiates it
    (`use_cross_attn=False` everywhere); the parameter tree is kept, the forward is not implemented. """

    def __init__(self, emb_dim, head_dim, kv_dim, num_heads=8, dropout=0.):
        super().__init__(emb_dim=emb_dim, head_dim=head_dim, num_heads=num_heads, dropout=dropout,
                         self_attn=False, kv_dim=kv_dim)

    def forward(self, enc_embs, query_embs, **kwargs):
        raise NotImplementedError("MultiHeadCrossAttention has no kernel: no shipped PlaySlot config uses cross-attention")


######################
# TRANSFORMER BLOCKS #
######################

def run_block_stack(blocks, x, slots_per_frame=0, cache=None):
    """ x (B, L, E) through a list of pre-LN blocks.  Inference with embed_dim 128 / MLP 512 / <= 128 tokens: ONE fused launch for
    the whole stack (csrc/tblock_tc.cu, 4-CTA clusters, activations on chip); otherwise block by block (4-5 launches each).
    `cache`: a DerivedWeights holding the packed operands of the stack """
    blocks = list(blocks)
    B, L, E = x.shape
    params = [p for b in blocks for p in b.parameters()]
    if ops.tblock_stack_applicable(blocks, L) and not AG.needs_grad(x, *params):
        cache = cache if cache is not None else blocks[0].__dict__.setdefault("_stack_cache", DerivedWeights())
        pack = cache.get(params, lambda: ops.pack_tblock_stack(blocks))
        return ops.tblock_stack(x.reshape(B * L, E), pack, B, L, slots_per_frame).view(B, L, E)
    for blk in blocks:
        x = blk.run(x, slots_per_frame=slots_per_frame)
    return x


class TransformerBlock(nn.Module):
    """ Common base (attention.py:380-421): owns `ln_mlp` (eps 1e-6) and `mlp` """

    def __init__(self, embed_dim, head_dim, num_heads, mlp_size, self_attn=True):
        super().__init__()
        self.embed_dim = embed_dim
        self.mlp_size = mlp_size
        self.head_dim = head_dim
        self.num_heads = num_heads
        self.self_attn = self_attn
        assert num_heads >= 1
        self.ln_mlp = nn.LayerNorm(embed_dim, eps=1e-6)
        self.mlp = MLP(in_dim=embed_dim, hidden_dim=mlp_size)

    def forward(self, inputs):
        raise NotImplementedError("Base Class does not implement 'forward' function...")

    def get_attention_masks(self, reshape=None):
        return self.attn.get_attention_masks(reshape=reshape)

    def run(self, x, slots_per_frame=0):
        """ pre-LN block on x (B, L, E): y = x + MHSA(LN(x)); out = y + MLP(LN(y)).  4-5 kernel launches. """
        y = self.attn.run(x, ln=(self.ln_att.weight, self.ln_att.bias, self.ln_att.eps), res=x,
                          slots_per_frame=slots_per_frame)
        return self.mlp.run(y, ln=(self.ln_mlp.weight, self.ln_mlp.bias, self.ln_mlp.eps), res=y)


class TransformerEncoderBlock(TransformerBlock):
    """ Pre-LN encoder block (attention.py:425-475); used as SAVi's transition module and inside InvDynS """

    def __init__(self, embed_dim, head_dim=32, num_heads=4, mlp_size=256, self_attn=True, project_out=False):
        super().__init__(embed_dim=embed_dim, head_dim=head_dim, num_heads=num_heads, mlp_size=mlp_size, self_attn=True)
        self.ln_att = nn.LayerNorm(embed_dim, eps=1e-6)
        self.attn = MultiHeadSelfAttention(emb_dim=embed_dim, head_dim=head_dim, num_heads=num_heads,
                                           project_out=project_out)

    def forward(self, inputs, mask=None):
        assert inputs.ndim == 3
        require_cuda(inputs, "TransformerEncoderBlock")
        x = f32c(inputs)
        return run_block_stack([self], x, self.attn.spf_of(mask, x.shape[1]))


class TransformerDecoder(TransformerBlock):
    """ Decoder block (attention.py:479-565); only the self-attention flavour (use_cross_attn=False) that the
    cOCVP / Markov dynamics models instantiate has a kernel. """

    def __init__(self, embed_dim, head_dim, num_heads, mlp_size, kv_dim=None, dropout=0,
                 use_cross_attn=False, project_out=False):
        super().__init__(embed_dim=embed_dim, head_dim=head_dim, num_heads=num_heads, mlp_size=mlp_size, self_attn=False)
        self.use_cross_attn = use_cross_attn
        self.ln_att = nn.LayerNorm(embed_dim, eps=1e-6)
        self.attn = MultiHeadSelfAttention(emb_dim=embed_dim, head_dim=head_dim, num_heads=num_heads,
                                           dropout=dropout, project_out=project_out)
        if use_cross_attn:
            if kv_dim is None:
                raise ValueError(f"If {use_cross_attn = }, 'kv_dim' must be provided...")
            self.ln_cross_att_q = nn.LayerNorm(embed_dim, eps=1e-6)
            self.ln_cross_att_kv = nn.LayerNorm(kv_dim, eps=1e-6)
            self.cross_attn = MultiHeadCrossAttention(emb_dim=embed_dim, head_dim=head_dim, num_heads=num_heads,
                                                      kv_dim=kv_dim, dropout=dropout)

    def forward(self, queries, feats=None, self_attn_mask=None, cross_attn_mask=None, slots_per_frame=None):
        assert queries.ndim == 3
        if self.use_cross_attn:
            raise NotImplementedError("TransformerDecoder with cross-attention has no kernel (unused by PlaySlot configs)")
        require_cuda(queries, "TransformerDecoder")
        x = f32c(queries)
        spf = slots_per_frame if slots_per_frame is not None else self.attn.spf_of(self_attn_mask, x.shape[1])
        return run_block_stack([self], x, spf)
