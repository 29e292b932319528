"""
Slot dynamics models -- mirror of src/models/Predictors/DynamicsModels.py: they take past slots plus latent
action prototypes / variability embeddings and predict the slots of the next frame.

Kernel mapping for one `forward_single`: ONE ps_linear embeds [slots | protos | vars] and adds the
sinusoidal time encoding in its epilogue; each transformer layer is ps_linear(LN+QKV) -> ps_mha (block-causal
staircase evaluated in-kernel, -1e9 fill) -> ps_linear(out_proj + residual) -> ps_linear(LN + MLP-1 + GELU) ->
ps_linear(MLP-2 + residual); the head is ps_linear(LN fused).
"""
import torch
import torch.nn as nn

from ... import autograd as AG
from .._common import DerivedWeights, require_cuda, f32c
from ..BlocksUtils.attention import TransformerDecoder, run_block_stack
from ..BlocksUtils.model_blocks import SlotPositionalEncoding
from ..BlocksUtils.model_utils import build_slot_causal_mask

__all__ = ["MarkovTransformerDynamicsModel", "SlotGPTDymamicsModel"]


class BaseSlotDynamicsModel(nn.Module):
    """ Embeddings, conditioning and regression head shared by the dynamics models (DynamicsModels.py:24-148) """

    CONDITION_MODES = ["sum", "concat"]

    def __init__(self, slot_dim=64, embed_dim=128, num_actions=None, action_dim=None,
                 condition_mode="sum", use_variability=False, residual=False, **kwargs):
        super().__init__()
        if condition_mode not in self.CONDITION_MODES:
            raise NameError(f"{condition_mode = } not in {self.CONDITION_MODES = }...")
        self.slot_dim = slot_dim
        self.num_actions = num_actions
        self.action_dim = action_dim
        self.condition_mode = condition_mode
        self.use_variability = use_variability
        self.residual = residual
        if condition_mode == "sum":
            predictor_dim = embed_dim
        else:
            predictor_dim = embed_dim + action_dim * (2 if use_variability else 1)
        self.predictor_dim = predictor_dim
        self.tok_emb = nn.Linear(slot_dim, embed_dim)
        if condition_mode == "sum":
            self.action_emb = nn.Linear(action_dim, embed_dim)
        if self.use_variability and condition_mode == "sum":
            self.variability_emb = nn.Linear(action_dim, embed_dim)
        self.regression_head = nn.Sequential(nn.LayerNorm(predictor_dim), nn.Linear(predictor_dim, slot_dim))
        self._emb_cache = DerivedWeights()

    def forward(self, slots, action_protos, action_vars=None, num_seed=None, num_preds=None,
                autoregressive=None, get_pred_only=False, **kwargs):
        """ autoregressive roll-out or a single teacher-forced step (DynamicsModels.py:91-116) """
        if autoregressive is None:
            raise ValueError(f"{autoregressive = } must be True/False...")
        if autoregressive:
            return self.forward_autoregressive(slots=slots, action_protos=action_protos, action_vars=action_vars,
                                               num_seed=num_seed, num_preds=num_preds, get_pred_only=get_pred_only)
        return self.forward_single(slots=slots, action_protos=action_protos, action_vars=action_vars)

    def condition_slots(self, tokens, action_protos, action_vars=None):
        """ reference API (DynamicsModels.py:119-148): tokens are already-embedded slots.  The fused path below
        (`_embed`) does embedding + conditioning (+ time encoding) in one GEMM instead. """
        assert action_protos is not None, "'action_protos' not provided..."
        if self.condition_mode == "sum":
            out = tokens + AG.linear(f32c(action_protos), self.action_emb.weight, self.action_emb.bias)
            if self.use_variability:
                assert action_vars is not None, "'action_vars' not provided..."
                out = out + AG.linear(f32c(action_vars), self.variability_emb.weight, self.variability_emb.bias)
            return out
        parts = [tokens, action_protos] + ([action_vars] if self.use_variability else [])
        return torch.cat(parts, dim=-1)

    # -- kernel path ------------------------------------------------------------------------------------------
    def _embed(self, slots, action_protos, action_vars, pe_table=None, n_frames=1):
        """ (B, F, S, D) -> (B*F*S, predictor_dim) conditioned tokens (+ pe_table[frame]) """
        B, Fr, S, _ = slots.shape
        M = B * Fr * S
        use_var = self.use_variability
        if self.condition_mode == "sum":
            srcs = [self.tok_emb.weight, self.tok_emb.bias, self.action_emb.weight, self.action_emb.bias]
            if use_var:
                srcs += [self.variability_emb.weight, self.variability_emb.bias]

            def build():
                ws = [self.tok_emb.weight.detach(), self.action_emb.weight.detach()]
                b = self.tok_emb.bias.detach() + self.action_emb.bias.detach()
                if use_var:
                    ws.append(self.variability_emb.weight.detach())
                    b = b + self.variability_emb.bias.detach()
                return torch.cat(ws, dim=1).contiguous(), b.contiguous()
            if torch.is_grad_enabled() and any(t.requires_grad for t in srcs):
                # training: concatenate the live parameters so that autograd reaches the three embeddings
                w = torch.cat([self.tok_emb.weight, self.action_emb.weight] + ([self.variability_emb.weight] if use_var else []), dim=1)
                b = self.tok_emb.bias + self.action_emb.bias + (self.variability_emb.bias if use_var else 0)
            else:
                w, b = self._emb_cache.get(srcs, build)
            parts = [slots, action_protos] + ([action_vars] if use_var else [])
            x = torch.cat([p.reshape(M, -1) for p in parts], dim=1)              # (M, D + A [+ A])
            if pe_table is not None:
                return AG.linear(x, w, b, res=pe_table, res_div=S, res_mod=n_frames)
            return AG.linear(x, w, b)
        tok = AG.linear(slots.reshape(M, -1), self.tok_emb.weight, self.tok_emb.bias)
        parts = [tok, action_protos.reshape(M, -1)] + ([action_vars.reshape(M, -1)] if use_var else [])
        x = torch.cat(parts, dim=1)
        if pe_table is not None:
            x = (x.view(B, Fr, S, -1) + pe_table[:Fr].view(1, Fr, 1, -1)).reshape(M, -1)
        return x

    def _head(self, tokens):
        ln, fc = self.regression_head[0], self.regression_head[1]
        return AG.linear(tokens, fc.weight, fc.bias, ln=(ln.weight, ln.bias, ln.eps))

    def _check_inputs(self, slots, action_protos, action_vars):
        if action_protos is None:
            raise ValueError("'action_protos' were not provided")
        S = slots.shape[2]
        # the reference adds (B, F, 1, E) action embeddings to (B, F, S, E) tokens by broadcasting
        # (prior samples of the single-action model, LatentActionPredictors.py:511-532): expand explicitly
        if action_protos.dim() == 4 and action_protos.shape[2] == 1 and S > 1:
            action_protos = action_protos.expand(-1, -1, S, -1)
        if action_vars is not None and action_vars.dim() == 4 and action_vars.shape[2] == 1 and S > 1:
            action_vars = action_vars.expand(-1, -1, S, -1)
        if slots.shape[:-1] != action_protos.shape[:-1]:
            raise ValueError(f"{slots.shape =} != {action_protos.shape = }")
        if self.use_variability and action_vars is None:
            raise ValueError(f"If {self.use_variability = }, latents must be provided...")
        require_cuda(slots, type(self).__name__)
        return f32c(slots), f32c(action_protos), (f32c(action_vars) if action_vars is not None else None)


class MarkovTransformerDynamicsModel(BaseSlotDynamicsModel):
    """ Markovian variant: s_t depends on s_{t-1} and a_t only (DynamicsModels.py:153-317); no mask, no time PE """

    def __init__(self, slot_dim=64, embed_dim=128, num_actions=None, action_dim=None, condition_mode="sum",
                 use_variability=False, residual=False, head_dim=32, num_heads=4, mlp_size=512, num_layers=4, **kwargs):
        super().__init__(slot_dim=slot_dim, embed_dim=embed_dim, num_actions=num_actions, action_dim=action_dim,
                         condition_mode=condition_mode, use_variability=use_variability, residual=residual)
        self.dynamics_model = nn.ModuleList([
                TransformerDecoder(embed_dim=self.predictor_dim, head_dim=head_dim, num_heads=num_heads,
                                   mlp_size=mlp_size, kv_dim=None, use_cross_attn=False, project_out=True)
                for _ in range(num_layers)])

    def forward_autoregressive(self, slots, action_protos, action_vars=None, num_seed=None, num_preds=None,
                               get_pred_only=False, **kwargs):
        assert num_seed is not None and num_preds is not None, "num_seed / num_preds must be provided in AR mode..."
        if self.use_variability:
            assert action_vars is not None, "variability latents required but not provided"
        preds, cur = [], slots[:, :1]
        for i in range(num_seed + num_preds - 1):
            nxt = self.forward_single(slots=cur, action_protos=action_protos[:, i:i+1],
                                      action_vars=action_vars[:, i:i+1] if self.use_variability else None)
            if not get_pred_only or i >= num_seed - 1:
                preds.append(nxt)
            cur = slots[:, i+1:i+2] if i < num_seed - 1 else nxt
        return torch.cat(preds, dim=1)

    def forward_single(self, slots, action_protos, action_vars=None, **kwargs):
        slots, action_protos, action_vars = self._check_inputs(slots, action_protos, action_vars)
        B, Fr, S, D = slots.shape
        t = self._embed(slots, action_protos, action_vars).view(B * Fr, S, -1)
        t = run_block_stack(self.dynamics_model, t)
        out = self._head(t).view(B, Fr, S, D)
        return out + slots if self.residual else out


class SlotGPTDymamicsModel(BaseSlotDynamicsModel):
    """ cOCVP: GPT-like block-causal transformer over (time x slot) tokens with a sliding window of
    `context_length` frames (DynamicsModels.py:321-550).  (Class name spelled as upstream.) """

    def __init__(self, slot_dim=64, embed_dim=128, num_actions=None, action_dim=None, condition_mode="sum",
                 use_variability=False, residual=False, head_dim=32, num_heads=4, mlp_size=512, num_layers=4,
                 context_length=8, pos_enc_dropout=0.1, **kwargs):
        if context_length is None:
            raise ValueError("'SlotGPTDymamicsModel' must receive 'context_length' param...")
        super().__init__(slot_dim=slot_dim, embed_dim=embed_dim, num_actions=num_actions, action_dim=action_dim,
                         condition_mode=condition_mode, use_variability=use_variability, residual=residual)
        self.context_length = context_length
        self.slot_causal_mask = None
        self.slots_in_causal_mask = None
        self.pos_emb = SlotPositionalEncoding(d_model=self.predictor_dim, max_len=context_length, dropout=pos_enc_dropout)
        self.dynamics_model = nn.ModuleList([
                TransformerDecoder(embed_dim=self.predictor_dim, head_dim=head_dim, num_heads=num_heads,
                                   mlp_size=mlp_size, kv_dim=None, use_cross_attn=False, project_out=True)
                for _ in range(num_layers)])

    def forward_autoregressive(self, slots, action_protos, action_vars=None, num_seed=None, num_preds=None,
                               get_pred_only=False, **kwargs):
        """ the whole (<= context_length)-frame window is recomputed every step, as upstream
        (DynamicsModels.py:396-454): the window slides, so positions shift and there is no KV cache. """
        assert num_seed is not None and num_preds is not None, "num_seed / num_preds must be provided in AR mode..."
        preds = []
        steps = num_preds if get_pred_only else num_seed + num_preds - 1
        cur = slots[:, :num_seed] if get_pred_only else slots[:, :1]
        for i in range(steps):
            upto = num_seed + i if get_pred_only else i + 1
            a = action_protos[:, :upto]
            v = action_vars[:, :upto] if action_vars is not None else None
            cur, a, v = self.enforce_window(slots=cur, action_protos=a, action_vars=v)
            nxt = self.forward_single(slots=cur, action_protos=a, action_vars=v)[:, -1:]
            preds.append(nxt)
            hook = getattr(self, "step_hook", None)
            if hook is not None:
                hook(i, nxt)          # pipeline.seed_rollout: decode this frame on a side stream while the next step runs
            if not get_pred_only and i < num_seed - 1:
                cur = slots[:, :upto + 1]
            else:
                cur = torch.cat([cur, nxt], dim=1)
        return torch.cat(preds, dim=1)

    def forward_single(self, slots, action_protos, action_vars=None, causal_mask=None):
        slots, action_protos, action_vars = self._check_inputs(slots, action_protos, action_vars)
        B, Fr, S, D = slots.shape
        if Fr > self.context_length:
            raise ValueError(f"num_frames = {Fr} cannot be > {self.context_length = }")
        pe = f32c(self.pos_emb.table())
        t = self._embed(slots, action_protos, action_vars, pe_table=pe, n_frames=Fr)
        if self.training and self.pos_emb.dropout.p > 0:
            # train-mode dropout after the positional encoding (model_blocks.py:211): the same nn.Dropout call on the same
            # (B, F, S, E) tensor, device RNG, like the reference on CUDA
            t = self.pos_emb.dropout(t.view(B, Fr, S, -1))
        t = t.reshape(B, Fr * S, -1)
        t = run_block_stack(self.dynamics_model, t, slots_per_frame=S)     # one fused launch for the L layers when applicable
        out = self._head(t).view(B, Fr, S, D)
        return out + slots if self.residual else out

    def enforce_window(self, slots, action_protos, action_vars=None):
        """ keep at most the last `context_length` frames (DynamicsModels.py:519-527) """
        if slots.shape[1] > self.context_length:
            c = self.context_length
            slots, action_protos = slots[:, -c:], action_protos[:, -c:]
            if action_vars is not None:
                action_vars = action_vars[:, -c:]
        return slots, action_protos, action_vars

    def get_slot_causal_mask(self, num_slots, num_frames, device):
        """ explicit staircase mask (API parity, DynamicsModels.py:530-550); the kernels derive it from indices """
        if self.slot_causal_mask is None or self.slots_in_causal_mask != num_slots:
            self.slots_in_causal_mask = num_slots
            self.slot_causal_mask = build_slot_causal_mask(num_slots=num_slots, seq_len=self.context_length, device=device)
        n = int(num_frames * num_slots)
        if self.slot_causal_mask.shape[0] < n:
            raise ValueError(f"{self.slot_causal_mask.shape = } too small for #frames & #slots")
        return self.slot_causal_mask[:n, :n]
