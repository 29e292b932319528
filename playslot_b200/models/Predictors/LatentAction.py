"""
Inverse-dynamics (InvDyn) modules -- mirror of src/models/Predictors/LatentAction.py.
Given slots [s_1..s_T] they infer the latent action between consecutive frames: a codebook prototype plus a
continuous variability embedding.

Kernel mapping: slot encoder / transformer / heads -> ps_linear + ps_mha; mean-var difference, Gaussian
sample, VQ lookup, straight-through prototype, variability and VQ loss -> ps_invdyn_tail (one launch).
The Gaussian noise is drawn on the host with the same call as the reference (CPU generator,
LatentAction.py:73) so both consume identical random streams.
"""
import torch
import torch.nn as nn

from ... import ops
from ... import autograd as AG
from .._common import DerivedWeights, require_cuda, f32c
from ..BlocksUtils.attention import TransformerEncoderBlock
from ..BlocksUtils.quantization import VectorQuantizer, EmaVectorQuantizer

__all__ = ["VQSingleSlotLatentAction", "VQSimpleMLPSlotLatentAction"]


class BaseSlotLatentAction(nn.Module):
    """ Shared heads and codebook (LatentAction.py:24-125): mean_fc, variance_fc, quantizer, slot_encoder """

    def __init__(self, slot_dim, emb_dim, action_dim, num_actions, use_ema_vq=False, **kwargs):
        super().__init__()
        self.slot_dim = slot_dim
        self.emb_dim = emb_dim
        self.action_dim = action_dim
        self.num_actions = num_actions
        self.use_ema_vq = use_ema_vq
        self.mean_fc = nn.Linear(emb_dim, action_dim)
        self.variance_fc = nn.Linear(emb_dim, action_dim)
        quantizer_cls = EmaVectorQuantizer if use_ema_vq else VectorQuantizer
        self.quantizer = quantizer_cls(num_embs=num_actions, emb_dim=action_dim, nbooks=1)
        self._setup_slot_encoder()
        self._head_cache = DerivedWeights()

    # -- pieces kept for API parity ---------------------------------------------------------------------------
    def sample(self, mean, var, eps=1e-6):
        noise = torch.randn(mean.size(), dtype=torch.float32).to(mean.device)
        return noise * torch.sqrt(var + eps) + mean

    def comput_action_dist(self, tokens):
        """ mean / |variance| heads and their temporal difference / sum (LatentAction.py:76-87) """
        mv = self._heads(f32c(tokens))
        A = self.action_dim
        mean, var = mv[..., :A], mv[..., A:].abs()
        return mean[:, 1:] - mean[:, :-1], var[:, 1:] + var[:, :-1]

    def compute_actions(self, slots):
        return self(slots)

    def get_action(self, action_idx=None, shape=None):
        """ codebook prototypes for given (or uniformly sampled) indices (LatentAction.py:94-113) """
        assert shape is not None, "A shape argument must be specified..."
        device = self.mean_fc.weight.device
        if action_idx is None:
            action_idx = torch.randint(low=0, high=self.quantizer.num_embs, size=shape, device=device)
        else:
            action_idx = torch.tensor(action_idx, device=device).expand(shape)
        return self.quantizer.get_codebook_entry(action_idx), action_idx

    def decompose_action_latent(self, action_latent):
        """ latent -> (prototype, variability) (LatentAction.py:115-125) """
        protos, _, _ = self.quantizer(action_latent)
        variability, _ = self.quantizer.get_variability(z=action_latent, action_embs=protos)
        return protos, variability

    # -- kernel path ----------------------------------------------------------------------------------------
    def _codebook(self):
        return f32c(self.quantizer.get_codewords().detach())

    def _heads(self, tokens):
        """ one GEMM for both heads: (..., E) -> (..., 2A) = [mean_fc | variance_fc] """
        srcs = [self.mean_fc.weight, self.mean_fc.bias, self.variance_fc.weight, self.variance_fc.bias]
        if torch.is_grad_enabled() and (tokens.requires_grad or any(t.requires_grad for t in srcs)):
            w = torch.cat([self.mean_fc.weight, self.variance_fc.weight], 0)         # training: live parameters
            b = torch.cat([self.mean_fc.bias, self.variance_fc.bias], 0)
            return AG.linear(tokens, w, b)
        w, b = self._head_cache.get(srcs, lambda: (
            torch.cat([self.mean_fc.weight.detach(), self.variance_fc.weight.detach()], 0).contiguous(),
            torch.cat([self.mean_fc.bias.detach(), self.variance_fc.bias.detach()], 0).contiguous()))
        return ops.linear(tokens, w, b)

    def _tail(self, mv, B, T, R, per_slot):
        """ mv (B, T, R, 2A) -> output dict of `forward` """
        A = self.action_dim
        shape = (B, T - 1, A) if not per_slot else (B, T - 1, R, A)
        static = getattr(self, "_static_noise", None)
        if static is not None:
            # CUDA-graph replay (pipeline.GraphedSeedRollout): the host draws the noise with the same call as below
            # and uploads it into this fixed device buffer before every replay
            assert static.numel() == B * (T - 1) * R * A, "static noise buffer does not match the captured shape"
            noise = static
        else:
            noise = torch.randn(shape, dtype=torch.float32).to(mv.device)   # same RNG call as the reference
        codebook = self.quantizer.get_codewords()
        out = AG.invdyn_tail(mv, f32c(noise.view(B, T - 1, R, A)), codebook, per_slot)
        idx = out["idx"]
        variability = out["variability"]
        if self.training:
            self._train_side_effects(out["z"].detach(), idx)
            if not per_slot:
                # the reference looks the codeword up a second time for the variability (forward_one inside get_variability,
                # quantization.py:365-368): its usage meter is updated a second time per forward ...
                vq = self.quantizer.vq if self.use_ema_vq else self.quantizer
                if self.use_ema_vq and self.quantizer.update:
                    # ... and in train mode with EMA the second lookup sees the ALREADY-UPDATED codebook; z keeps its
                    # gradient, the codeword does not
                    _, zq2, idx2, _ = ops.vq_lookup(f32c(out["z"].detach()), self._codebook())
                    variability = out["z"] - zq2
                    vq.track_assigment(idx2.reshape(-1, 1).detach(), 0)
                else:
                    vq.track_assigment(idx.reshape(-1, 1).detach(), 0)
        if per_slot:
            res = {
                "action_dist": out["action_dist"],                            # (B, T-1, 2, S, A)
                "sampled_latent_action": out["z"],
                "action_variability": variability,
                "action_proto": out["proto"],
                "action_idxs": idx.view(B, T - 1, R, 1),
            }
        else:
            res = {
                "action_dist": out["action_dist"].view(B, T - 1, 2, A),
                "sampled_latent_action": out["z"].view(B, T - 1, A),
                "action_variability": variability.view(B, T - 1, A),
                "action_proto": out["proto"].view(B, T - 1, A),
                "action_idxs": idx.view(B, T - 1, 1),
            }
        res["vq_losses"] = {"quant_loss": out["quant"].mean(), "commit_loss": out["commit"].mean()}
        return res

    def _train_side_effects(self, z, idx):
        """ train-mode bookkeeping of the reference: codeword-usage trackers and the EMA codebook update """
        vq = self.quantizer.vq if self.use_ema_vq else self.quantizer
        vq.track_assigment(idx.reshape(-1, 1).detach(), 0)
        if self.use_ema_vq and self.quantizer.update:
            self.quantizer.ema_update(z.squeeze(2) if z.shape[2] == 1 else z, idx)


class VQSingleSlotLatentAction(BaseSlotLatentAction):
    """ InvDynS (LatentAction.py:129-249): an [ACT] token gathers all slots of a frame through a transformer
    encoder; one latent action per time step. """

    def __init__(self, slot_dim, emb_dim, action_dim, num_actions, num_layers, num_heads, head_dim, mlp_dim,
                 use_ema_vq=False, **kwargs):
        super().__init__(slot_dim=slot_dim, emb_dim=emb_dim, action_dim=action_dim, num_actions=num_actions,
                         use_ema_vq=use_ema_vq)
        self.act_token = nn.Parameter(torch.zeros(1, 1, emb_dim))
        self.transformer = nn.Sequential(*[
                TransformerEncoderBlock(embed_dim=emb_dim, head_dim=head_dim, num_heads=num_heads,
                                        mlp_size=mlp_dim, project_out=True)
                for _ in range(num_layers)])

    def _setup_slot_encoder(self):
        self.slot_encoder = nn.Sequential(nn.LayerNorm(self.slot_dim), nn.Linear(self.slot_dim, self.emb_dim))

    def forward(self, slots):
        """ slots (B, T, S, D) -> dict(action_dist (B,T-1,2,A), sampled_latent_action, action_variability,
        action_proto (B,T-1,A), action_idxs (B,T-1,1), vq_losses) """
        if len(slots.shape) != 4:
            raise ValueError(f"{slots.shape = } must be (B, N, num_slots, slot_dim)")
        require_cuda(slots, "VQSingleSlotLatentAction")
        slots = f32c(slots)
        B, T, S, _ = slots.shape
        ln, fc = self.slot_encoder[0], self.slot_encoder[1]
        emb = AG.linear(slots.view(B * T, S, -1), fc.weight, fc.bias, ln=(ln.weight, ln.bias, ln.eps))
        if torch.is_grad_enabled() and (emb.requires_grad or self.act_token.requires_grad):
            tokens = torch.cat([self.act_token.view(1, 1, -1).expand(B * T, 1, -1), emb], dim=1).contiguous()
        else:
            tokens = torch.empty(B * T, S + 1, self.emb_dim, device=slots.device, dtype=torch.float32)
            tokens[:, 0] = self.act_token.detach().view(1, -1)
            tokens[:, 1:] = emb
        for blk in self.transformer:
            tokens = blk.run(tokens)
        act = tokens[:, 0].contiguous()                                       # (B*T, E)
        mv = self._heads(act).view(B, T, 1, -1)
        return self._tail(mv, B, T, 1, per_slot=False)


class VQSimpleMLPSlotLatentAction(BaseSlotLatentAction):
    """ InvDynM (LatentAction.py:254-359): per-slot MLP, one latent action per slot and time step """

    def __init__(self, slot_dim, emb_dim, hidden_dim, action_dim, num_actions, use_ema_vq=False, **kwargs):
        self.hidden_dim = hidden_dim
        super().__init__(slot_dim=slot_dim, emb_dim=emb_dim, hidden_dim=hidden_dim, action_dim=action_dim,
                         num_actions=num_actions, use_ema_vq=use_ema_vq, **kwargs)

    def _setup_slot_encoder(self):
        self.slot_encoder = nn.Sequential(
                nn.Linear(self.slot_dim, self.hidden_dim),
                nn.ReLU(),
                nn.LayerNorm(self.hidden_dim),
                nn.Linear(self.hidden_dim, self.emb_dim)
            )

    def forward(self, slots):
        """ slots (B, T, S, D) -> dict(action_dist (B,T-1,2,S,A), ..., action_proto (B,T-1,S,A),
        action_idxs (B,T-1,S,1), vq_losses) """
        assert len(slots.shape) == 4, f"{slots.shape = } must be (B, N, num_slots, slot_dim)"
        require_cuda(slots, "VQSimpleMLPSlotLatentAction")
        slots = f32c(slots)
        B, T, S, _ = slots.shape
        fc1, ln, fc2 = self.slot_encoder[0], self.slot_encoder[2], self.slot_encoder[3]
        h = AG.linear(slots, fc1.weight, fc1.bias, act=ops.ACT_RELU)
        e = AG.linear(h, fc2.weight, fc2.bias, ln=(ln.weight, ln.bias, ln.eps))
        mv = self._heads(e).view(B, T, S, -1)
        return self._tail(mv, B, T, S, per_slot=True)
