"""
Action-conditional OCVP -- mirror of src/models/Predictors/ActionConditional_OCVP.py: ground-truth robot actions
are embedded with one Linear layer and condition the same dynamics models as PlaySlot (no variability).
"""
import torch.nn as nn

from ... import ops
from .._common import require_cuda, no_grad_only, f32c
from .DynamicsModels import SlotGPTDymamicsModel, MarkovTransformerDynamicsModel

_DYNAMICS = {
    "MarkovTransformerDynamicsModel": MarkovTransformerDynamicsModel,
    "SlotGPTDymamicsModel": SlotGPTDymamicsModel,
}


class ActionCondtionalOCVP(nn.Module):
    """ (class name spelled as upstream) ctor args: slot_dim, raw_action_dim, action_embed_dim, condition_mode,
    autoregressive_dynamics, DynamicsModel (ActionConditional_OCVP.py:16-60) """

    def __init__(self, slot_dim, raw_action_dim, action_embed_dim, condition_mode, autoregressive_dynamics, DynamicsModel):
        super().__init__()
        self.slot_dim = slot_dim
        self.raw_action_dim = raw_action_dim
        self.action_embed_dim = action_embed_dim
        self.condition_mode = condition_mode
        self.autoregressive_dynamics = autoregressive_dynamics
        self.action_encoder = nn.Sequential(nn.Linear(raw_action_dim, action_embed_dim))
        self.dynamics_model = self._get_dynamics_model(dynamics_model_params=DynamicsModel)

    def _get_dynamics_model(self, dynamics_model_params):
        if "model_name" not in dynamics_model_params.keys():
            raise KeyError("DynamicsModel params do not have 'model_name'...")
        name = dynamics_model_params.get("model_name", None)
        params = dynamics_model_params.get("model_params", None)
        if name not in _DYNAMICS:
            raise NameError(f"Upsi, Dynamics module name {name} not recognized...")
        return _DYNAMICS[name](slot_dim=self.slot_dim, num_actions=4, action_dim=self.action_embed_dim,
                               condition_mode=self.condition_mode, use_variability=False,
                               action_is_one_hot=False, **params)

    def forward(self, slot_history, actions, num_seed, num_preds, **kwargs):
        """ slot_history (B, >= ns+np-1, S, D), actions (B, >= ns+np-1, raw_action_dim) ->
        (pred_slots (B, ns+np-1, S, D), {"pred_slots": ...}) (ActionConditional_OCVP.py:97-128) """
        n = num_seed + num_preds - 1
        n_frames, S = slot_history.shape[1], slot_history.shape[2]
        if n_frames < n:
            raise ValueError(f"num_frames = {n_frames} too small for {(num_seed, num_preds) = }")
        if actions.shape[1] < n:
            raise ValueError(f"num_actions = {actions.shape[1]} too small for {(num_seed, num_preds) = }")
        require_cuda(slot_history, "ActionCondtionalOCVP")
        no_grad_only("ActionCondtionalOCVP", slot_history, self.action_encoder[0].weight)
        slot_history = slot_history[:, :n]
        actions = f32c(actions[:, :n])
        # raw action dims are tiny (e.g. 4): the embedding GEMM needs K % 4 == 0, pad with zero columns
        lin = self.action_encoder[0]
        w, a = lin.weight, actions
        if w.shape[1] % 4:
            padk = 4 - w.shape[1] % 4
            w = nn.functional.pad(w.detach(), (0, padk)).contiguous()
            a = nn.functional.pad(a, (0, padk)).contiguous()
        emb = ops.linear(a, w, lin.bias)                                    # (B, n, A)
        emb = emb.unsqueeze(2).repeat(1, 1, S, 1)
        pred = self.dynamics_model(slots=slot_history, action_protos=emb, action_vars=None, num_seed=num_seed,
                                   num_preds=num_preds, autoregressive=self.autoregressive_dynamics)
        return pred, {"pred_slots": pred}
