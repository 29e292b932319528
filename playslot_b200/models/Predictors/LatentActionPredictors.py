"""
PlaySlot predictors -- mirror of src/models/Predictors/LatentActionPredictors.py: an inverse-dynamics module
(`latent_action`) infers latent actions from consecutive slot sets and a conditional dynamics model
(`dynamics_model`, cOCVP) rolls the slots forward.  Orchestration only; all arithmetic happens in the two
sub-modules' kernels.
"""
import torch
import torch.nn as nn

from .DynamicsModels import SlotGPTDymamicsModel, MarkovTransformerDynamicsModel
from .LatentAction import VQSingleSlotLatentAction, VQSimpleMLPSlotLatentAction

__all__ = ["SlotLatentPredictor", "SlotSingleActionPredictor"]

_DYNAMICS = {
    "MarkovTransformerDynamicsModel": MarkovTransformerDynamicsModel,
    "SlotGPTDymamicsModel": SlotGPTDymamicsModel,
}


def _per_slot(t, num_slots):
    """ (B, T, A) -> (B, T, S, A): the single-action model shares one action among the slots of a frame """
    return t.unsqueeze(2).repeat(1, 1, num_slots, 1)


class BaseSlotLatentAction(nn.Module):
    """ Common constructor args and roll-out logic (LatentActionPredictors.py:27-247) """

    LATENT_ACTION_MODELS = {}

    def __init__(self, slot_dim, num_actions, action_dim, condition_mode, use_variability, autoregressive_dynamics):
        super().__init__()
        self.num_actions = num_actions
        self.action_dim = action_dim
        self.slot_dim = slot_dim
        self.condition_mode = condition_mode
        self.use_variability = use_variability
        self.autoregressive_dynamics = autoregressive_dynamics

    def _get_dynamics_model(self, dynamics_model_params):
        if "model_name" not in dynamics_model_params.keys():
            raise KeyError("DynamicsModel params do not have 'model_name'...")
        name = dynamics_model_params.get("model_name", None)
        params = dynamics_model_params.get("model_params", None)
        if name not in _DYNAMICS:
            raise NameError(f"Upsi, Dynamics module name {name} not recognized...")
        return _DYNAMICS[name](slot_dim=self.slot_dim, num_actions=self.num_actions, action_dim=self.action_dim,
                               condition_mode=self.condition_mode, use_variability=self.use_variability, **params)

    def _get_latent_action(self, latent_action_params):
        if "model_name" not in latent_action_params.keys():
            raise KeyError("LatentAction params do not have 'model_name'...")
        name = latent_action_params.get("model_name", None)
        params = latent_action_params.get("model_params", None)
        if name not in self.LATENT_ACTION_MODELS:
            raise NameError(f"Latent Action model {name} is not supported...")
        return self.LATENT_ACTION_MODELS[name](slot_dim=self.slot_dim, num_actions=self.num_actions,
                                               action_dim=self.action_dim, **params)

    batch_prior_samples = True      # run the random-action roll-outs of forward_multiple_samples as one batched roll-out

    @torch.no_grad()
    def forward_multiple_samples(self, slot_history, num_samples, num_seed, num_preds, use_posterior=False):
        """
        N future roll-outs (LatentActionPredictors.py:91-163): optionally one with the posterior actions inferred
        by InvDyn from the whole clip, the others with random codebook actions.
        slot_history (B, >= num_seed [+ num_preds], S, D) -> ((B, num_samples, num_preds, S, D), prior action idx)
        """
        n_frames, S = slot_history.shape[1], slot_history.shape[2]
        if use_posterior and n_frames < num_preds + num_seed:
            raise ValueError(f"If {use_posterior = }, num_frames = {n_frames} must be >= {num_seed + num_preds}")
        if not use_posterior and n_frames < num_seed:
            raise ValueError(f"num_frames = {n_frames} must be >= {num_seed = }")
        seed = slot_history[:, :num_seed]
        rollouts, prior_actions = [], []
        if use_posterior:
            inferred = self.latent_action.compute_actions(slot_history)
            protos = inferred.get("action_proto")
            variab = inferred.get("action_variability") if self.use_variability else None
            if protos.dim() == 3:
                protos = _per_slot(protos, S)
            pred, _ = self.autoregressive_inference(seed_slots=seed, action_protos=protos, action_vars=variab, N=num_preds)
            rollouts.append(pred)
        n_prior = num_samples - 1 if use_posterior else num_samples
        if n_prior > 1 and self.batch_prior_samples:
            # The prior samples are independent roll-outs: draw their random actions (and variability latents) first, in the
            # reference's order -- the dynamics model consumes no random numbers in between -- then run ONE roll-out over
            # the batch axis n_prior * B.  The roll-out is latency-bound at batch B, so ten samples cost little more than one.
            draws = []
            for _ in range(n_prior):
                protos, idx = self.get_random_actions(slots=seed, num_preds=num_preds)
                variab = self.get_random_latents(slots=seed, num_preds=num_preds, src_action_idx=idx) if self.use_variability else None
                draws.append((protos, idx, variab))
            B = seed.shape[0]
            pred, _ = self.autoregressive_inference(
                seed_slots=seed.repeat(n_prior, 1, 1, 1),
                action_protos=torch.cat([d[0] for d in draws], dim=0),
                action_vars=torch.cat([d[2] for d in draws], dim=0) if self.use_variability else None, N=num_preds)
            for i, d in enumerate(draws):
                rollouts.append(pred[i * B:(i + 1) * B])
                prior_actions.append(d[1])
        else:
            for _ in range(n_prior):
                pred, idx = self.autoregressive_inference(seed_slots=seed, action_protos=None, action_vars=None, N=num_preds)
                rollouts.append(pred)
                prior_actions.append(idx)
        all_pred = torch.stack(rollouts, dim=1)
        if len(prior_actions) > 0:
            prior_actions = torch.stack(prior_actions, dim=1)
        return all_pred, prior_actions

    @torch.no_grad()
    def autoregressive_inference(self, seed_slots, action_protos=None, action_vars=None, N=10):
        """ N-step roll-out from the seed slots (LatentActionPredictors.py:166-226); random actions / variability
        embeddings are drawn when none are given. """
        B, n_seed, S, D = seed_slots.shape
        if action_protos is None:
            action_protos, action_idx = self.get_random_actions(slots=seed_slots, num_preds=N)
        else:
            action_idx = None
            if action_protos.dim() != 4:
                action_protos = action_protos.unsqueeze(-2).repeat(1, 1, S, 1)
        if not self.use_variability:
            action_vars = None
        elif action_vars is None:
            assert action_idx is not None
            action_vars = self.get_random_latents(slots=seed_slots, num_preds=N, src_action_idx=action_idx)
        elif action_vars.dim() != 4:
            action_vars = action_vars.unsqueeze(-2).repeat(1, 1, S, 1)
        pred = self.dynamics_model(slots=seed_slots, action_protos=action_protos, action_vars=action_vars,
                                   num_seed=n_seed, num_preds=N, autoregressive=True, get_pred_only=True)
        return pred.reshape(B, N, S, D), action_idx

    def get_actions_emb(self, action_idx, size=(1,)):
        """ codebook embeddings for indices given as int / list / tensor (LatentActionPredictors.py:229-247) """
        shape = None
        if isinstance(action_idx, int):
            action_idx = [action_idx]
        elif torch.is_tensor(action_idx):
            shape = action_idx.shape
            action_idx = action_idx.flatten().tolist()
        if not isinstance(action_idx, (list, tuple)):
            raise TypeError(f"{type(action_idx) = } must be a list/tuple...")
        emb = torch.stack([self.latent_action.get_action(action_idx=i, shape=size)[0] for i in action_idx])
        if shape is not None:
            emb = emb.reshape(*shape, -1)
        return emb

    def _random_latents(self, slots, size, src_action_idx):
        """ variability = (target codeword - source codeword) * U(0, 0.5); RNG calls as upstream (:383-397) """
        target_idx = torch.randint(low=0, high=self.num_actions, size=size).flatten()
        src = self.get_actions_emb(action_idx=src_action_idx, size=(1,)).reshape(*size, -1)
        tgt = self.get_actions_emb(action_idx=target_idx, size=(1,)).reshape(*size, -1)
        alpha = torch.rand(size=size, device=slots.device) / 2
        return (tgt - src) * alpha.unsqueeze(-1)


class SlotLatentPredictor(BaseSlotLatentAction):
    """ PlaySlot with InvDynM: one latent action per slot (LatentActionPredictors.py:252-398) """

    LATENT_ACTION_MODELS = {"VQSimpleMLPSlotLatentAction": VQSimpleMLPSlotLatentAction}

    def __init__(self, slot_dim, num_actions, action_dim, condition_mode, use_variability,
                 autoregressive_dynamics, LatentAction, DynamicsModel):
        super().__init__(slot_dim=slot_dim, num_actions=num_actions, action_dim=action_dim,
                         use_variability=use_variability, condition_mode=condition_mode,
                         autoregressive_dynamics=autoregressive_dynamics)
        self.latent_action = self._get_latent_action(latent_action_params=LatentAction)
        self.dynamics_model = self._get_dynamics_model(dynamics_model_params=DynamicsModel)

    def forward(self, slots, num_seed, num_preds, **kwargs):
        """ training-style pass: posterior actions from all frames, then prediction of frames 2..ns+np.
        slots (B, >= ns+np, S, D) -> (pred_slots (B, ns+np-1, S, D), dict) """
        n_frames = slots.shape[1]
        if n_frames < num_seed + num_preds:
            raise ValueError(f"num_frames = {n_frames} must be >= {num_seed + num_preds = }")
        slots = slots[:, :num_seed + num_preds]
        out = self.latent_action(slots)
        protos = out.pop("action_proto")
        variab = out.get("action_variability", None)
        pred = self.dynamics_model(slots=slots[:, :-1], action_protos=protos,
                                   action_vars=variab if self.use_variability else None,
                                   num_seed=num_seed, num_preds=num_preds, autoregressive=self.autoregressive_dynamics)
        return pred, {"pred_slots": pred, **out}

    def get_random_actions(self, slots, num_preds):
        B, n_seed, S, _ = slots.shape
        return self.latent_action.get_action(shape=(B, n_seed + num_preds, S))

    def get_random_latents(self, slots, num_preds, src_action_idx=0):
        B, n_seed, S, _ = slots.shape
        return self._random_latents(slots, (B, n_seed + num_preds, S), src_action_idx)


class SlotSingleActionPredictor(BaseSlotLatentAction):
    """ PlaySlot with InvDynS: a single latent action per time step, shared by all slots
    (LatentActionPredictors.py:403-532) """

    CONDITION_MODES = ["sum", "concat"]
    LATENT_ACTION_MODELS = {"VQSingleSlotLatentAction": VQSingleSlotLatentAction}

    def __init__(self, slot_dim, num_actions, action_dim, condition_mode, use_variability,
                 autoregressive_dynamics, LatentAction, DynamicsModel):
        super().__init__(slot_dim=slot_dim, num_actions=num_actions, action_dim=action_dim,
                         condition_mode=condition_mode, use_variability=use_variability,
                         autoregressive_dynamics=autoregressive_dynamics)
        self.latent_action = self._get_latent_action(latent_action_params=LatentAction)
        self.dynamics_model = self._get_dynamics_model(dynamics_model_params=DynamicsModel)

    def forward(self, slots, num_seed, num_preds, **kwargs):
        """ slots (B, ns+np, S, D) -> (pred_slots (B, ns+np-1, S, D), dict); extra kwargs (imgs=, actions=) are
        accepted and ignored like upstream (04_train_PlaySlot.py:78-84). """
        out = self.latent_action(slots)
        protos = out.pop("action_proto")
        variab = out.get("action_variability", None)
        S = slots.shape[2]
        protos = _per_slot(protos, S)
        if variab is not None:
            variab = _per_slot(variab, S)
        pred = self.dynamics_model(slots=slots[:, :-1], action_protos=protos,
                                   action_vars=variab if self.use_variability else None,
                                   num_seed=num_seed, num_preds=num_preds, autoregressive=self.autoregressive_dynamics)
        return pred, {"pred_slots": pred, **out}

    def get_random_actions(self, slots, num_preds):
        B, n_seed, S, _ = slots.shape
        emb, idx = self.latent_action.get_action(shape=(B, n_seed + num_preds))
        return _per_slot(emb, S), idx

    def get_random_latents(self, slots, num_preds, src_action_idx=0):
        B, n_seed, _, _ = slots.shape
        return self._random_latents(slots, (B, n_seed + num_preds, 1), src_action_idx)
