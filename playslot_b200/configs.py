"""
Hyper-parameters of the shipped PlaySlot experiments (experiments/<name>/predictors/PlaySlot/experiment_params.json
in the reference), restated so that bench.py / smoke() / tests can build the models on a machine where the
reference tree is absent.  tests/test_host_cpu.py checks them against the reference's JSON files when
/root/reference is mounted.
"""
import copy


def _savi(num_slots, enc_channels, mlp_hidden, initializer, resolution=(64, 64), downsample_decoder=False):
    res = list(resolution)
    return {
        "model_name": "SAVi",
        "model_params": {
            "num_slots": num_slots, "slot_dim": 128, "num_iterations_first": 3, "num_iterations": 1,
            "in_channels": 3, "mlp_hidden": mlp_hidden, "mlp_encoder_dim": 128, "initializer": initializer,
            "transition_module_params": {"model_name": "TransformerBlock", "num_heads": 4, "head_dim": 32, "mlp_size": 512},
            "encoder": {"encoder_name": "ConvEncoder",
                        "encoder_params": {"num_channels": [enc_channels] * 4, "kernel_size": 5, "resolution": res,
                                           "downsample_encoder": False, "downsample": 2}},
            "decoder": {"decoder_name": "ConvDecoder",
                        "decoder_params": {"num_channels": [64] * 4, "kernel_size": 5, "resolution": list(res),
                                           "downsample_decoder": downsample_decoder, "upsample": 1}},
        },
    }


def _single_action(num_actions, action_dim, context_length=None):
    dyn = {"residual": False, "embed_dim": 256, "head_dim": 64, "num_heads": 8, "mlp_size": 1024, "num_layers": 4,
           "pos_enc_dropout": 0}
    if context_length is not None:
        dyn = {"residual": False, "embed_dim": 256, "head_dim": 64, "num_heads": 8, "mlp_size": 1024, "num_layers": 4,
               "context_length": context_length, "pos_enc_dropout": 0.0}
    return {
        "predictor_name": "SlotSingleAction",
        "predictor_params": {
            "num_actions": num_actions, "action_dim": action_dim, "slot_dim": 128, "condition_mode": "sum",
            "use_variability": True, "autoregressive_dynamics": True,
            "LatentAction": {"model_name": "VQSingleSlotLatentAction",
                             "model_params": {"emb_dim": 256, "num_layers": 4, "num_heads": 4, "head_dim": 64,
                                              "mlp_dim": 1024, "use_ema_vq": True}},
            "DynamicsModel": {"model_name": "SlotGPTDymamicsModel", "model_params": dyn},
        },
    }


def _multi_action():
    return {
        "predictor_name": "SlotLatentPredictor",
        "predictor_params": {
            "num_actions": 5, "action_dim": 8, "slot_dim": 128, "condition_mode": "sum",
            "use_variability": False, "autoregressive_dynamics": True,
            "LatentAction": {"model_name": "VQSimpleMLPSlotLatentAction",
                             "model_params": {"emb_dim": 128, "hidden_dim": 256, "use_ema_vq": True}},
            "DynamicsModel": {"model_name": "SlotGPTDymamicsModel",
                              "model_params": {"residual": False, "embed_dim": 128, "head_dim": 64, "num_heads": 8,
                                               "mlp_size": 512, "num_layers": 8, "pos_enc_dropout": 0.0}},
        },
    }


_EXPERIMENTS = {
    "BlockPush": lambda: {"model": _savi(8, 32, 256, "Learned"), "predictor": _single_action(8, 16)},
    "ButtonPress": lambda: {"model": _savi(4, 32, 256, "Learned"), "predictor": _single_action(8, 24)},
    "GridShapes5Objs": lambda: {"model": _savi(6, 64, 128, "LearnedRandom"), "predictor": _multi_action()},
    "Sketchy": lambda: {"model": _savi(8, 32, 256, "Learned", downsample_decoder=True), "predictor": _single_action(9, 16, context_length=10)},
}


def experiment(name, num_slots=None, resolution=None):
    """ {"model": ..., "predictor": ...} of a shipped experiment; optional overrides for the synthetic
    high-resolution / many-slot case (BASELINE.json configs[4]). """
    if name not in _EXPERIMENTS:
        raise KeyError(f"unknown experiment {name!r}; available: {sorted(_EXPERIMENTS)}")
    cfg = copy.deepcopy(_EXPERIMENTS[name]())
    mp = cfg["model"]["model_params"]
    if num_slots is not None:
        mp["num_slots"] = num_slots
    if resolution is not None:
        mp["encoder"]["encoder_params"]["resolution"] = list(resolution)
        mp["decoder"]["decoder_params"]["resolution"] = list(resolution)
    return cfg


def build_models(cfg, device="cuda"):
    """ (SAVi, predictor) built like lib/setup_model.py:29-116 builds them, in eval mode on `device` """
    from .models.SAVi import SAVi
    from .models.Predictors.LatentActionPredictors import SlotSingleActionPredictor, SlotLatentPredictor
    savi = SAVi(**copy.deepcopy(cfg["model"]["model_params"]))
    name = cfg["predictor"]["predictor_name"]
    cls = {"SlotSingleAction": SlotSingleActionPredictor, "SlotLatentPredictor": SlotLatentPredictor}[name]
    pred = cls(**copy.deepcopy(cfg["predictor"]["predictor_params"]))
    return savi.eval().to(device), pred.eval().to(device)
