""" Loading the committed golden fixtures (made by oracle/make_golden.py from the real reference). """
import os

import torch

from synth import synth_state_dict, synth_video, synth_noise

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PIPELINE_CASES = ["blockpush", "blockpush_sharp", "buttonpress", "gridshapes5", "sketchy_sharp",
                  "gridshapes5_bench", "blockpush_bench"]          # *_bench: the benchmark length, 6 seed + 15 predicted frames


def load_case(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    scfg, pcfg = fx["savi_cfg"], fx["predictor_cfg"]
    B, ns, npred, wseed = fx["B"], fx["num_seed"], fx["num_preds"], fx["weight_seed"]
    T = ns + npred
    H, W = scfg["encoder"]["encoder_params"]["resolution"]
    S, D, A = scfg["num_slots"], scfg["slot_dim"], pcfg["action_dim"]
    ssd = synth_state_dict(fx["savi_manifest"], wseed, fx["sharpen"])
    psd = synth_state_dict(fx["predictor_manifest"], wseed + 100)
    video = synth_video(B, T, H, W, wseed)
    per_slot = pcfg["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
    noise = synth_noise((B, T - 1, S, A) if per_slot else (B, T - 1, A), wseed)
    init_noise = synth_noise((B, S, D), wseed + 1) if scfg["initializer"] == "LearnedRandom" else None
    # the fixtures pin the RNG streams they were generated with
    cs = fx["video_checksum"]
    assert abs(float(video.double().sum()) - cs[0]) < 1e-6 * abs(cs[0]), "synthetic video RNG stream differs from the fixture"
    assert abs(float(noise.double().sum()) - fx["noise_checksum"]) < 1e-6 * max(1.0, abs(fx["noise_checksum"]))
    return fx, ssd, psd, video, noise, init_noise
