"""
Pins the CPU oracle (oracle/playslot_oracle.py) against the golden vectors that
oracle/make_golden.py produced by running the REAL reference modules
(/root/reference/src/models via lib.setup_model) in the build container.
CPU-only; tolerance = fp32 noise (the oracle follows the reference's op order).
"""
import os

import pytest
import torch

from golden_util import GOLDEN_DIR, PIPELINE_CASES, load_case
from synth import synth_state_dict, rel_l2, max_rel
from oracle import playslot_oracle as O

TOL = 5e-5      # oracle vs reference: same algorithm -> fp32 rounding only (sharpened softmaxes amplify it ~10x)


def _check(name, a, b, tol=TOL):
    r, m = rel_l2(a, b), max_rel(a, b)
    assert r <= tol and m <= 5 * tol, f"{name}: rel_l2={r:.3e} max_rel={m:.3e}"


@pytest.mark.parametrize("case", PIPELINE_CASES)
def test_pipeline_oracle_matches_reference(case):
    fx, ssd, psd, video, noise, init_noise = load_case(case)
    scfg, pcfg = fx["savi_cfg"], fx["predictor_cfg"]
    ns, npred, B = fx["num_seed"], fx["num_preds"], fx["B"]
    psd = dict(psd)
    E = psd["dynamics_model.tok_emb.weight"].shape[0]
    ctx = pcfg["DynamicsModel"]["model_params"].get("context_length", 8)
    psd["dynamics_model.pos_emb.pe"] = O.sinusoid_pe(ctx, E).view(1, ctx, 1, E)

    with torch.no_grad():
        sh, sa_masks = O.savi_encode_video(ssd, scfg, video, ns + npred, init_noise, return_masks=True)
        _check("slot_history", sh, fx["slot_history"])
        _check("sa_masks", sa_masks, fx["sa_masks_last"])
        pred, act = O.posterior_rollout(psd, pcfg, sh, ns, npred, noise)
        assert torch.equal(act["action_idxs"].reshape(-1), fx["action_idxs"].reshape(-1)), "VQ indices differ"
        _check("action_proto", act["action_proto"], fx["action_proto"])
        _check("action_variability", act["action_variability"], fx["action_variability"])
        _check("action_dist", act["action_dist"], fx["action_dist"])
        _check("pred_slots", pred, fx["pred_slots"])
        S, D = pred.shape[2:]
        k = fx["decode_last"]
        recon, recons, masks = O.decode(ssd, scfg, pred[:, -k:].reshape(B * k, S, D))
        _check("recon", recon, fx["recon"])
        _check("masks", masks, fx["masks"])
        _check("recons_m0", recons[0], fx["recons_m0"])


@pytest.mark.parametrize("key", ["S8", "S4"])
def test_slot_attention_oracle_matches_reference(key):
    fx = torch.load(os.path.join(GOLDEN_DIR, "slot_attention.pt"), weights_only=False)[key]
    sd = synth_state_dict(fx["manifest"], fx["weight_seed"], fx["sharpen"])
    S = int(key[1:])
    g = torch.Generator().manual_seed(fx["input_seed"])
    feats = torch.randn(2, 4096, 128, generator=g)
    slots = torch.randn(2, S, 128, generator=g)
    assert abs(float(feats.double().sum()) - fx["feats_checksum"]) < 1e-6 * max(1.0, abs(fx["feats_checksum"]))
    with torch.no_grad():
        o3, m3 = O.slot_attention(sd, feats, slots, 3)
        o1, m1 = O.slot_attention(sd, feats, slots, 1)
    _check("slots_step0", o3, fx["slots_step0"])
    _check("masks_step0", m3, fx["masks_step0"])
    _check("slots_step1", o1, fx["slots_step1"])
    _check("masks_step1", m1, fx["masks_step1"])


def test_sinusoid_pe_matches_reference_buffer():
    # the reference re-injects its freshly built `pos_emb.pe` on checkpoint load (setup_model.py:277-279);
    # the oracle's closed form must equal the buffer stored in the fixture run.
    pe = O.sinusoid_pe(8, 256)
    assert pe.shape == (8, 256)
    assert float(pe[0, 1]) == 1.0 and float(pe[0, 0]) == 0.0


def test_aux_predictors_oracle_matches_reference():
    """ rows a12 / a13: OCVPSeq + PredictorWrapper and ActionCondtionalOCVP against the reference-made fixture """
    aux = torch.load(os.path.join(GOLDEN_DIR, "aux_predictors.pt"), weights_only=False)
    g = torch.Generator().manual_seed(4242)
    sh = torch.randn(2, 14, 8, 128, generator=g)
    actions = torch.randn(2, 14, 4, generator=g)
    fx = aux["ocvp_seq"]
    assert abs(float(sh.double().sum()) - fx["sh_checksum"]) < 1e-6 * max(1.0, abs(fx["sh_checksum"]))
    sd = synth_state_dict(fx["manifest"], fx["weight_seed"])
    sd["predictor.pe.pe"] = O.sinusoid_pe(14, fx["cfg"]["token_dim"]).view(1, 14, 1, -1)
    pp = fx["prediction_params"]
    with torch.no_grad():
        step = lambda buf: O.ocvp_seq(sd, fx["cfg"], buf)
        _check("ocvp single", step(sh[:, :5]), fx["single"])
        _check("ocvp full", O.wrapper_rollout(step, sh, pp["num_context"], pp["num_preds"], pp["input_buffer_size"], False), fx["full"])
        _check("ocvp pred_only", O.wrapper_rollout(step, sh, pp["num_context"], pp["num_preds"], pp["input_buffer_size"], True), fx["pred_only"])
    fx = aux["action_cond"]
    sd = synth_state_dict(fx["manifest"], fx["weight_seed"])
    ctx = fx["cfg"]["DynamicsModel"]["model_params"].get("context_length", 8)
    E = fx["cfg"]["DynamicsModel"]["model_params"]["embed_dim"]
    sd["dynamics_model.pos_emb.pe"] = O.sinusoid_pe(ctx, E).view(1, ctx, 1, E)
    with torch.no_grad():
        _check("action_cond", O.action_conditional(sd, fx["cfg"], sh, actions, fx["num_seed"], fx["num_preds"]), fx["pred_slots"])
