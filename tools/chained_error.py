"""
Measured error of the CHAINED path (our slots -> our InvDyn / roll-out -> our decoder) against the reference-made golden
fixtures, per case, beside the stage-wise numbers (every stage fed with the reference's tensors).  Run on the GPU box:

    python tools/chained_error.py > gpurun_out/chained_error.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_util import PIPELINE_CASES, load_case              # noqa: E402
from build_util import make_savi, make_predictor, FixedRandn   # noqa: E402
from synth import rel_l2, max_rel                              # noqa: E402


def err(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return {"rel_l2": rel_l2(a, b), "max_rel": max_rel(a, b)}


def main():
    dev = "cuda"
    out = {}
    for case in PIPELINE_CASES:
        fx, ssd, psd, video, noise, init_noise = load_case(case)
        scfg, pcfg = fx["savi_cfg"], fx["predictor_cfg"]
        B, ns, npred = fx["B"], fx["num_seed"], fx["num_preds"]
        S, D = scfg["num_slots"], scfg["slot_dim"]
        savi = make_savi(scfg, ssd, dev)
        pred = make_predictor(fx["predictor_name"], pcfg, psd, dev)
        k = fx["decode_last"]
        with torch.no_grad():
            if init_noise is not None:
                with FixedRandn(init_noise):
                    sh = savi(video.to(dev), num_imgs=ns + npred, decode=False)["slot_history"]
            else:
                sh = savi(video.to(dev), num_imgs=ns + npred, decode=False)["slot_history"]
            with FixedRandn(noise):
                ps_stage, _ = pred.forward_multiple_samples(fx["slot_history"].to(dev), 1, ns, npred, use_posterior=True)
            with FixedRandn(noise):
                act = pred.latent_action.compute_actions(sh)
            with FixedRandn(noise):
                ps_chain, _ = pred.forward_multiple_samples(sh, 1, ns, npred, use_posterior=True)
            recon_c, (_, masks_c) = savi.decode(ps_chain[:, 0, -k:].reshape(B * k, S, D))
            recon_s, (_, masks_s) = savi.decode(fx["pred_slots"][:, -k:].reshape(B * k, S, D).to(dev))
        out[case] = {
            "frames": f"{ns}+{npred}", "B": B,
            "slot_history": err(sh, fx["slot_history"]),
            "vq_indices_equal_chained": bool(torch.equal(act["action_idxs"].cpu().reshape(-1), fx["action_idxs"].reshape(-1))),
            "vq_margin_min": fx["vq_margin_min"],
            "pred_slots_stagewise": err(ps_stage[:, 0], fx["pred_slots"]),
            "pred_slots_chained": err(ps_chain[:, 0], fx["pred_slots"]),
            "pred_slots_chained_last_step": err(ps_chain[:, 0, -1], fx["pred_slots"][:, -1]),
            "recon_stagewise": err(recon_s, fx["recon"]), "masks_stagewise": err(masks_s, fx["masks"]),
            "recon_chained": err(recon_c, fx["recon"]), "masks_chained": err(masks_c, fx["masks"]),
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
