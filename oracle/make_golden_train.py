"""
Golden vectors of ONE training step (forward + backward) of the joint InvDyn + cOCVP model -- runs ONLY in the build
container, where the reference is mounted at /root/reference.  TEST INFRASTRUCTURE.

Imports the unmodified reference (`lib.setup_model`), loads the seeded synthetic weights of tests/synth.py, and runs
what `Trainer.forward_loss_metric` runs (src/04_train_PlaySlot.py:32-122) on a small synthetic batch in TRAIN mode:
frozen SAVi without gradients, predictor forward (EMA codebook update included), decoder with gradients, the
PredImgMSE + PredSlotMSE + VQ losses (src/lib/loss.py:201-238,319-340), `loss.backward()`.
Stored under tests/golden/train_<case>.pt: the loss terms, pred_slots, the updated EMA buffers, the gradient norm of
every predictor parameter and the full gradient of every parameter with at most 4096 elements.

Usage:  python oracle/make_golden_train.py
"""
import copy
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PLAYSLOT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.dont_write_bytecode = True
from synth import synth_state_dict, synth_video, synth_noise  # noqa: E402

CASES = {
    # name: (experiment, B, num_seed, num_preds, weight seed)
    "buttonpress": ("ButtonPress", 2, 3, 3, 21),
    "gridshapes5": ("GridShapes5Objs", 2, 2, 3, 22),
}


def main():
    os.chdir(REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    from lib.setup_model import setup_model, setup_predictor
    from lib.loss import PredImgMSELoss, PredSlotMSELoss, VQLoss
    for name, (exp, B, ns, npred, wseed) in CASES.items():
        j = json.load(open(os.path.join(REF, "experiments", exp, "predictors", "PlaySlot", "experiment_params.json")))
        scfg = copy.deepcopy(j["model"]["model_params"])
        pcfg = copy.deepcopy(j["predictor"]["predictor_params"])
        savi = setup_model(copy.deepcopy(j["model"])).eval()
        pred = setup_predictor(copy.deepcopy(j))
        for p_ in savi.parameters():
            p_.requires_grad_(False)
        sman = {k: list(v.shape) for k, v in savi.state_dict().items()}
        pman = {k: list(v.shape) for k, v in pred.state_dict().items()}
        ssd = synth_state_dict(sman, wseed)
        psd = synth_state_dict(pman, wseed + 100)
        psd["dynamics_model.pos_emb.pe"] = pred.state_dict()["dynamics_model.pos_emb.pe"].clone()
        savi.load_state_dict(ssd, strict=True)
        pred.load_state_dict(psd, strict=True)
        pred.train()

        T = ns + npred
        H, W = scfg["encoder"]["encoder_params"]["resolution"]
        S, D, A = scfg["num_slots"], scfg["slot_dim"], pcfg["action_dim"]
        video = synth_video(B, T, H, W, wseed)
        per_slot = pcfg["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
        noise = synth_noise((B, T - 1, S, A) if per_slot else (B, T - 1, A), wseed)
        init_noise = synth_noise((B, S, D), wseed + 1) if scfg["initializer"] == "LearnedRandom" else None

        real_randn = torch.randn
        with torch.no_grad():
            if init_noise is not None:
                torch.randn = lambda *a, **k: init_noise.clone()
            try:
                sh = savi(video, num_imgs=T, decode=False)["slot_history"]
            finally:
                torch.randn = real_randn
        torch.randn = lambda *a, **k: noise.clone()
        try:
            pred_slots, out = pred(imgs=video, slots=sh, actions=None, num_seed=ns, num_preds=npred)
        finally:
            torch.randn = real_randn
        n = pred_slots.shape[1]
        pred_imgs, _ = savi.decode(pred_slots.reshape(B * n, S, D))
        pred_imgs = pred_imgs.view(B, n, *pred_imgs.shape[1:])
        l_img = PredImgMSELoss()(pred_imgs=pred_imgs, target_imgs=video[:, 1:T])
        l_slot = PredSlotMSELoss()(preds=pred_slots, targets=sh[:, 1:T])
        vq = VQLoss(beta=0.25)(**out["vq_losses"])
        loss = l_img + l_slot + vq["vq_quant_loss"] + vq["vq_commit_loss"]
        pred_slots.retain_grad()
        loss.backward()

        grads_full, grad_norms = {}, {}
        for k, p_ in pred.named_parameters():
            if p_.grad is None:
                grad_norms[k] = None
                continue
            grad_norms[k] = float(p_.grad.double().norm())
            if p_.numel() <= 4096:
                grads_full[k] = p_.grad.clone()
        fx = {
            "case": name, "experiment": exp, "B": B, "num_seed": ns, "num_preds": npred, "weight_seed": wseed,
            "savi_cfg": scfg, "predictor_cfg": pcfg, "predictor_name": j["predictor"]["predictor_name"],
            "savi_manifest": sman, "predictor_manifest": pman,
            "slot_history": sh.clone(), "pred_slots": pred_slots.detach().clone(),
            "d_pred_slots": pred_slots.grad.clone(),
            "loss": float(loss), "pred_img_mse": float(l_img), "pred_slot_mse": float(l_slot),
            "quant": float(vq["vq_quant_loss"]), "commit_beta": float(vq["vq_commit_loss"]),
            "grad_norms": grad_norms, "grads_full": grads_full,
            "buffers_after": {k: v.clone() for k, v in pred.state_dict().items() if "quantizer" in k},
            "action_idxs": out["action_idxs"].clone(),
        }
        path = os.path.join(REPO, "tests", "golden", f"train_{name}.pt")
        torch.save(fx, path)
        nz = sum(1 for v in grad_norms.values() if v)
        print(f"{name}: loss {float(loss):.6f} (img {float(l_img):.5f} slot {float(l_slot):.5f} quant {float(vq['vq_quant_loss']):.5f}) "
              f"{nz}/{len(grad_norms)} params with gradient, {len(grads_full)} stored in full, {os.path.getsize(path) / 1e6:.2f} MB")


SAVI_CASES = {
    # name: (experiment, B, T, weight seed)   -- SAVi training step of src/02_train_savi.py:21-70
    "savi_blockpush":   ("BlockPush",       2, 3, 61),
    "savi_gridshapes5": ("GridShapes5Objs", 2, 2, 62),
}


def main_savi():
    """ reference SAVi training step: model(videos) with decode, MSE(recons_imgs, videos), backward; stores the loss, the gradient
    norm of every parameter and the full gradient of the small ones """
    os.chdir(REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    from lib.setup_model import setup_model
    for name, (exp, B, T, wseed) in SAVI_CASES.items():
        j = json.load(open(os.path.join(REF, "experiments", exp, "predictors", "PlaySlot", "experiment_params.json")))
        scfg = copy.deepcopy(j["model"]["model_params"])
        savi = setup_model(copy.deepcopy(j["model"])).train()
        sman = {k: list(v.shape) for k, v in savi.state_dict().items()}
        savi.load_state_dict(synth_state_dict(sman, wseed), strict=True)
        H, W = scfg["encoder"]["encoder_params"]["resolution"]
        S, D = scfg["num_slots"], scfg["slot_dim"]
        video = synth_video(B, T, H, W, wseed)
        init_noise = synth_noise((B, S, D), wseed + 1) if scfg["initializer"] == "LearnedRandom" else None
        real_randn = torch.randn
        if init_noise is not None:
            torch.randn = lambda *a, **k: init_noise.clone()
        try:
            out = savi(video, num_imgs=T)
        finally:
            torch.randn = real_randn
        loss = torch.nn.functional.mse_loss(out["recons_imgs"], video.clamp(0, 1))
        loss.backward()
        grad_norms, grads_full = {}, {}
        for k, p_ in savi.named_parameters():
            grad_norms[k] = None if p_.grad is None else float(p_.grad.double().norm())
            if p_.grad is not None and p_.numel() <= 8192:
                grads_full[k] = p_.grad.clone()
        fx = {"case": name, "experiment": exp, "B": B, "T": T, "weight_seed": wseed, "savi_cfg": scfg, "savi_manifest": sman,
              "loss": float(loss), "slot_history": out["slot_history"].detach().clone(),
              "recon_last": out["recons_imgs"][:, -1].detach().clone(), "masks_last": out["masks"][:, -1].detach().clone(),
              "grad_norms": grad_norms, "grads_full": grads_full}
        path = os.path.join(REPO, "tests", "golden", f"train_{name}.pt")
        torch.save(fx, path)
        print(f"{name}: loss {float(loss):.6f}, {sum(1 for v in grad_norms.values() if v)}/{len(grad_norms)} params with gradient, "
              f"{len(grads_full)} stored in full, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    if "--savi" in sys.argv:
        main_savi()
    else:
        main()
