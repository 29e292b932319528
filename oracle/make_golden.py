"""
Golden-vector generator  --  runs ONLY in the build container, where the
reference is mounted at /root/reference.  TEST INFRASTRUCTURE.

It imports the *unmodified* reference modules (`lib.setup_model.setup_model /
setup_predictor`), loads the seeded synthetic weights of `tests/synth.py` into
them, runs the seed+rollout path of `src/05_evaluate_PlaySlot.py:112-139`
(`savi(videos, decode=False)` -> `forward_multiple_samples(use_posterior=True)`
-> `savi.decode`) on seeded synthetic video and stores the outputs as small
fixtures under tests/golden/.  The fixtures carry the config dicts and the
state_dict manifests (key -> shape) so that nothing has to be read from
/root/reference at test time.

Usage:  python oracle/make_golden.py            (writes tests/golden/*.pt)
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


def _import_reference():
    os.chdir(REF)                                   # CONFIG.py:11 resolves paths from cwd
    sys.path.insert(0, os.path.join(REF, "src"))
    from lib.setup_model import setup_model, setup_predictor
    return setup_model, setup_predictor


CASES = {
    # name: (experiment, B, num_seed, num_preds, weight seed, sharpen, overrides)
    "blockpush":        ("BlockPush",       2, 6, 5, 11, 1.0, {}),
    "blockpush_sharp":  ("BlockPush",       1, 6, 3, 12, 6.0, {}),
    "buttonpress":      ("ButtonPress",     2, 4, 3, 13, 1.0, {}),
    "gridshapes5":      ("GridShapes5Objs", 2, 6, 4, 14, 1.0, {}),
    "sketchy_sharp":    ("Sketchy",         1, 6, 6, 15, 4.0, {}),
    # the benchmark length (BASELINE.json configs[2]: 6 seed + 15 predicted frames): errors compound over 15 roll-out steps
    "gridshapes5_bench": ("GridShapes5Objs", 2, 6, 15, 16, 1.0, {"decode_last": 1}),
    "blockpush_bench":   ("BlockPush",       2, 6, 15, 17, 1.0, {"decode_last": 1}),
}
DECODE_LAST = 2       # decode only the last 2 predicted frames of every video (fixture size)


def manifest_of(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def run_case(name, setup_model, setup_predictor):
    exp, B, ns, npred, wseed, sharpen, over = CASES[name]
    decode_last = over.get("decode_last", DECODE_LAST)
    j = json.load(open(os.path.join(REF, "experiments", exp, "predictors", "PlaySlot", "experiment_params.json")))
    scfg = copy.deepcopy(j["model"]["model_params"])
    pcfg = copy.deepcopy(j["predictor"]["predictor_params"])
    savi = setup_model(copy.deepcopy(j["model"])).eval()
    pred = setup_predictor(copy.deepcopy(j)).eval()
    sman, pman = manifest_of(savi), manifest_of(pred)

    ssd = synth_state_dict(sman, wseed, sharpen)
    psd = synth_state_dict(pman, wseed + 100)
    psd["dynamics_model.pos_emb.pe"] = pred.state_dict()["dynamics_model.pos_emb.pe"].clone()
    savi.load_state_dict(ssd, strict=True)
    pred.load_state_dict(psd, strict=True)

    T = ns + npred
    H, W = scfg["encoder"]["encoder_params"]["resolution"]
    S, D, A = scfg["num_slots"], scfg["slot_dim"], pcfg["action_dim"]
    video = synth_video(B, T, H, W, wseed)
    per_slot = pcfg["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
    noise_shape = (B, T - 1, S, A) if per_slot else (B, T - 1, A)
    noise = synth_noise(noise_shape, wseed)
    init_noise = synth_noise((B, S, D), wseed + 1) if scfg["initializer"] == "LearnedRandom" else None

    with torch.no_grad():
        if init_noise is not None:
            # LearnedRandom draws torch.randn on the module's device (initializers.py:123): inject
            real_randn = torch.randn
            torch.randn = lambda *a, **k: init_noise.clone()
            try:
                out = savi(video, num_imgs=T, decode=False)
            finally:
                torch.randn = real_randn
        else:
            out = savi(video, num_imgs=T, decode=False)
        sh = out["slot_history"]
        sa_masks = savi.slot_attention.get_attention_masks().clone()

        # InvDyn.sample draws torch.randn(mean.size()) on the CPU generator (LatentAction.py:73): inject
        real_randn = torch.randn
        torch.randn = lambda *a, **k: noise.clone()
        try:
            act = pred.latent_action.compute_actions(sh)
        finally:
            torch.randn = real_randn
        torch.randn = lambda *a, **k: noise.clone()
        try:
            pred_slots, _ = pred.forward_multiple_samples(sh, 1, ns, npred, use_posterior=True)
        finally:
            torch.randn = real_randn
        pred_slots = pred_slots[:, 0]
        dec_in = pred_slots[:, -decode_last:].reshape(B * decode_last, S, D)
        recon, (recons, masks) = savi.decode(dec_in)

        # margins of the VQ argmin (for the exact index comparison)
        cb = pred.latent_action.quantizer.get_codewords()
        z = act["sampled_latent_action"].reshape(-1, A)
        dist = (z.pow(2).sum(1, keepdim=True) - 2 * z @ cb.t() + cb.t().pow(2).sum(0, keepdim=True))
        top2 = dist.topk(2, dim=1, largest=False).values
        margin = (top2[:, 1] - top2[:, 0]).min().item()

    fx = {
        "case": name, "experiment": exp, "B": B, "num_seed": ns, "num_preds": npred,
        "weight_seed": wseed, "sharpen": sharpen, "decode_last": decode_last,
        "savi_cfg": scfg, "predictor_cfg": pcfg,
        "predictor_name": j["predictor"]["predictor_name"],
        "savi_manifest": sman, "predictor_manifest": pman,
        "video_checksum": [float(video.double().sum()), float(video[0, 0, 0, 0, :4].double().sum())],
        "noise_checksum": float(noise.double().sum()),
        "slot_history": sh.clone(), "sa_masks_last": sa_masks,
        "action_proto": act["action_proto"].clone(), "action_variability": act["action_variability"].clone(),
        "action_idxs": act["action_idxs"].clone(), "action_dist": act["action_dist"].clone(),
        "sampled_latent_action": act["sampled_latent_action"].clone(),
        "vq_margin_min": margin,
        "pred_slots": pred_slots.clone(),
        "recon": recon.clone(), "masks": masks.clone(), "recons_m0": recons[0].clone(),
    }
    return fx


def run_slot_attention_cases(setup_model):
    """ Stand-alone SlotAttention (config-2 shapes, reduced batch): 3 iterations, S in {4, 8}. """
    sys.path.insert(0, os.path.join(REF, "src"))
    from models.BlocksUtils.attention import SlotAttention
    res = {}
    for S, sharpen in ((8, 1.0), (4, 6.0)):
        sa = SlotAttention(dim_feats=128, dim_slots=128, num_slots=S, num_iters_first=3, num_iters=1, mlp_hidden=256).eval()
        man = {"slot_attention." + k: list(v.shape) for k, v in sa.state_dict().items()}
        sd = synth_state_dict(man, 21 + S, sharpen)
        sa.load_state_dict({k[len("slot_attention."):]: v for k, v in sd.items()}, strict=True)
        g = torch.Generator().manual_seed(100 + S)
        feats = torch.randn(2, 4096, 128, generator=g)
        slots = torch.randn(2, S, 128, generator=g)
        with torch.no_grad():
            o3 = sa(feats, slots, step=0); m3 = sa.get_attention_masks().clone()
            o1 = sa(feats, slots, step=1); m1 = sa.get_attention_masks().clone()
        res[f"S{S}"] = {"manifest": man, "weight_seed": 21 + S, "sharpen": sharpen, "input_seed": 100 + S,
                        "feats_checksum": float(feats.double().sum()),
                        "slots_step0": o3.clone(), "masks_step0": m3, "slots_step1": o1.clone(), "masks_step1": m1}
    return res


def run_aux_cases(setup_predictor):
    """ rows a12 / a13: OCVPSeq behind PredictorWrapper and ActionCondtionalOCVP, built by the reference factory """
    res = {}
    g = torch.Generator().manual_seed(4242)
    sh = torch.randn(2, 14, 8, 128, generator=g)
    actions = torch.randn(2, 14, 4, generator=g)
    # OCVP-Seq (experiments/BlockPush/predictors/OCVP) with a sliding buffer shorter than the roll-out
    j = json.load(open(os.path.join(REF, "experiments", "BlockPush", "predictors", "OCVP", "experiment_params.json")))
    j["prediction_params"]["input_buffer_size"] = 8
    pred = setup_predictor(copy.deepcopy(j)).eval()
    man = manifest_of(pred)
    sd = synth_state_dict(man, 31)
    sd["predictor.pe.pe"] = pred.state_dict()["predictor.pe.pe"].clone()
    pred.load_state_dict(sd, strict=True)
    with torch.no_grad():
        full, _ = pred(sh, get_pred_only=False)
        only, _ = pred(sh, get_pred_only=True)
        single = pred.predictor(sh[:, :5])
    res["ocvp_seq"] = {"manifest": man, "weight_seed": 31, "cfg": j["predictor"]["predictor_params"],
                       "prediction_params": j["prediction_params"], "input_seed": 4242,
                       "full": full.clone(), "pred_only": only.clone(), "single": single.clone(),
                       "sh_checksum": float(sh.double().sum())}
    # action-conditional OCVP (experiments/BlockPush/predictors/ActCondOCVP)
    j = json.load(open(os.path.join(REF, "experiments", "BlockPush", "predictors", "ActCondOCVP", "experiment_params.json")))
    pred = setup_predictor(copy.deepcopy(j)).eval()
    man = manifest_of(pred)
    sd = synth_state_dict(man, 32)
    sd["dynamics_model.pos_emb.pe"] = pred.state_dict()["dynamics_model.pos_emb.pe"].clone()
    pred.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out, _ = pred(sh, actions, num_seed=6, num_preds=8)
    res["action_cond"] = {"manifest": man, "weight_seed": 32, "cfg": j["predictor"]["predictor_params"], "input_seed": 4242,
                          "num_seed": 6, "num_preds": 8, "pred_slots": out.clone()}
    return res


class _RecordDraws:
    """ records every torch.randint / torch.rand result the reference draws while sampling prior actions, in call order """

    def __enter__(self):
        self.randint, self.rand = [], []
        self._ri, self._r = torch.randint, torch.rand

        def ri(*a, **k):
            t = self._ri(*a, **k)
            self.randint.append(t.clone().cpu())
            return t

        def r(*a, **k):
            t = self._r(*a, **k)
            self.rand.append(t.clone().cpu())
            return t
        torch.randint, torch.rand = ri, r
        return self

    def __exit__(self, *exc):
        torch.randint, torch.rand = self._ri, self._r
        return False


def run_prior_cases(setup_predictor):
    """ forward_multiple_samples with PRIOR samples (random codebook actions + random variability latents,
    LatentActionPredictors.py:148-157,383-396,511-532): the reference's draws are recorded so that the CUDA path can replay
    them; single-action model with variability (BlockPush) and per-slot model without (GridShapes5Objs) """
    res = {}
    for exp, wseed in (("BlockPush", 41), ("GridShapes5Objs", 42)):
        j = json.load(open(os.path.join(REF, "experiments", exp, "predictors", "PlaySlot", "experiment_params.json")))
        pred = setup_predictor(copy.deepcopy(j)).eval()
        man = manifest_of(pred)
        sd = synth_state_dict(man, wseed)
        sd["dynamics_model.pos_emb.pe"] = pred.state_dict()["dynamics_model.pos_emb.pe"].clone()
        pred.load_state_dict(sd, strict=True)
        S = j["model"]["model_params"]["num_slots"]
        A = j["predictor"]["predictor_params"]["action_dim"]
        B, ns, npred, nsamp = 2, 6, 4, 3
        g = torch.Generator().manual_seed(wseed + 5)
        sh = torch.randn(B, ns + npred, S, 128, generator=g)
        per_slot = j["predictor"]["predictor_params"]["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
        noise = synth_noise((B, ns + npred - 1, S, A) if per_slot else (B, ns + npred - 1, A), wseed)
        torch.manual_seed(wseed)
        real_randn = torch.randn
        torch.randn = lambda *a, **k: noise.clone()
        try:
            with torch.no_grad(), _RecordDraws() as rec:
                prior, prior_idx = pred.forward_multiple_samples(sh, nsamp, ns, npred, use_posterior=False)
                n_prior_only = (len(rec.randint), len(rec.rand))
                mixed, mixed_idx = pred.forward_multiple_samples(sh, nsamp, ns, npred, use_posterior=True)
        finally:
            torch.randn = real_randn
        res[exp] = {"manifest": man, "weight_seed": wseed, "cfg": j["predictor"]["predictor_params"],
                    "predictor_name": j["predictor"]["predictor_name"], "input_seed": wseed + 5,
                    "B": B, "num_seed": ns, "num_preds": npred, "num_samples": nsamp, "num_slots": S,
                    "randint_draws": rec.randint, "rand_draws": rec.rand, "draws_prior_only": n_prior_only,
                    "prior_pred_slots": prior.clone(), "prior_action_idx": prior_idx.clone(),
                    "mixed_pred_slots": mixed.clone(), "mixed_action_idx": mixed_idx.clone(),
                    "sh_checksum": float(sh.double().sum())}
    return res


def closed_loop_policy(slots_t, Wp):
    """ deterministic stand-in for the behaviour model of 11_evaluate_behavior_on_simulation.py:151 (models.Downstream is outside
    the hot path): latent action = tanh(mean over slots (or per slot) of slots_t @ Wp) """
    if Wp.dim() == 3:                                  # per-slot latent actions (SlotLatentPredictor)
        return torch.tanh(torch.einsum("bsd,sda->bsa", slots_t, Wp))
    return torch.tanh(slots_t.mean(dim=1) @ Wp)


def run_closed_loop_cases(setup_model, setup_predictor):
    """ the latent-imagination loop of 11_evaluate_behavior_on_simulation.py:128-183 with the reference modules: SAVi on the first
    frame (decode=True), then per step policy -> InvDyn.decompose_action_latent -> cOCVP.enforce_window -> forward_single ->
    savi.decode(pred_slots[:, 0]); 11 steps, so the 8-frame window slides """
    res = {}
    for exp, wseed in (("BlockPush", 51), ("GridShapes5Objs", 52)):
        j = json.load(open(os.path.join(REF, "experiments", exp, "predictors", "PlaySlot", "experiment_params.json")))
        scfg = copy.deepcopy(j["model"]["model_params"])
        pcfg = copy.deepcopy(j["predictor"]["predictor_params"])
        savi = setup_model(copy.deepcopy(j["model"])).eval()
        pred = setup_predictor(copy.deepcopy(j)).eval()
        sman, pman = manifest_of(savi), manifest_of(pred)
        ssd = synth_state_dict(sman, wseed)
        psd = synth_state_dict(pman, wseed + 100)
        psd["dynamics_model.pos_emb.pe"] = pred.state_dict()["dynamics_model.pos_emb.pe"].clone()
        savi.load_state_dict(ssd, strict=True)
        pred.load_state_dict(psd, strict=True)
        S, D, A = scfg["num_slots"], scfg["slot_dim"], pcfg["action_dim"]
        per_slot = pcfg["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
        B, steps = 2, 11
        g = torch.Generator().manual_seed(wseed + 7)
        init_img = torch.rand(B, 3, 64, 64, generator=g)
        Wp = torch.randn((S, D, A) if per_slot else (D, A), generator=g) / D ** 0.5
        init_noise = synth_noise((B, S, D), wseed + 1) if scfg["initializer"] == "LearnedRandom" else None
        InvDyn, cOCVP = pred.latent_action, pred.dynamics_model
        with torch.no_grad():
            real_randn = torch.randn
            if init_noise is not None:
                torch.randn = lambda *a, **k: init_noise.clone()
            try:
                out = savi(init_img.unsqueeze(1), num_imgs=1, decode=True)
            finally:
                torch.randn = real_randn
            cur_slots = out["slot_history"]
            all_pred_slots, protos_hist, vars_hist, idx_hist = [cur_slots], [], [], []
            imgs, masks = [out["recons_imgs"][:, 0]], [out["masks"][:, 0]]
            for i in range(steps):
                # the reference indexes the (window-cropped) slot tensor with the step counter (`cur_slots[:, i:i+1]`), which is the
                # newest frame until the window is full; the stand-in policy reads the newest frame throughout
                latent = closed_loop_policy(all_pred_slots[-1][:, 0] if all_pred_slots[-1].dim() == 4 else all_pred_slots[-1], Wp)
                proto, var = InvDyn.decompose_action_latent(latent)
                protos_hist.append(proto)
                vars_hist.append(var)
                cp, cv = torch.stack(protos_hist, dim=1), torch.stack(vars_hist, dim=1)
                cur_slots, cp, cv = cOCVP.enforce_window(slots=cur_slots, action_protos=cp, action_vars=cv)
                if len(cp.shape) == 3:
                    cp = cp.unsqueeze(2).expand(-1, -1, S, -1)
                    cv = cv.unsqueeze(2).expand(-1, -1, S, -1)
                ps = cOCVP.forward_single(slots=cur_slots, action_protos=cp, action_vars=cv)[:, -1:]
                all_pred_slots.append(ps)
                cur_slots = torch.cat(all_pred_slots, dim=1)
                img, (_, msk) = savi.decode(ps[:, 0])
                imgs.append(img)
                masks.append(msk)
        res[exp] = {"savi_cfg": scfg, "predictor_cfg": pcfg, "predictor_name": j["predictor"]["predictor_name"],
                    "savi_manifest": sman, "predictor_manifest": pman, "weight_seed": wseed, "input_seed": wseed + 7,
                    "B": B, "steps": steps, "per_slot": per_slot,
                    "pred_slots": torch.cat(all_pred_slots, dim=1).clone(), "protos": torch.stack(protos_hist, dim=1).clone(),
                    "vars": torch.stack(vars_hist, dim=1).clone(),
                    "imgs_last3": torch.stack(imgs[-3:], dim=1).clone(), "masks_last": masks[-1].clone(),
                    "init_recon": imgs[0].clone()}
    return res


def main():
    setup_model, setup_predictor = _import_reference()
    outdir = os.path.join(REPO, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    if os.environ.get("GOLDEN_ONLY_LOOP"):
        cl = run_closed_loop_cases(setup_model, setup_predictor)
        path = os.path.join(outdir, "closed_loop.pt")
        torch.save(cl, path)
        print(f"closed_loop: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
        return
    if os.environ.get("GOLDEN_ONLY_PRIOR"):
        pr = run_prior_cases(setup_predictor)
        path = os.path.join(outdir, "prior_samples.pt")
        torch.save(pr, path)
        print(f"prior_samples: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
        return
    only = os.environ.get("GOLDEN_ONLY_CASES")
    cases = [c for c in CASES if (not only or c in only.split(","))]
    for name in ([] if os.environ.get("GOLDEN_ONLY_AUX") else cases):
        fx = run_case(name, setup_model, setup_predictor)
        path = os.path.join(outdir, f"{name}.pt")
        torch.save(fx, path)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB), vq margin {fx['vq_margin_min']:.3e}, "
              f"mask max {float(fx['masks'].max()):.3f}, sa mask max {float(fx['sa_masks_last'].max()):.3f}")
    if only:
        return
    aux = run_aux_cases(setup_predictor)
    path = os.path.join(outdir, "aux_predictors.pt")
    torch.save(aux, path)
    print(f"aux_predictors: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    if os.environ.get("GOLDEN_ONLY_AUX"):
        return
    pr = run_prior_cases(setup_predictor)
    path = os.path.join(outdir, "prior_samples.pt")
    torch.save(pr, path)
    print(f"prior_samples: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    cl = run_closed_loop_cases(setup_model, setup_predictor)
    path = os.path.join(outdir, "closed_loop.pt")
    torch.save(cl, path)
    print(f"closed_loop: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")
    sa = run_slot_attention_cases(setup_model)
    path = os.path.join(outdir, "slot_attention.pt")
    torch.save(sa, path)
    print(f"slot_attention: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
