""" development diagnostic: decode / encode error of each conv mode against the golden fixtures and an fp64 oracle """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import PIPELINE_CASES, load_case
from build_util import make_savi
from synth import rel_l2, max_rel
from oracle import playslot_oracle as O

for case in PIPELINE_CASES:
    fx, ssd, psd, video, noise, init_noise = load_case(case)
    scfg = fx["savi_cfg"]
    B, k = fx["B"], fx["decode_last"]
    S, D = scfg["num_slots"], scfg["slot_dim"]
    slots = fx["pred_slots"][:, -k:].reshape(B * k, S, D)
    ssd64 = {kk: v.double() for kk, v in ssd.items()}
    with torch.no_grad():
        r64 = O.decode(ssd64, scfg, slots.double())
        frames = video[:, :2].reshape(-1, *video.shape[2:])
        e64 = O.encode(ssd64, frames.double())
    print(f"== {case}: golden(fp32 ref) vs fp64: recon {rel_l2(fx['recon'], r64[0]):.2e}/{max_rel(fx['recon'], r64[0]):.2e} "
          f"masks {rel_l2(fx['masks'], r64[2]):.2e}/{max_rel(fx['masks'], r64[2]):.2e}")
    for mode in ("fp32", "fp16x3", "fp16", "bf16"):
        savi = make_savi(scfg, ssd, "cuda"); savi.conv_mode = mode
        with torch.no_grad():
            recon, (recons, masks) = savi.decode(slots.cuda())
            enc = savi.encode(frames.cuda())
        print(f"   {mode:7s} vs fp64: recon {rel_l2(recon.cpu(), r64[0]):.2e}/{max_rel(recon.cpu(), r64[0]):.2e}  "
              f"masks {rel_l2(masks.cpu(), r64[2]):.2e}/{max_rel(masks.cpu(), r64[2]):.2e}  "
              f"recons {rel_l2(recons.cpu(), r64[1]):.2e}/{max_rel(recons.cpu(), r64[1]):.2e}  "
              f"encode {rel_l2(enc.cpu(), e64):.2e}/{max_rel(enc.cpu(), e64):.2e}")
