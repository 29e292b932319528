""" development tool: error of slot_history / pred_slots / decode against the reference goldens for each
    combination of arithmetic modes (which tensor-core path contributes how much) """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import load_case
from build_util import make_savi, FixedRandn
from synth import rel_l2, max_rel
from playslot_b200 import ops

case = sys.argv[1] if len(sys.argv) > 1 else "blockpush_sharp"
fx, ssd, psd, video, noise, init_noise = load_case(case)
scfg = fx["savi_cfg"]
ns, npred = fx["num_seed"], fx["num_preds"]
for conv in ("fp32", "fp16x3"):
    for lin in ("fp32", "fp16x3"):
        for sa in ("fp32", "tc"):
            savi = make_savi(scfg, ssd, "cuda")
            savi.conv_mode = conv
            ops.set_linear_mode(lin)
            savi.slot_attention.mode = sa
            with torch.no_grad():
                if init_noise is not None:
                    with FixedRandn(init_noise):
                        out = savi(video.cuda(), num_imgs=ns + npred, decode=False)
                else:
                    out = savi(video.cuda(), num_imgs=ns + npred, decode=False)
            sh = out["slot_history"].cpu()
            ref = fx["slot_history"]
            per_t = [max_rel(sh[:, t], ref[:, t]) for t in (0, 1, 5, 10, ref.shape[1] - 1)]
            print(f"{case}: conv={conv:6s} linear={lin:6s} sa={sa:4s}: slot_history rel_l2={rel_l2(sh, ref):.2e} max_rel={max_rel(sh, ref):.2e} "
                  f"| max_rel at t=0,1,5,10,last: " + " ".join(f"{e:.1e}" for e in per_t))
