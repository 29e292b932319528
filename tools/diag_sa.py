""" development tool: SlotAttention tc vs fp32 kernels vs the fp64 oracle on the features of a golden case """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import load_case
from build_util import make_savi
from synth import rel_l2, max_rel
from playslot_b200 import ops
from oracle import playslot_oracle as O

case = sys.argv[1] if len(sys.argv) > 1 else "sketchy_sharp"
fx, ssd, psd, video, noise, init_noise = load_case(case)
scfg = fx["savi_cfg"]
savi = make_savi(scfg, ssd, "cuda")
savi.conv_mode = "fp32"; ops.set_linear_mode("fp32")
with torch.no_grad():
    feats = savi.encode(video[:, 0].cuda())
    slots0 = savi.initializer(batch_size=video.shape[0]).detach().float().contiguous()
    sd = {k: v for k, v in ssd.items() if k.startswith("slot_attention.")}
    sd64 = {k: v.double() for k, v in sd.items()}
    for iters in (1, 3):
        ro64, rm64 = O.slot_attention(sd64, feats.cpu().double(), slots0.cpu().double(), iters)
        ro32, rm32 = O.slot_attention(sd, feats.cpu(), slots0.cpu(), iters)
        print(f"iters={iters}: oracle fp32 vs fp64: slots {max_rel(ro32, ro64.float()):.2e} masks {max_rel(rm32, rm64.float()):.2e}")
        sa = savi.slot_attention
        sa.num_iters_first = iters
        for mode in ("fp32", "tc"):
            sa.mode = mode
            o = sa(feats, slots0, step=0)
            mk = sa.get_attention_masks()
            print(f"   {mode:4s} vs fp64: slots max_rel {max_rel(o.cpu(), ro64.float()):.2e} rel_l2 {rel_l2(o.cpu(), ro64.float()):.2e} | "
                  f"masks max_rel {max_rel(mk.cpu(), rm64.float()):.2e} rel_l2 {rel_l2(mk.cpu(), rm64.float()):.2e}")
    # logits scale
    w = sa._weights()
    x = feats.double()
    xh = torch.nn.functional.layer_norm(x, (128,), sd64["slot_attention.norm_input.weight"].cuda(), sd64["slot_attention.norm_input.bias"].cuda(), 1e-5)
    k = xh @ sd64["slot_attention.to_k.weight"].cuda().T
    q = torch.nn.functional.layer_norm(slots0.double(), (128,), sd64["slot_attention.norm_slot.weight"].cuda(), sd64["slot_attention.norm_slot.bias"].cuda(), 1e-5) @ sd64["slot_attention.to_q.weight"].cuda().T
    dots = torch.einsum("bsd,bnd->bsn", q, k) * 128 ** -0.5
    print("logits: max|.|", float(dots.abs().max()), "std", float(dots.std()), " feats std", float(feats.std()), "max", float(feats.abs().max()))
