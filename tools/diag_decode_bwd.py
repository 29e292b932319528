""" development tool: stage-by-stage check of the decoder data gradient against torch autograd (fp64, on the GPU) """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from build_util import make_savi
from synth import synth_state_dict, rel_l2, max_rel
from playslot_b200 import ops, configs

cfg = configs.experiment("ButtonPress")["model"]["model_params"]
import copy
from playslot_b200.models.SAVi import SAVi
m = SAVi(**copy.deepcopy(cfg))
man = {k: list(v.shape) for k, v in m.state_dict().items()}
ssd = synth_state_dict(man, 21)
savi = make_savi(cfg, ssd, "cuda")
S, D, H, W = cfg["num_slots"], 128, 64, 64
g = lambda *s, seed=0: torch.randn(*s, generator=torch.Generator().manual_seed(seed))
slots = g(3, S, D, seed=11).cuda()
w_recon = g(3, 3, 64, 64, seed=12).cuda()

def unpack(act, n):
    full = act.view(torch.float16).float().view(2, n, H + 4, W + 4, 64)
    v = full[0] + full[1] / 2048.0
    return v[:, 2:H + 2, 2:W + 2].permute(0, 3, 1, 2).contiguous()

# reference in fp64 torch
sd64 = {k: v.double().cuda() for k, v in ssd.items()}
sl = slots.double().clone().requires_grad_(True)
pos = savi.decoder_pos_embedding.table(channels_last=False).double()
z = sl.reshape(-1, D, 1, 1).expand(-1, D, H, W) + pos[None]
hs = []
h = z
for i in (0, 2, 4, 6):
    h = F.relu(F.conv2d(h, sd64[f"decoder.decoder.{i}.block.0.weight"], sd64[f"decoder.decoder.{i}.block.0.bias"], padding=2))
    h.retain_grad(); hs.append(h)
o = F.conv2d(h, sd64["decoder.decoder.7.weight"], sd64["decoder.decoder.7.bias"], padding=1).view(3, S, 4, H, W)
o.retain_grad()
rgb, al = o[:, :, :3], o[:, :, 3:]
mk = torch.softmax(al, dim=1)
recon = (rgb * mk).sum(1)
(recon * w_recon.double()).sum().backward()

# ours, stage by stage
with torch.no_grad():
    r, rs, ms, saved = savi._decode_impl(slots, keep=True)
    m0, m1, acts, h4 = saved[0]
    n_img = 3 * S
    print("fwd recon", rel_l2(r.cpu(), recon.float().cpu()))
    for i in range(3):
        print(f"fwd h{i+1}", rel_l2(unpack(acts[i], n_img).cpu(), hs[i].float().cpu()))
    print("fwd h4", rel_l2(h4.cpu(), hs[3].float().cpu()))
    g_o = ops.dec_composite_bwd(rs, ms, w_recon.contiguous(), None, None)
    print("g_o", rel_l2(g_o.cpu(), o.grad.float().view(n_img, 4, H, W).cpu()), max_rel(g_o.cpu(), o.grad.float().view(n_img, 4, H, W).cpu()))
    blocks = savi.decoder.conv_blocks()
    wfin = savi._final_packed(savi.decoder.final_conv())
    ga = ops.conv_tc_workspace(n_img, H, W, "cuda"); gb = ops.conv_tc_workspace(n_img, H, W, "cuda")
    ops.dec_final_dgrad(g_o, wfin, h4, ga, 3)
    ref4 = (hs[3].grad * (hs[3] > 0)).float()
    print("g_pre4", rel_l2(unpack(ga, n_img).cpu(), ref4.cpu()), max_rel(unpack(ga, n_img).cpu(), ref4.cpu()))
    ops.conv5x5_tc_dgrad(ga, savi._tc_weight_flipped(blocks[3]), acts[2], n_img, H, W, 3, act_out=gb)
    ref3 = (hs[2].grad * (hs[2] > 0)).float()
    print("g_pre3", rel_l2(unpack(gb, n_img).cpu(), ref3.cpu()), max_rel(unpack(gb, n_img).cpu(), ref3.cpu()))
    ops.conv5x5_tc_dgrad(gb, savi._tc_weight_flipped(blocks[2]), acts[1], n_img, H, W, 3, act_out=ga)
    ref2 = (hs[1].grad * (hs[1] > 0)).float()
    print("g_pre2", rel_l2(unpack(ga, n_img).cpu(), ref2.cpu()), max_rel(unpack(ga, n_img).cpu(), ref2.cpu()))
    tokens = torch.empty(n_img, H * W, 64, device="cuda")
    ops.conv5x5_tc_dgrad(ga, savi._tc_weight_flipped(blocks[1]), acts[0], n_img, H, W, 3, out_tokens=tokens)
    ref1 = (hs[0].grad * (hs[0] > 0)).float().permute(0, 2, 3, 1).reshape(n_img, H * W, 64)
    print("g_pre1", rel_l2(tokens.cpu(), ref1.cpu()), max_rel(tokens.cpu(), ref1.cpu()))
    wsum, _, k, _ = savi._decoder_constants()
    dG = ops.dec_expand_bwd(tokens, n_img, H, W, k)
    dsl = ops.linear(dG, wsum.t().contiguous(), mode="fp32").view(3, S, D)
    print("d_slots", rel_l2(dsl.cpu(), sl.grad.float().cpu()), max_rel(dsl.cpu(), sl.grad.float().cpu()))
    # reference dG from ref1
    dsl_ref_from_ref1 = ops.linear(ops.dec_expand_bwd(ref1.contiguous(), n_img, H, W, k), wsum.t().contiguous(), mode="fp32").view(3, S, D)
    print("d_slots from reference g_pre1", rel_l2(dsl_ref_from_ref1.cpu(), sl.grad.float().cpu()))

    # ---- stage-wise references computed from OUR inputs and OUR relu masks (isolates each kernel)
    Wt = lambda i: sd64[f"decoder.decoder.{i}.block.0.weight"]
    def stage(g_in_ours, w, mask_ours):
        return (F.conv_transpose2d(g_in_ours.double(), w, padding=2) * (mask_ours > 0)).float()
    ga2 = ops.conv_tc_workspace(n_img, H, W, "cuda"); gb2 = ops.conv_tc_workspace(n_img, H, W, "cuda")
    ops.dec_final_dgrad(g_o, wfin, h4, ga2, 3)
    gp4 = unpack(ga2, n_img)
    ref = (F.conv_transpose2d(g_o.double(), sd64["decoder.decoder.7.weight"], padding=1) * (h4 > 0)).float()
    print("stage final_dgrad", rel_l2(gp4.cpu(), ref.cpu()), max_rel(gp4.cpu(), ref.cpu()))
    ops.conv5x5_tc_dgrad(ga2, savi._tc_weight_flipped(blocks[3]), acts[2], n_img, H, W, 3, act_out=gb2)
    gp3 = unpack(gb2, n_img); ref = stage(gp4, Wt(6), unpack(acts[2], n_img))
    print("stage conv4 dgrad", rel_l2(gp3.cpu(), ref.cpu()), max_rel(gp3.cpu(), ref.cpu()))
    ops.conv5x5_tc_dgrad(gb2, savi._tc_weight_flipped(blocks[2]), acts[1], n_img, H, W, 3, act_out=ga2)
    gp2 = unpack(ga2, n_img); ref = stage(gp3, Wt(4), unpack(acts[1], n_img))
    print("stage conv3 dgrad", rel_l2(gp2.cpu(), ref.cpu()), max_rel(gp2.cpu(), ref.cpu()))
    ops.conv5x5_tc_dgrad(ga2, savi._tc_weight_flipped(blocks[1]), acts[0], n_img, H, W, 3, out_tokens=tokens)
    ref = stage(gp2, Wt(2), unpack(acts[0], n_img)).permute(0, 2, 3, 1).reshape(n_img, H * W, 64)
    print("stage conv2 dgrad", rel_l2(tokens.cpu(), ref.cpu()), max_rel(tokens.cpu(), ref.cpu()))
