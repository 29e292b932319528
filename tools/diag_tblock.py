""" development tool: ps_tblock_stack against the fp64 oracle over a sweep of (n_seq, L), repeated to expose races """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from playslot_b200 import ops
from playslot_b200.models.BlocksUtils.attention import TransformerDecoder
from oracle import playslot_oracle as O

torch.manual_seed(0)
n_layers, H, dh, spf = 8, 8, 64, 6
blocks = [TransformerDecoder(embed_dim=128, head_dim=dh, num_heads=H, mlp_size=512, project_out=True) for _ in range(n_layers)]
for b in blocks:
    for name, prm in b.named_parameters():
        with torch.no_grad():
            if prm.dim() == 1:
                prm.copy_(torch.randn_like(prm) * 0.1 + (1.0 if "ln_" in name and name.endswith("weight") else 0.0))
            else:
                prm.copy_(torch.randn_like(prm) * prm.shape[1] ** -0.5)
sds = [{"blk." + k: v.detach().double() for k, v in b.state_dict().items()} for b in blocks]
blocks = [b.cuda().eval() for b in blocks]
ops.set_linear_mode("fp16x3")
pack = ops.pack_tblock_stack(blocks)
for L in (36, 42, 48, 6, 12):
    for n_seq in (1, 2, 3, 8, 13, 64):
        g = torch.Generator().manual_seed(L * 100 + n_seq)
        x = torch.randn(n_seq, L, 128, generator=g)
        fr = torch.arange(L) // spf
        mask = (fr[None, :] <= fr[:, None]).to(torch.float64)
        ref = x.double()
        for sd in sds:
            ref = O.transformer_block(sd, "blk", ref, H, dh, mask)
        errs = []
        outs = []
        for rep in range(6):
            with torch.no_grad():
                y = ops.tblock_stack(x.cuda().view(n_seq * L, 128), pack, n_seq, L, spf).view(n_seq, L, 128).cpu()
            outs.append(y)
            errs.append(float((y.double() - ref).abs().max() / ref.abs().max()))
        same = all(torch.equal(outs[0], o) for o in outs[1:])
        print(f"L={L:3d} n_seq={n_seq:3d}  max_rel err over 6 reps: min {min(errs):.2e} max {max(errs):.2e}  bit-identical repeats: {same}", flush=True)
