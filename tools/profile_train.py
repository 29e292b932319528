""" development tool: per-kernel device time of one training step (CUDA events around every launch of the C-ABI library) """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import configs, ops, train

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = configs.experiment("ButtonPress")
torch.manual_seed(0)
savi, pred = configs.build_models(cfg, "cuda")
for p in savi.parameters():
    p.requires_grad_(False)
opt = train.make_optimizer(pred)
video = torch.rand(B, 14, 3, 64, 64, device="cuda")
for _ in range(3):
    train.train_step(savi, pred, opt, video, 6, 8)
torch.cuda.synchronize()
ops.profile_start()
train.train_step(savi, pred, opt, video, 6, 8)
prof = ops.profile_stop()
tot = sum(v["ms"] for v in prof.values())
print(f"B={B}: sum of library kernel times {tot:.1f} ms over {sum(v['calls'] for v in prof.values())} launches")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:18]:
    print(f"  {k:26s} {v['ms']:8.2f} ms  {v['calls']:5d} calls")
