import os, sys
sys.path.insert(0, "/root/repo")
import torch
from playslot_b200.models.BlocksUtils.attention import SlotAttention
m = SlotAttention(128, 128, 8, 3, 1, 256).eval().cuda()
SlotAttention.TOKENS_PER_CTA = 1024
feats = torch.randn(64, 4096, 128, device="cuda"); slots = torch.randn(64, 8, 128, device="cuda")
with torch.no_grad():
    for _ in range(2):
        m(feats, slots, step=1)
    torch.cuda.synchronize()
    os.environ["PLAYSLOT_SA_TRACE"] = "1"
    m(feats, slots, step=1)
