import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops
dev = "cuda"; M = 3072
torch.manual_seed(0)
for (N, K, ln, act) in ((1536, 128, True, 0), (128, 512, False, 0), (512, 128, True, 2)):
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5; bias = torch.randn(N, device=dev)
    lnp = (torch.ones(K, device=dev), torch.zeros(K, device=dev), 1e-6) if ln else None
    for _ in range(3):
        ops.linear(x, w, bias, ln=lnp, act=act)
qkv = torch.randn(M, 1536, device=dev)
for _ in range(3):
    ops.mha(qkv, 64, 48, 8, 64, 6)
torch.cuda.synchronize()
print("done")
