"""error of the fp16x3 tensor-core convolution and of the FFMA convolution against fp64, at several activation magnitudes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from playslot_b200 import ops
DEV = "cuda"
def g(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale
def err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm()), float((a - b).abs().max() / b.abs().max())
n, H, W = 4, 64, 64
for xs in [1.0, 1e-2, 1e-4, 30.0]:
    for relu_in in [False, True]:
        x = g(n, 64, H, W, seed=1, scale=xs)
        if relu_in: x = F.relu(x)
        w, b = g(64, 64, 5, 5, seed=2, scale=1600 ** -0.5), g(64, seed=3, scale=0.1 * xs)
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=2)
        ref32 = F.conv2d(x, w, b, padding=2)
        a0 = ops.conv_tc_workspace(n, H, W, DEV)
        ops.nchw_to_split(x.to(DEV), a0, 3)
        out = torch.empty(n, 64, H, W, device=DEV)
        ops.conv5x5_tc(a0, ops.pack_conv_weight_tc(w.to(DEV), 3), b.to(DEV), n, H, W, False, 3, out_nchw=out)
        y = ops.conv2d(x.to(DEV), ops.pack_conv_weight(w.to(DEV)), b.to(DEV), 5, False)
        print("xscale %g relu_in %d: tc %.2e/%.2e  ffma %.2e/%.2e  torch-cpu-fp32 %.2e/%.2e" % ((xs, relu_in) + err(out, ref) + err(y, ref) + err(ref32, ref)))
