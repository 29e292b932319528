""" development tool: fused encoder MLP kernel vs the unfused chain (2x linear_tc + sa_prepare), same process, median of reps """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops

n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 21 * 4096
K1 = int(sys.argv[2]) if len(sys.argv) > 2 else 64
x = torch.randn(n_tok, 64, device="cuda")
gam, bet = torch.ones(K1, device="cuda"), torch.zeros(K1, device="cuda")
w1, b1 = torch.randn(128, K1, device="cuda") * K1 ** -0.5, torch.zeros(128, device="cuda")
w2, b2 = torch.randn(128, 128, device="cuda") * 0.09, torch.zeros(128, device="cuda")
xs = x[:, :K1].contiguous() if K1 < 64 else x


def timed(fn, reps=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def unfused():
    h = ops.linear(xs, w1, b1, ln=(gam, bet, 1e-5), act=ops.ACT_RELU)
    return ops.sa_prepare(ops.linear(h, w2, b2), 1e-5)


f = timed(lambda: ops.enc_mlp_planes(x, K1, (gam, bet, 1e-5), (w1, b1), (w2, b2), 1e-5))
u = timed(unfused)
gb = n_tok * (K1 * 4 + 512) / 1e9
print(f"n_tok {n_tok} K1 {K1}: fused min {f[0]:.3f} med {f[1]:.3f} ms ({gb / f[1] * 1e3:.0f} GB/s algorithmic) | unfused min {u[0]:.3f} med {u[1]:.3f} ms | ratio {u[1] / f[1]:.2f}x")
