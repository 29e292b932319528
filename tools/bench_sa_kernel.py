""" development tool: back-to-back launch timing of the SlotAttention streaming kernels (one event pair around 20 launches),
    B=64, N=4096, S=8; inputs (134 MB) exceed the 126 MB L2 """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops
from playslot_b200.models.BlocksUtils.attention import SlotAttention

B, N, S = 64, 4096, 8
torch.manual_seed(0)
m = SlotAttention(128, 128, S, 3, 1, 256).eval().cuda()
feats = torch.randn(B, N, 128, device="cuda"); slots = torch.randn(B, S, 128, device="cuda")
w = m._weights()
A = torch.empty(B, S, 128, device="cuda"); a0 = torch.empty(B, S, device="cuda"); a1 = torch.empty(B, S, device="cuda")
ops.sa_update2(None, 1, slots, None, w, A, a0, 256, 1e-5, False, True)
planes = ops.sa_prepare(feats, 1e-5)
masks = torch.empty(B, S, N, device="cuda")
by = B * N * 128 * 4


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


CH2 = ops.sa_tc2_chunks(B, N)
partial = torch.empty(B, CH2, ops.sa_partial_floats(S), device="cuda")
t2 = timeit(lambda: ops.sa_main_tc2(planes, 0, B, N, A, a0, partial, None, CH2, 1e-8))
t2m = timeit(lambda: ops.sa_main_tc2(planes, 0, B, N, A, a0, partial, masks, CH2, 1e-8))
print(f"tc2 (balanced tile partition, {CH2} partial slots per frame): {t2:6.1f} us ({by / t2 / 1e3:6.0f} GB/s)  +masks {t2m:6.1f} us")
for tokens in (256, 512, 1024, 2048):
    CH = N // tokens
    partial = torch.empty(B, CH, ops.sa_partial_floats(S), device="cuda")
    t1 = timeit(lambda: ops.sa_main_tc(feats, A, a0, partial, None, CH, 1e-5, 1e-8))
    print(f"tokens/item {tokens:5d}: tc1 {t1:6.1f} us ({by / t1 / 1e3:6.0f} GB/s)")
tp = timeit(lambda: ops.sa_prepare(feats, 1e-5))
print(f"sa_prepare {tp:.1f} us ({2 * by / tp / 1e3:.0f} GB/s read+write)")
partial = torch.empty(B, 16, ops.sa_partial_floats(S), device="cuda")
out = torch.empty_like(slots)
tu = timeit(lambda: ops.sa_update2(partial, 16, slots, out, w, A, a0, 256, 1e-5, True, True))
print(f"sa_update2 {tu:.1f} us")
