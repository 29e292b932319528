""" development tool: back-to-back timing (one event pair around 50 launches) of the kernels of one cOCVP layer at the
    GridShapes5Objs rollout shape: 64 videos x 8 frames x 6 slots = 3072 tokens, E=128, 8 heads x 64, MLP 512 """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops

torch.manual_seed(0)
dev = "cuda"


def timeit(fn, reps=50):
    """ device time per launch: `reps` launches captured in a CUDA graph (no host launch cost between them) """
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


M = 3072
g, b = torch.ones(128, device=dev), torch.zeros(128, device=dev)
for (N, K, ln, act, res) in ((1536, 128, True, 0, False), (128, 512, False, 0, True), (512, 128, True, 2, False), (128, 512, False, 0, True), (128, 136, False, 0, False)):
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * K ** -0.5; bias = torch.randn(N, device=dev)
    r = torch.randn(M, N, device=dev) if res else None
    lnp = (g[:K].contiguous() if K <= 128 else torch.ones(K, device=dev), b[:K].contiguous() if K <= 128 else torch.zeros(K, device=dev), 1e-6) if ln else None
    for mode in ("fp16x3", "fp32"):
        t = timeit(lambda: ops.linear(x, w, bias, ln=lnp, act=act, res=r, mode=mode))
        print(f"linear M={M} N={N:5d} K={K:4d} ln={int(ln)} act={act} res={int(res)} {mode:7s}: {t:6.1f} us  {2.0 * M * N * K / t / 1e6:7.1f} TFLOP/s")
qkv = torch.randn(M, 1536, device=dev)
res = torch.randn(M, 512, device=dev)
t = timeit(lambda: ops.mha(qkv, 64, 48, 8, 64, 6))
print(f"mha 64 seq x 48 tok x 8 heads x 64: {t:6.1f} us")
qkv2 = torch.randn(64 * 6, 384, device=dev)
t = timeit(lambda: ops.mha(qkv2, 64, 6, 4, 32, 0))
print(f"mha (transition) 64 seq x 6 tok x 4 heads x 32: {t:6.1f} us")
