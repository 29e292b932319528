""" development tool: per-node cost of a CUDA graph of dependent tiny kernels (the floor every launch of the step pays) """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from playslot_b200 import ops
sys.path.insert(0, os.path.join(ROOT, "tools"))

def graph_time(fn, reps=200):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

z = torch.randn(256, device="cuda")
print("tiny act kernel (256 elements), dependent chain: %.2f us per node" % graph_time(lambda: ops.act_fwd(z, 1)))
x = torch.randn(128, 128, device="cuda"); w = torch.randn(128, 128, device="cuda") * 0.1; b = torch.zeros(128, device="cuda")
print("linear_tc 128x128x128 (one CTA): %.2f us per node" % graph_time(lambda: ops.linear(x, w, b, mode="fp16x3")))
x = torch.randn(3072, 128, device="cuda")
print("linear_tc 3072x128x128: %.2f us per node" % graph_time(lambda: ops.linear(x, w, b, mode="fp16x3")))
x = torch.randn(3072, 512, device="cuda"); w = torch.randn(128, 512, device="cuda") * 0.05
print("linear_tc 3072x128x512: %.2f us per node" % graph_time(lambda: ops.linear(x, w, b, mode="fp16x3")))
print("linear fp32 3072x128x512: %.2f us per node" % graph_time(lambda: ops.linear(x, w, b, mode="fp32")))
