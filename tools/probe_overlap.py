"""
Does the latency-bound cOCVP step chain run concurrently with the persistent decoder kernels?  Two CUDA graphs (10 roll-out steps;
4 frame decodes of 64 x 6 slot-images) replayed alone and together on two streams, for several SM reservations.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from playslot_b200 import configs, ops          # noqa: E402


def capture(fn, stream):
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        fn()
    return g


def timed(graphs_streams, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs = []
    for _ in range(reps):
        for g, s in graphs_streams:
            with torch.cuda.stream(s):
                g.replay()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    cfg = configs.experiment("GridShapes5Objs")
    torch.manual_seed(0)
    savi, pred = configs.build_models(cfg, "cuda")
    B, S, D, A = 64, 6, 128, 8
    slots = torch.randn(B, 8, S, D, device="cuda")
    protos = torch.randn(B, 8, S, A, device="cuda")
    dec_in = torch.randn(B, S, D, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    dyn = pred.dynamics_model

    @torch.no_grad()
    def chain():
        for _ in range(10):
            dyn.forward_single(slots=slots, action_protos=protos, action_vars=None)

    g_chain = capture(chain, s1)
    print("chain alone (10 steps): %.3f ms" % timed([(g_chain, s1)]), flush=True)
    for r in (0, 8, 16, 32, 64):
        @torch.no_grad()
        def dec():
            with ops.sm_reserve(r):
                for _ in range(4):
                    savi.decode(dec_in)
        g_dec = capture(dec, s2)
        a = timed([(g_dec, s2)])
        b = timed([(g_chain, s1), (g_dec, s2)])
        print(f"reserve {r:2d}: decode alone (4 frames) {a:.3f} ms | chain + decode together {b:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
