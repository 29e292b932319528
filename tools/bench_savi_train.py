"""
SAVi training step (src/02_train_savi.py:21-70; BlockPush experiment: batch 32, 8 frames, 64x64) -- ours vs the unmodified reference
on the same GPU (eager PyTorch, default flags) and per-kernel shares of our step.   python tools/bench_savi_train.py [--batch 32]
"""
import argparse
import contextlib
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from playslot_b200 import configs, ops, train          # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--experiment", default="BlockPush")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    cfg = configs.experiment(args.experiment)
    torch.manual_seed(0)
    savi, _ = configs.build_models(cfg, "cuda")
    opt = torch.optim.Adam(savi.parameters(), lr=1e-4)
    video = torch.rand(args.batch, args.frames, 3, 64, 64, device="cuda")
    ms = timed(lambda: train.savi_train_step(savi, opt, video), args.reps)
    # the same step captured in one CUDA graph (train.GraphedSaviTrainStep)
    graphed_ms, graphed_err = None, None
    try:
        import copy as _copy
        savi_g = _copy.deepcopy(savi)
        opt_g = torch.optim.Adam(savi_g.parameters(), lr=1e-4, capturable=True)
        gstep = train.GraphedSaviTrainStep(savi_g, opt_g, args.batch, args.frames)
        graphed_ms = timed(lambda: gstep(video), max(args.reps, 10))
        launches = gstep.launches_per_replay
        del gstep, savi_g, opt_g
    except Exception as e:
        graphed_err = repr(e)[:300]
    ops.profile_start()
    train.savi_train_step(savi, opt, video)
    prof = ops.profile_stop()
    tot = sum(v["ms"] for v in prof.values())
    out = {"workload": f"SAVi training step, {args.experiment}, {args.batch} videos x {args.frames} frames", "ours_ms": ms,
           "ours_graphed_ms": graphed_ms, "graphed_error": graphed_err,
           "ours_videos_per_s": args.batch / ms * 1e3,
           "kernel_ms": {k: round(v["ms"], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]}, "profiled_kernel_ms": tot}
    try:
        sys.path.insert(0, ROOT)
        from oracle.install_ref import load_reference
        with contextlib.redirect_stdout(sys.stderr):
            setup_model, _, experiment_json = load_reference()
            j = experiment_json(args.experiment)
            torch.manual_seed(0)
            ref = setup_model(copy.deepcopy(j["model"])).train().cuda()
        ropt = torch.optim.Adam(ref.parameters(), lr=1e-4)

        def ref_step():
            ropt.zero_grad()
            o = ref(video, num_imgs=args.frames)
            loss = torch.nn.functional.mse_loss(o["recons_imgs"], video.clamp(0, 1))
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05)
            ropt.step()
        rms = timed(ref_step, args.reps)
        out["reference_torch_cuda_ms"] = rms
        out["speedup_vs_reference_torch_cuda"] = rms / ms
    except Exception as e:
        out["reference_error"] = repr(e)[:200]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
