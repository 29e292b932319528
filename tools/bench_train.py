#!/usr/bin/env python
"""
BASELINE.json configs[3]: joint cOCVP + InvDyn training step (forward + backward + gradient all-reduce + clip + Adam),
ButtonPress config, synthetic video, batch sharded over the GPUs (weak scaling: --batch videos per GPU).
  python tools/bench_train.py [--batch 8] [--steps 10]                       (1 GPU)
  python -m torch.distributed.run --nproc-per-node N ... tools/bench_train.py (N GPUs, NCCL)
Prints one JSON line on rank 0.  Arithmetic: fp32 linears / attention, fp16x3 tensor-core decoder convolutions.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from playslot_b200 import configs, ops, shard, train

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--experiment", default="ButtonPress")
ap.add_argument("--num-seed", type=int, default=6)
ap.add_argument("--num-preds", type=int, default=8)
ap.add_argument("--graph", action="store_true", help="capture the whole step in a CUDA graph (single GPU)")
ap.add_argument("--linear-mode", default="fp32", choices=["fp32", "fp16x3"], help="arithmetic of the predictor Linears (with --graph)")
args = ap.parse_args()
rank, world, local = shard.world()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = configs.experiment(args.experiment)
torch.manual_seed(0)
savi, pred = configs.build_models(cfg, "cuda")
for p in savi.parameters():
    p.requires_grad_(False)
opt = train.make_optimizer(pred, capturable=args.graph)
bucket = train.GradientBucket(pred.parameters())
T = args.num_seed + args.num_preds
g = torch.Generator().manual_seed(7 + rank)
video = torch.rand(args.batch, T, 3, 64, 64, generator=g).cuda()


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


if args.graph:
    graphed = train.GraphedTrainStep(savi, pred, opt, args.batch, args.num_seed, args.num_preds, linear_mode=args.linear_mode)
    step = lambda: graphed(video)
else:
    step = lambda: train.train_step(savi, pred, opt, video, args.num_seed, args.num_preds, bucket=bucket)
for _ in range(args.warmup):
    loss, terms = step()
sync()
ops.reset_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(args.steps):
    loss, terms = step()
e1.record()
sync()
wall = time.perf_counter() - t0
ms = shard.max_over_ranks(max(e0.elapsed_time(e1), 1e3 * wall), device="cuda") / args.steps
if rank == 0:
    n_params = sum(p.numel() for p in pred.parameters() if p.requires_grad)
    print(json.dumps({
        "metric": "train_step_videos_per_sec", "value": args.batch * world / (ms * 1e-3), "unit": "videos/s", "n_gpus": world,
        "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "scaling": "weak",
        "frames_per_sec": args.batch * world * T / (ms * 1e-3),
        "config": {"workload": f"{args.experiment} joint InvDyn+cOCVP training step, {args.batch} videos/GPU, {args.num_seed}+{args.num_preds} frames, "
                               "SAVi frozen (encode without gradients), decoder data gradient, MSE(img)+MSE(slots)+VQ, clip 0.05, Adam",
                   "dtype": "f32 (decoder convolutions fp16x3 on tcgen05)", "cuda_graph": bool(args.graph), "linear_mode": args.linear_mode if args.graph else "fp32"},
        "loss": float(loss), "trainable_params": n_params, "allreduce_bytes_per_step": 4 * n_params if world > 1 else 0,
        "gpu_launches_per_step": ops.launch_count() / args.steps}))
if world > 1:
    dist.destroy_process_group()
