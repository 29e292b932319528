"""
bench.py --mode train: BASELINE.json configs[3] -- the joint cOCVP + InvDyn training step of src/04_train_PlaySlot.py:32-122
(SAVi frozen: encode without gradients; predictor forward with BPTT; decoder with its data gradient; PredImgMSE + PredSlotMSE +
VQ losses; backward; NCCL gradient all-reduce; clip_grad_norm_(0.05); Adam) on the ButtonPress experiment, 6 seed + 8 predicted
frames, synthetic video, --batch videos per GPU (default 8 = the JSON's training.batch_size; weak scaling over the GPUs).

The whole step -- both all-reduces included -- is ONE CUDA graph per rank (train.GraphedTrainStep).  Same JSON-line contract as
the rollout mode; metric = training videos per second over all GPUs.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EXPERIMENT, NUM_SEED, NUM_PREDS = "ButtonPress", 6, 8
METRIC, UNIT = "train_step_videos_per_sec", "videos/s"


def _config(batch, world, arith):
    return {"workload": f"{EXPERIMENT} joint InvDyn + cOCVP training step (BASELINE.json configs[3]): {batch} videos/GPU x {world} GPU, "
                        f"{NUM_SEED} seed + {NUM_PREDS} pred frames, 64x64 RGB, 4 slots; SAVi frozen, decoder data gradient, "
                        "PredImgMSE + PredSlotMSE + VQ, clip 0.05, Adam",
            "name": "train_buttonpress", "batch_per_gpu": batch, "num_seed": NUM_SEED, "num_preds": NUM_PREDS,
            "arithmetic": arith,
            "l2_policy": "per-step working set (saved activations of the decoder: > 1 GB) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"videos sharded over {world} GPU(s); one NCCL all-reduce of the flat gradient bucket + one of the EMA-VQ "
                           "statistics per step, both inside the CUDA graph"}


def run_train(args):
    import torch.distributed as dist
    from bench import ClockSampler
    from playslot_b200 import configs, ops, shard, train

    rank, world, local = shard.world()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch or 8
    T = NUM_SEED + NUM_PREDS
    cfg = configs.experiment(EXPERIMENT)
    torch.manual_seed(0)                                         # the same initial weights on every rank
    savi, pred = configs.build_models(cfg, "cuda")
    for p in savi.parameters():
        p.requires_grad_(False)
    arith = args.train_arith
    savi.conv_mode = "fp16" if arith == "bf16" else "fp16x3"
    opt = train.make_optimizer(pred, capturable=True)
    graphed = train.GraphedTrainStep(savi, pred, opt, B, NUM_SEED, NUM_PREDS, linear_mode=arith)
    g = torch.Generator().manual_seed(7 + rank)
    video_host = torch.rand(B, T, 3, 64, 64, generator=g).pin_memory()
    video = video_host.cuda()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        loss, _ = graphed(video)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = graphed(video)
    e1.record()
    barrier()
    ms_total = shard.max_over_ranks(e0.elapsed_time(e1), device="cuda")
    # end to end: the batch comes from pinned host memory and the loss goes back to the host, every step
    losses = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss, _ = graphed(video_host)
        losses.append(float(loss))                               # device -> host read of the step's result (synchronises)
    barrier()
    ms_e2e = shard.max_over_ranks(1e3 * (time.perf_counter() - t0), device="cuda")
    clocks = sampler.stop() if rank == 0 else None

    # the gradient all-reduce alone (nothing overlaps it inside the step, so this is its exposed time)
    ar_ms = 0.0
    if world > 1:
        b = graphed.bucket
        for _ in range(3):
            dist.all_reduce(b.flat)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            dist.all_reduce(b.flat)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = shard.max_over_ranks(a0.elapsed_time(a1) / 20, device="cuda")
        b.flat.zero_()
    n_params = sum(p.numel() for p in pred.parameters() if p.requires_grad)
    if rank == 0:
        ms = ms_total / args.steps
        value = B * world / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32 (Linear layers fp32 FFMA; decoder convolutions fp16x3 on tcgen05)",
                      "fp16x3": "f32 (Linear forward / data-gradient GEMMs and decoder convolutions as fp16x3 tensor-core products)",
                      "bf16": "bf16 (Linear forward / data-gradient GEMMs: one bf16 tensor-core product, fp32 accumulate, fp32 master "
                              "weights and weight gradients; decoder convolutions: one fp16 product; InvDyn tail fp32)"}[arith],
            "data": "synthetic", "config": _config(B, world, arith), "clocks": clocks,
            "frames_per_sec": value * T,
            "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": video_host.numel() * 4 * world, "d2h_bytes_per_step": 4 * world,
                    "api": "GraphedTrainStep(video_pinned_host) -> loss.item() every step"},
            "gpu_launches": int(graphed.launches_per_replay * args.steps),
            "allreduce": {"bytes_per_step": 4 * (n_params + 1) if world > 1 else 0, "alone_ms": ar_ms,
                          "note": "NCCL all-reduce of the flat fp32 gradient bucket timed alone, back to back; inside the step it follows "
                                  "the backward (nothing left to overlap), so this is its exposed time"},
            "trainable_params": n_params, "loss_first_last": [losses[0], losses[-1]],
        }
        if world == 1 and not args.skip_extras:
            try:
                from oracle.install_ref import available
                if available():
                    tcb = torch_cuda_train_baseline(B)
                    line["torch_cuda_baseline"] = tcb
                    line["vs_torch_cuda"] = {"device_resident": value / tcb["value"], "e2e": line["e2e"]["value"] / tcb["value"],
                                             "note": "our videos/s / the reference's PyTorch-CUDA training step, same GPU, same batch"}
            except Exception as e:
                line["torch_cuda_baseline"] = {"error": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # destroy_process_group() after graph-captured collectives blocks in the NCCL watchdog shutdown (measured: the process hangs
        # until killed); the results are out, leave without the teardown
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def reference_train_step(device, B):
    """ step(seed) -> seconds: one training step of the UNMODIFIED reference (modules + lib/loss.py from oracle/_ref; frozen SAVi,
    predictor forward, decode, PredImgMSE + PredSlotMSE + VQ losses, backward, clip, Adam -- 04_train_PlaySlot.py:32-122,
    basePredictorTrainer.py:320-333) on `device`, eager PyTorch with default flags """
    import contextlib
    import copy
    from oracle.install_ref import load_reference
    T = NUM_SEED + NUM_PREDS
    with contextlib.redirect_stdout(sys.stderr):
        setup_model, setup_predictor, experiment_json = load_reference()
        from lib.loss import PredImgMSELoss, PredSlotMSELoss, VQLoss
        j = experiment_json(EXPERIMENT)
        torch.manual_seed(0)
        savi = setup_model(copy.deepcopy(j["model"])).eval().to(device)
        pred = setup_predictor(copy.deepcopy(j)).train().to(device)
    for p in savi.parameters():
        p.requires_grad_(False)
    opt = torch.optim.Adam(pred.parameters(), lr=1e-4)
    l_img, l_slot, l_vq = PredImgMSELoss(), PredSlotMSELoss(), VQLoss(beta=0.25)
    cuda = str(device).startswith("cuda")

    def step(seed):
        g = torch.Generator().manual_seed(seed)
        video = torch.rand(B, T, 3, 64, 64, generator=g).to(device)
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            sh = savi(video, num_imgs=T, decode=False)["slot_history"]
        pred_slots, out = pred(imgs=video, slots=sh, actions=None, num_seed=NUM_SEED, num_preds=NUM_PREDS)
        n, S, D = pred_slots.shape[1:]
        imgs, _ = savi.decode(pred_slots.reshape(B * n, S, D))
        vq = l_vq(**out["vq_losses"])
        loss = l_img(pred_imgs=imgs.view(B, n, *imgs.shape[1:]), target_imgs=video[:, 1:T]) + l_slot(preds=pred_slots, targets=sh[:, 1:T]) \
            + vq["vq_quant_loss"] + vq["vq_commit_loss"]
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(pred.parameters(), 0.05)
        opt.step()
        if cuda:
            torch.cuda.synchronize()
        return time.perf_counter() - t0
    return step


def torch_cuda_train_baseline(B, reps=10):
    """ the reference's training step on the same GPU (the PyTorch-CUDA denominator of the training number) """
    step = reference_train_step("cuda", B)
    ts = sorted(step(300 + i) for i in range(2 + reps))[: reps]     # drop the two slowest (warm-up: cuDNN autotune, allocator)
    ms = 1e3 * ts[len(ts) // 2]
    torch.cuda.empty_cache()
    return {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_min": 1e3 * ts[0], "ms_max": 1e3 * ts[-1], "timed_reps": reps,
            "batch": B, "kind": "unmodified reference modules + lib/loss.py (oracle/_ref) on cuda:0: eager ATen / cuDNN / cuBLAS, "
                                "default flags (TF32 convolutions), fp32; wall clock with synchronise, median"}


def run_train_reference(args):
    """ the unmodified reference's training step (modules + lib/loss.py from oracle/_ref) on the host cores, bounded sample """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.install_ref import available
    if not available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref did not travel with the snapshot"}))
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.ref_batch or 2
    T = NUM_SEED + NUM_PREDS
    step = reference_train_step("cpu", B)

    for i in range(args.warmup):
        step(100 + i)
    tot = sum(step(200 + i) for i in range(args.steps))
    val = B * args.steps / tot
    cfg = _config(args.batch or 8, args.gpus, "f32")
    cfg["ref_batch"] = B
    cfg["sample"] = f"bounded sample: {B} videos per step on the host cores; videos/s is normalised per video"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": f"{B} videos x {T} frames per step, {args.steps} steps, {cores} threads; the unmodified reference modules "
                                   "and lib/loss.py (oracle/_ref): forward + backward + clip + Adam, torch CPU fp32"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
