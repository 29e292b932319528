#!/usr/bin/env python
"""
bench.py -- PlaySlot seed + rollout throughput (frames/s) on N B200s, with the roofline of the dominant
kernel, the SlotAttention roofline (BASELINE.json metric) and the CPU baseline beside it.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun launches N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the unmodified reference from oracle/_ref)

A "step" = one pass of the seed+rollout path of src/05_evaluate_PlaySlot.py:112-139 over one batch of
synthetic video per GPU: savi(videos, decode=False) -> forward_multiple_samples(posterior) -> savi.decode.
Workload (BASELINE.json configs[2]): GridShapes5Objs config, 64 videos per GPU, 6 seed + 15 predicted
frames, 64x64 RGB, random-init weights, fp32.  Videos shard over the GPUs (weak scaling, no collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC, UNIT = "seed_rollout_frames_per_sec", "frames/s"

# --config: the workload.  "gridshapes5" is BASELINE.json configs[2], the configuration the headline metric is quoted on (and
# the default); the shipped 32-channel-encoder experiments and configs[4] (Sketchy-shaped 128x128, 16 slots, 6 + 30 frames)
# are reported with the same code path.
WORKLOADS = {
    "gridshapes5":   {"experiment": "GridShapes5Objs", "batch": 64, "ns": 6, "np": 15, "res": 64, "slots": None, "ref_batch": 4,
                      "what": "BASELINE.json configs[2]"},
    "blockpush":     {"experiment": "BlockPush", "batch": 64, "ns": 6, "np": 15, "res": 64, "slots": None, "ref_batch": 4,
                      "what": "BlockPush experiment (BASELINE.json configs[0] shapes), 8 slots, 32-channel encoder"},
    "buttonpress":   {"experiment": "ButtonPress", "batch": 64, "ns": 6, "np": 15, "res": 64, "slots": None, "ref_batch": 4,
                      "what": "ButtonPress experiment, 4 slots, 32-channel encoder"},
    "sketchy":       {"experiment": "Sketchy", "batch": 64, "ns": 6, "np": 15, "res": 64, "slots": None, "ref_batch": 4,
                      "what": "Sketchy experiment, 8 slots, 32-channel encoder, context 10"},
    "sketchy_hires": {"experiment": "Sketchy", "batch": 8, "ns": 6, "np": 30, "res": 128, "slots": 16, "ref_batch": 1,
                      "what": "BASELINE.json configs[4]: Sketchy-shaped 128x128 video, 16 slots, 6 + 30 frames, context 10"},
}
WORKLOAD = "gridshapes5"
EXPERIMENT, NUM_SEED, NUM_PREDS, RES, NUM_SLOTS = "GridShapes5Objs", 6, 15, 64, None


def set_workload(name):
    global WORKLOAD, EXPERIMENT, NUM_SEED, NUM_PREDS, RES, NUM_SLOTS
    w = WORKLOADS[name]
    WORKLOAD, EXPERIMENT, NUM_SEED, NUM_PREDS, RES, NUM_SLOTS = name, w["experiment"], w["ns"], w["np"], w["res"], w["slots"]
    return w


def experiment_cfg():
    from playslot_b200 import configs
    return configs.experiment(EXPERIMENT, num_slots=NUM_SLOTS, resolution=(RES, RES) if RES != 64 else None)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_burst": j["bf16_tflops"], "bf16_sustained": j["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """ samples SM clocks and throttle reasons with nvidia-smi while the timed region runs """
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # median of the upper half = clock under load (idle samples before/after the region drag the median down)
        under_load = sm[len(sm) // 2:] if sm else []
        med = under_load[len(under_load) // 2] if under_load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm

def reference_models(experiment, device, seed=0):
    """ the UNMODIFIED reference modules (oracle/_ref, installed by oracle/install_ref.py) built by the reference's own
    lib/setup_model.py:29-116 from the shipped experiment_params.json, random-init, eval mode """
    import contextlib
    import copy
    from oracle.install_ref import load_reference
    with contextlib.redirect_stdout(sys.stderr):                   # the reference's builders print their configuration
        setup_model, setup_predictor, experiment_json = load_reference()
        j = experiment_json(experiment)
        if NUM_SLOTS is not None:
            j["model"]["model_params"]["num_slots"] = NUM_SLOTS
        if RES != 64:
            j["model"]["model_params"]["encoder"]["encoder_params"]["resolution"] = [RES, RES]
            j["model"]["model_params"]["decoder"]["decoder_params"]["resolution"] = [RES, RES]
        torch.manual_seed(seed)
        savi = setup_model(copy.deepcopy(j["model"])).eval().to(device)
        pred = setup_predictor(copy.deepcopy(j)).eval().to(device)
    return savi, pred


@torch.no_grad()
def reference_pipeline(savi, pred, video, ns, npred):
    """ src/05_evaluate_PlaySlot.py:112-139 through the reference's public API """
    out = savi(video, num_imgs=ns + npred, decode=False)
    pred_slots, _ = pred.forward_multiple_samples(out["slot_history"], 1, ns, npred, use_posterior=True)
    B, n, P, S, D = pred_slots.shape
    imgs, _ = savi.decode(pred_slots.reshape(B * n * P, S, D))
    return imgs


def reference_available():
    from oracle.install_ref import available
    return available()


def oracle_models(cfg, seed=0):
    """ random-init weights of the architecture (our module mirror's initialiser), as CPU state dicts """
    from playslot_b200 import configs
    torch.manual_seed(seed)
    savi, pred = configs.build_models(cfg, "cpu")
    return ({k: v.detach() for k, v in savi.state_dict().items()}, {k: v.detach() for k, v in pred.state_dict().items()})


def oracle_step(cfg, ssd, psd, B, seed):
    from oracle import playslot_oracle as O
    scfg, pcfg = cfg["model"]["model_params"], cfg["predictor"]["predictor_params"]
    g = torch.Generator().manual_seed(seed)
    T = NUM_SEED + NUM_PREDS
    H, W = scfg["encoder"]["encoder_params"]["resolution"]
    S, D, A = scfg["num_slots"], scfg["slot_dim"], pcfg["action_dim"]
    video = torch.rand(B, T, 3, H, W, generator=g)
    per_slot = pcfg["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
    noise = torch.randn((B, T - 1, S, A) if per_slot else (B, T - 1, A), generator=g)
    init_noise = torch.randn(B, S, D, generator=g) if scfg["initializer"] == "LearnedRandom" else None
    t0 = time.perf_counter()
    with torch.no_grad():
        O.seed_rollout_pipeline(ssd, scfg, psd, pcfg, video, NUM_SEED, NUM_PREDS, noise, init_noise)
    return time.perf_counter() - t0


def cpu_reference_timer(cores):
    """ returns (step(B, seed) -> seconds, kind, description): the unmodified reference on the host cores when oracle/_ref
    travelled with the snapshot, otherwise the oracle port """
    torch.set_num_threads(cores)
    if reference_available():
        savi, pred = reference_models(EXPERIMENT, "cpu")

        def step(B, seed):
            g = torch.Generator().manual_seed(seed)
            video = torch.rand(B, NUM_SEED + NUM_PREDS, 3, RES, RES, generator=g)
            t0 = time.perf_counter()
            reference_pipeline(savi, pred, video, NUM_SEED, NUM_PREDS)
            return time.perf_counter() - t0
        return step, "reference", "the unmodified reference modules (oracle/_ref, lib/setup_model.py builders), torch CPU fp32"
    cfg = experiment_cfg()
    ssd, psd = oracle_models(cfg)
    return (lambda B, seed: oracle_step(cfg, ssd, psd, B, seed)), "port", "torch CPU fp32 oracle port (oracle/_ref absent)"


def torch_cuda_baseline(B, reps=10):
    """ the 10x target's denominator ('GPU reference baseline'): the unmodified reference modules
    .to('cuda'), eager, PyTorch's default flags (cuDNN TF32 convolutions on, fp32 matmuls), CUDA-event timed """
    savi, pred = reference_models(EXPERIMENT, "cuda")
    T = NUM_SEED + NUM_PREDS
    video = torch.rand(B, T, 3, RES, RES, device="cuda")
    ts = []
    for i in range(2 + reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reference_pipeline(savi, pred, video, NUM_SEED, NUM_PREDS)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    ms = ts[len(ts) // 2]
    del savi, pred
    torch.cuda.empty_cache()
    return {"value": B * T / ms * 1e3, "unit": UNIT, "ms_per_step": ms, "ms_min": ts[0], "ms_max": ts[-1], "timed_reps": reps,
            "warmup_reps": 2, "batch": B,
            "kind": "unmodified reference modules (oracle/_ref) on cuda:0: eager ATen / cuDNN / cuBLAS, default flags (TF32 "
                    "convolutions, fp32 matmuls); median of the timed reps"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    step, kind, what = cpu_reference_timer(cores)
    B = args.ref_batch
    for i in range(args.warmup):
        step(1, 100 + i)
    times = [step(B, 200 + i) for i in range(args.steps)]
    tot = sum(times)
    frames = B * (NUM_SEED + NUM_PREDS) * args.steps
    val = frames / tot
    cfg = workload_config(args.batch, args.gpus)                   # the workload of our arm ...
    cfg["ref_batch"] = B                                           # ... and what this arm actually ran per step
    cfg["sample"] = (f"bounded sample: {B} videos x {NUM_SEED + NUM_PREDS} frames per step on the host cores (the GPU arm runs "
                     f"{args.batch} videos per GPU per step); frames/s is normalised per frame")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{B} videos x {NUM_SEED + NUM_PREDS} frames per step, {args.steps} steps, {cores} threads; {what}"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(batch_per_gpu, n_gpus):
    slots = NUM_SLOTS or experiment_cfg()["model"]["model_params"]["num_slots"]
    return {"workload": f"{EXPERIMENT} seed+rollout: {batch_per_gpu} videos/GPU x {n_gpus} GPU, {NUM_SEED} seed + {NUM_PREDS} pred frames, "
                        f"{RES}x{RES} RGB, {slots} slots, encode(decode=False) -> InvDyn posterior + cOCVP rollout -> decode",
            "name": WORKLOAD, "what": WORKLOADS[WORKLOAD]["what"],
            "batch_per_gpu": batch_per_gpu, "num_seed": NUM_SEED, "num_preds": NUM_PREDS,
            "l2_policy": "per-step working set (>2 GB of activations) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"videos sharded over {n_gpus} GPU(s), no data-path collective"}


# ------------------------------------------------------------------------------------------------ GPU arm

def ncu_traffic(kernel):
    """ dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the committed ncu --set full capture
    (profiles/r01_traffic.json), or None """
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(kernel)
    except Exception:
        return None


def conv_flops(meta):
    n, cin, cout, H, W, k = meta
    return 2.0 * n * H * W * cin * cout * k * k


def sa_main_bytes(meta):
    B, N, S, D, masks = meta
    return B * N * D * 4 + (B * S * N * 4 if masks else 0) + 2 * B * S * (D + 1) * 4


def slot_attention_config2(pk):
    """ BASELINE.json configs[1]: SlotAttention forward, B=64, N=4096, D=128, 3 iterations, S in {8, 4}; plus the streaming
    kernel of one iteration timed back to back (one CUDA-event pair around 20 launches: an event pair around a single
    ~30 us launch would add its own ~5 us) on inputs larger than the L2 """
    from playslot_b200 import ops
    from playslot_b200.models.BlocksUtils.attention import SlotAttention
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for S in (8, 4):
        torch.manual_seed(S)
        m = SlotAttention(128, 128, S, 3, 1, 256).eval().cuda()
        feats = torch.randn(64, 4096, 128, device="cuda")
        slots = torch.randn(64, S, 128, device="cuda")
        with torch.no_grad():
            for _ in range(3):
                m(feats, slots, step=0)
            ts = []
            for _ in range(10):
                flush.zero_()                                   # flush L2 between timed iterations
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                m(feats, slots, step=0)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            byts = 64 * 4096 * 128 * 4 + 2 * 64 * S * 128 * 4 + 64 * S * 4096 * 4 + 720e3
            res = {"forward_ms": ms, "algorithmic_MB": byts / 1e6, "achieved_GBs": byts / ms / 1e6,
                   "frac": byts / ms / 1e6 / pk["hbm_gbs"], "frames_per_s": 64 / ms * 1e3,
                   "note": "whole 3-iteration forward (feature pre-normalisation pass + 3 streaming passes + 4 slot updates) against "
                           "the read-the-features-ONCE bound of SURVEY.md 8(d)"}
            if m.mode == "tc":
                # one iteration's streaming kernel, back to back
                w = m._weights()
                A = torch.empty(64, S, 128, device="cuda"); a0 = torch.empty(64, S, device="cuda")
                ops.sa_update2(None, 1, slots, None, w, A, a0, 256, 1e-5, False, True)
                planes = ops.sa_prepare(feats, 1e-5)
                CH = ops.sa_tc2_chunks(64, 4096)                # partial slots per frame of the balanced tile partition
                partial = torch.empty(64, CH, ops.sa_partial_floats(S), device="cuda")
                time.sleep(1.0)                                 # a kernel timed alone: let the clocks recover from the step loop
                masks = torch.empty(64, S, 4096, device="cuda")
                for with_masks in (False, True):
                    fn = lambda: ops.sa_main_tc2(planes, 0, 64, 4096, A, a0, partial, masks if with_masks else None, CH, 1e-8)
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) / 20 * 1e3
                    kb = 64 * 4096 * 128 * 4 + (64 * S * 4096 * 4 if with_masks else 0) + 2 * 64 * S * 129 * 4
                    res["iteration_kernel" + ("_with_masks" if with_masks else "")] = {
                        "kernel": "sa_main_tc2", "us": us, "algorithmic_MB": kb / 1e6, "achieved_GBs": kb / us / 1e3,
                        "frac": kb / us / 1e3 / pk["hbm_gbs"]}
            if reference_available():
                # the same call on the UNMODIFIED reference module (oracle/_ref), eager PyTorch-CUDA with default flags, same weights
                try:
                    from oracle.install_ref import load_reference
                    load_reference()
                    from models.BlocksUtils.attention import SlotAttention as RefSlotAttention
                    r = RefSlotAttention(128, 128, S, 3, 1, 256).eval().cuda()
                    r.load_state_dict(m.state_dict())
                    for _ in range(3):
                        ref_out = r(feats, slots, step=0)
                    ours_out = m(feats, slots, step=0)
                    tr = []
                    for _ in range(10):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        r(feats, slots, step=0)
                        e1.record()
                        torch.cuda.synchronize()
                        tr.append(e0.elapsed_time(e1))
                    tr.sort()
                    res["reference_torch_cuda_ms"] = tr[len(tr) // 2]
                    res["vs_reference_torch_cuda"] = tr[len(tr) // 2] / ms
                    res["max_rel_err_vs_reference_torch_cuda"] = float((ours_out - ref_out).abs().max() / ref_out.abs().max())
                    del r
                except Exception as e:                          # the baseline leg never takes the bench line down
                    res["reference_torch_cuda_ms"] = None
                    res["reference_error"] = repr(e)[:200]
        out[f"S{S}"] = res
    # "encode-only" variant of configs[1]: savi.encode(x) of one frame batch = CNN encoder + position embedding +
    # encoder MLP + the 3-iteration SlotAttention, BlockPush (8 slots) and ButtonPress (4 slots) shapes, 64 frames of 64 x 64
    try:
        out["encode_only"] = encode_only_sweep(flush)
    except Exception as e:
        out["encode_only"] = {"error": repr(e)[:300]}
    return out


@torch.no_grad()
def encode_only_sweep(flush):
    import contextlib
    import copy
    from playslot_b200 import configs
    res = {}
    for exp in ("BlockPush", "ButtonPress"):
        cfg = configs.experiment(exp)
        torch.manual_seed(0)
        savi, _ = configs.build_models(cfg, "cuda")
        savi.eval()
        x = torch.rand(64, 3, 64, 64, device="cuda")
        S = cfg["model"]["model_params"]["num_slots"]

        def ours():
            return savi_encode(savi, x)
        for _ in range(3):
            o = ours()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ours(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        entry = {"slots": S, "frames": 64, "ours_ms": ts[len(ts) // 2], "ours_frames_per_s": 64 / ts[len(ts) // 2] * 1e3}
        if reference_available():
            from oracle.install_ref import load_reference
            with contextlib.redirect_stdout(sys.stderr):
                setup_model, _, experiment_json = load_reference()
                rsavi = setup_model(copy.deepcopy(experiment_json(exp)["model"])).eval().cuda()
            rsavi.load_state_dict(savi.state_dict())
            for _ in range(3):
                ro = savi_encode(rsavi, x)
            tr = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); savi_encode(rsavi, x); e1.record()
                torch.cuda.synchronize()
                tr.append(e0.elapsed_time(e1))
            tr.sort()
            entry["reference_torch_cuda_ms"] = tr[len(tr) // 2]
            entry["vs_reference_torch_cuda"] = tr[len(tr) // 2] / entry["ours_ms"]
            entry["max_rel_err_vs_reference_torch_cuda"] = float((o - ro).abs().max() / ro.abs().max())
            del rsavi
        res[exp] = entry
        del savi
    torch.cuda.empty_cache()
    return res


def savi_encode(savi, x):
    """ one frame batch through the public API, SAVi.forward with one frame and decode=False (SAVi.py:135-204): CNN encoder,
    position embedding, encoder MLP, the 3 SlotAttention iterations of a first frame (and the transition block, negligible) --
    the same call for our mirror and for the reference module """
    return savi(x.unsqueeze(1), num_imgs=1, decode=False)["slot_history"]


def run_ours(args):
    import torch.distributed as dist
    from playslot_b200 import configs, ops, shard
    from playslot_b200.pipeline import seed_rollout, seed_rollout_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    cfg = experiment_cfg()
    torch.manual_seed(0)                                           # same random-init weights on every rank
    savi, pred = configs.build_models(cfg, "cuda")
    savi.conv_mode = args.conv_mode
    ops.set_linear_mode(args.conv_mode)                            # one arithmetic mode for every tensor-core path
    savi.slot_attention.mode = "fp32" if args.conv_mode == "fp32" else "tc"
    B, T = args.batch, NUM_SEED + NUM_PREDS
    g = torch.Generator().manual_seed(1234 + rank)                 # a different synthetic shard per rank
    video_host = torch.rand(B, T, 3, RES, RES, generator=g).pin_memory()
    video = video_host.cuda()
    out_host = torch.empty(B, NUM_PREDS, 3, RES, RES, pin_memory=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        return shard.max_over_ranks(ms, device="cuda")

    from playslot_b200.pipeline import GraphedSeedRollout
    graphed = None if args.eager else GraphedSeedRollout(savi, pred, B, NUM_SEED, NUM_PREDS, resolution=(RES, RES))
    step = (lambda: graphed(video)) if graphed else (lambda: seed_rollout(savi, pred, video, NUM_SEED, NUM_PREDS))
    step_host = (lambda: graphed.run_host(video_host, out_host)) if graphed else \
        (lambda: seed_rollout_host(savi, pred, video_host, NUM_SEED, NUM_PREDS, out_host))
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident timing -------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ops.reset_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = graphed.launches_per_replay * args.steps if graphed else ops.launch_count()

    # ---- end-to-end timing: pinned host video in, predicted frames back to the host, every step ------------
    for _ in range(2):
        step_host()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e2.record()
    for _ in range(args.steps):
        step_host()
    e3.record()
    barrier()
    wall_e2e = time.perf_counter() - t_host0
    ms_e2e_serial = max_over_ranks(max(e2.elapsed_time(e3), 1e3 * wall_e2e))
    ms_e2e, e2e_mode = ms_e2e_serial, "one blocking call per batch: upload, kernels, download, synchronise"
    if graphed is not None:
        # the streaming API (HostPipeline.submit / wait): the same per-batch uploads and downloads from / to pinned host
        # memory, on copy streams, overlapping the neighbouring batches' kernels -- host wall clock around the whole loop
        pipe = graphed.host_pipeline()
        for _ in range(2):
            pipe.submit(video_host, out_host)
        pipe.wait()
        barrier()
        t_host0 = time.perf_counter()
        for _ in range(args.steps):
            pipe.submit(video_host, out_host)
        pipe.wait()
        wall_pipe = time.perf_counter() - t_host0
        barrier()
        ms_e2e = max_over_ranks(1e3 * wall_pipe)
        e2e_mode = "HostPipeline.submit per batch + wait: uploads / downloads on copy streams overlap neighbouring batches"
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel device times: the same step once more, eagerly, with CUDA events around every launch (the graph
    #      replays exactly these kernels; events cannot be placed inside a captured graph) -----------------------------
    ops.profile_start()
    seed_rollout(savi, pred, video, NUM_SEED, NUM_PREDS)
    prof = ops.profile_stop()
    barrier()

    frames_per_step = B * T * world
    value = frames_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = frames_per_step * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (by summed device time over the timed region) ------------------------
    total_prof = sum(v["ms"] for v in prof.values()) or 1.0
    shares = {k: round(v["ms"] / total_prof, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    roofline = None
    if dom[0].startswith("conv"):
        fl = sum(conv_flops(m) for m in dom[1]["meta"])
        ach = fl / (dom[1]["ms"] * 1e-3) / 1e12
        note = ("algorithmic FLOPs = 2*n*H*W*Cin*Cout*k*k of the launches in the timed region (the fp16x3 mode issues 3x that on "
                "the tensor pipe, padded-grid junk columns add 6%); peak = " + pk["source"] + " sustained bf16")
        roofline = {"kernel": dom[0], "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["bf16_sustained"], "traffic": None, "note": note, "share_of_step": shares[dom[0]],
                    "launches": dom[1]["calls"], "avg_launch_ms": dom[1]["ms"] / dom[1]["calls"]}
        if dom[0].startswith("conv5x5_tc_p3"):
            roofline["issued_tensor_TFLOPs"] = 3 * ach * (68.0 / 64.0)
        tr = ncu_traffic(dom[0])
        if tr:
            n_imgs = sum(m[0] for m in dom[1]["meta"]) / max(dom[1]["calls"], 1)
            roofline["traffic"] = tr["dram_bytes_per_launch"] / tr["images"] * n_imgs
            roofline["traffic_note"] = (f"DRAM read+write of a {tr['images']}-image launch under ncu ({tr['dram_bytes_per_launch'] / 1e9:.2f} GB vs "
                                        f"{tr['algorithmic_bytes'] / 1e9:.2f} GB algorithmic: one read + one write of the split planes), scaled to the "
                                        f"average {n_imgs:.0f} images per launch of the timed region")
    sa_name = next((n for n in ("sa_main_tc2", "sa_main_tc", "sa_main") if n in prof), "sa_main")
    sa = prof.get(sa_name)
    roofline_sa = None
    if sa:
        by = sum(sa_main_bytes(m) for m in sa["meta"])
        ach = by / (sa["ms"] * 1e-3) / 1e9
        tr_sa = ncu_traffic(sa_name)
        roofline_sa = {"kernel": sa_name, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                       "frac": ach / pk["hbm_gbs"], "traffic": tr_sa["dram_bytes_per_launch"] if tr_sa else None, "launches": sa["calls"],
                       "avg_launch_us": 1e3 * sa["ms"] / sa["calls"], "share_of_step": shares.get(sa_name),
                       "note": "one streaming pass of a SlotAttention iteration over 64 frames x 4096 tokens x 512 bytes (pre-normalised fp16 hi/lo "
                               "planes, read once) + masks; per-launch CUDA-event pairs inside the step add ~5 us to each ~30 us launch: "
                               "slot_attention_config2.*.iteration_kernel holds the back-to-back timing of the same kernel"}
        if roofline is None:
            roofline = roofline_sa

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp16x3": "f32 (convolutions, linears and SlotAttention contractions as fp16x3 split tensor-core products: 22-bit mantissa, fp32 accumulate)",
                  "fp16": "fp16 tensor-core products + f32", "bf16": "bf16 tensor-core products + f32", "fp32": "f32 (CUDA cores)"}[args.conv_mode],
        "data": "synthetic", "config": dict(workload_config(B, world), conv_mode=args.conv_mode, linear_mode=ops.linear_mode(),
                                            slot_attention_mode=savi.slot_attention.mode),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": video_host.numel() * 4 * world,
                "d2h_bytes_per_step": out_host.numel() * 4 * world, "ms_per_step": ms_e2e / args.steps,
                "api": e2e_mode, "serial_ms_per_step": ms_e2e_serial / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline, "roofline_slot_attention": roofline_sa,
        "kernel_time_shares": shares, "peaks": pk,
    }
    if world == 1 and not args.skip_extras:
        if WORKLOAD == "gridshapes5":
            line["slot_attention_config2"] = slot_attention_config2(pk)
        if args.conv_mode == "fp16x3" and graphed is not None:
            # The reference's PyTorch-CUDA path runs its convolutions in TF32 (10-bit mantissa, cuDNN default).  The same
            # workload with single fp16 products (11-bit significand; stated tolerance 1e-2) is the
            # precision-matched comparison to `torch_cuda_baseline`; the headline `value` stays the fp32-grade mode.
            try:
                savi.conv_mode = "fp16"
                ops.set_linear_mode("fp16")
                g16 = GraphedSeedRollout(savi, pred, B, NUM_SEED, NUM_PREDS, resolution=(RES, RES))
                for _ in range(3):
                    g16(video)
                torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(args.steps):
                    g16(video)
                f1.record()
                torch.cuda.synchronize()
                ms16 = f0.elapsed_time(f1) / args.steps
                line["tf32_class_mode"] = {"conv_mode": "fp16", "linear_mode": "fp16", "ms_per_step": ms16,
                                           "value": B * T / ms16 * 1e3, "unit": UNIT,
                                           "note": "single fp16 tensor-core products, fp32 accumulate: the precision class of the "
                                                   "TF32 convolutions torch_cuda_baseline runs"}
                del g16
            except Exception as e:                                     # informational only
                line["tf32_class_mode"] = {"error": repr(e)[:200]}
            finally:
                savi.conv_mode = args.conv_mode
                ops.set_linear_mode(args.conv_mode)
        cores = os.cpu_count() or 1
        try:
            if not reference_available():
                raise FileNotFoundError("oracle/_ref did not travel with the snapshot")
            tcb = torch_cuda_baseline(B)
            line["torch_cuda_baseline"] = tcb
            line["vs_torch_cuda"] = {"device_resident": value / tcb["value"], "e2e": e2e_value / tcb["value"],
                                     "target": 10.0, "note": "our frames/s / the reference's PyTorch-CUDA frames/s, same GPU, same batch"}
            if "value" in line.get("tf32_class_mode", {}):
                line["vs_torch_cuda"]["tf32_class_mode"] = line["tf32_class_mode"]["value"] / tcb["value"]
        except Exception as e:                                         # informational only
            line["torch_cuda_baseline"] = {"error": repr(e)[:200]}
        step_cpu, kind, what = cpu_reference_timer(cores)
        step_cpu(1, 7)                                                  # warm-up
        bs = args.ref_batch
        dt = step_cpu(bs, 8)
        line["cpu_baseline"] = {"value": bs * T / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"{bs} videos x {T} frames of the same workload, one pass after a 1-video warm-up "
                                          f"({dt:.1f} s), {cores} threads; {what}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="gridshapes5", choices=sorted(WORKLOADS), help="workload (default: BASELINE.json configs[2])")
    ap.add_argument("--mode", default="rollout", choices=["rollout", "train"],
                    help="rollout = seed + rollout inference (the headline metric); train = BASELINE.json configs[3], the joint "
                         "InvDyn + cOCVP training step on ButtonPress with the NCCL gradient all-reduce")
    ap.add_argument("--batch", type=int, default=None, help="videos per GPU (default: the workload's)")
    ap.add_argument("--ref-batch", type=int, default=None, help="videos per step of the CPU arm (bounded sample)")
    ap.add_argument("--train-arith", default="fp32", choices=["fp32", "fp16x3", "bf16"],
                    help="--mode train: arithmetic of the predictor's Linear layers (bf16 = one bf16 tensor-core product, fp32 master "
                         "weights; the decoder convolutions then run single fp16 products)")
    ap.add_argument("--conv-mode", default=os.environ.get("PLAYSLOT_CONV_MODE", "fp16x3"), choices=["fp16x3", "fp16", "bf16", "fp32"],
                    help="arithmetic of the 64->64 5x5 convolutions (fp16x3 = split fp16 tensor-core products, fp32-grade)")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--skip-extras", action="store_true", help="skip the config-2 SlotAttention sweep and the CPU baseline")
    args = ap.parse_args()
    if args.mode == "train":
        from bench_train import run_train, run_train_reference
        (run_train_reference if args.impl == "reference" else run_train)(args)
        return
    w = set_workload(args.config)
    if args.batch is None:
        args.batch = w["batch"]
    if args.ref_batch is None:
        args.ref_batch = w["ref_batch"]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
