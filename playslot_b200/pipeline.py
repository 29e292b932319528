"""
The seed + rollout path a user of the reference runs (src/05_evaluate_PlaySlot.py:112-139, `forward_model`):
    savi(videos, num_imgs=ns+np, decode=False) -> predictor.forward_multiple_samples(slot_history, num_samples,
    ns, np, use_posterior=True) -> savi.decode(pred_slots)
exposed as one call, plus a host-buffer variant used for end-to-end timing.
"""
import os

import torch

from . import ops


# Measured on B200 (tools/bench_overlap.py, tools/probe_overlap.py; profiles/r02_overlap_probe.md): NO gain -- the roll-out kernels
# (96-144 CTAs each) are not purely latency-bound; confined to the 8-32 SMs a persistent decoder kernel can spare they slow down
# 3-9x and become the critical path (56.1 ms sequential vs 55.9-58.2 ms overlapped).  Kept as an option, off by default.
OVERLAP_DECODE = os.environ.get("PLAYSLOT_OVERLAP_DECODE", "0") != "0"
SM_RESERVE = int(os.environ.get("PLAYSLOT_SM_RESERVE", "16"))


@torch.no_grad()
def seed_rollout(savi, predictor, videos, num_seed, num_preds, num_samples=1, use_posterior=True, decode=True, overlap=None,
                 **init_kwargs):
    """
    videos (B, >= num_seed+num_preds, 3, H, W) on the GPU ->
      dict(slot_history (B, ns+np, S, D), pred_slots (B, num_samples, np, S, D), pred_imgs (B, num_samples, np, 3, H, W),
           pred_objs, pred_masks)

    overlap (default: off, see OVERLAP_DECODE; posterior roll-out with num_samples == 1 only): the roll-out is a chain of small kernels
    and the decoder a throughput-bound stack of persistent convolutions, so predicted frame i is decoded on a side stream
    while the cOCVP step of frame i+1 runs on the main stream; the persistent decoder kernels leave SM_RESERVE SMs free for
    it (ops.sm_reserve).  Same kernels on the same data as the sequential order: results are bit-identical.
    """
    T = num_seed + num_preds
    out = savi(videos, num_imgs=T, decode=False, **init_kwargs)
    slot_history = out["slot_history"]
    dyn = getattr(predictor, "dynamics_model", None)
    if overlap is None:
        overlap = OVERLAP_DECODE
    overlap = bool(overlap and decode and num_samples == 1 and use_posterior and dyn is not None and num_preds > 1)
    if not overlap:
        pred_slots, actions = predictor.forward_multiple_samples(
            slot_history=slot_history, num_samples=num_samples, num_seed=num_seed, num_preds=num_preds, use_posterior=use_posterior)
        res = {"slot_history": slot_history, "pred_slots": pred_slots, "actions": actions}
        if decode:
            B, n, P, S, D = pred_slots.shape
            imgs, (objs, masks) = savi.decode(pred_slots.reshape(B * n * P, S, D))
            res["pred_imgs"] = imgs.view(B, n, P, *imgs.shape[1:])
            res["pred_objs"] = objs.view(B, n, P, *objs.shape[1:])
            res["pred_masks"] = masks.view(B, n, P, *masks.shape[1:])
        return res

    main = torch.cuda.current_stream()
    side = _side_stream(videos.device)
    capturing = torch.cuda.is_current_stream_capturing()
    B, S, D = slot_history.shape[0], slot_history.shape[2], slot_history.shape[3]
    frames = []                                                  # per predicted frame: (recon, recons, masks), made on `side`
    side.wait_stream(main)

    def decode_step(i, nxt):
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)
        last = i == num_preds - 1
        with torch.cuda.stream(side), ops.sm_reserve(0 if last else SM_RESERVE):
            recon, (recons, masks) = savi.decode(nxt.reshape(B, S, D))
        if not capturing:
            for t in (nxt,):
                t.record_stream(side)
        frames.append((recon, recons, masks))

    dyn.step_hook = decode_step
    try:
        pred_slots, actions = predictor.forward_multiple_samples(
            slot_history=slot_history, num_samples=1, num_seed=num_seed, num_preds=num_preds, use_posterior=True)
    finally:
        dyn.step_hook = None
    main.wait_stream(side)
    res = {"slot_history": slot_history, "pred_slots": pred_slots, "actions": actions}
    for key, j in (("pred_imgs", 0), ("pred_objs", 1), ("pred_masks", 2)):
        parts = [f[j] for f in frames]
        if not capturing:
            for t in parts:
                t.record_stream(main)
        res[key] = torch.stack(parts, dim=1).unsqueeze(1)        # (B, 1, P, ...)
    return res


_SIDE = {}


def _side_stream(device):
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=key)
    return _SIDE[key]


@torch.no_grad()
def seed_rollout_host(savi, predictor, videos_host, num_seed, num_preds, out_host=None, device="cuda", **kw):
    """
    Host-buffer entry point: `videos_host` is a (pinned) CPU tensor; the predicted frames come back in
    `out_host` (pinned CPU, (B, np, 3, H, W)).  Both copies are stream-ordered with the kernels.
    """
    v = videos_host.to(device, non_blocking=True)
    res = seed_rollout(savi, predictor, v, num_seed, num_preds, **kw)
    imgs = res["pred_imgs"][:, 0]
    if out_host is None:
        out_host = torch.empty(imgs.shape, dtype=imgs.dtype, pin_memory=True)
    out_host.copy_(imgs, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_host, res


class PinnedRing:
    """ two pinned host staging buffers for a small per-step upload (the InvDyn noise): the host may run several steps
    ahead of the device, so a buffer is rewritten only after the asynchronous copy that last read it has completed """

    def __init__(self, shape, dtype=torch.float32, depth=2):
        self.bufs = [torch.empty(shape, dtype=dtype, pin_memory=True) for _ in range(depth)]
        self.done = [None] * depth
        self.n = 0

    def upload(self, host_tensor, device_dst):
        i = self.n % len(self.bufs)
        self.n += 1
        if self.done[i] is not None:
            self.done[i].synchronize()
        self.bufs[i].copy_(host_tensor)
        device_dst.copy_(self.bufs[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.done[i] = ev


class GraphedSeedRollout:
    """
    `seed_rollout` for a fixed batch shape, captured once in a CUDA graph and replayed: the ~800 kernel launches of
    a step (15 sequential cOCVP steps x 8 layers on <= 48 tokens, 21 SlotAttention frames, ...) are latency-bound
    and a Python / ctypes launch loop cannot feed them fast enough.  Everything data-dependent on the host stays
    outside the graph: the video is copied into a static device buffer and the InvDyn noise is drawn on the host
    with the reference's own call (torch.randn on the CPU generator, LatentAction.py:73) and uploaded before each
    replay, so results are identical to the eager path under the same seed.
    """

    def __init__(self, savi, predictor, batch, num_seed, num_preds, resolution=(64, 64), num_samples=1, device="cuda", warmup=2,
                 video_dtype=torch.float32):
        """ video_dtype=torch.uint8: the clip arrives as 0..255 bytes (a quarter of the PCIe / HBM traffic of fp32 frames); the
        x / 255 normalisation of the reference's ToTensor runs inside the first encoder convolution """
        if num_samples != 1:
            raise NotImplementedError("the graphed path covers the posterior rollout (num_samples=1); prior samples draw "
                                      "random codebook indices on the host and use the eager path")
        self.savi, self.predictor = savi, predictor
        self.num_seed, self.num_preds = num_seed, num_preds
        T = num_seed + num_preds
        H, W = resolution
        la = predictor.latent_action
        S, A = savi.num_slots, la.action_dim
        per_slot = type(la).__name__ == "VQSimpleMLPSlotLatentAction"
        self.noise_shape = (batch, T - 1, S, A) if per_slot else (batch, T - 1, A)
        self.video = torch.zeros(batch, T, 3, H, W, device=device, dtype=video_dtype)
        self.noise = torch.zeros(self.noise_shape, device=device)
        self._noise_ring = PinnedRing(self.noise_shape)
        la._static_noise = self.noise
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):                       # allocates workspaces, packs weights, configures kernels
                    seed_rollout(savi, predictor, self.video, num_seed, num_preds)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            from . import ops
            n0 = ops.launch_count()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = seed_rollout(savi, predictor, self.video, num_seed, num_preds)
            self.launches_per_replay = ops.launch_count() - n0        # playslot_b200 kernels inside the graph
        finally:
            la._static_noise = None
        self._la = la

    def draw_noise(self):
        """ the reference's RNG call for the Gaussian latent sample (host generator) """
        return torch.randn(self.noise_shape, dtype=torch.float32)

    @torch.no_grad()
    def __call__(self, videos, noise=None):
        """ videos: (B, T, 3, H, W) on the host (pinned) or on the device; returns the static output dict """
        self.video.copy_(videos, non_blocking=True)
        self._noise_ring.upload(self.draw_noise() if noise is None else noise, self.noise)
        self.graph.replay()
        return self.out

    @torch.no_grad()
    def run_host(self, videos_host, out_host):
        """ end-to-end: pinned host video in, predicted frames back in `out_host` (pinned) """
        res = self(videos_host)
        out_host.copy_(res["pred_imgs"][:, 0], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def host_pipeline(self):
        """ a HostPipeline over this graph: submit(video_host, out_host) per batch, wait() at the end """
        return HostPipeline(self)


class HostPipeline:
    """
    Streaming host <-> device path for evaluation / serving loops over a GraphedSeedRollout (what the reference's
    DataLoader + `.to(device)` + `.cpu()` loop does, 05_evaluate_PlaySlot.py:75-108): every batch is uploaded from pinned host
    memory and its predicted frames are downloaded into pinned host memory, but the PCIe copies run on their own streams
    through staging buffers, so the upload of batch i+1 and the download of batch i-1 overlap the kernels of batch i.

        pipe = graphed.host_pipeline()
        for video_host, out_host in batches:        # pinned tensors
            pipe.submit(video_host, out_host)       # returns at once
        pipe.wait()                                 # every out_host is complete
    """

    def __init__(self, graphed):
        self.g = graphed
        self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
        self.video_stage = torch.empty_like(graphed.video)
        self.noise_stage = torch.empty_like(graphed.noise)
        self.out_stage = torch.empty_like(graphed.out["pred_imgs"][:, 0])
        self._noise_pinned = [torch.empty(graphed.noise_shape, pin_memory=True) for _ in range(2)]
        self._noise_done = [None, None]
        self._stage_free = None         # main stream has copied the staging buffers into the graph's static inputs
        self._d2h_done = None           # the download of the previous batch has left out_stage
        self._n = 0

    @torch.no_grad()
    def submit(self, video_host, out_host, noise=None):
        g, main = self.g, torch.cuda.current_stream()
        b = self._n & 1
        if self._noise_done[b] is not None:
            self._noise_done[b].synchronize()                      # the upload two batches ago has read this pinned buffer
        self._noise_pinned[b].copy_(g.draw_noise() if noise is None else noise)
        with torch.cuda.stream(self.h2d):
            if self._stage_free is not None:
                self.h2d.wait_event(self._stage_free)
            self.video_stage.copy_(video_host, non_blocking=True)
            self.noise_stage.copy_(self._noise_pinned[b], non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(self.h2d)
        self._noise_done[b] = uploaded
        main.wait_event(uploaded)
        g.video.copy_(self.video_stage, non_blocking=True)         # device-to-device, microseconds
        g.noise.copy_(self.noise_stage, non_blocking=True)
        self._stage_free = torch.cuda.Event()
        self._stage_free.record(main)
        g.graph.replay()
        if self._d2h_done is not None:
            main.wait_event(self._d2h_done)
        self.out_stage.copy_(g.out["pred_imgs"][:, 0], non_blocking=True)
        computed = torch.cuda.Event()
        computed.record(main)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(computed)
            out_host.copy_(self.out_stage, non_blocking=True)
            self._d2h_done = torch.cuda.Event()
            self._d2h_done.record(self.d2h)
        self._n += 1

    def wait(self):
        self.h2d.synchronize()
        torch.cuda.current_stream().synchronize()
        self.d2h.synchronize()


class ClosedLoopStepper:
    """
    Latent imagination, one frame at a time -- the loop of src/11_evaluate_behavior_on_simulation.py:128-183 around an external
    behaviour model:

        stepper = ClosedLoopStepper(savi, predictor, batch)
        first = stepper.reset(init_img)                      # savi(init_img.unsqueeze(1), num_imgs=1, decode=True)
        for i in range(n):
            latent = policy(stepper.slots[:, -1:])[:, 0]     # (B, A) or (B, S, A), by the caller
            out = stepper.step(latent)                       # decompose_action_latent -> enforce_window -> forward_single -> decode

    Single-step latency is what matters for planning: with graphs=True every window length (1 .. context_length frames) gets
    its own CUDA graph (captured lazily at first use) of decompose + cOCVP step + decode of the predicted frame, so a step is
    two small device copies and one graph replay.  Same kernels as the eager calls: results are bit-identical.
    """

    def __init__(self, savi, predictor, batch, graphs=True):
        self.savi, self.predictor = savi, predictor
        self.inv, self.dyn = predictor.latent_action, predictor.dynamics_model
        self.B, self.S, self.D = batch, savi.num_slots, savi.slot_dim
        self.A = self.inv.action_dim
        self.ctx = self.dyn.context_length
        self.per_slot = type(self.inv).__name__ == "VQSimpleMLPSlotLatentAction"
        self.graphs_enabled = graphs
        self._graphs = {}
        self.slots = None
        self._protos, self._vars = [], []

    @torch.no_grad()
    def reset(self, init_img, **init_kwargs):
        """ init_img (B, 3, H, W) -> dict(slots (B, 1, S, D), recons_imgs, recons_objs, masks) of the first frame """
        out = self.savi(init_img.unsqueeze(1), num_imgs=1, decode=True, **init_kwargs)
        self.slots = out["slot_history"]
        self._protos, self._vars = [], []
        return {"slots": self.slots, "recons_imgs": out["recons_imgs"][:, 0], "recons_objs": out["recons_objs"][:, 0], "masks": out["masks"][:, 0]}

    def _body(self, latent, slots_win, protos_prev, vars_prev):
        """ decompose the new latent action, append it to the (window-cropped) history, predict and decode the next frame """
        proto, var = self.inv.decompose_action_latent(latent)
        protos = torch.cat([protos_prev, proto.unsqueeze(1)], dim=1) if protos_prev is not None else proto.unsqueeze(1)
        vars_ = torch.cat([vars_prev, var.unsqueeze(1)], dim=1) if vars_prev is not None else var.unsqueeze(1)
        cp, cv = protos, vars_
        if cp.dim() == 3:                                    # single-action model: one action per frame, shared by the slots
            cp = cp.unsqueeze(2).expand(-1, -1, self.S, -1)
            cv = cv.unsqueeze(2).expand(-1, -1, self.S, -1)
        nxt = self.dyn.forward_single(slots=slots_win, action_protos=cp, action_vars=cv)[:, -1:]
        img, (objs, masks) = self.savi.decode(nxt[:, 0])
        return proto, var, nxt, img, objs, masks

    @torch.no_grad()
    def step(self, latent_action):
        """ latent_action (B, A) [or (B, S, A) for the per-slot model] -> dict(pred_slots (B, 1, S, D), pred_img, pred_objs, pred_masks) """
        assert self.slots is not None, "call reset(init_img) first"
        # window bookkeeping as upstream: keep the last `ctx` frames of slots and of the action history (incl. the new action)
        slots_win = self.slots[:, -self.ctx:]
        keep = self.ctx - 1
        protos_prev = torch.stack(self._protos[-keep:], dim=1) if (self._protos and keep > 0) else None
        vars_prev = torch.stack(self._vars[-keep:], dim=1) if (self._vars and keep > 0) else None
        # the reference crops slots and actions to the same length: frames and actions pair up from the newest backwards
        Fr = slots_win.shape[1]
        if protos_prev is not None and protos_prev.shape[1] > Fr - 1:
            protos_prev, vars_prev = protos_prev[:, -(Fr - 1):] if Fr > 1 else None, vars_prev[:, -(Fr - 1):] if Fr > 1 else None
        latent_action = latent_action.contiguous().float()
        if not self.graphs_enabled:
            proto, var, nxt, img, objs, masks = self._body(latent_action, slots_win.contiguous(), protos_prev, vars_prev)
        else:
            g = self._graphs.get(Fr)
            if g is None:
                g = self._capture(Fr, latent_action, slots_win, protos_prev, vars_prev)
                self._graphs[Fr] = g
            g["latent"].copy_(latent_action)
            g["slots"].copy_(slots_win)
            if protos_prev is not None:
                g["protos"].copy_(protos_prev)
                g["vars"].copy_(vars_prev)
            g["graph"].replay()
            proto, var, nxt, img, objs, masks = (t.clone() for t in g["out"])
        self._protos.append(proto)
        self._vars.append(var)
        self.slots = torch.cat([self.slots, nxt], dim=1)
        return {"pred_slots": nxt, "pred_img": img, "pred_objs": objs, "pred_masks": masks, "action_proto": proto, "action_var": var}

    def _capture(self, Fr, latent, slots_win, protos_prev, vars_prev):
        st = {"latent": latent.clone(), "slots": slots_win.contiguous().clone(),
              "protos": None if protos_prev is None else protos_prev.clone(), "vars": None if vars_prev is None else vars_prev.clone()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body(st["latent"], st["slots"], st["protos"], st["vars"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["out"] = self._body(st["latent"], st["slots"], st["protos"], st["vars"])
        st["graph"] = graph
        return st
