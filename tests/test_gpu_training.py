"""
GPU parity tests of the training path (pytest -m gpu): backward kernels against torch autograd in fp64, the decoder
data gradient against the differentiable CPU oracle, and one full training-step forward + backward against vectors
produced by the real reference in train mode (oracle/make_golden_train.py).
Tolerances: forward quantities 1e-4 (as everywhere); gradients 2e-4 rel-L2 / max-rel against fp64 autograd, 5e-4 against
the reference's own fp32 backward (two fp32 evaluations of a 13-step BPTT chain).
"""
import copy
import os

import pytest
import torch
import torch.nn.functional as F

from golden_util import GOLDEN_DIR
from build_util import make_savi, make_predictor, FixedRandn
from synth import synth_state_dict, synth_video, synth_noise, rel_l2, max_rel
from oracle import playslot_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def check(name, got, ref, tol):
    got, ref = got.detach().cpu(), ref.detach().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    r, m = rel_l2(got, ref), max_rel(got, ref)
    assert r <= tol and m <= tol, f"{name}: rel_l2={r:.3e} max_rel={m:.3e} (tol {tol:.0e})"


def g(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("M,N,K", [(37, 48, 20), (256, 256, 128), (1000, 1024, 256), (416, 16, 256)])
@pytest.mark.parametrize("mode", ["plain", "ln_gelu", "relu_res", "abs"])
def test_linear_backward(M, N, K, mode):
    from playslot_b200 import autograd as AG, ops
    x, w, b = g(M, K, seed=1), g(N, K, seed=2, scale=K ** -0.5), g(N, seed=3, scale=0.1)
    gam, bet, res = 1 + 0.1 * g(K, seed=4), 0.1 * g(K, seed=5), g(M, N, seed=6)
    gy = g(M, N, seed=7)
    leaves = [t.clone().to(DEV).requires_grad_(True) for t in (x, w, b, gam, bet, res)]
    xd, wd, bd, gd, btd, rd = leaves
    ref_leaves = [t.clone().double().requires_grad_(True) for t in (x, w, b, gam, bet, res)]
    xr, wr, br, gr, btr, rr = ref_leaves
    if mode == "plain":
        y = AG.linear(xd, wd, bd)
        yr = F.linear(xr, wr, br)
    elif mode == "ln_gelu":
        y = AG.linear(xd, wd, bd, ln=(gd, btd, 1e-6), act=ops.ACT_GELU)
        yr = F.gelu(F.linear(F.layer_norm(xr, (K,), gr, btr, 1e-6), wr, br))
    elif mode == "relu_res":
        y = AG.linear(xd, wd, None, act=ops.ACT_RELU, res=rd)
        yr = F.relu(F.linear(xr, wr)) + rr
    else:
        y = AG.linear(xd, wd, bd, act=ops.ACT_ABS)
        yr = F.linear(xr, wr, br).abs()
    check("y", y, yr.float(), 1e-5)
    y.backward(gy.to(DEV))
    yr.backward(gy.double())
    for name, a, r in zip(("dx", "dw", "db", "dgamma", "dbeta", "dres"), leaves, ref_leaves):
        if r.grad is None:
            assert a.grad is None or float(a.grad.abs().max()) == 0.0, name
            continue
        check(f"{name}[{mode}]", a.grad, r.grad.float(), 2e-5)


@pytest.mark.parametrize("n_seq,L,H,dh,spf", [(3, 9, 4, 64, 0), (2, 64, 8, 64, 8), (5, 5, 4, 32, 0), (2, 32, 8, 64, 4), (1, 96, 8, 64, 12)])
def test_mha_backward(n_seq, L, H, dh, spf):
    from playslot_b200 import autograd as AG
    inner = H * dh
    qkv = g(n_seq * L, 3 * inner, seed=L)
    gy = g(n_seq * L, inner, seed=L + 1)
    qd = qkv.clone().to(DEV).requires_grad_(True)
    out = AG.mha(qd, n_seq, L, H, dh, spf)
    out.backward(gy.to(DEV))
    qr = qkv.clone().double().requires_grad_(True)
    q, k, v = [t.view(n_seq, L, H, dh).transpose(1, 2) for t in qr.split(inner, dim=1)]
    dots = q @ k.transpose(-1, -2) * dh ** -0.5
    if spf:
        mask = O.block_causal_mask(L // spf, spf)
        dots = dots.masked_fill(mask[None, None] == 0, -1e9)
    ref = (torch.softmax(dots, -1) @ v).transpose(1, 2).reshape(n_seq * L, inner)
    ref.backward(gy.double())
    check("mha out", out, ref.float(), 1e-5)
    check("mha dqkv", qd.grad, qr.grad.float(), 2e-5)


@pytest.mark.parametrize("per_slot,R", [(False, 1), (True, 5)])
def test_invdyn_tail_backward(per_slot, R):
    from playslot_b200 import autograd as AG
    B, T, A, K = 3, 6, 16, 8
    mv = g(B, T, R, 2 * A, seed=1)
    noise = g(B, T - 1, R, A, seed=2)
    cb = g(K, A, seed=3, scale=0.6)
    gw = [g(B, T - 1, 2, R, A, seed=4), g(B, T - 1, R, A, seed=5), g(B, T - 1, R, A, seed=6), g(B, T - 1, R, A, seed=7), g(B, T - 1, R, seed=8)]
    mvd = mv.clone().to(DEV).requires_grad_(True)
    cbd = cb.clone().to(DEV).requires_grad_(True)
    out = AG.invdyn_tail(mvd, noise.to(DEV), cbd, per_slot)
    loss = sum((o * w.to(DEV)).sum() for o, w in zip((out["action_dist"], out["z"], out["proto"], out["variability"], out["quant"]), gw))
    loss = loss + 0.25 * out["commit"].sum()
    loss.backward()
    # fp64 autograd restatement with the reference's stop-gradients (quantization.py:294-322, LatentAction.py:236-239,344-348)
    mvr = mv.clone().double().requires_grad_(True)
    cbr = cb.clone().double().requires_grad_(True)
    mean, var = mvr[..., :A], mvr[..., A:].abs()
    m, v = mean[:, 1:] - mean[:, :-1], var[:, 1:] + var[:, :-1]
    z = noise.double() * torch.sqrt(v + 1e-6) + m
    zq, idx, _ = O.vq_lookup(cbr, z)
    proto = z + (zq - z).detach()
    # InvDynS: get_variability re-runs forward_one (quantization.py:365-368), z_q = onehot @ E stays differentiable w.r.t. E
    variab = (z - proto) if per_slot else (z - zq)
    quant = ((zq.detach() - z) ** 2).mean(-1)
    commit = ((zq - z.detach()) ** 2).mean(-1)
    dist = torch.stack([m, v], dim=2)
    lr = sum((o * w.double()).sum() for o, w in zip((dist, z, proto, variab, quant), gw)) + 0.25 * commit.sum()
    lr.backward()
    assert torch.equal(out["idx"].cpu().reshape(-1), idx.reshape(-1))
    check("d_mv", mvd.grad, mvr.grad.float(), 2e-5)
    check("d_codebook", cbd.grad, cbr.grad.float(), 2e-5)


def _load_train_case(name):
    fx = torch.load(os.path.join(GOLDEN_DIR, f"train_{name}.pt"), weights_only=False)
    ssd = synth_state_dict(fx["savi_manifest"], fx["weight_seed"])
    psd = synth_state_dict(fx["predictor_manifest"], fx["weight_seed"] + 100)
    scfg, pcfg = fx["savi_cfg"], fx["predictor_cfg"]
    B, T = fx["B"], fx["num_seed"] + fx["num_preds"]
    H, W = scfg["encoder"]["encoder_params"]["resolution"]
    S, D, A = scfg["num_slots"], scfg["slot_dim"], pcfg["action_dim"]
    video = synth_video(B, T, H, W, fx["weight_seed"])
    per_slot = pcfg["LatentAction"]["model_name"] == "VQSimpleMLPSlotLatentAction"
    noise = synth_noise((B, T - 1, S, A) if per_slot else (B, T - 1, A), fx["weight_seed"])
    return fx, ssd, psd, video, noise


def _unpack_split(act, n, H, W):
    full = act.view(torch.float16).float().view(2, n, H + 4, W + 4, 64)
    return (full[0] + full[1] / 2048.0)[:, 2:H + 2, 2:W + 2].permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("gscale", [1.0, 1e-7])
@pytest.mark.parametrize("case", ["buttonpress", "gridshapes5"])
def test_decode_data_gradient(case, gscale):
    """
    d(recon, recons, masks)/d(slots) of SAVi.decode.  A ReLU network's gradient is discontinuous in the activation
    signs, and two forwards of different rounding disagree on the sign of a few of the 4.7 M pre-activations (each
    flip changes the gradient by O(1e-3) relative, measured).  The exact check therefore evaluates the decoder in
    fp64 with the ReLU masks of OUR forward frozen in (kernel arithmetic, tolerance 1e-4); the plain differentiable
    oracle is checked at 1e-2.  gscale = 1e-7 exercises loss-sized gradients (the power-of-two gradient scaling).
    """
    fx, ssd, psd, video, noise = _load_train_case(case)
    scfg = fx["savi_cfg"]
    savi = make_savi(scfg, ssd, DEV)
    for p in savi.parameters():
        p.requires_grad_(False)
    S, D, H, W = scfg["num_slots"], scfg["slot_dim"], 64, 64
    M = 3
    slots = g(M, S, D, seed=11)
    w_recon, w_recons, w_masks = g(M, 3, H, W, seed=12) * gscale, g(M, S, 3, H, W, seed=13, scale=0.1) * gscale, g(M, S, 1, H, W, seed=14) * gscale
    sd_ = slots.clone().to(DEV).requires_grad_(True)
    recon, (recons, masks) = savi.decode(sd_)
    ((recon * w_recon.to(DEV)).sum() + (recons * w_recons.to(DEV)).sum() + (masks * w_masks.to(DEV)).sum()).backward()

    # fp64 evaluation with our forward's ReLU masks
    with torch.no_grad():
        _, _, _, saved = savi._decode_impl(slots.to(DEV), keep=True)
    _, _, acts, h4 = saved[0]
    if isinstance(h4, tuple):                   # ("planes", ...): the last block's activation is kept in the split-plane format too
        h4 = _unpack_split(h4[1], M * S, H, W)
    mk = [(_unpack_split(a, M * S, H, W) > 0).double() for a in acts] + [(h4 > 0).double()]
    sd64 = {k: v.double().to(DEV) for k, v in ssd.items()}
    sr = slots.clone().double().to(DEV).requires_grad_(True)
    pos = savi.decoder_pos_embedding.table(channels_last=False).double()
    h = sr.reshape(-1, D, 1, 1).expand(-1, D, H, W) + pos[None]
    for li, i in enumerate((0, 2, 4, 6)):
        h = F.conv2d(h, sd64[f"decoder.decoder.{i}.block.0.weight"], sd64[f"decoder.decoder.{i}.block.0.bias"], padding=2) * mk[li]
    o = F.conv2d(h, sd64["decoder.decoder.7.weight"], sd64["decoder.decoder.7.bias"], padding=1).view(M, S, 4, H, W)
    r_recons, r_masks = o[:, :, :3], torch.softmax(o[:, :, 3:], dim=1)
    r_recon = (r_recons * r_masks).sum(1)
    ((r_recon * w_recon.double().to(DEV)).sum() + (r_recons * w_recons.double().to(DEV)).sum() + (r_masks * w_masks.double().to(DEV)).sum()).backward()
    check("recon", recon, r_recon.float(), 1e-4)
    check("d_slots (frozen ReLU masks, fp64)", sd_.grad, sr.grad.float(), 1e-4)

    # plain differentiable oracle: agreement up to the ReLU sign flips
    so = slots.clone().double().requires_grad_(True)
    ssd64 = {k: v.double() for k, v in ssd.items()}
    o_recon, o_recons, o_masks = O.decode(ssd64, scfg, so)
    ((o_recon * w_recon.double()).sum() + (o_recons * w_recons.double()).sum() + (o_masks * w_masks.double()).sum()).backward()
    check("d_slots (oracle)", sd_.grad, so.grad.float(), 1e-2)

    # recon only (the training loss): the other two outputs receive no gradient at all
    sd2 = slots.clone().to(DEV).requires_grad_(True)
    recon2, _ = savi.decode(sd2)
    (recon2 * w_recon.to(DEV)).sum().backward()
    so2 = slots.clone().double().requires_grad_(True)
    (O.decode(ssd64, scfg, so2)[0] * w_recon.double()).sum().backward()
    check("d_slots (recon only, oracle)", sd2.grad, so2.grad.float(), 1e-2)


@pytest.fixture
def train_linear_mode_guard():
    from playslot_b200 import autograd as AG
    old = AG.TRAIN_LINEAR_MODE
    yield AG
    AG.TRAIN_LINEAR_MODE = old


@pytest.mark.parametrize("linear_mode", ["fp32", "fp16x3"])
@pytest.mark.parametrize("case", ["buttonpress", "gridshapes5"])
def test_training_step_matches_reference_golden(case, linear_mode, train_linear_mode_guard):
    """ forward + backward of src/04_train_PlaySlot.py:32-122 in train mode against the real reference; "fp16x3" = forward and
    data-gradient GEMMs of the Linear layers on the tensor cores """
    from playslot_b200 import train
    train_linear_mode_guard.TRAIN_LINEAR_MODE = linear_mode
    fx, ssd, psd, video, noise = _load_train_case(case)
    scfg, pcfg = fx["savi_cfg"], fx["predictor_cfg"]
    ns, npred, B = fx["num_seed"], fx["num_preds"], fx["B"]
    savi = make_savi(scfg, ssd, DEV)
    for p in savi.parameters():
        p.requires_grad_(False)
    pred = make_predictor(fx["predictor_name"], pcfg, psd, DEV).train()
    sh = fx["slot_history"].to(DEV)
    with FixedRandn(noise):
        pred_slots, out = pred(slots=sh, num_seed=ns, num_preds=npred)
    assert torch.equal(out["action_idxs"].cpu().reshape(-1), fx["action_idxs"].reshape(-1)), "VQ indices differ"
    check("pred_slots", pred_slots, fx["pred_slots"], 1e-4)
    pred_slots.retain_grad()
    S, D = pred_slots.shape[-2:]
    n = pred_slots.shape[1]
    pred_imgs, _ = savi.decode(pred_slots.reshape(B * n, S, D))
    pred_imgs = pred_imgs.view(B, n, *pred_imgs.shape[1:])
    T = ns + npred
    loss, terms = train.predictor_losses(pred_slots, sh[:, 1:T], pred_imgs, video[:, 1:T].to(DEV), out["vq_losses"])
    for k, ref in (("pred_img_mse", fx["pred_img_mse"]), ("pred_slot_mse", fx["pred_slot_mse"]), ("quant", fx["quant"])):
        assert abs(float(terms[k]) - ref) <= 1e-4 * abs(ref), f"{k}: {float(terms[k])} vs {ref}"
    assert abs(float(loss.detach()) - fx["loss"]) <= 1e-4 * abs(fx["loss"])
    loss.backward()
    check("d loss / d pred_slots", pred_slots.grad, fx["d_pred_slots"], 5e-4)
    named = dict(pred.named_parameters())
    worst = 0.0
    floor = 2e-6 * max(v for v in fx["grad_norms"].values() if v)      # analytically-zero gradients (e.g. mean_fc.bias) are rounding noise
    for k, ref_norm in fx["grad_norms"].items():
        if ref_norm is None:
            assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, f"{k}: unexpected gradient"
            continue
        assert named[k].grad is not None, f"{k}: missing gradient"
        got = float(named[k].grad.double().norm())
        worst = max(worst, abs(got - ref_norm) / max(ref_norm, 1e-12))
        assert abs(got - ref_norm) <= 5e-4 * ref_norm + floor, f"{k}: grad norm {got} vs {ref_norm}"
    for k, ref in fx["grads_full"].items():
        if fx["grad_norms"][k] > 100 * floor:
            check(f"grad {k}", named[k].grad, ref, 5e-4)
    sd_after = pred.state_dict()
    for k, ref in fx["buffers_after"].items():
        check(f"buffer {k} after the EMA update", sd_after[k], ref, 1e-5)


def test_train_step_runs_and_learns():
    """ three optimisation steps of playslot_b200.train.train_step on a fixed batch reduce the loss """
    from playslot_b200 import train
    fx, ssd, psd, video, noise = _load_train_case("buttonpress")
    savi = make_savi(fx["savi_cfg"], ssd, DEV)
    for p in savi.parameters():
        p.requires_grad_(False)
    pred = make_predictor(fx["predictor_name"], fx["predictor_cfg"], psd, DEV)
    opt = train.make_optimizer(pred, lr=1e-4)
    bucket = train.GradientBucket(pred.parameters())
    losses = []
    for _ in range(3):
        with FixedRandn(noise):
            l, _ = train.train_step(savi, pred, opt, video.to(DEV), fx["num_seed"], fx["num_preds"], bucket=bucket, clip_gradient=0.05)
        losses.append(float(l))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0], losses


def test_graphed_train_step_equals_eager():
    """ train.GraphedTrainStep (one CUDA graph per optimisation step) follows the eager train_step: same losses and the same
    parameters after three steps, up to the summation-order noise of the atomics in the weight-gradient kernels """
    import copy
    from playslot_b200 import configs, train
    cfg = configs.experiment("ButtonPress")
    torch.manual_seed(0)
    savi, pred_a = configs.build_models(cfg, DEV)
    for p in savi.parameters():
        p.requires_grad_(False)
    pred_b = copy.deepcopy(pred_a)
    B, ns, npred = 2, 3, 3
    gen = torch.Generator().manual_seed(11)
    vids = [torch.rand(B, ns + npred, 3, 64, 64, generator=gen).to(DEV) for _ in range(3)]
    opt_b = train.make_optimizer(pred_b, capturable=True)
    graphed = train.GraphedTrainStep(savi, pred_b, opt_b, B, ns, npred)
    noises = [torch.randn(graphed.noise_shape, generator=gen) for _ in range(3)]
    opt_a = train.make_optimizer(pred_a)
    la = pred_a.latent_action
    losses_a, losses_b = [], []
    for v, n in zip(vids, noises):
        la._static_noise = n.to(DEV)
        try:
            loss, _ = train.train_step(savi, pred_a, opt_a, v, ns, npred)
        finally:
            la._static_noise = None
        losses_a.append(float(loss))
        losses_b.append(float(graphed(v, noise=n)[0]))
    for i, (a, b) in enumerate(zip(losses_a, losses_b)):
        assert abs(a - b) <= 2e-5 * abs(a), f"step {i}: eager loss {a} vs graphed {b}"
    for (name, pa), pb in zip(pred_a.named_parameters(), pred_b.parameters()):
        if pa.requires_grad:
            assert float((pa.detach() - pb.detach()).abs().max()) <= 2e-5 * max(1.0, float(pa.detach().abs().max())), name
    for (name, ba), bb in zip(pred_a.named_buffers(), pred_b.buffers()):
        assert float((ba.float() - bb.float()).abs().max()) <= 2e-5 * max(1.0, float(ba.float().abs().max())), name


def test_eager_forward_after_graphed_steps_sees_new_weights():
    """ GraphedTrainStep replays update the parameters on the device without bumping their Python-side `_version`; every
    derived-weight cache (packed tensor-core copies, concatenated q/k/v / embedding / head matrices) must still be rebuilt:
    an eager no_grad forward after the replays equals the forward of a freshly deep-copied model """
    import copy
    from playslot_b200 import configs, train
    cfg = configs.experiment("ButtonPress")
    torch.manual_seed(0)
    savi, pred = configs.build_models(cfg, DEV)
    for p in savi.parameters():
        p.requires_grad_(False)
    B, ns, npred = 2, 3, 3
    gen = torch.Generator().manual_seed(12)
    sh = torch.randn(B, ns + npred, 4, 128, generator=gen).to(DEV)
    noise = torch.randn(B, ns + npred - 1, cfg["predictor"]["predictor_params"]["action_dim"], generator=gen)
    pred.eval()
    with torch.no_grad(), FixedRandnLocal(noise):
        before, _ = pred.forward_multiple_samples(sh, 1, ns, npred, use_posterior=True)      # fills every cache
    opt = train.make_optimizer(pred, lr=1e-2, capturable=True)
    graphed = train.GraphedTrainStep(savi, pred, opt, B, ns, npred, clip_gradient=None)
    for _ in range(3):
        graphed(torch.rand(B, ns + npred, 3, 64, 64, generator=gen).to(DEV))
    torch.cuda.synchronize()
    pred.eval()
    fresh = copy.deepcopy(pred)                                   # new tensor objects: no cache can survive the copy
    with torch.no_grad(), FixedRandnLocal(noise):
        after, _ = pred.forward_multiple_samples(sh, 1, ns, npred, use_posterior=True)
    with torch.no_grad(), FixedRandnLocal(noise):
        want, _ = fresh.forward_multiple_samples(sh, 1, ns, npred, use_posterior=True)
    assert float((after - before).abs().max()) > 1e-4, "three optimiser steps at lr 1e-2 must change the prediction"
    assert torch.equal(after, want), "eager inference after graph replays ran on stale derived weights"


class FixedRandnLocal:
    def __init__(self, value):
        self.value = value

    def __enter__(self):
        self._orig = torch.randn
        torch.randn = lambda *a, **k: self.value.clone().to(k["device"]) if k.get("device") is not None else self.value.clone()
        return self

    def __exit__(self, *exc):
        torch.randn = self._orig
        return False


def test_training_step_bf16_mode_stated_tolerance(train_linear_mode_guard):
    """ the 16-bit tensor-core training mode of BASELINE.json configs[3] (Linear forward / data-gradient GEMMs: one bf16 product,
    fp32 accumulate, fp32 master weights, fp32 weight gradients and InvDyn tail; decoder convolutions: one fp16 product) against
    the fp32 training step (itself pinned to the reference by test_training_step_matches_reference_golden), on a batch large enough
    for every Linear to take the tensor-core path (>= 128 rows).  Stated tolerances: loss 2e-2, gradient norms 1e-1, gradient
    direction (cosine over all parameters) >= 0.98 """
    from playslot_b200 import configs, train
    import copy
    cfg = configs.experiment("ButtonPress")
    torch.manual_seed(0)
    savi, pred_a = configs.build_models(cfg, DEV)
    for p in savi.parameters():
        p.requires_grad_(False)
    pred_b = copy.deepcopy(pred_a)
    B, ns, npred = 16, 6, 4
    gen = torch.Generator().manual_seed(21)
    video = torch.rand(B, ns + npred, 3, 64, 64, generator=gen).to(DEV)
    noise = torch.randn(B, ns + npred - 1, cfg["predictor"]["predictor_params"]["action_dim"], generator=gen)
    res = {}
    for name, pred, lin, conv in (("fp32", pred_a, "fp32", "fp16x3"), ("bf16", pred_b, "bf16", "fp16")):
        train_linear_mode_guard.TRAIN_LINEAR_MODE = lin
        savi.conv_mode = conv
        pred.train()
        with FixedRandn(noise):
            loss, terms, _ = train.forward_loss(savi, pred, video, ns, npred)
        loss.backward()
        res[name] = (float(loss), {k: p.grad.detach().double().flatten() for k, p in pred.named_parameters() if p.grad is not None})
    la, ga = res["fp32"]
    lb, gb = res["bf16"]
    assert abs(la - lb) <= 2e-2 * abs(la), f"loss {lb} (bf16 mode) vs {la} (fp32)"
    va, vb = torch.cat([ga[k] for k in ga]), torch.cat([gb[k] for k in ga])
    cos = float((va @ vb) / (va.norm() * vb.norm()))
    worst = max(abs(float(gb[k].norm() - ga[k].norm())) / float(ga[k].norm()) for k in ga if float(ga[k].norm()) > 1e-3 * float(va.norm()))
    print(f"bf16 training mode: loss {lb:.6f} vs {la:.6f}, gradient cosine {cos:.5f}, worst gradient-norm error {worst:.3e}")
    assert cos >= 0.98 and worst <= 1e-1, (cos, worst)
