"""
CPU oracle for the PlaySlot per-frame slot pipeline  --  TEST INFRASTRUCTURE ONLY.

This file is a functional, stateless restatement (torch fp32/fp64 on the CPU) of
the arithmetic the reference performs on the hot path (SURVEY.md section 8a,
Appendix B).  It is NOT the product: only `tests/`, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it, and
only as the checker / the CPU arm.  Nothing under `playslot_b200/` imports it.

Parity pinning: every function below is checked against outputs of the real
reference modules (imported from /root/reference in the build container by
`oracle/make_golden.py`); the resulting vectors are committed under
`tests/golden/` and `tests/test_oracle_golden.py` re-checks the oracle against
them on any machine.  Status: PINNED ("Oracle").

All functions take a flat `sd` dict {reference state_dict key -> tensor} plus a
key prefix, so the oracle consumes exactly the checkpoint layout of the
reference (`SAVi.state_dict()`, predictor `state_dict()`).

Reference citations are `src/models/...:line` in /root/reference.
"""

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# small helpers
# --------------------------------------------------------------------------- #

def _lin(sd, p, x, bias=True):
    """ y = x W^T (+ b) with W = sd[p + '.weight'] (nn.Linear layout (out, in)). """
    w = sd[p + ".weight"]
    b = sd[p + ".bias"] if bias and (p + ".bias") in sd else None
    return F.linear(x, w, b)


def _ln(sd, p, x, eps):
    """ LayerNorm over the last axis with affine params sd[p + '.weight'/'.bias']. """
    w = sd[p + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[p + ".bias"], eps)


def position_grid(H, W, dtype=torch.float32):
    """
    4-channel coordinate grid [Y, X, 1-Y, 1-X], each (H, W).
    Follows BlocksUtils/model_utils.py:12-34 (linspace(-1,1), meshgrid 'ij',
    concat with 1-grid; computed in float64 numpy then cast to float32).
    """
    ys = torch.linspace(-1.0, 1.0, H, dtype=torch.float64)
    xs = torch.linspace(-1.0, 1.0, W, dtype=torch.float64)
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    g = torch.stack([Y, X], dim=0).to(torch.float32)
    g = torch.cat([g, 1.0 - g], dim=0)           # (4, H, W)
    return g.to(dtype)


def soft_position_embedding(sd, p, H, W):
    """
    Weight-only positional embedding  Conv1x1_{4->C}(grid)  -> (C, H, W).
    BlocksUtils/model_blocks.py:85-119.
    """
    w = sd[p + ".projection.weight"]             # (C, 4, 1, 1)
    b = sd[p + ".projection.bias"]
    g = position_grid(H, W, w.dtype).to(w.device)
    return torch.einsum("ck,khw->chw", w[:, :, 0, 0], g) + b[:, None, None]


# --------------------------------------------------------------------------- #
# SAVi encoder  (SAVi.py:207-217, EncodersDecoders/encoders.py:77-104)
# --------------------------------------------------------------------------- #

def encode(sd, imgs, n_layers=4):
    """ imgs (M,3,H,W) -> features (M, H*W, D_enc). """
    h = imgs
    for i in range(n_layers):
        w = sd[f"encoder.encoder.{i}.block.0.weight"]
        b = sd[f"encoder.encoder.{i}.block.0.bias"]
        h = F.conv2d(h, w, b, stride=1, padding=w.shape[-1] // 2)
        if i < n_layers - 1:                     # no activation after the last conv
            h = F.relu(h)
    M, C, H, W = h.shape
    h = h.permute(0, 2, 3, 1) + soft_position_embedding(sd, "encoder_pos_embedding", H, W).permute(1, 2, 0)
    h = h.reshape(M, H * W, C)
    h = _ln(sd, "encoder_mlp.0", h, 1e-5)
    h = F.relu(_lin(sd, "encoder_mlp.1", h))
    return _lin(sd, "encoder_mlp.3", h)


# --------------------------------------------------------------------------- #
# SlotAttention  (BlocksUtils/attention.py:76-119)
# --------------------------------------------------------------------------- #

def slot_attention(sd, feats, slots, n_iters, p="slot_attention", eps=1e-8):
    """
    feats (B,N,Df), slots (B,S,D) -> (slots (B,S,D), attention masks (B,S,N)).
    The masks are the softmax-over-slots + eps tensor of the LAST iteration,
    before the renormalisation over locations (attention.py:107-109).
    """
    B, S, D = slots.shape
    x = _ln(sd, p + ".norm_input", feats, 1e-5)
    k = _lin(sd, p + ".to_k", x, bias=False)
    v = _lin(sd, p + ".to_v", x, bias=False)
    w_ih, w_hh = sd[p + ".gru.weight_ih"], sd[p + ".gru.weight_hh"]
    b_ih, b_hh = sd[p + ".gru.bias_ih"], sd[p + ".gru.bias_hh"]
    masks = None
    for _ in range(n_iters):
        h_prev = slots
        q = _lin(sd, p + ".to_q", _ln(sd, p + ".norm_slot", slots, 1e-5), bias=False)
        dots = torch.matmul(q, k.transpose(1, 2)) * (D ** -0.5)          # (B,S,N)
        attn = torch.softmax(dots, dim=1) + eps
        masks = attn
        attn = attn / attn.sum(dim=-1, keepdim=True)
        upd = torch.matmul(attn, v)                                        # (B,S,D)
        # GRUCell, gate order r,z,n (torch.nn.GRUCell)
        gi = F.linear(upd, w_ih, b_ih)
        gh = F.linear(h_prev, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, dim=-1)
        h_r, h_z, h_n = gh.chunk(3, dim=-1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        slots = (1.0 - z) * n + z * h_prev
        m = _ln(sd, p + ".norm_mlp", slots, 1e-5)
        slots = slots + _lin(sd, p + ".mlp.2", F.relu(_lin(sd, p + ".mlp.0", m)))
    return slots, masks


# --------------------------------------------------------------------------- #
# pre-LN transformer block  (attention.py:204-316, 425-475, 479-565)
# --------------------------------------------------------------------------- #

def block_causal_mask(n_frames, n_slots, dtype=torch.float32):
    """ M[i,j] = 1 iff frame(j) <= frame(i)   (model_utils.py:115-124). """
    f = torch.arange(n_frames * n_slots) // n_slots
    return (f[None, :] <= f[:, None]).to(dtype)


def transformer_block(sd, p, x, n_heads, head_dim, mask=None):
    """
    x (B,L,E).  y = x + MHSA(LN_1e-6(x)); out = y + MLP_gelu(LN_1e-6(y)).
    out_proj is applied iff its weights exist in `sd` (attention.py:177,195-198).
    Masked logits are filled with -1e9 (attention.py:235).
    """
    B, L, E = x.shape
    h = _ln(sd, p + ".ln_att", x, 1e-6)
    q = _lin(sd, p + ".attn.q", h, bias=False).view(B, L, n_heads, head_dim).transpose(1, 2)
    k = _lin(sd, p + ".attn.k", h, bias=False).view(B, L, n_heads, head_dim).transpose(1, 2)
    v = _lin(sd, p + ".attn.v", h, bias=False).view(B, L, n_heads, head_dim).transpose(1, 2)
    dots = torch.matmul(q, k.transpose(-1, -2)) * (head_dim ** -0.5)
    if mask is not None:
        dots = dots.masked_fill(mask[None, None] == 0, -1e9)
    a = torch.softmax(dots, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2).reshape(B, L, n_heads * head_dim)
    if (p + ".attn.out_proj.0.weight") in sd:
        o = _lin(sd, p + ".attn.out_proj.0", o)
    y = x + o
    m = _ln(sd, p + ".ln_mlp", y, 1e-6)
    m = _lin(sd, p + ".mlp.mlp.2", F.gelu(_lin(sd, p + ".mlp.mlp.0", m)))
    return y + m


# --------------------------------------------------------------------------- #
# SAVi recurrent loop  (SAVi.py:135-204)
# --------------------------------------------------------------------------- #

def initial_slots(sd, B, num_slots, init_noise=None):
    """
    Learned: repeat the learned slots (initializers.py:54-83).
    LearnedRandom: mu + sigma * noise (initializers.py:87-124); the noise has to
    be injected because the reference draws it on the module's device.
    """
    if "initializer.slots" in sd:
        return sd["initializer.slots"].repeat(B, 1, 1)
    mu, sigma = sd["initializer.slots_mu"], sd["initializer.slots_sigma"]
    D = mu.shape[-1]
    if init_noise is None:
        raise ValueError("LearnedRandom initializer: pass init_noise (B,S,D)")
    return mu.expand(B, num_slots, D) + sigma.expand(B, num_slots, D) * init_noise


def savi_encode_video(sd, cfg, video, num_imgs, init_noise=None, return_masks=False):
    """
    video (B,T,3,H,W) -> slot_history (B,num_imgs,S,D) [, last-frame masks (B,S,N)].
    cfg = SAVi `model_params` dict of the experiment JSON.
    """
    B = video.shape[0]
    S = cfg["num_slots"]
    it_first = cfg.get("num_iterations_first") or cfg["num_iterations"]
    it_rest = cfg["num_iterations"]
    tp = cfg["transition_module_params"]
    n_layers = len(cfg["encoder"]["encoder_params"]["num_channels"])
    pred = initial_slots(sd, B, S, init_noise)
    hist, masks = [], None
    for t in range(num_imgs):
        feats = encode(sd, video[:, t], n_layers)
        slots, masks = slot_attention(sd, feats, pred, it_first if t == 0 else it_rest)
        hist.append(slots)
        if tp.get("model_name") == "TransformerBlock":
            pred = transformer_block(sd, "transition_module", slots, tp["num_heads"], tp["head_dim"])
        else:
            pred = slots
    hist = torch.stack(hist, dim=1)
    return (hist, masks) if return_masks else hist


# --------------------------------------------------------------------------- #
# spatial-broadcast decoder + alpha compositing (SAVi.py:220-248, decoders.py:89-113)
# --------------------------------------------------------------------------- #

def decode(sd, cfg, slots):
    """
    slots (M,S,D) -> recon (M,3,H,W), recons (M,S,3,H,W), masks (M,S,1,H,W).
    `recons` is returned un-masked, as the reference does.
    """
    M, S, D = slots.shape
    H, W = cfg["decoder"]["decoder_params"]["resolution"]
    n_hidden = len(cfg["decoder"]["decoder_params"]["num_channels"])
    up = cfg["decoder"]["decoder_params"].get("upsample")
    pe = soft_position_embedding(sd, "decoder_pos_embedding", H, W)        # (D,H,W)
    h = slots.reshape(M * S, D, 1, 1) + pe[None]
    idx = 0
    for i in range(n_hidden):
        w = sd[f"decoder.decoder.{idx}.block.0.weight"]
        b = sd[f"decoder.decoder.{idx}.block.0.bias"]
        h = F.relu(F.conv2d(h, w, b, padding=w.shape[-1] // 2))
        idx += 1
        if isinstance(up, int) and i < n_hidden - 1:                      # Upsample module slot
            if up != 1:
                h = F.interpolate(h, scale_factor=up, mode="nearest")
            idx += 1
    w, b = sd[f"decoder.decoder.{idx}.weight"], sd[f"decoder.decoder.{idx}.bias"]
    o = F.conv2d(h, w, b, padding=1)
    o = o.reshape(M, S, o.shape[1], o.shape[2], o.shape[3])
    rgb, alpha = o[:, :, :-1], o[:, :, -1:]
    masks = torch.softmax(alpha, dim=1)
    recon = (rgb * masks).sum(dim=1)
    return recon, rgb, masks


# --------------------------------------------------------------------------- #
# VQ  (BlocksUtils/quantization.py:18-20, 256-322, 338-372)
# --------------------------------------------------------------------------- #

def vq_lookup(codebook, z):
    """
    z (...,A), codebook (K,A) -> (z_q exact codewords, idx int64 (...,1), dist (n,K)).
    dist = |z|^2 - 2 z C^T + |C|^2, argmin with first-index tie-break.
    """
    zf = z.reshape(-1, z.shape[-1])
    dist = zf.pow(2).sum(1, keepdim=True) - 2 * zf @ codebook.t() + codebook.t().pow(2).sum(0, keepdim=True)
    idx = torch.argmin(dist, dim=1)
    z_q = codebook[idx].view(z.shape)
    return z_q, idx.view(*z.shape[:-1], 1), dist


def vq_forward(codebook, z):
    """ returns (proto = z + (z_q - z), quant_loss, commit_loss, idx) per vector. """
    z_q, idx, _ = vq_lookup(codebook, z)
    quant = ((z_q - z) ** 2).mean(-1)
    commit = ((z_q - z) ** 2).mean(-1)
    proto = z + (z_q - z)                      # straight-through value (not bit-equal to z_q)
    return proto, quant, commit, idx


# --------------------------------------------------------------------------- #
# InvDyn  (Predictors/LatentAction.py)
# --------------------------------------------------------------------------- #

def _action_tail(sd, p, tokens, noise, per_slot):
    """
    mean/var heads, temporal difference, Gaussian sample, VQ, variability.
    LatentAction.py:72-87 (sample / comput_action_dist), :228-249, :340-359.
    tokens: (B,T,E) [InvDynS] or (B,T,S,E) [InvDynM]; noise matches (B,T-1[,S],A).
    """
    mean = _lin(sd, p + ".mean_fc", tokens)
    var = _lin(sd, p + ".variance_fc", tokens).abs()
    m = mean[:, 1:] - mean[:, :-1]
    v = var[:, 1:] + var[:, :-1]
    z = noise * torch.sqrt(v + 1e-6) + m
    codebook = sd[p + ".quantizer.vq.embeddings.0.weight"] if (p + ".quantizer.vq.embeddings.0.weight") in sd \
        else sd[p + ".quantizer.embeddings.0.weight"]
    proto, quant, commit, idx = vq_forward(codebook, z)
    if per_slot:
        variability = z - proto                # InvDynM: uses the STE protos (LatentAction.py:344-348)
    else:
        variability = z - codebook[idx.squeeze(-1)]   # InvDynS: re-looked-up exact codeword (:236-239)
    return {
        "action_dist": torch.stack([m, v], dim=2),
        "sampled_latent_action": z,
        "action_variability": variability,
        "action_proto": proto,
        "action_idxs": idx,
        "vq_losses": {"quant_loss": quant.mean(), "commit_loss": commit.mean()},
    }


def invdyn_single(sd, cfg, slots, noise, p="latent_action"):
    """
    InvDynS = VQSingleSlotLatentAction.forward (LatentAction.py:193-249).
    slots (B,T,S,D), noise (B,T-1,A).  cfg = LatentAction.model_params.
    """
    B, T, S, D = slots.shape
    e = _lin(sd, p + ".slot_encoder.1", _ln(sd, p + ".slot_encoder.0", slots, 1e-5))   # (B,T,S,E)
    E = e.shape[-1]
    act = sd[p + ".act_token"].expand(B, T, 1, E)
    tok = torch.cat([act, e], dim=2).reshape(B * T, S + 1, E)
    for l in range(cfg["num_layers"]):
        tok = transformer_block(sd, f"{p}.transformer.{l}", tok, cfg["num_heads"], cfg["head_dim"])
    out = tok[:, 0].reshape(B, T, E)
    return _action_tail(sd, p, out, noise, per_slot=False)


def invdyn_multi(sd, cfg, slots, noise, p="latent_action"):
    """
    InvDynM = VQSimpleMLPSlotLatentAction.forward (LatentAction.py:308-359).
    slots (B,T,S,D), noise (B,T-1,S,A).
    """
    h = F.relu(_lin(sd, p + ".slot_encoder.0", slots))
    e = _lin(sd, p + ".slot_encoder.3", _ln(sd, p + ".slot_encoder.2", h, 1e-5))
    return _action_tail(sd, p, e, noise, per_slot=True)


# --------------------------------------------------------------------------- #
# cOCVP = SlotGPTDymamicsModel (Predictors/DynamicsModels.py:321-550)
# --------------------------------------------------------------------------- #

def sinusoid_pe(max_len, E, dtype=torch.float32):
    """ model_blocks.py:178-190 """
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, E, 2) * (-math.log(10000.0) / E))
    pe = torch.zeros(max_len, E)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def cocvp_step(sd, cfg, slots, protos, variab=None, p="dynamics_model"):
    """
    forward_single (DynamicsModels.py:457-516): slots (B,F,S,D), protos/variab
    (B,F,S,A) -> predictions (B,F,S,D).  cfg = DynamicsModel.model_params.
    """
    B, Fr, S, D = slots.shape
    t = _lin(sd, p + ".tok_emb", slots) + _lin(sd, p + ".action_emb", protos)
    if variab is not None and (p + ".variability_emb.weight") in sd:
        t = t + _lin(sd, p + ".variability_emb", variab)
    pe = sd[p + ".pos_emb.pe"][0, :Fr]                   # (F,1,E)
    t = (t + pe[None]).reshape(B, Fr * S, -1)
    mask = block_causal_mask(Fr, S, t.dtype).to(t.device)
    for l in range(cfg["num_layers"]):
        t = transformer_block(sd, f"{p}.dynamics_model.{l}", t, cfg["num_heads"], cfg["head_dim"], mask)
    out = _lin(sd, p + ".regression_head.1", _ln(sd, p + ".regression_head.0", t, 1e-5))
    out = out.reshape(B, Fr, S, D)
    if cfg.get("residual", False):
        out = out + slots
    return out


def cocvp_rollout(sd, cfg, seed_slots, protos, variab, num_preds, p="dynamics_model"):
    """
    forward_autoregressive(get_pred_only=True) (DynamicsModels.py:396-454):
    seed_slots (B,ns,S,D); protos/variab (B,>=ns+np-1,S,A) -> (B,np,S,D).
    """
    ctx = cfg.get("context_length", 8)
    ns = seed_slots.shape[1]
    cur = seed_slots
    preds = []
    for i in range(num_preds):
        a = protos[:, :ns + i]
        v = variab[:, :ns + i] if variab is not None else None
        if cur.shape[1] > ctx:
            cur, a = cur[:, -ctx:], a[:, -ctx:]
            v = v[:, -ctx:] if v is not None else None
        nxt = cocvp_step(sd, cfg, cur, a, v, p)[:, -1:]
        preds.append(nxt)
        cur = torch.cat([cur, nxt], dim=1)
    return torch.cat(preds, dim=1)


def cocvp_train_unroll(sd, cfg, slots, protos, variab, num_seed, num_preds, p="dynamics_model"):
    """
    forward_autoregressive(get_pred_only=False): teacher-forced seed steps, then
    fed-back predictions; returns (B, ns+np-1, S, D)  (DynamicsModels.py:431-453).
    """
    ctx = cfg.get("context_length", 8)
    cur = slots[:, :1]
    preds = []
    for i in range(num_seed + num_preds - 1):
        a = protos[:, :i + 1]
        v = variab[:, :i + 1] if variab is not None else None
        if cur.shape[1] > ctx:
            cur, a = cur[:, -ctx:], a[:, -ctx:]
            v = v[:, -ctx:] if v is not None else None
        nxt = cocvp_step(sd, cfg, cur, a, v, p)[:, -1:]
        preds.append(nxt)
        cur = slots[:, :i + 2] if i < num_seed - 1 else torch.cat([cur, nxt], dim=1)
    return torch.cat(preds, dim=1)


# --------------------------------------------------------------------------- #
# OCVP-Seq + PredictorWrapper, action-conditional OCVP  
# --------------------------------------------------------------------------- #

def ocvp_seq(sd, cfg, inputs, p="predictor"):
    """
    OCVPSeq.forward (Predictors/transformers.py:231-340): inputs (B,F,S,D) -> (B,F,S,D).
    cfg = predictor_params (token_dim, hidden_dim, num_layers, n_heads, residual).
    """
    B, Fr, S, D = inputs.shape
    E, H = cfg["token_dim"], cfg["n_heads"]
    t = _lin(sd, p + ".mlp_in", inputs) + sd[p + ".pe.pe"][:, :Fr]
    for l in range(cfg["num_layers"]):
        o = transformer_block(sd, f"{p}.transformer_encoders.{l}.object_encoder_block", t.reshape(B * Fr, S, E), H, E // H)
        o = o.reshape(B, Fr, S, E).transpose(1, 2).reshape(B * S, Fr, E)
        o = transformer_block(sd, f"{p}.transformer_encoders.{l}.time_encoder_block", o, H, E // H)
        t = o.reshape(B, S, Fr, E).transpose(1, 2)
    out = _lin(sd, p + ".mlp_out", t)
    return out + inputs if cfg.get("residual", False) else out


def wrapper_rollout(step_fn, slot_history, num_context, num_preds, buffer_size, get_pred_only):
    """ PredictorWrapper.forward in eval mode (Predictors/predictor_wrappers.py:42-75): no teacher forcing """
    start = num_context if get_pred_only else 1
    n = num_preds if get_pred_only else num_context + num_preds - 1
    buf = slot_history[:, :start]
    preds = []
    for _ in range(n):
        cur = step_fn(buf)[:, -1]
        buf = torch.cat([buf, cur.unsqueeze(1)], dim=1)
        if buf.shape[1] > buffer_size:
            buf = buf[:, buf.shape[1] - buffer_size:]
        preds.append(cur)
    return torch.stack(preds, dim=1)


def action_conditional(sd, cfg, slot_history, actions, num_seed, num_preds):
    """ ActionCondtionalOCVP.forward (Predictors/ActionConditional_OCVP.py:97-128); cfg = predictor_params """
    n = num_seed + num_preds - 1
    S = slot_history.shape[2]
    emb = _lin(sd, "action_encoder.0", actions[:, :n]).unsqueeze(2).repeat(1, 1, S, 1)
    dm = cfg["DynamicsModel"]["model_params"]
    return cocvp_train_unroll(sd, dm, slot_history[:, :n], emb, None, num_seed, num_preds)


# --------------------------------------------------------------------------- #
# whole seed + rollout pipeline  (05_evaluate_PlaySlot.py:112-139)
# --------------------------------------------------------------------------- #

def posterior_rollout(psd, pcfg, slot_history, num_seed, num_preds, noise):
    """
    forward_multiple_samples(num_samples=1, use_posterior=True)
    (LatentActionPredictors.py:91-163).  pcfg = `predictor_params`.
    Returns (pred_slots (B,np,S,D), invdyn output dict).
    """
    S = slot_history.shape[2]
    la, dm = pcfg["LatentAction"], pcfg["DynamicsModel"]
    if la["model_name"] == "VQSingleSlotLatentAction":
        out = invdyn_single(psd, la["model_params"], slot_history, noise)
        protos = out["action_proto"].unsqueeze(2).repeat(1, 1, S, 1)
        variab = out["action_variability"].unsqueeze(2).repeat(1, 1, S, 1)
    else:
        out = invdyn_multi(psd, la["model_params"], slot_history, noise)
        protos, variab = out["action_proto"], out["action_variability"]
    if not pcfg["use_variability"]:
        variab = None
    pred = cocvp_rollout(psd, dm["model_params"], slot_history[:, :num_seed], protos, variab, num_preds)
    return pred, out


def seed_rollout_pipeline(ssd, scfg, psd, pcfg, video, num_seed, num_preds, noise, init_noise=None):
    """ encode -> InvDyn + rollout -> decode; returns dict of all stage outputs. """
    B = video.shape[0]
    sh = savi_encode_video(ssd, scfg, video, num_seed + num_preds, init_noise)
    pred, act = posterior_rollout(psd, pcfg, sh, num_seed, num_preds, noise)
    recon, rgb, masks = decode(ssd, scfg, pred.reshape(B * num_preds, *pred.shape[2:]))
    return {"slot_history": sh, "pred_slots": pred, "actions": act,
            "recon": recon, "recons": rgb, "masks": masks}
