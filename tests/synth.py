"""
Deterministic synthetic weights and inputs for the parity tests.

The reference ships no weights (.MISSING_LARGE_BLOBS) and its own initialiser
zeroes every bias, which would hide bias bugs.  `synth_state_dict` fills a
state_dict *manifest* (key -> shape) with seeded values whose scale keeps
activations O(1); it is used both by `oracle/make_golden.py` (loading them into
the real reference modules) and by the tests (loading them into ours), so no
weight tensor has to be committed.
"""
import zlib

import torch


def _gen(seed, name):
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))


def synth_state_dict(manifest, seed, sharpen=1.0, skip=("pos_emb.pe", "pe.pe")):
    """ manifest: {key: shape}.  Keys ending in one of `skip` are left out (buffers built by the module). """
    sd = {}
    for name, shape in manifest.items():
        if any(name.endswith(s) for s in skip):
            continue
        shape = tuple(shape)
        g = _gen(seed, name)
        if name.endswith("ema_count"):
            t = 1.0 + torch.rand(shape, generator=g)
        elif "embeddings" in name or name.endswith("ema_weight"):
            t = 0.6 * torch.randn(shape, generator=g)
        elif name.endswith(".bias") or name.endswith("bias_ih") or name.endswith("bias_hh"):
            t = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:                         # LayerNorm gains
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("initializer.slots") or name.endswith("slots_mu") or name.endswith("act_token"):
            t = torch.randn(shape, generator=g)
        elif name.endswith("slots_sigma"):
            t = 0.5 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) / (fan_in ** 0.5)
            if len(shape) == 4:
                t = t * 1.6                            # keep ReLU conv stacks from shrinking
        if sharpen != 1.0:
            if name.endswith("slot_attention.to_q.weight") or name.endswith("slot_attention.to_k.weight"):
                t = t * sharpen
            if "decoder.decoder." in name and len(shape) == 4 and shape[0] == 4:   # final RGB+alpha conv
                t = t * sharpen
        sd[name] = t.contiguous()
    return sd


def synth_video(B, T, H, W, seed):
    g = torch.Generator().manual_seed(seed + 7919)
    return torch.rand(B, T, 3, H, W, generator=g)


def synth_noise(shape, seed):
    g = torch.Generator().manual_seed(seed + 104729)
    return torch.randn(*shape, generator=g)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    """ max |a-b| normalised by max |b|  (the parity metric used with rel_l2) """
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
