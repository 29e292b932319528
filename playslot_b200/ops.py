"""
Thin tensor-level wrappers around the C ABI (include/playslot_b200.h).

PyTorch is used for device memory, streams and (later) autograd plumbing only: every function here checks
its arguments, allocates the outputs with torch.empty and launches hand-written kernels on the current
CUDA stream.  CPU tensors are rejected -- there is no fallback.
"""
import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_GELU, ACT_ABS = 0, 1, 2, 3


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor")
    if not t.is_cuda:
        raise _lib.PlaySlotKernelError(
            f"{name} lives on {t.device}: playslot_b200 runs on CUDA (sm_100a) only and has no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


# --- optional per-kernel device timing (CUDA events on the launching stream); used by bench.py only --------
_PROFILE = None


def profile_start():
    """ start collecting (name, start_event, end_event, meta) for every launch made through this module """
    global _PROFILE
    _PROFILE = []


def profile_stop():
    """ stop collecting; returns {name: {"ms": total, "calls": n, "meta": [..]}} (synchronises the device) """
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, meta in rec or []:
        d = out.setdefault(name, {"ms": 0.0, "calls": 0, "meta": []})
        d["ms"] += e0.elapsed_time(e1)
        d["calls"] += 1
        if meta is not None:
            d["meta"].append(meta)
    return out


class _timed:
    def __init__(self, name, meta=None):
        self.name, self.meta = name, meta

    def __enter__(self):
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.e1.record()
            _PROFILE.append((self.name, self.e0, self.e1, self.meta))
        return False


def launch_count():
    return int(_lib.load().ps_launch_count())


def reset_launch_count():
    _lib.load().ps_reset_launch_count()


# ------------------------------------------------------------------------------------------------ dense

def linear(x, weight, bias=None, ln=None, act=ACT_NONE, res=None, res_div=1, res_mod=None, out=None):
    """
    y = act(LN(x) @ weight.T + bias) + res_rows.   x: (..., K) contiguous; weight (N, K).
    ln = (gamma, beta, eps) fuses a LayerNorm over K in front.  `res` is a 2-D (R, N) tensor; output
    row m receives res[(m // res_div) % res_mod] (defaults: plain residual of the same shape).
    """
    lib = _lib.load()
    _req(x, "x"); _req(weight, "weight")
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise ValueError(f"linear: weight {tuple(weight.shape)} does not match input features {K}")
    M = x.numel() // K
    x2 = x.view(M, K)
    if out is None:
        out = torch.empty(*x.shape[:-1], N, device=x.device, dtype=torch.float32)
    y2 = out.view(M, N)
    g = b = None
    eps = 0.0
    if ln is not None:
        g, b, eps = ln
        _req(g, "ln gamma"); _req(b, "ln beta")
    if bias is not None:
        _req(bias, "bias")
    ldres = 0
    if res is not None:
        _req(res, "res")
        res = res.view(-1, N)
        ldres = N
        if res_mod is None:
            res_mod = res.shape[0]
    with _timed("linear", (M, N, K)):
        st = lib.ps_linear(_ptr(x2), K, _ptr(weight), _ptr(bias), _ptr(g), _ptr(b), float(eps),
                           _ptr(res), ldres, int(res_div), int(res_mod or 0), _ptr(y2), N, M, N, K, int(act), _stream())
    _lib.check(st, "ps_linear")
    return out


def mha(qkv, n_seq, L, n_heads, head_dim, slots_per_frame=0, res=None):
    """ qkv: (n_seq*L, 3*n_heads*head_dim) -> (n_seq*L, n_heads*head_dim) (+ res of the same shape) """
    lib = _lib.load()
    _req(qkv, "qkv")
    inner = n_heads * head_dim
    qkv2 = qkv.view(n_seq * L, 3 * inner)
    out = torch.empty(n_seq * L, inner, device=qkv.device, dtype=torch.float32)
    if res is not None:
        _req(res, "res")
        assert res.numel() == out.numel()
    with _timed("mha"):
        st = lib.ps_mha(_ptr(qkv2), 3 * inner, _ptr(res), _ptr(out), inner, n_seq, L, n_heads, head_dim, int(slots_per_frame), _stream())
    _lib.check(st, "ps_mha")
    return out


def pack_conv_weight(w):
    """ nn.Conv2d weight (Cout, Cin, K, K) -> packed (Cin, K*K, Cout) used by ps_conv2d / ps_dec_composite """
    Cout, Cin, K, _ = w.shape
    return w.detach().permute(1, 2, 3, 0).reshape(Cin, K * K, Cout).contiguous()


def conv2d(x, w_packed, bias, ksize, relu, out_nhwc=False, addend=None):
    """ x (n, Cin, H, W) -> (n, Cout, H, W) or, with out_nhwc, (n, H, W, Cout) (+ addend (H*W, Cout)) """
    lib = _lib.load()
    _req(x, "x"); _req(w_packed, "w_packed")
    n, Cin, H, W = x.shape
    Cout = w_packed.shape[-1]
    assert w_packed.shape[0] == Cin and w_packed.shape[1] == ksize * ksize
    shape = (n, H, W, Cout) if out_nhwc else (n, Cout, H, W)
    y = torch.empty(shape, device=x.device, dtype=torch.float32)
    MAXI = 65535
    for i0 in range(0, n, MAXI):
        i1 = min(n, i0 + MAXI)
        with _timed("conv2d_%dx%d_%dto%d" % (ksize, ksize, Cin, Cout), (i1 - i0, Cin, Cout, H, W, ksize)):
            st = lib.ps_conv2d(_ptr(x[i0:i1]), _ptr(w_packed), _ptr(bias), _ptr(addend), _ptr(y[i0:i1]),
                               i1 - i0, Cin, Cout, H, W, ksize, int(relu), int(out_nhwc), _stream())
        _lib.check(st, "ps_conv2d")
    return y


# ------------------------------------------------------------------------------------------------ SlotAttention

def sa_partial_floats(S):
    return int(_lib.load().ps_sa_partial_floats(S))


def sa_main(x, A, a0, a1, partial, masks, n_chunks, ln_eps, attn_eps):
    lib = _lib.load()
    _req(x, "feats"); _req(A, "A"); _req(partial, "partial")
    B, N, D = x.shape
    S = A.shape[1]
    with _timed("sa_main", (B, N, S, D, masks is not None)):
        st = lib.ps_sa_main(_ptr(x), _ptr(A), _ptr(a0), _ptr(a1), _ptr(partial), _ptr(masks), B, N, S, D, n_chunks,
                            float(ln_eps), float(attn_eps), _stream())
    _lib.check(st, "ps_sa_main")


def sa_update(partial, n_chunks, slots_in, slots_out, w, A, a0, a1, mlp_hidden, scale, ln_eps, do_update, do_prep):
    """ w: dict of prepared SlotAttention weights (see models.BlocksUtils.attention.SlotAttention._weights) """
    lib = _lib.load()
    B, S, D = slots_in.shape
    with _timed("sa_update"):
      st = lib.ps_sa_update(_ptr(partial), n_chunks, _ptr(slots_in), _ptr(slots_out),
                          _ptr(w["ln_in_g"]), _ptr(w["ln_in_b"]), _ptr(w["WvT"]),
                          _ptr(w["WihT"]), _ptr(w["WhhT"]), _ptr(w["b_ih"]), _ptr(w["b_hh"]),
                          _ptr(w["ln_mlp_g"]), _ptr(w["ln_mlp_b"]),
                          _ptr(w["W1T"]), _ptr(w["b1"]), _ptr(w["W2T"]), _ptr(w["b2"]),
                          _ptr(w["ln_slot_g"]), _ptr(w["ln_slot_b"]), _ptr(w["WqT"]), _ptr(w["Wk"]),
                          _ptr(A), _ptr(a0), _ptr(a1), B, S, D, int(mlp_hidden), float(scale), float(ln_eps),
                          int(do_update), int(do_prep), _stream())
    _lib.check(st, "ps_sa_update")


# ------------------------------------------------------------------------------------------------ decoder

def dec_expand(G, P, n_img, C, H, W, ksize):
    lib = _lib.load()
    _req(G, "G"); _req(P, "P")
    out = torch.empty(n_img, C, H, W, device=G.device, dtype=torch.float32)
    with _timed("dec_expand"):
        st = lib.ps_dec_expand(_ptr(G), _ptr(P), _ptr(out), n_img, C, H, W, ksize, _stream())
    _lib.check(st, "ps_dec_expand")
    return out


def dec_composite(h, w_packed, bias, M, S):
    lib = _lib.load()
    _req(h, "h"); _req(w_packed, "w_packed"); _req(bias, "bias")
    _, Cin, H, W = h.shape
    recon = torch.empty(M, 3, H, W, device=h.device, dtype=torch.float32)
    recons = torch.empty(M, S, 3, H, W, device=h.device, dtype=torch.float32)
    masks = torch.empty(M, S, 1, H, W, device=h.device, dtype=torch.float32)
    MAXM = 65535
    for m0 in range(0, M, MAXM):
        m1 = min(M, m0 + MAXM)
        with _timed("dec_composite"):
            st = lib.ps_dec_composite(_ptr(h[m0 * S:m1 * S]), _ptr(w_packed), _ptr(bias), _ptr(recon[m0:m1]),
                                      _ptr(recons[m0:m1]), _ptr(masks[m0:m1]), m1 - m0, S, Cin, H, W, _stream())
        _lib.check(st, "ps_dec_composite")
    return recon, recons, masks


# ------------------------------------------------------------------------------------------------ InvDyn tail

def invdyn_tail(mv, noise, codebook, per_slot, eps=1e-6):
    """
    mv (B, T, R, 2A), noise (B, T-1, R, A), codebook (K, A) ->
    dict(action_dist (B,T-1,2,R,A), z, proto, variability (B,T-1,R,A), idx int64 (B,T-1,R), loss (B,T-1,R))
    """
    lib = _lib.load()
    _req(mv, "mv"); _req(noise, "noise"); _req(codebook, "codebook")
    B, T, R, A2 = mv.shape
    A = A2 // 2
    K = codebook.shape[0]
    dev = mv.device
    dist = torch.empty(B, T - 1, 2, R, A, device=dev, dtype=torch.float32)
    z = torch.empty(B, T - 1, R, A, device=dev, dtype=torch.float32)
    proto = torch.empty_like(z)
    var = torch.empty_like(z)
    idx = torch.empty(B, T - 1, R, device=dev, dtype=torch.int64)
    loss = torch.empty(B, T - 1, R, device=dev, dtype=torch.float32)
    st = lib.ps_invdyn_tail(_ptr(mv), _ptr(noise), _ptr(codebook), _ptr(dist), _ptr(z), _ptr(proto), _ptr(var),
                            _ptr(idx), _ptr(loss), B, T, R, A, K, int(per_slot), float(eps), _stream())
    _lib.check(st, "ps_invdyn_tail")
    return {"action_dist": dist, "z": z, "proto": proto, "variability": var, "idx": idx, "loss": loss}


def vq_lookup(z, codebook):
    """ z (..., A) -> (proto = z + (z_q - z), z_q, idx int64 (..., 1), loss (...)) """
    lib = _lib.load()
    _req(z, "z"); _req(codebook, "codebook")
    A = z.shape[-1]
    n = z.numel() // A
    K = codebook.shape[0]
    proto = torch.empty_like(z)
    zq = torch.empty_like(z)
    idx = torch.empty(*z.shape[:-1], 1, device=z.device, dtype=torch.int64)
    loss = torch.empty(z.shape[:-1], device=z.device, dtype=torch.float32)
    st = lib.ps_vq_lookup(_ptr(z), _ptr(codebook), _ptr(proto), _ptr(zq), _ptr(idx), _ptr(loss), n, A, K, _stream())
    _lib.check(st, "ps_vq_lookup")
    return proto, zq, idx, loss
