// SlotAttention iteration (reference: src/models/BlocksUtils/attention.py:92-119), fp32, re-associated so
// that k and v are never materialised:
//     dots[s,j]   = scale * q_s . (W_k xhat_j)      = rstd_j * (A_s . (x_j - mu_j)) + a0_s
//     updates[s]  = W_v * sum_j attn_n[s,j] xhat_j  = W_v * (gamma o U_s + beta * rs_s) / rs_s
// with xhat = LayerNorm(x), A_s = scale * (q_s W_k) o gamma, a0_s = scale * (q_s W_k) . beta,
// w[s,j] = attn[s,j] * rstd_j, U_s = sum_j w[s,j] (x_j - mu_j), rs_s = sum_j attn[s,j].
// Tokens are centred (x - mu) before both contractions so that no large terms have to cancel afterwards.
// One iteration = ps_sa_main (streams the feature tokens once: logits, softmax over slots, +eps, masks,
// weighted sums as per-chunk partials) followed by ps_sa_update (reduces the partials, renormalises,
// v-projection, GRUCell, residual MLP, and the q / A preparation of the next iteration).
#include "common.cuh"
#include "../../include/playslot_b200.h"

#define SA_D 128          // feature dim == slot dim == 128 (all shipped configs)
#define SA_TT 128         // tokens per shared-memory tile
#define SA_XS 144         // row stride (floats) of the token tile: conflict-free LDS.128 for both passes

struct SaMainParams {
    const float* x;       // [B][N][128]
    const float* A;       // [B][S][128]
    const float* a0;      // [B][S]
    const float* a1;      // [B][S]
    float* partial;       // [B][CH][S*128 + 2*S]
    float* masks;         // [B][S][N] or null
    int B, N, S, CH, tpc;
    float ln_eps, attn_eps;
};

template <int SMAX>
__global__ __launch_bounds__(256, 1) void sa_main_kernel(SaMainParams p) {
    extern __shared__ __align__(16) float sm[];
    float* xs = sm;                                   // [2][SA_TT][SA_XS]
    float* As = xs + 2 * SA_TT * SA_XS;               // [SMAX][128]
    float* ws = As + SMAX * SA_D;                     // [SA_TT][SMAX]
    float* mus = ws + SA_TT * SMAX;                   // [SA_TT] token means
    float* red = mus + SA_TT;                         // [8][SMAX]
    __shared__ float s_a0[SMAX];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int S = p.S;
    const int tok0 = chunk * p.tpc;
    const int ntok = min(p.tpc, p.N - tok0);
    const int ntiles = (ntok + SA_TT - 1) / SA_TT;
    const float* xb = p.x + ((size_t)b * p.N + tok0) * SA_D;

    // folded query matrix of this frame
    for (int idx = tid; idx < SMAX * SA_D; idx += 256) {
        const int s = idx / SA_D;
        As[idx] = (s < S) ? p.A[((size_t)b * S + s) * SA_D + (idx - s * SA_D)] : 0.f;
    }
    if (tid < SMAX) s_a0[tid] = (tid < S) ? p.a0[b * S + tid] : 0.f;

    auto load_tile = [&](int t, int buf) {
        float* dst = xs + buf * SA_TT * SA_XS;
        const int base_tok = t * SA_TT;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int idx = tid + 256 * i;
            const int row = idx >> 5, c4 = idx & 31;
            float* d = dst + row * SA_XS + c4 * 4;
            if (base_tok + row < ntok) cp_async16(d, xb + (size_t)(base_tok + row) * SA_D + c4 * 4);
            else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };

    // pass-A mapping: 4 threads per token pair;   pass-B mapping: one float4 channel group per lane
    const int jp = tid >> 2, q = tid & 3;
    const int c4b = lane;
    float U[SMAX][4];
#pragma unroll
    for (int s = 0; s < SMAX; ++s) { U[s][0] = U[s][1] = U[s][2] = U[s][3] = 0.f; }
    float rs_acc[SMAX];
#pragma unroll
    for (int s = 0; s < SMAX; ++s) rs_acc[s] = 0.f;

    load_tile(0, 0);
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) { load_tile(t + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* xt = xs + buf * SA_TT * SA_XS;

        // ---------------- pass A: LayerNorm statistics, logits, softmax over slots ----------------
        {
            const float* r0 = xt + jp * SA_XS;
            const float* r1 = xt + (jp + 64) * SA_XS;
            float sum0 = 0.f, sum1 = 0.f;
            float dot0[SMAX], dot1[SMAX];
#pragma unroll
            for (int s = 0; s < SMAX; ++s) { dot0[s] = 0.f; dot1[s] = 0.f; }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = (4 * i + q) * 4;
                const float4 v0 = *reinterpret_cast<const float4*>(r0 + c);
                const float4 v1 = *reinterpret_cast<const float4*>(r1 + c);
                sum0 += (v0.x + v0.y) + (v0.z + v0.w);
                sum1 += (v1.x + v1.y) + (v1.z + v1.w);
            }
            sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
            sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
            const float mean0 = sum0 * (1.0f / SA_D), mean1 = sum1 * (1.0f / SA_D);
            float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = (4 * i + q) * 4;
                float4 v0 = *reinterpret_cast<const float4*>(r0 + c);
                float4 v1 = *reinterpret_cast<const float4*>(r1 + c);
                v0.x -= mean0; v0.y -= mean0; v0.z -= mean0; v0.w -= mean0;
                v1.x -= mean1; v1.y -= mean1; v1.z -= mean1; v1.w -= mean1;
                sq0 = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, sq0))));
                sq1 = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, sq1))));
#pragma unroll
                for (int s = 0; s < SMAX; ++s) {
                    const float4 a = *reinterpret_cast<const float4*>(As + s * SA_D + c);
                    dot0[s] = fmaf(a.x, v0.x, fmaf(a.y, v0.y, fmaf(a.z, v0.z, fmaf(a.w, v0.w, dot0[s]))));
                    dot1[s] = fmaf(a.x, v1.x, fmaf(a.y, v1.y, fmaf(a.z, v1.z, fmaf(a.w, v1.w, dot1[s]))));
                }
            }
            sq0 += __shfl_xor_sync(0xffffffffu, sq0, 1); sq0 += __shfl_xor_sync(0xffffffffu, sq0, 2);
            sq1 += __shfl_xor_sync(0xffffffffu, sq1, 1); sq1 += __shfl_xor_sync(0xffffffffu, sq1, 2);
            const float rstd0 = 1.0f / sqrtf(sq0 * (1.0f / SA_D) + p.ln_eps);
            const float rstd1 = 1.0f / sqrtf(sq1 * (1.0f / SA_D) + p.ln_eps);
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                dot0[s] += __shfl_xor_sync(0xffffffffu, dot0[s], 1); dot0[s] += __shfl_xor_sync(0xffffffffu, dot0[s], 2);
                dot1[s] += __shfl_xor_sync(0xffffffffu, dot1[s], 1); dot1[s] += __shfl_xor_sync(0xffffffffu, dot1[s], 2);
            }
            // logits and softmax over the slot axis
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                if (s < S) {
                    dot0[s] = rstd0 * dot0[s] + s_a0[s];
                    dot1[s] = rstd1 * dot1[s] + s_a0[s];
                    mx0 = fmaxf(mx0, dot0[s]); mx1 = fmaxf(mx1, dot1[s]);
                }
            }
            float den0 = 0.f, den1 = 0.f;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                if (s < S) {
                    dot0[s] = expf(dot0[s] - mx0); den0 += dot0[s];
                    dot1[s] = expf(dot1[s] - mx1); den1 += dot1[s];
                } else { dot0[s] = 0.f; dot1[s] = 0.f; }
            }
            const int j0 = t * SA_TT + jp, j1 = j0 + 64;
            const float val0 = (j0 < ntok) ? 1.f : 0.f, val1 = (j1 < ntok) ? 1.f : 0.f;
            const float inv0 = 1.0f / den0, inv1 = 1.0f / den1;
#pragma unroll
            for (int s = 0; s < SMAX; ++s) {
                if (s < S) {
                    dot0[s] = (dot0[s] * inv0 + p.attn_eps) * val0;     // softmax(dim=slots) + eps
                    dot1[s] = (dot1[s] * inv1 + p.attn_eps) * val1;
                }
            }
            if (q == 0) {
#pragma unroll
                for (int s = 0; s < SMAX; ++s) rs_acc[s] += dot0[s] + dot1[s];
            }
            if (q == 0) {
                mus[jp] = mean0;
#pragma unroll
                for (int s = 0; s < SMAX; ++s) ws[jp * SMAX + s] = dot0[s] * rstd0;
            } else if (q == 1) {
                mus[jp + 64] = mean1;
#pragma unroll
                for (int s = 0; s < SMAX; ++s) ws[(jp + 64) * SMAX + s] = dot1[s] * rstd1;
            } else if (p.masks) {
                const int j = (q == 2) ? j0 : j1;
                if (j < ntok) {
                    float* mrow = p.masks + (size_t)b * S * p.N + tok0 + j;
#pragma unroll
                    for (int s = 0; s < SMAX; ++s)
                        if (s < S) mrow[(size_t)s * p.N] = (q == 2) ? dot0[s] : dot1[s];
                }
            }
        }
        __syncthreads();

        // ---------------- pass B: U[s][c] += w[s][j] * x[j][c] ----------------
        {
#pragma unroll 4
            for (int j = warp; j < SA_TT; j += 8) {
                float4 xv = *reinterpret_cast<const float4*>(xt + j * SA_XS + c4b * 4);
                const float mu = mus[j];
                xv.x -= mu; xv.y -= mu; xv.z -= mu; xv.w -= mu;
                const float* wr = ws + j * SMAX;
#pragma unroll
                for (int s4 = 0; s4 < SMAX; s4 += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(wr + s4);
                    U[s4 + 0][0] = fmaf(w.x, xv.x, U[s4 + 0][0]); U[s4 + 0][1] = fmaf(w.x, xv.y, U[s4 + 0][1]);
                    U[s4 + 0][2] = fmaf(w.x, xv.z, U[s4 + 0][2]); U[s4 + 0][3] = fmaf(w.x, xv.w, U[s4 + 0][3]);
                    U[s4 + 1][0] = fmaf(w.y, xv.x, U[s4 + 1][0]); U[s4 + 1][1] = fmaf(w.y, xv.y, U[s4 + 1][1]);
                    U[s4 + 1][2] = fmaf(w.y, xv.z, U[s4 + 1][2]); U[s4 + 1][3] = fmaf(w.y, xv.w, U[s4 + 1][3]);
                    U[s4 + 2][0] = fmaf(w.z, xv.x, U[s4 + 2][0]); U[s4 + 2][1] = fmaf(w.z, xv.y, U[s4 + 2][1]);
                    U[s4 + 2][2] = fmaf(w.z, xv.z, U[s4 + 2][2]); U[s4 + 2][3] = fmaf(w.z, xv.w, U[s4 + 2][3]);
                    U[s4 + 3][0] = fmaf(w.w, xv.x, U[s4 + 3][0]); U[s4 + 3][1] = fmaf(w.w, xv.y, U[s4 + 3][1]);
                    U[s4 + 3][2] = fmaf(w.w, xv.z, U[s4 + 3][2]); U[s4 + 3][3] = fmaf(w.w, xv.w, U[s4 + 3][3]);
                }
            }
        }
        __syncthreads();
    }

    // ---------------- CTA reduction and partial write-out ----------------
    float* scratch = xs;                                // [8][SMAX][128], the token tiles are dead now
#pragma unroll
    for (int s = 0; s < SMAX; ++s)
        *reinterpret_cast<float4*>(scratch + (warp * SMAX + s) * SA_D + c4b * 4) =
            make_float4(U[s][0], U[s][1], U[s][2], U[s][3]);
#pragma unroll
    for (int s = 0; s < SMAX; ++s) {
        const float r = warp_sum(rs_acc[s]);             // lanes with q != 0 hold zeros
        if (lane == 0) red[warp * SMAX + s] = r;
    }
    __syncthreads();
    const int pstride = S * SA_D + S;
    float* pout = p.partial + ((size_t)b * p.CH + chunk) * pstride;
    for (int idx = tid; idx < S * SA_D; idx += 256) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += scratch[(w * SMAX) * SA_D + (idx / SA_D) * SA_D + (idx % SA_D)];
        pout[idx] = v;
    }
    if (tid < S) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w * SMAX + tid];
        pout[S * SA_D + tid] = v;
    }
}

template <int SMAX>
static int launch_sa_main(const SaMainParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(2 * SA_TT * SA_XS + SMAX * SA_D + SA_TT * SMAX + SA_TT + 8 * SMAX) * sizeof(float);
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(sa_main_kernel<SMAX>, smem, smem_cfg));
    sa_main_kernel<SMAX><<<dim3(p.CH, p.B), 256, smem, st>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_sa_partial_floats(int S) { return S * SA_D + S; }

extern "C" int ps_sa_main(const float* x, const float* A, const float* a0, const float* a1, float* partial,
                          float* masks, int B, int N, int S, int dim, int n_chunks, float ln_eps,
                          float attn_eps, void* stream) {
    PS_REQUIRE(x && A && a0 && a1 && partial, "ps_sa_main: null pointer");
    PS_REQUIRE(dim == SA_D, "ps_sa_main: feature/slot dim must be %d (got %d)", SA_D, dim);
    PS_REQUIRE(S >= 1 && S <= 16, "ps_sa_main: 1 <= num_slots <= 16 (got %d)", S);
    PS_REQUIRE(B >= 0 && N > 0 && n_chunks > 0, "ps_sa_main: bad shape");
    if (B == 0) return PS_OK;
    PS_REQUIRE(B <= 65535, "ps_sa_main: batch too large for one launch");
    PS_REQUIRE(ps_aligned16(x) && ps_aligned16(A), "ps_sa_main: pointers must be 16-byte aligned");
    SaMainParams p;
    p.x = x; p.A = A; p.a0 = a0; p.a1 = a1; p.partial = partial; p.masks = masks;
    p.B = B; p.N = N; p.S = S; p.CH = n_chunks; p.tpc = (N + n_chunks - 1) / n_chunks;
    p.ln_eps = ln_eps; p.attn_eps = attn_eps;
    PS_REQUIRE((size_t)(n_chunks - 1) * p.tpc < (size_t)N, "ps_sa_main: n_chunks=%d leaves empty chunks for N=%d", n_chunks, N);
    cudaStream_t st = (cudaStream_t)stream;
    if (S <= 4) return launch_sa_main<4>(p, st);
    if (S <= 8) return launch_sa_main<8>(p, st);
    return launch_sa_main<16>(p, st);
}

// ------------------------------------------------------------------------------------------------
// slot update: partial reduction, renormalisation, W_v, GRUCell, residual MLP, next-iteration q/A prep
// ------------------------------------------------------------------------------------------------
struct SaUpdParams {
    const float* partial; int CH;
    const float* slots_in; float* slots_out;
    const float* ln_in_g; const float* ln_in_b;
    const float* WvT;
    const float* WihT; const float* WhhT; const float* b_ih; const float* b_hh;
    const float* ln_mlp_g; const float* ln_mlp_b;
    const float* W1T; const float* b1; const float* W2T; const float* b2;
    const float* ln_slot_g; const float* ln_slot_b;
    const float* WqT; const float* Wk;
    float* A; float* a0; float* a1;
    int B, S, Hm;
    float scale, ln_eps;
    int do_update, do_prep;
};

#define SU_RS 4           // slot rows per CTA
#define SU_THREADS 384

// acc[r] += sum_k Wt[k*ldw + o] * in[k*SU_RS + r]
__device__ __forceinline__ void su_matvec(const float* __restrict__ Wt, int ldw, int K, const float* in,
                                          int o, float acc[SU_RS]) {
#pragma unroll 8
    for (int k = 0; k < K; ++k) {
        const float w = __ldg(Wt + (size_t)k * ldw + o);
        const float4 x = *reinterpret_cast<const float4*>(in + k * SU_RS);
        acc[0] = fmaf(w, x.x, acc[0]); acc[1] = fmaf(w, x.y, acc[1]);
        acc[2] = fmaf(w, x.z, acc[2]); acc[3] = fmaf(w, x.w, acc[3]);
    }
}

// LayerNorm of the SU_RS rows stored as in[d*SU_RS + r] (d < 128) -> out (same layout); warps 0..3 work
__device__ __forceinline__ void su_layernorm(const float* in, float* out, const float* g, const float* b,
                                             float eps, int warp, int lane) {
    if (warp < SU_RS) {
        float v[4], s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = in[(lane + 32 * i) * SU_RS + warp]; s += v[i]; }
        const float mean = warp_sum(s) * (1.0f / SA_D);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; ss = fmaf(d, d, ss); }
        const float rstd = 1.0f / sqrtf(warp_sum(ss) * (1.0f / SA_D) + eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = lane + 32 * i;
            out[d * SU_RS + warp] = (v[i] - mean) * rstd * g[d] + b[d];
        }
    }
}

__global__ __launch_bounds__(SU_THREADS) void sa_update_kernel(SaUpdParams p) {
    __shared__ __align__(16) float s_h[SA_D * SU_RS];       // previous slots (GRU hidden), then new slots
    __shared__ __align__(16) float s_u[SA_D * SU_RS];       // normalised updates in feature space / scratch
    __shared__ __align__(16) float s_v[SA_D * SU_RS];       // updates after W_v / LayerNorm outputs
    __shared__ __align__(16) float s_gi[3 * SA_D * SU_RS];
    __shared__ __align__(16) float s_gh[3 * SA_D * SU_RS];
    __shared__ __align__(16) float s_hid[512 * SU_RS];
    __shared__ float s_rs[SU_RS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, r0 = blockIdx.x * SU_RS;
    const int S = p.S;
    const int nr = min(SU_RS, S - r0);

    for (int idx = tid; idx < SA_D * SU_RS; idx += SU_THREADS) {
        const int r = idx / SA_D, d = idx - r * SA_D;
        s_h[d * SU_RS + r] = (r < nr) ? p.slots_in[((size_t)b * S + r0 + r) * SA_D + d] : 0.f;
    }
    __syncthreads();

    if (p.do_update) {
        const int pstride = S * SA_D + S;
        const float* pb = p.partial + (size_t)b * p.CH * pstride;
        if (tid < SU_RS) {
            float v = 0.f;
            if (tid < nr)
                for (int c = 0; c < p.CH; ++c) v += pb[(size_t)c * pstride + S * SA_D + r0 + tid];
            s_rs[tid] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < SA_D * SU_RS; idx += SU_THREADS) {
            const int r = idx / SA_D, c = idx - r * SA_D;
            float v = 0.f;
            if (r < nr) {
                float u = 0.f;
                for (int ch = 0; ch < p.CH; ++ch) u += pb[(size_t)ch * pstride + (r0 + r) * SA_D + c];
                const float rs = s_rs[r];
                v = (p.ln_in_g[c] * u + p.ln_in_b[c] * rs) / rs;
            }
            s_u[c * SU_RS + r] = v;
        }
        __syncthreads();
        // v projection
        if (tid < SA_D) {
            float acc[SU_RS] = {0.f, 0.f, 0.f, 0.f};
            su_matvec(p.WvT, SA_D, SA_D, s_u, tid, acc);
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) s_v[tid * SU_RS + r] = acc[r];
        }
        __syncthreads();
        // GRU gates: gi = W_ih upd + b_ih, gh = W_hh h + b_hh   (gate order r, z, n)
        {
            float ai[SU_RS], ah[SU_RS];
            const float bi = p.b_ih[tid], bh = p.b_hh[tid];
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) { ai[r] = bi; ah[r] = bh; }
            su_matvec(p.WihT, 3 * SA_D, SA_D, s_v, tid, ai);
            su_matvec(p.WhhT, 3 * SA_D, SA_D, s_h, tid, ah);
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) { s_gi[tid * SU_RS + r] = ai[r]; s_gh[tid * SU_RS + r] = ah[r]; }
        }
        __syncthreads();
        for (int idx = tid; idx < SA_D * SU_RS; idx += SU_THREADS) {
            const int d = idx / SU_RS, r = idx - d * SU_RS;
            const float rg = sigmoidf_(s_gi[d * SU_RS + r] + s_gh[d * SU_RS + r]);
            const float zg = sigmoidf_(s_gi[(SA_D + d) * SU_RS + r] + s_gh[(SA_D + d) * SU_RS + r]);
            const float ng = tanhf(s_gi[(2 * SA_D + d) * SU_RS + r] + rg * s_gh[(2 * SA_D + d) * SU_RS + r]);
            s_h[idx] = (1.0f - zg) * ng + zg * s_h[idx];
        }
        __syncthreads();
        // residual MLP
        su_layernorm(s_h, s_v, p.ln_mlp_g, p.ln_mlp_b, p.ln_eps, warp, lane);
        __syncthreads();
        for (int o = tid; o < p.Hm; o += SU_THREADS) {
            float acc[SU_RS];
            const float bb = p.b1[o];
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) acc[r] = bb;
            su_matvec(p.W1T, p.Hm, SA_D, s_v, o, acc);
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) s_hid[o * SU_RS + r] = fmaxf(acc[r], 0.f);
        }
        __syncthreads();
        if (tid < SA_D) {
            float acc[SU_RS];
            const float bb = p.b2[tid];
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) acc[r] = bb;
            su_matvec(p.W2T, SA_D, p.Hm, s_hid, tid, acc);
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) s_h[tid * SU_RS + r] += acc[r];
        }
        __syncthreads();
        for (int idx = tid; idx < SA_D * SU_RS; idx += SU_THREADS) {
            const int r = idx / SA_D, d = idx - r * SA_D;
            if (r < nr) p.slots_out[((size_t)b * S + r0 + r) * SA_D + d] = s_h[d * SU_RS + r];
        }
    }

    if (p.do_prep) {
        // q = LN_slot(slots) W_q^T ; qk = q W_k ; A = scale * qk o gamma_in ; a1 = sum A ; a0 = scale * qk . beta_in
        su_layernorm(s_h, s_v, p.ln_slot_g, p.ln_slot_b, p.ln_eps, warp, lane);
        __syncthreads();
        if (tid < SA_D) {
            float acc[SU_RS] = {0.f, 0.f, 0.f, 0.f};
            su_matvec(p.WqT, SA_D, SA_D, s_v, tid, acc);
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) s_u[tid * SU_RS + r] = acc[r];
        }
        __syncthreads();
        if (tid < SA_D) {
            float acc[SU_RS] = {0.f, 0.f, 0.f, 0.f};
            su_matvec(p.Wk, SA_D, SA_D, s_u, tid, acc);       // W_k is (d, c): column c of q W_k
            const float g = p.ln_in_g[tid], be = p.ln_in_b[tid];
#pragma unroll
            for (int r = 0; r < SU_RS; ++r) {
                const float a = p.scale * acc[r] * g;
                s_gi[tid * SU_RS + r] = a;                       // A
                s_gh[tid * SU_RS + r] = p.scale * acc[r] * be;   // a0 terms
                if (r < nr) p.A[((size_t)b * S + r0 + r) * SA_D + tid] = a;
            }
        }
        __syncthreads();
        if (warp < SU_RS && warp < nr) {
            float s1 = 0.f, s0 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s1 += s_gi[(lane + 32 * i) * SU_RS + warp];
                s0 += s_gh[(lane + 32 * i) * SU_RS + warp];
            }
            s1 = warp_sum(s1); s0 = warp_sum(s0);
            if (lane == 0) { p.a1[b * S + r0 + warp] = s1; p.a0[b * S + r0 + warp] = s0; }   // a1 = sum_c A (diagnostic only)
        }
    }
}

extern "C" int ps_sa_update(const float* partial, int n_chunks, const float* slots_in, float* slots_out,
                            const float* ln_in_g, const float* ln_in_b, const float* WvT,
                            const float* WihT, const float* WhhT, const float* b_ih, const float* b_hh,
                            const float* ln_mlp_g, const float* ln_mlp_b,
                            const float* W1T, const float* b1, const float* W2T, const float* b2,
                            const float* ln_slot_g, const float* ln_slot_b, const float* WqT, const float* Wk,
                            float* A, float* a0, float* a1,
                            int B, int S, int dim, int mlp_hidden, float scale, float ln_eps,
                            int do_update, int do_prep, void* stream) {
    PS_REQUIRE(dim == SA_D, "ps_sa_update: slot dim must be %d (got %d)", SA_D, dim);
    PS_REQUIRE(S >= 1 && S <= 16, "ps_sa_update: 1 <= num_slots <= 16 (got %d)", S);
    PS_REQUIRE(mlp_hidden >= 1 && mlp_hidden <= 512, "ps_sa_update: mlp_hidden must be <= 512 (got %d)", mlp_hidden);
    PS_REQUIRE(slots_in, "ps_sa_update: null slots");
    PS_REQUIRE(!do_update || (partial && slots_out && WvT && WihT && WhhT && b_ih && b_hh && W1T && W2T && b1 && b2 && ln_mlp_g && ln_mlp_b),
               "ps_sa_update: update requested with null weights");
    PS_REQUIRE(!do_prep || (A && a0 && a1 && WqT && Wk && ln_slot_g && ln_slot_b), "ps_sa_update: prep requested with null buffers");
    PS_REQUIRE(ln_in_g && ln_in_b, "ps_sa_update: null norm_input params");
    if (B == 0) return PS_OK;
    PS_REQUIRE(B <= 65535, "ps_sa_update: batch too large for one launch");
    SaUpdParams p;
    p.partial = partial; p.CH = n_chunks; p.slots_in = slots_in; p.slots_out = slots_out;
    p.ln_in_g = ln_in_g; p.ln_in_b = ln_in_b; p.WvT = WvT; p.WihT = WihT; p.WhhT = WhhT; p.b_ih = b_ih; p.b_hh = b_hh;
    p.ln_mlp_g = ln_mlp_g; p.ln_mlp_b = ln_mlp_b; p.W1T = W1T; p.b1 = b1; p.W2T = W2T; p.b2 = b2;
    p.ln_slot_g = ln_slot_g; p.ln_slot_b = ln_slot_b; p.WqT = WqT; p.Wk = Wk; p.A = A; p.a0 = a0; p.a1 = a1;
    p.B = B; p.S = S; p.Hm = mlp_hidden; p.scale = scale; p.ln_eps = ln_eps; p.do_update = do_update; p.do_prep = do_prep;
    dim3 grid((S + SU_RS - 1) / SU_RS, B);
    sa_update_kernel<<<grid, SU_THREADS, 0, (cudaStream_t)stream>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// slot update, latency-optimised variant (ps_sa_update2).  Same mathematics as ps_sa_update with two weight-only
// re-associations done once by the caller (exact up to fp32 rounding):
//   W_iv  = W_ih W_v                       gi = W_iv u + b_ih            (drops the separate v projection)
//   M_A   = scale * diag(gamma_in) W_k^T W_q,  v0 = scale * beta_in^T W_k^T W_q
//                                          A = M_A LN_slot(slots),  a0 = v0 . LN_slot(slots)
// One CTA per (frame, group of 4 slots), 768 threads; every mat-vec is split over K so that each thread issues
// all of its (<= 32) weight loads before the first FMA: a stage costs about one L2 round trip.
// ------------------------------------------------------------------------------------------------
#define SU2_THREADS 768

struct SaUpd2Params {
    const float* partial; int CH;
    const float* slots_in; float* slots_out;
    const float* ln_in_g; const float* ln_in_b;
    const float* WivT; const float* WhhT; const float* b_ih; const float* b_hh;      // [128][384]
    const float* ln_mlp_g; const float* ln_mlp_b;
    const float* W1T; const float* b1; const float* W2T; const float* b2;            // [128][Hm], [Hm][128]
    const float* ln_slot_g; const float* ln_slot_b;
    const float* MAT; const float* v0;                                               // [128 e][128 c], [128]
    float* A; float* a0;
    int B, S, Hm;
    float ln_eps;
    int do_update, do_prep;
};

// partial[(ks*N + o)*R + r] = sum_{k in slice ks} Wt[k*ld + o] * in[k*R + r]   for all (o, ks) tasks
template <int KS, int R>
__device__ __forceinline__ void su2_matvec(const float* __restrict__ Wt, int ld, int K, int N, const float* in,
                                           float* part, int tid) {
    const int kper = K / KS;
    for (int task = tid; task < N * KS; task += SU2_THREADS) {
        const int o = task % N, ks = task / N;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        for (int k0 = ks * kper; k0 < (ks + 1) * kper; k0 += 32) {
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = __ldg(Wt + (size_t)(k0 + j) * ld + o);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
#pragma unroll
                for (int r4 = 0; r4 < R; r4 += 4) {
                    const float4 x = *reinterpret_cast<const float4*>(in + (k0 + j) * R + r4);
                    acc[r4] = fmaf(w[j], x.x, acc[r4]); acc[r4 + 1] = fmaf(w[j], x.y, acc[r4 + 1]);
                    acc[r4 + 2] = fmaf(w[j], x.z, acc[r4 + 2]); acc[r4 + 3] = fmaf(w[j], x.w, acc[r4 + 3]);
                }
            }
        }
#pragma unroll
        for (int r4 = 0; r4 < R; r4 += 4)
            *reinterpret_cast<float4*>(part + (size_t)task * R + r4) = make_float4(acc[r4], acc[r4 + 1], acc[r4 + 2], acc[r4 + 3]);
    }
}

template <int R>
__device__ __forceinline__ float su2_sum(const float* part, int N, int KS, int o, int r) {
    float v = 0.f;
    for (int ks = 0; ks < KS; ++ks) v += part[(size_t)(ks * N + o) * R + r];
    return v;
}

// LayerNorm of the R rows stored as in[d*R + r] (d < 128) -> out (same layout); warps 0..R-1 work
template <int R>
__device__ __forceinline__ void su2_layernorm(const float* in, float* out, const float* g, const float* b,
                                              float eps, int warp, int lane) {
    if (warp < R) {
        float v[4], s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = in[(lane + 32 * i) * R + warp]; s += v[i]; }
        const float mean = warp_sum(s) * (1.0f / SA_D);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; ss = fmaf(d, d, ss); }
        const float rstd = 1.0f / sqrtf(warp_sum(ss) * (1.0f / SA_D) + eps);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int d = lane + 32 * i;
            out[d * R + warp] = (v[i] - mean) * rstd * g[d] + b[d];
        }
    }
}

// R slot rows of one frame per CTA (4, or 8 for the 5- and 6-slot configurations: see ps_sa_update2).
template <int R>
__global__ __launch_bounds__(SU2_THREADS) void sa_update2_kernel(SaUpd2Params p) {
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // PDL: the streaming pass that follows may start its feature loads (see slot_attention_tc.cu)
    extern __shared__ __align__(16) float su2_smem[];
    float* s_h = su2_smem;                        // [128][R] hidden state / new slots, [d][r]
    float* s_u = s_h + SA_D * R;                  // [128][R] normalised updates / LayerNorm outputs
    float* s_hid = s_u + SA_D * R;                // [512][R] MLP hidden
    float* s_p0 = s_hid + 512 * R;                // [2 * 384][R] mat-vec partials
    float* s_p1 = s_p0 + 2 * 384 * R;
    __shared__ float s_rs[R];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, r0 = blockIdx.x * R;
    const int S = p.S;
    const int nr = min(R, S - r0);

    for (int idx = tid; idx < SA_D * R; idx += SU2_THREADS) {
        const int r = idx / SA_D, d = idx - r * SA_D;
        s_h[d * R + r] = (r < nr) ? p.slots_in[((size_t)b * S + r0 + r) * SA_D + d] : 0.f;
    }
    if (p.do_update) {
        const int pstride = S * SA_D + S;
        const float* pb = p.partial + (size_t)b * p.CH * pstride;
        if (tid < R) {
            float v = 0.f;
            if (tid < nr)
                for (int c = 0; c < p.CH; ++c) v += pb[(size_t)c * pstride + S * SA_D + r0 + tid];
            s_rs[tid] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < SA_D * R; idx += SU2_THREADS) {
            const int r = idx / SA_D, c = idx - r * SA_D;
            float v = 0.f;
            if (r < nr) {
                float u = 0.f;
                for (int ch = 0; ch < p.CH; ++ch) u += pb[(size_t)ch * pstride + (r0 + r) * SA_D + c];
                const float rs = s_rs[r];
                v = (p.ln_in_g[c] * u + p.ln_in_b[c] * rs) / rs;
            }
            s_u[c * R + r] = v;
        }
        __syncthreads();
        // GRU gates (gate order r, z, n): gi = W_iv u + b_ih, gh = W_hh h + b_hh
        su2_matvec<2, R>(p.WivT, 3 * SA_D, SA_D, 3 * SA_D, s_u, s_p0, tid);
        su2_matvec<2, R>(p.WhhT, 3 * SA_D, SA_D, 3 * SA_D, s_h, s_p1, tid);
        __syncthreads();
        for (int idx = tid; idx < SA_D * R; idx += SU2_THREADS) {
            const int d = idx / R, r = idx - d * R;
            const float gir = su2_sum<R>(s_p0, 384, 2, d, r) + p.b_ih[d], ghr = su2_sum<R>(s_p1, 384, 2, d, r) + p.b_hh[d];
            const float giz = su2_sum<R>(s_p0, 384, 2, SA_D + d, r) + p.b_ih[SA_D + d], ghz = su2_sum<R>(s_p1, 384, 2, SA_D + d, r) + p.b_hh[SA_D + d];
            const float gin = su2_sum<R>(s_p0, 384, 2, 2 * SA_D + d, r) + p.b_ih[2 * SA_D + d], ghn = su2_sum<R>(s_p1, 384, 2, 2 * SA_D + d, r) + p.b_hh[2 * SA_D + d];
            const float rg = sigmoidf_(gir + ghr), zg = sigmoidf_(giz + ghz);
            const float ng = tanhf(gin + rg * ghn);
            s_h[idx] = (1.0f - zg) * ng + zg * s_h[idx];
        }
        __syncthreads();
        su2_layernorm<R>(s_h, s_u, p.ln_mlp_g, p.ln_mlp_b, p.ln_eps, warp, lane);
        __syncthreads();
        su2_matvec<2, R>(p.W1T, p.Hm, SA_D, p.Hm, s_u, s_p0, tid);           // Hm <= 384: 2 x Hm x R floats fit s_p0
        __syncthreads();
        for (int idx = tid; idx < p.Hm * R; idx += SU2_THREADS) {
            const int o = idx / R, r = idx - o * R;
            s_hid[idx] = fmaxf(su2_sum<R>(s_p0, p.Hm, 2, o, r) + p.b1[o], 0.f);
        }
        __syncthreads();
        su2_matvec<4, R>(p.W2T, SA_D, p.Hm, SA_D, s_hid, s_p1, tid);
        __syncthreads();
        for (int idx = tid; idx < SA_D * R; idx += SU2_THREADS) {
            const int d = idx / R, r = idx - d * R;
            const float v = s_h[idx] + su2_sum<R>(s_p1, SA_D, 4, d, r) + p.b2[d];
            s_h[idx] = v;
            if (r < nr) p.slots_out[((size_t)b * S + r0 + r) * SA_D + d] = v;
        }
        __syncthreads();
    } else {
        __syncthreads();
    }
    if (p.do_prep) {
        su2_layernorm<R>(s_h, s_u, p.ln_slot_g, p.ln_slot_b, p.ln_eps, warp, lane);
        __syncthreads();
        su2_matvec<4, R>(p.MAT, SA_D, SA_D, SA_D, s_u, s_p0, tid);
        if (warp >= 16 && warp < 16 + nr) {                               // a0 = v0 . n, one warp per slot row
            const int r = warp - 16;
            float s0 = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) s0 = fmaf(p.v0[lane + 32 * i], s_u[(lane + 32 * i) * R + r], s0);
            s0 = warp_sum(s0);
            if (lane == 0) p.a0[b * S + r0 + r] = s0;
        }
        __syncthreads();
        for (int idx = tid; idx < SA_D * R; idx += SU2_THREADS) {
            const int c = idx / R, r = idx - c * R;
            if (r < nr) p.A[((size_t)b * S + r0 + r) * SA_D + c] = su2_sum<R>(s_p0, SA_D, 4, c, r);
        }
    }
}

template <int R>
static int launch_sa_update2(const SaUpd2Params& p, cudaStream_t st) {
    const size_t smem = (size_t)(2 * SA_D * R + 512 * R + 2 * 2 * 384 * R) * sizeof(float);
    static PsSmemConfig smem_cfg = {};
    if (smem > 48 * 1024) PS_CUDA(ps_ensure_smem(sa_update2_kernel<R>, smem, smem_cfg));
    dim3 grid((p.S + R - 1) / R, p.B);
    sa_update2_kernel<R><<<grid, SU2_THREADS, smem, st>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_sa_update2(const float* partial, int n_chunks, const float* slots_in, float* slots_out,
                             const float* ln_in_g, const float* ln_in_b,
                             const float* WivT, const float* WhhT, const float* b_ih, const float* b_hh,
                             const float* ln_mlp_g, const float* ln_mlp_b,
                             const float* W1T, const float* b1, const float* W2T, const float* b2,
                             const float* ln_slot_g, const float* ln_slot_b, const float* MAT, const float* v0,
                             float* A, float* a0, int B, int S, int dim, int mlp_hidden, float ln_eps,
                             int do_update, int do_prep, void* stream) {
    PS_REQUIRE(dim == SA_D, "ps_sa_update2: slot dim must be %d (got %d)", SA_D, dim);
    PS_REQUIRE(S >= 1 && S <= 16, "ps_sa_update2: 1 <= num_slots <= 16 (got %d)", S);
    PS_REQUIRE(mlp_hidden >= 32 && mlp_hidden <= 384 && mlp_hidden % 128 == 0,
               "ps_sa_update2: mlp_hidden must be 128, 256 or 384 (got %d)", mlp_hidden);
    PS_REQUIRE(slots_in, "ps_sa_update2: null slots");
    PS_REQUIRE(!do_update || (partial && slots_out && WivT && WhhT && b_ih && b_hh && W1T && W2T && b1 && b2 && ln_mlp_g && ln_mlp_b && ln_in_g && ln_in_b),
               "ps_sa_update2: update requested with null weights");
    PS_REQUIRE(!do_prep || (A && a0 && MAT && v0 && ln_slot_g && ln_slot_b), "ps_sa_update2: prep requested with null buffers");
    if (B == 0) return PS_OK;
    PS_REQUIRE(B <= 65535, "ps_sa_update2: batch too large for one launch");
    SaUpd2Params p;
    p.partial = partial; p.CH = n_chunks; p.slots_in = slots_in; p.slots_out = slots_out;
    p.ln_in_g = ln_in_g; p.ln_in_b = ln_in_b; p.WivT = WivT; p.WhhT = WhhT; p.b_ih = b_ih; p.b_hh = b_hh;
    p.ln_mlp_g = ln_mlp_g; p.ln_mlp_b = ln_mlp_b; p.W1T = W1T; p.b1 = b1; p.W2T = W2T; p.b2 = b2;
    p.ln_slot_g = ln_slot_g; p.ln_slot_b = ln_slot_b; p.MAT = MAT; p.v0 = v0; p.A = A; p.a0 = a0;
    p.B = B; p.S = S; p.Hm = mlp_hidden; p.ln_eps = ln_eps; p.do_update = do_update; p.do_prep = do_prep;
    // 5 or 6 slots: one 8-row CTA per frame instead of a full 4-row CTA plus one that re-reads all the weights for 1 - 2 rows
    // (GridShapes5Objs, 64 frames: 24 -> 20 us per launch).  With 8 slots two 4-row CTAs per frame are faster than one 8-row CTA
    // (measured: 0.221 vs 0.231 ms for the 3-iteration forward) -- the kernel needs the parallelism as much as the L2 bytes.
    if (S == 5 || S == 6) return launch_sa_update2<8>(p, (cudaStream_t)stream);
    return launch_sa_update2<4>(p, (cudaStream_t)stream);
}
