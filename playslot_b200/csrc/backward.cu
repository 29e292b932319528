// Backward kernels of the joint InvDyn + cOCVP training step (src/04_train_PlaySlot.py:32-122,
// src/base/basePredictorTrainer.py:320-333): what loss.backward() runs through in the reference --
//   nn.LayerNorm / nn.Linear / ReLU / GELU / |.| of the predictor (weight and data gradients),
//   multi-head attention with the block-causal mask (BlocksUtils/attention.py:204-316),
//   the InvDyn tail (LatentAction.py:72-87,228-249; quantization.py:294-322) and
//   the data gradient of the frozen SAVi decoder (SAVi.py:220-248, decoders.py:89-113).
// Training-time forwards that must keep intermediates (un-fused LayerNorm, activation) live here too.
// All fp32; the 64->64 convolution data gradients re-use the tcgen05 kernel of conv_tc.cu with flipped weights.
#include "common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_fp16.h>

// ------------------------------------------------------------------------------------------------
// LayerNorm forward (keeps mean / rstd) and backward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                    const float* __restrict__ b, float* __restrict__ y,
                                                    float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                    int M, int K, float eps) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * 8 + warp;
    if (m >= M) return;
    const float* xr = x + (size_t)m * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s += xr[k];
    const float mean = warp_sum(s) / (float)K;
    float ss = 0.f;
    for (int k = lane; k < K; k += 32) { const float d = xr[k] - mean; ss = fmaf(d, d, ss); }
    const float rstd = 1.0f / sqrtf(warp_sum(ss) / (float)K + eps);
    for (int k = lane; k < K; k += 32) y[(size_t)m * K + k] = (xr[k] - mean) * rstd * g[k] + b[k];
    if (lane == 0) { mean_o[m] = mean; rstd_o[m] = rstd; }
}

// dx = rstd * (gy - mean_k(gy) - xhat * mean_k(gy * xhat)),  gy = dy * gamma;  dgamma += dy * xhat, dbeta += dy
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                    const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
                                                    const float* __restrict__ g, float* __restrict__ dx,
                                                    float* __restrict__ dg, float* __restrict__ db, int M, int K, int rows_per_cta) {
    extern __shared__ float s_acc[];                 // [2][K]
    for (int k = threadIdx.x; k < 2 * K; k += 256) s_acc[k] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_begin = blockIdx.x * rows_per_cta, m_end = min(M, m_begin + rows_per_cta);
    for (int m = m_begin + warp; m < m_end; m += 8) {
        const float mean = mean_i[m], rstd = rstd_i[m];
        const float* xr = x + (size_t)m * K;
        const float* dr = dy + (size_t)m * K;
        float s1 = 0.f, s2 = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float xh = (xr[k] - mean) * rstd, gy = dr[k] * g[k];
            s1 += gy; s2 = fmaf(gy, xh, s2);
        }
        s1 = warp_sum(s1) / (float)K; s2 = warp_sum(s2) / (float)K;
        for (int k = lane; k < K; k += 32) {
            const float xh = (xr[k] - mean) * rstd, d = dr[k];
            dx[(size_t)m * K + k] = rstd * (d * g[k] - s1 - xh * s2);
            atomicAdd(&s_acc[k], d * xh);
            atomicAdd(&s_acc[K + k], d);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += 256) {
        atomicAdd(&dg[k], s_acc[k]);
        atomicAdd(&db[k], s_acc[K + k]);
    }
}

extern "C" int ps_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean,
                                float* rstd, int M, int K, void* stream) {
    PS_REQUIRE(x && gamma && beta && y && mean && rstd, "ps_layernorm_fwd: null pointer");
    PS_REQUIRE(M >= 0 && K > 0, "ps_layernorm_fwd: bad shape");
    if (M == 0) return PS_OK;
    ln_fwd_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, M, K, eps);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                float* dx, float* dgamma, float* dbeta, int M, int K, void* stream) {
    PS_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "ps_layernorm_bwd: null pointer");
    PS_REQUIRE(M >= 0 && K > 0 && K <= 4096, "ps_layernorm_bwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    PS_CUDA(cudaMemsetAsync(dgamma, 0, K * sizeof(float), st));
    PS_CUDA(cudaMemsetAsync(dbeta, 0, K * sizeof(float), st));
    if (M == 0) return PS_OK;
    // 8 rows per pass (one per warp); enough CTAs to fill the GPU at the training batch sizes (M = 256 .. a few thousand rows)
    int rows_per_cta = ((M + 147) / 148 + 7) / 8 * 8;
    if (rows_per_cta < 8) rows_per_cta = 8;
    if (rows_per_cta > 64) rows_per_cta = 64;
    ln_bwd_kernel<<<(M + rows_per_cta - 1) / rows_per_cta, 256, 2 * K * sizeof(float), st>>>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, K, rows_per_cta);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradient of a Linear layer:  dW[n][k] = sum_m dY[m][n] X[m][k],  db[n] = sum_m dY[m][n]
// 64x64 (n,k) tiles, the token axis split over gridDim.z with atomic accumulation into zeroed outputs
// ------------------------------------------------------------------------------------------------
#define WG_T 64
#define WG_M 16
__global__ void __launch_bounds__(256) linear_wgrad_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ X, int ldx,
                                                          float* __restrict__ dW, float* __restrict__ db, int M, int N, int K, int m_per_split) {
    __shared__ __align__(16) float As[WG_M][WG_T + 4];       // dY rows [m][n]
    __shared__ __align__(16) float Bs[WG_M][WG_T + 4];       // X rows  [m][k]
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * WG_T, k0 = blockIdx.y * WG_T;
    const int m_begin = blockIdx.z * m_per_split, m_end = min(M, m_begin + m_per_split);
    const int lr = tid >> 4, lc = (tid & 15) * 4;            // load: 16 rows x 16 float4 columns
    const int tx = tid & 15, ty = tid >> 4;                  // compute: 4x4 outputs per thread
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m0 = m_begin; m0 < m_end; m0 += WG_M) {
        const int m = m0 + lr;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < m_end) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (n0 + lc + j < N) a[j] = dY[(size_t)m * lddy + n0 + lc + j];
                if (k0 + lc + j < K) b[j] = X[(size_t)m * ldx + k0 + lc + j];
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { As[lr][lc + j] = a[j]; Bs[lr][lc + j] = b[j]; }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < WG_M; ++mm) {
            const float4 av = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[mm][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
            }
            if (tx == 0 && blockIdx.y == 0) { bsum[0] += ar[0]; bsum[1] += ar[1]; bsum[2] += ar[2]; bsum[3] += ar[3]; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(&dW[(size_t)n * K + k], acc[i][j]);
        }
        if (db && tx == 0 && blockIdx.y == 0) atomicAdd(&db[n], bsum[i]);
    }
}

static int linear_wgrad_run(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int N, int K,
                            bool accumulate, void* stream);

extern "C" int ps_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int N, int K,
                               void* stream) {
    return linear_wgrad_run(dY, lddy, X, ldx, dW, db, M, N, K, false, stream);
}

// dW += dY^T X, db += colsum(dY): the same kernel without clearing the outputs -- a weight that is used several times in one
// backward pass (the 8 roll-out steps of the training script share every predictor weight) accumulates straight into its
// gradient buffer instead of through one autograd add kernel per use
extern "C" int ps_linear_wgrad_acc(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int N, int K,
                                   void* stream) {
    return linear_wgrad_run(dY, lddy, X, ldx, dW, db, M, N, K, true, stream);
}

static int linear_wgrad_run(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int N, int K,
                            bool accumulate, void* stream) {
    PS_REQUIRE(dY && X && dW, "ps_linear_wgrad: null pointer");
    PS_REQUIRE(M >= 0 && N > 0 && K > 0, "ps_linear_wgrad: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) {
        PS_CUDA(cudaMemsetAsync(dW, 0, (size_t)N * K * sizeof(float), st));
        if (db) PS_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
    }
    if (M == 0) return PS_OK;
    const int tiles = ((N + WG_T - 1) / WG_T) * ((K + WG_T - 1) / WG_T);
    int splits = (2 * 148 + tiles - 1) / tiles;
    const int max_splits = (M + 4 * WG_M - 1) / (4 * WG_M);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int m_per_split = ((M + splits - 1) / splits + WG_M - 1) / WG_M * WG_M;
    splits = (M + m_per_split - 1) / m_per_split;
    dim3 grid((N + WG_T - 1) / WG_T, (K + WG_T - 1) / WG_T, splits);
    PS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "ps_linear_wgrad: grid too large");
    linear_wgrad_kernel<<<grid, 256, 0, st>>>(dY, lddy, X, ldx, dW, db, M, N, K, m_per_split);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// activations (act: 1 ReLU, 2 GELU exact-erf, 3 |x|)
// ------------------------------------------------------------------------------------------------
__global__ void act_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, long long n, int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = z[i];
        y[i] = act == 1 ? fmaxf(v, 0.f) : (act == 2 ? gelu_erf(v) : fabsf(v));
    }
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ dz, long long n, int act) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = z[i];
        float d;
        if (act == 1) d = v > 0.f ? 1.f : 0.f;
        else if (act == 2) d = 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v);
        else d = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
        dz[i] = dy[i] * d;
    }
}
static inline unsigned ew_blocks(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}
extern "C" int ps_act_fwd(const float* z, float* y, long long n, int act, void* stream) {
    PS_REQUIRE(z && y && act >= 1 && act <= 3, "ps_act_fwd: bad arguments");
    if (n == 0) return PS_OK;
    act_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(z, y, n, act);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
extern "C" int ps_act_bwd(const float* dy, const float* z, float* dz, long long n, int act, void* stream) {
    PS_REQUIRE(dy && z && dz && act >= 1 && act <= 3, "ps_act_bwd: bad arguments");
    if (n == 0) return PS_OK;
    act_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, z, dz, n, act);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-head self attention backward (sequences of at most 96 tokens); one CTA per (sequence, head)
// ------------------------------------------------------------------------------------------------
struct MhaBwdParams {
    const float* qkv; int ld;
    const float* dout; int ldo;
    float* dqkv;                 // same layout as qkv
    int n_seq, L, H, dh, spf;
    float scale;
};

__global__ void __launch_bounds__(256) mha_bwd_kernel(MhaBwdParams p) {
    extern __shared__ __align__(16) float sm[];
    const int seq = blockIdx.x, h = blockIdx.y;
    const int L = p.L, dh = p.dh, inner = p.H * dh;
    float* Qs = sm;                       // [L][dh+1]
    float* Ks = Qs + L * (dh + 1);
    float* Vs = Ks + L * (dh + 1);
    float* Ds = Vs + L * (dh + 1);        // dO
    float* Ps = Ds + L * (dh + 1);        // [L][L] probabilities
    float* Ss = Ps + L * L;               // [L][L] dS (already scaled)
    const float* base = p.qkv + (size_t)seq * L * p.ld;
    const float* dob = p.dout + (size_t)seq * L * p.ldo;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int idx = tid; idx < L * dh; idx += 256) {
        const int j = idx / dh, d = idx - j * dh;
        Qs[j * (dh + 1) + d] = base[(size_t)j * p.ld + h * dh + d];
        Ks[j * (dh + 1) + d] = base[(size_t)j * p.ld + inner + h * dh + d];
        Vs[j * (dh + 1) + d] = base[(size_t)j * p.ld + 2 * inner + h * dh + d];
        Ds[j * (dh + 1) + d] = dob[(size_t)j * p.ldo + h * dh + d];
    }
    __syncthreads();
    // phase 1: one warp per query row: P[i][:], dS[i][:] and dQ[i][:]
    for (int i = w; i < L; i += 8) {
        const float* qr = Qs + i * (dh + 1);
        const float* dr = Ds + i * (dh + 1);
        float mx = -INFINITY;
        for (int j = lane; j < L; j += 32) {
            const float* kr = Ks + j * (dh + 1);
            float dot = 0.f;
            for (int d = 0; d < dh; ++d) dot = fmaf(qr[d], kr[d], dot);
            dot *= p.scale;
            if (p.spf > 0 && (j / p.spf) > (i / p.spf)) dot = -1e9f;
            Ps[i * L + j] = dot;
            mx = fmaxf(mx, dot);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < L; j += 32) { const float e = expf(Ps[i * L + j] - mx); Ps[i * L + j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float delta = 0.f;
        for (int j = lane; j < L; j += 32) {
            const float pj = Ps[i * L + j] * inv;
            const float* vr = Vs + j * (dh + 1);
            float dp = 0.f;
            for (int d = 0; d < dh; ++d) dp = fmaf(dr[d], vr[d], dp);
            Ps[i * L + j] = pj;
            Ss[i * L + j] = dp;
            delta = fmaf(pj, dp, delta);
        }
        delta = warp_sum(delta);
        for (int j = lane; j < L; j += 32) {
            const bool masked = p.spf > 0 && (j / p.spf) > (i / p.spf);
            Ss[i * L + j] = masked ? 0.f : Ps[i * L + j] * (Ss[i * L + j] - delta) * p.scale;
        }
        __syncwarp();
        for (int d = lane; d < dh; d += 32) {
            float acc = 0.f;
            for (int j = 0; j < L; ++j) acc = fmaf(Ss[i * L + j], Ks[j * (dh + 1) + d], acc);
            p.dqkv[((size_t)seq * L + i) * p.ld + h * dh + d] = acc;
        }
    }
    __syncthreads();
    // phase 2: dK[j][d] = sum_i dS[i][j] Q[i][d],  dV[j][d] = sum_i P[i][j] dO[i][d]
    for (int idx = tid; idx < L * dh; idx += 256) {
        const int j = idx / dh, d = idx - j * dh;
        float dk = 0.f, dv = 0.f;
        for (int i = 0; i < L; ++i) {
            dk = fmaf(Ss[i * L + j], Qs[i * (dh + 1) + d], dk);
            dv = fmaf(Ps[i * L + j], Ds[i * (dh + 1) + d], dv);
        }
        p.dqkv[((size_t)seq * L + j) * p.ld + inner + h * dh + d] = dk;
        p.dqkv[((size_t)seq * L + j) * p.ld + 2 * inner + h * dh + d] = dv;
    }
}

extern "C" int ps_mha_bwd(const float* qkv, int ld, const float* dout, int ldo, float* dqkv, int n_seq, int L, int n_heads,
                          int head_dim, int slots_per_frame, void* stream) {
    PS_REQUIRE(qkv && dout && dqkv, "ps_mha_bwd: null pointer");
    PS_REQUIRE(n_seq >= 0 && L > 0 && n_heads > 0 && head_dim > 0, "ps_mha_bwd: bad shape");
    if (n_seq == 0) return PS_OK;
    MhaBwdParams p;
    p.qkv = qkv; p.ld = ld; p.dout = dout; p.ldo = ldo; p.dqkv = dqkv; p.n_seq = n_seq; p.L = L; p.H = n_heads; p.dh = head_dim;
    p.spf = slots_per_frame; p.scale = 1.0f / sqrtf((float)head_dim);
    const size_t smem = (size_t)(4 * L * (head_dim + 1) + 2 * L * L) * sizeof(float);
    PS_REQUIRE(smem <= 200 * 1024, "ps_mha_bwd: %d tokens x head_dim %d exceed the shared-memory budget of the backward kernel", L, head_dim);
    static PsSmemConfig smem_cfg = {};
    if (smem > 48 * 1024) PS_CUDA(ps_ensure_smem(mha_bwd_kernel, smem, smem_cfg, 200 * 1024));
    PS_REQUIRE(n_heads <= 65535, "ps_mha_bwd: too many heads");
    mha_bwd_kernel<<<dim3(n_seq, n_heads), 256, smem, (cudaStream_t)stream>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// InvDyn tail backward.  Forward (invdyn.cu): m = mu[t+1]-mu[t], v = |nu[t+1]|+|nu[t]|, z = noise*sqrt(v+eps)+m,
// proto = z + sg(zq - z), variability = z - zq (single-action model; 0-gradient in the per-slot model),
// quant = mean_A (sg(zq) - z)^2.  Gradients flow to mv = [mu | nu] only (codebook handled by the caller).
// ------------------------------------------------------------------------------------------------
struct TailBwdParams {
    const float* mv; const float* noise; const float* z; const long long* idx; const float* codebook;
    const float* g_dist; const float* g_z; const float* g_proto; const float* g_var; const float* g_quant;   // any may be null
    float* d_mv;
    int B, T, R, A, K, per_slot;
    float eps;
};

__device__ __forceinline__ void tail_grads(const TailBwdParams& p, int b, int t, int r, int a, float& dm, float& dv) {
    const int A = p.A;
    const size_t n = ((size_t)b * (p.T - 1) + t) * p.R + r;
    const float* mv0 = p.mv + (((size_t)b * p.T + t) * p.R + r) * 2 * A;
    const float* mv1 = p.mv + (((size_t)b * p.T + t + 1) * p.R + r) * 2 * A;
    const float v = fabsf(mv1[A + a]) + fabsf(mv0[A + a]);
    float gz = 0.f;
    if (p.g_z) gz += p.g_z[n * A + a];
    if (p.g_proto) gz += p.g_proto[n * A + a];
    if (p.g_var && !p.per_slot) gz += p.g_var[n * A + a];
    if (p.g_quant) {
        const float zq = p.codebook[(size_t)p.idx[n] * A + a];
        gz += p.g_quant[n] * 2.0f * (p.z[n * A + a] - zq) / (float)A;
    }
    dm = gz; dv = gz * p.noise[n * A + a] * 0.5f / sqrtf(v + p.eps);
    if (p.g_dist) {
        dm += p.g_dist[((((size_t)b * (p.T - 1) + t) * 2 + 0) * p.R + r) * A + a];
        dv += p.g_dist[((((size_t)b * (p.T - 1) + t) * 2 + 1) * p.R + r) * A + a];
    }
}

__global__ void invdyn_tail_bwd_kernel(TailBwdParams p) {
    const long long total = (long long)p.B * p.T * p.R * p.A;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int a = (int)(i % p.A);
    const int r = (int)((i / p.A) % p.R);
    const int t = (int)((i / ((long long)p.A * p.R)) % p.T);
    const int b = (int)(i / ((long long)p.A * p.R * p.T));
    float dmu = 0.f, dnu = 0.f;
    const float raw = p.mv[(((size_t)b * p.T + t) * p.R + r) * 2 * p.A + p.A + a];
    const float sgn = raw > 0.f ? 1.f : (raw < 0.f ? -1.f : 0.f);
    if (t >= 1) { float dm, dv; tail_grads(p, b, t - 1, r, a, dm, dv); dmu += dm; dnu += dv * sgn; }      // frame t is the "next" frame of step t-1
    if (t < p.T - 1) { float dm, dv; tail_grads(p, b, t, r, a, dm, dv); dmu -= dm; dnu += dv * sgn; }       // and the "current" frame of step t
    float* o = p.d_mv + (((size_t)b * p.T + t) * p.R + r) * 2 * p.A;
    o[a] = dmu; o[p.A + a] = dnu;
}

extern "C" int ps_invdyn_tail_bwd(const float* mv, const float* noise, const float* z, const long long* idx, const float* codebook,
                                  const float* g_dist, const float* g_z, const float* g_proto, const float* g_var,
                                  const float* g_quant, float* d_mv, int B, int T, int R, int A, int K, int per_slot,
                                  float eps, void* stream) {
    PS_REQUIRE(mv && noise && z && idx && codebook && d_mv, "ps_invdyn_tail_bwd: null pointer");
    PS_REQUIRE(T >= 2 && R >= 1 && A >= 1, "ps_invdyn_tail_bwd: bad shape");
    const long long total = (long long)B * T * R * A;
    if (total == 0) return PS_OK;
    TailBwdParams p;
    p.mv = mv; p.noise = noise; p.z = z; p.idx = idx; p.codebook = codebook; p.g_dist = g_dist; p.g_z = g_z; p.g_proto = g_proto;
    p.g_var = g_var; p.g_quant = g_quant; p.d_mv = d_mv; p.B = B; p.T = T; p.R = R; p.A = A; p.K = K; p.per_slot = per_slot; p.eps = eps;
    invdyn_tail_bwd_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// decoder data gradient, compositing stage (SAVi.py:226-234): recon = sum_s rgb_s mask_s, mask = softmax_s(alpha)
//   g_rgb = g_recons + g_recon * mask;  g_mask = g_masks + sum_c g_recon_c rgb_c;  g_alpha = mask * (g_mask - sum_s mask g_mask)
// followed by the data gradient of the last 3x3 convolution (64 <- 4 channels) and of the ReLU in front of it,
// written in the zero-bordered NHWC split format consumed by the tcgen05 convolution kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_composite_bwd_kernel(const float* __restrict__ recons, const float* __restrict__ masks,
                                                               const float* __restrict__ g_recon, const float* __restrict__ g_recons,
                                                               const float* __restrict__ g_masks, float* __restrict__ g_o,
                                                               int M, int S, int HW) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * HW) return;
    const int m = (int)(i / HW), px = (int)(i % HW);
    float gr[3] = {0.f, 0.f, 0.f};
    if (g_recon)
        for (int c = 0; c < 3; ++c) gr[c] = g_recon[((size_t)m * 3 + c) * HW + px];
    float dotsum = 0.f;
    for (int s = 0; s < S; ++s) {
        const size_t ms = (size_t)m * S + s;
        float gm = g_masks ? g_masks[ms * HW + px] : 0.f;
        for (int c = 0; c < 3; ++c) gm = fmaf(gr[c], recons[(ms * 3 + c) * HW + px], gm);
        dotsum = fmaf(masks[ms * HW + px], gm, dotsum);
    }
    for (int s = 0; s < S; ++s) {
        const size_t ms = (size_t)m * S + s;
        const float mk = masks[ms * HW + px];
        float gm = g_masks ? g_masks[ms * HW + px] : 0.f;
        for (int c = 0; c < 3; ++c) {
            gm = fmaf(gr[c], recons[(ms * 3 + c) * HW + px], gm);
            g_o[(ms * 4 + c) * HW + px] = (g_recons ? g_recons[(ms * 3 + c) * HW + px] : 0.f) + gr[c] * mk;
        }
        g_o[(ms * 4 + 3) * HW + px] = mk * (gm - dotsum);
    }
}

// dh[img][ci][y][x] = sum_{co<4,ky,kx} w[co][ci][ky][kx] * g_o[img][co][y+1-ky][x+1-kx], masked by relu'(h), written as
// padded NHWC split (hi | lo) 16-bit planes.  w_packed is the forward packing [Cin=64][9][4].
__global__ void __launch_bounds__(256) dec_final_dgrad_kernel(const float* __restrict__ g_o, const float* __restrict__ w_packed,
                                                             const float* __restrict__ h, const uint16_t* __restrict__ h_hi,
                                                             uint16_t* __restrict__ out_hi,
                                                             uint16_t* __restrict__ out_lo, const float* __restrict__ g_scale,
                                                             int n_img, int H, int W) {
    // One thread per pixel, all 64 input channels: the 36 incoming gradient values of the pixel's 3x3 neighbourhood live in
    // registers, the weights are broadcast from shared memory (every lane reads the same float4), the ReLU mask reads of `h`
    // are coalesced along x, and a thread writes whole 32-byte sectors of its 128-byte NHWC rows.
    __shared__ __align__(16) float s_w[64 * 9 * 4];
    const float gs = g_scale ? *g_scale : 1.0f;          // power-of-two gradient scale (keeps the fp16 split in range)
    for (int i = threadIdx.x; i < 64 * 36; i += 256) s_w[i] = w_packed[i] * gs;
    __syncthreads();
    const int HW = H * W, Wp = W + 4, Hp = H + 4;
    const long long total = (long long)n_img * HW;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(pix % W), y = (int)((pix / W) % H);
        const int img = (int)(pix / HW);
        float4 go[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int yy = y + 1 - ky, xx = x + 1 - kx;
                const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
                const float* gp = g_o + (size_t)img * 4 * HW + yy * W + xx;
                go[ky * 3 + kx] = ok ? make_float4(gp[0], gp[HW], gp[2 * HW], gp[3 * HW]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        const float* hp = h ? h + (size_t)img * 64 * HW + y * W + x : nullptr;
        const size_t o = (((size_t)img * Hp + y + 2) * Wp + x + 2) * 64;
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4* wr = reinterpret_cast<const float4*>(s_w + (c0 + j) * 36);
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const float4 w4 = wr[t];
                    acc = fmaf(w4.x, go[t].x, acc); acc = fmaf(w4.y, go[t].y, acc);
                    acc = fmaf(w4.z, go[t].z, acc); acc = fmaf(w4.w, go[t].w, acc);
                }
                v[j] = acc;
                if (hp) v[j] = hp[(size_t)(c0 + j) * HW] > 0.f ? acc : 0.f;
            }
            if (!hp) {                 // saved activation as split planes (same position as the output row): pass where its hi part is positive
                const uint4* mk = reinterpret_cast<const uint4*>(h_hi + o + c0);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint4 m4 = __ldg(mk + g);
                    const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if ((mw[j] & 0x7FFFu) == 0u || (mw[j] & 0x8000u)) v[g * 8 + 2 * j] = 0.f;
                        if ((mw[j] & 0x7FFF0000u) == 0u || (mw[j] & 0x80000000u)) v[g * 8 + 2 * j + 1] = 0.f;
                    }
                }
            }
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                const float2 f = __half22float2(hh);
                const __half2 ll = __floats2half2_rn((v[2 * j] - f.x) * 2048.0f, (v[2 * j + 1] - f.y) * 2048.0f);
                hw[j] = *reinterpret_cast<const uint32_t*>(&hh); lw[j] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(out_hi + o + c0), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]),
                         "r"(hw[3]), "r"(hw[4]), "r"(hw[5]), "r"(hw[6]), "r"(hw[7]) : "memory");
            if (out_lo)
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(out_lo + o + c0), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]),
                             "r"(lw[3]), "r"(lw[4]), "r"(lw[5]), "r"(lw[6]), "r"(lw[7]) : "memory");
        }
    }
}

extern "C" int ps_dec_composite_bwd(const float* recons, const float* masks, const float* g_recon, const float* g_recons,
                                    const float* g_masks, float* g_o, int M, int S, int H, int W, void* stream) {
    PS_REQUIRE(recons && masks && g_o, "ps_dec_composite_bwd: null pointer");
    PS_REQUIRE(g_recon || g_recons || g_masks, "ps_dec_composite_bwd: no incoming gradient");
    const long long total = (long long)M * H * W;
    if (total == 0) return PS_OK;
    dec_composite_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(recons, masks, g_recon, g_recons, g_masks, g_o, M, S, H * W);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

static int dec_final_dgrad_run(const float* g_o, const float* w_packed, const float* h, const void* h_planes, void* act_out,
                               const float* g_scale, int n_img, int H, int W, int precision, void* stream);

extern "C" int ps_dec_final_dgrad(const float* g_o, const float* w_packed, const float* h, void* act_out, const float* g_scale,
                                  int n_img, int H, int W, int precision, void* stream) {
    PS_REQUIRE(h, "ps_dec_final_dgrad: null pointer");
    return dec_final_dgrad_run(g_o, w_packed, h, nullptr, act_out, g_scale, n_img, H, W, precision, stream);
}

// the same with the saved activation given as the padded split planes the tensor-core layers write (its hi plane carries the sign):
// the forward of the predictor's training step then never materialises the 64-channel fp32 NCHW activation
extern "C" int ps_dec_final_dgrad_planes(const float* g_o, const float* w_packed, const void* h_planes, void* act_out,
                                         const float* g_scale, int n_img, int H, int W, int precision, void* stream) {
    PS_REQUIRE(h_planes && ps_aligned16(h_planes), "ps_dec_final_dgrad_planes: null / unaligned planes");
    return dec_final_dgrad_run(g_o, w_packed, nullptr, h_planes, act_out, g_scale, n_img, H, W, precision, stream);
}

static int dec_final_dgrad_run(const float* g_o, const float* w_packed, const float* h, const void* h_planes, void* act_out,
                               const float* g_scale, int n_img, int H, int W, int precision, void* stream) {
    PS_REQUIRE(g_o && w_packed && act_out, "ps_dec_final_dgrad: null pointer");
    PS_REQUIRE(precision == 2 || precision == 3, "ps_dec_final_dgrad: precision must be 2 (fp16) or 3 (fp16x3)");
    if (n_img == 0) return PS_OK;
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    uint16_t* hi = reinterpret_cast<uint16_t*>(act_out);
    uint16_t* lo = precision == 3 ? hi + P * 64 : nullptr;
    const long long total = (long long)n_img * H * W;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    dec_final_dgrad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g_o, w_packed, h, reinterpret_cast<const uint16_t*>(h_planes), hi, lo,
                                                                               g_scale, n_img, H, W);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// gradient of the collapsed first decoder convolution: dG[img][cls(y)][cls(x)][c] = sum over the pixels of the border
// class of g[img][y][x][c] (g: fp32 NHWC tokens, already masked by relu').  One CTA per image.
__global__ void __launch_bounds__(256) dec_expand_bwd_kernel(const float* __restrict__ g, float* __restrict__ dG, const float* __restrict__ g_scale,
                                                            int H, int W, int pad) {
    // thread = (4 consecutive channels, column lane 0..15): float4 loads, the columns x = lane, lane + 16, ... of every row; interior
    // pixels of interior rows (most of the image) accumulate in registers, the border classes go through shared-memory atomics
    // (the first version walked the image with one 4-byte load and one dependent add per step: 470 us per 256 images)
    extern __shared__ float s_acc[];                 // [ncls*ncls][64]
    const int ncls = 2 * pad + 1;
    for (int i = threadIdx.x; i < ncls * ncls * 64; i += 256) s_acc[i] = 0.f;
    __syncthreads();
    const int img = blockIdx.x;
    const int c4 = (threadIdx.x & 15) * 4, xl = threadIdx.x >> 4;
    const float* gi = g + (size_t)img * H * W * 64 + c4;
    float4 mid_mid = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = 0; y < H; ++y) {
        const int cy = y < pad ? y : (y > H - 1 - pad ? 2 * pad - (H - 1 - y) : pad);
        float4 mid = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int x = xl; x < W; x += 16) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(gi + ((size_t)y * W + x) * 64));
            if (x < pad || x > W - 1 - pad) {
                const int cx = x < pad ? x : 2 * pad - (W - 1 - x);
                float* d = &s_acc[(cy * ncls + cx) * 64 + c4];
                atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w);
            } else {
                mid.x += v.x; mid.y += v.y; mid.z += v.z; mid.w += v.w;
            }
        }
        if (cy == pad) {
            mid_mid.x += mid.x; mid_mid.y += mid.y; mid_mid.z += mid.z; mid_mid.w += mid.w;
        } else {
            float* d = &s_acc[(cy * ncls + pad) * 64 + c4];
            atomicAdd(d, mid.x); atomicAdd(d + 1, mid.y); atomicAdd(d + 2, mid.z); atomicAdd(d + 3, mid.w);
        }
    }
    {
        float* d = &s_acc[(pad * ncls + pad) * 64 + c4];
        atomicAdd(d, mid_mid.x); atomicAdd(d + 1, mid_mid.y); atomicAdd(d + 2, mid_mid.z); atomicAdd(d + 3, mid_mid.w);
    }
    __syncthreads();
    const float inv = g_scale ? 1.0f / *g_scale : 1.0f;
    for (int i = threadIdx.x; i < ncls * ncls * 64; i += 256) dG[(size_t)img * ncls * ncls * 64 + i] = s_acc[i] * inv;
}

extern "C" int ps_dec_expand_bwd(const float* g_tokens, float* dG, const float* g_scale, int n_img, int H, int W, int ksize, void* stream) {
    PS_REQUIRE(g_tokens && dG, "ps_dec_expand_bwd: null pointer");
    PS_REQUIRE(ksize % 2 == 1 && H > ksize && W > ksize, "ps_dec_expand_bwd: unsupported geometry");
    if (n_img == 0) return PS_OK;
    const int ncls = ksize;
    dec_expand_bwd_kernel<<<n_img, 256, ncls * ncls * 64 * sizeof(float), (cudaStream_t)stream>>>(g_tokens, dG, g_scale, H, W, ksize / 2);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
