// InvDyn tail (reference: src/models/Predictors/LatentAction.py:72-87,228-249,340-359 and
// BlocksUtils/quantization.py:18-20,256-322,338-372): temporal difference of the mean / |variance| heads,
// Gaussian sample with host-provided noise, nearest-codeword VQ (first-index tie break), straight-through
// prototype value, variability embedding and the per-vector quantisation loss.  Always fp32.
#include "common.cuh"
#include "../../include/playslot_b200.h"

#define VQ_MAX_A 64
#define VQ_MAX_K 64

struct TailParams {
    const float* mv;        // [B][T][R][2A]  (mean | raw variance head)
    const float* noise;     // [B][T-1][R][A]
    const float* codebook;  // [K][A]
    float* dist_out;        // [B][T-1][2][R][A]   (action_dist: mean, var)
    float* z_out;           // [B][T-1][R][A]      sampled latent action
    float* proto_out;       // [B][T-1][R][A]      z + (z_q - z)
    float* var_out;         // [B][T-1][R][A]      variability
    long long* idx_out;     // [B][T-1][R]
    float* loss_out;        // [B][T-1][R]         mean_A (z_q - z)^2
    int B, T, R, A, K;
    int per_slot;
    float eps;
};

__global__ void invdyn_tail_kernel(TailParams p) {
    __shared__ float s_cb[VQ_MAX_K * VQ_MAX_A];
    __shared__ float s_cn[VQ_MAX_K];
    for (int i = threadIdx.x; i < p.K * p.A; i += blockDim.x) s_cb[i] = p.codebook[i];
    __syncthreads();
    for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
        float s = 0.f;
        for (int a = 0; a < p.A; ++a) s += s_cb[k * p.A + a] * s_cb[k * p.A + a];
        s_cn[k] = s;
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = p.B * (p.T - 1) * p.R;
    if (n >= total) return;
    const int r = n % p.R;
    const int t = (n / p.R) % (p.T - 1);
    const int b = n / (p.R * (p.T - 1));
    const int A = p.A;
    const float* mv0 = p.mv + (((size_t)b * p.T + t) * p.R + r) * 2 * A;
    const float* mv1 = p.mv + (((size_t)b * p.T + t + 1) * p.R + r) * 2 * A;
    const float* nz = p.noise + (size_t)n * A;
    float z[VQ_MAX_A];
    float zn = 0.f;
    float* dm = p.dist_out + ((((size_t)b * (p.T - 1) + t) * 2 + 0) * p.R + r) * A;
    float* dv = p.dist_out + ((((size_t)b * (p.T - 1) + t) * 2 + 1) * p.R + r) * A;
#pragma unroll 4
    for (int a = 0; a < A; ++a) {
        const float m = mv1[a] - mv0[a];
        const float v = fabsf(mv1[A + a]) + fabsf(mv0[A + a]);
        dm[a] = m; dv[a] = v;
        const float zz = nz[a] * sqrtf(v + p.eps) + m;
        z[a] = zz;
        zn += zz * zz;
        p.z_out[(size_t)n * A + a] = zz;
    }
    int best = 0;
    float bestd = INFINITY;
    for (int k = 0; k < p.K; ++k) {
        float dot = 0.f;
#pragma unroll 4
        for (int a = 0; a < A; ++a) dot = fmaf(z[a], s_cb[k * A + a], dot);
        const float d = zn - 2.0f * dot + s_cn[k];
        if (d < bestd) { bestd = d; best = k; }
    }
    float loss = 0.f;
#pragma unroll 4
    for (int a = 0; a < A; ++a) {
        const float zq = s_cb[best * A + a];
        const float diff = zq - z[a];
        const float proto = z[a] + diff;
        p.proto_out[(size_t)n * A + a] = proto;
        p.var_out[(size_t)n * A + a] = p.per_slot ? (z[a] - proto) : (z[a] - zq);
        loss += diff * diff;
    }
    p.idx_out[n] = best;
    p.loss_out[n] = loss / (float)A;
}

extern "C" int ps_invdyn_tail(const float* mv, const float* noise, const float* codebook, float* dist_out,
                              float* z_out, float* proto_out, float* var_out, long long* idx_out, float* loss_out,
                              int B, int T, int R, int A, int K, int per_slot, float eps, void* stream) {
    PS_REQUIRE(mv && noise && codebook && dist_out && z_out && proto_out && var_out && idx_out && loss_out,
               "ps_invdyn_tail: null pointer");
    PS_REQUIRE(A >= 1 && A <= VQ_MAX_A && K >= 1 && K <= VQ_MAX_K, "ps_invdyn_tail: action_dim <= %d and num_actions <= %d", VQ_MAX_A, VQ_MAX_K);
    PS_REQUIRE(T >= 2 && R >= 1 && B >= 0, "ps_invdyn_tail: need at least two frames");
    const int total = B * (T - 1) * R;
    if (total == 0) return PS_OK;
    TailParams p;
    p.mv = mv; p.noise = noise; p.codebook = codebook; p.dist_out = dist_out; p.z_out = z_out; p.proto_out = proto_out;
    p.var_out = var_out; p.idx_out = idx_out; p.loss_out = loss_out;
    p.B = B; p.T = T; p.R = R; p.A = A; p.K = K; p.per_slot = per_slot; p.eps = eps;
    invdyn_tail_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone VQ lookup
// ------------------------------------------------------------------------------------------------
__global__ void vq_lookup_kernel(const float* __restrict__ z, const float* __restrict__ codebook, float* proto_out,
                                 float* zq_out, long long* idx_out, float* loss_out, int n, int A, int K) {
    __shared__ float s_cb[VQ_MAX_K * VQ_MAX_A];
    __shared__ float s_cn[VQ_MAX_K];
    for (int i = threadIdx.x; i < K * A; i += blockDim.x) s_cb[i] = codebook[i];
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float s = 0.f;
        for (int a = 0; a < A; ++a) s += s_cb[k * A + a] * s_cb[k * A + a];
        s_cn[k] = s;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* zi = z + (size_t)i * A;
    float zn = 0.f;
    for (int a = 0; a < A; ++a) zn += zi[a] * zi[a];
    int best = 0;
    float bestd = INFINITY;
    for (int k = 0; k < K; ++k) {
        float dot = 0.f;
        for (int a = 0; a < A; ++a) dot = fmaf(zi[a], s_cb[k * A + a], dot);
        const float d = zn - 2.0f * dot + s_cn[k];
        if (d < bestd) { bestd = d; best = k; }
    }
    float loss = 0.f;
    for (int a = 0; a < A; ++a) {
        const float zq = s_cb[best * A + a];
        const float diff = zq - zi[a];
        proto_out[(size_t)i * A + a] = zi[a] + diff;
        zq_out[(size_t)i * A + a] = zq;
        loss += diff * diff;
    }
    idx_out[i] = best;
    loss_out[i] = loss / (float)A;
}

extern "C" int ps_vq_lookup(const float* z, const float* codebook, float* proto_out, float* zq_out,
                            long long* idx_out, float* loss_out, int n, int A, int K, void* stream) {
    PS_REQUIRE(z && codebook && proto_out && zq_out && idx_out && loss_out, "ps_vq_lookup: null pointer");
    PS_REQUIRE(A >= 1 && A <= VQ_MAX_A && K >= 1 && K <= VQ_MAX_K, "ps_vq_lookup: emb_dim <= %d and num_embs <= %d", VQ_MAX_A, VQ_MAX_K);
    if (n == 0) return PS_OK;
    vq_lookup_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(z, codebook, proto_out, zq_out, idx_out, loss_out, n, A, K);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
