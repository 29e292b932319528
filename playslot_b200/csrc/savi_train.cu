// Backward / training kernels of the SAVi decomposition model (src/02_train_savi.py:21-70: model(videos) -> MSE(recons, videos)
// -> loss.backward() through decoder, SlotAttention, transition block, encoder).  fp32 CUDA-core kernels: correctness first --
// the forward and the data gradients of the 5x5 convolutions stay on the tcgen05 kernels (conv_tc.cu); what is added here are
// the pieces that had no backward yet:
//   ps_sa_core_fwd / ps_sa_core_bwd : the attention core of one SlotAttention iteration (attention.py:97-110) on pre-projected
//                                     queries, with gradients w.r.t. the queries and the normalised features
//   ps_gru_gates_fwd / _bwd         : the element-wise part of nn.GRUCell (attention.py:112-115)
//   ps_conv5x5_planes_wgrad         : weight gradient of a 64 -> 64 5x5 convolution from padded NHWC split planes
//   ps_planes_colsum                : bias gradient (sum over positions) of a gradient in plane format
//   ps_conv_first_wgrad             : weight gradient of the first encoder convolution (Cin <= 4, fp32 NCHW images)
//   ps_conv3x3_final_wgrad          : weight / bias gradient of the decoder's last 3x3 convolution (64 -> 4)
//   ps_colsum                       : column sums of a row-major matrix (position-embedding / broadcast-constant gradients)
//   ps_nhwc_to_split_scaled         : fp32 NHWC gradient tokens -> padded split planes with a power-of-two scale
#include "common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int SC_D = 128;          // feature / slot width of the attention core
constexpr int SC_SMAX = 16;        // at most 16 slots
constexpr float LO_INV = 1.0f / 2048.0f;

// ------------------------------------------------------------------------------------------------ SlotAttention core
// dots[s][n] = A[s] . xn[n]   (A already carries the D^-1/2 scale and the key projection: A = scale * q W_k)
// p = softmax over s;  a = p + eps;  Z[s] = sum_n a[s][n];  t[s] = sum_n a[s][n] xn[n] / Z[s]
// One CTA per (frame, token chunk), 64 tokens at a time through shared memory:
//   phase 1: 4 lanes per token (32 features each) -> logits, softmax over the slots -> a[n][s] in shared memory
//   phase 2: thread = (feature f, slot half) accumulates sum_n a[n][s] xn[n][f] in registers      (no atomics: deterministic)
constexpr int SC_TB = 64;                         // tokens per block
constexpr int SC_XS = 132;                        // padded token row: quarter q of a row starts at q * 33 (conflict-free both ways)
__device__ __forceinline__ int sc_xi(int n, int f) { return n * SC_XS + (f >> 5) * 33 + (f & 31); }

__device__ __forceinline__ void sc_load_tile(float* s_x, const float* __restrict__ xn, size_t row0, int n_valid) {
    // 64 tokens x 128 features, coalesced float4 loads, scattered into the padded layout
    for (int i = threadIdx.x; i < SC_TB * 32; i += 256) {
        const int n = i >> 5, c4 = i & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < n_valid) v = *reinterpret_cast<const float4*>(xn + (row0 + n) * SC_D + c4 * 4);
        float* d = s_x + sc_xi(n, c4 * 4);
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}

__global__ void __launch_bounds__(256) sa_core_fwd_kernel(const float* __restrict__ A, const float* __restrict__ xn, float* __restrict__ part_t,
                                                         float* __restrict__ part_z, float* __restrict__ masks, int S, int N, int n_chunks,
                                                         int tok_per_chunk, float eps) {
    const int b = blockIdx.y, ch = blockIdx.x;
    extern __shared__ float sm[];
    float* s_x = sm;                               // [64][132]
    float* s_A = s_x + SC_TB * SC_XS;              // [S][128]
    float* s_a = s_A + SC_SMAX * SC_D;             // [64][16]
    for (int i = threadIdx.x; i < S * SC_D; i += 256) s_A[i] = A[(size_t)b * S * SC_D + i];
    const int q4 = threadIdx.x & 3, tl = threadIdx.x >> 2;
    const int f = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int s_half = (S + 1) >> 1, s_lo = half * s_half, s_hi = min(S, s_lo + s_half);
    float acc[SC_SMAX / 2], accz[SC_SMAX / 2];
#pragma unroll
    for (int i = 0; i < SC_SMAX / 2; ++i) { acc[i] = 0.f; accz[i] = 0.f; }
    const int n0 = ch * tok_per_chunk, n1 = min(N, n0 + tok_per_chunk);
    for (int nb = n0; nb < n1; nb += SC_TB) {
        const int nv = min(SC_TB, n1 - nb);
        __syncthreads();
        sc_load_tile(s_x, xn, (size_t)b * N + nb, nv);
        __syncthreads();
        {   // phase 1
            float d[SC_SMAX];
            float mx = -3.0e38f;
            const float* xr = s_x + tl * SC_XS + q4 * 33;
#pragma unroll
            for (int s = 0; s < SC_SMAX; ++s) {
                if (s < S) {
                    float v = 0.f;
                    const float* a = s_A + s * SC_D + q4 * 32;
#pragma unroll 8
                    for (int j = 0; j < 32; ++j) v = fmaf(a[j], xr[j], v);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    d[s] = v;
                    mx = fmaxf(mx, v);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int s = 0; s < SC_SMAX; ++s) if (s < S) { d[s] = expf(d[s] - mx); sum += d[s]; }
            const float inv = 1.0f / sum;
            if (q4 == 0) {
#pragma unroll
                for (int s = 0; s < SC_SMAX; ++s) {
                    if (s < S) {
                        const float a = tl < nv ? d[s] * inv + eps : 0.f;
                        s_a[tl * SC_SMAX + s] = a;
                        if (masks && tl < nv) masks[((size_t)b * S + s) * N + nb + tl] = a;
                    }
                }
            }
        }
        __syncthreads();
        // phase 2
        for (int n = 0; n < nv; ++n) {
            const float xv = s_x[sc_xi(n, f)];
#pragma unroll
            for (int i = 0; i < SC_SMAX / 2; ++i) {
                if (s_lo + i < s_hi) {
                    const float a = s_a[n * SC_SMAX + s_lo + i];
                    acc[i] = fmaf(a, xv, acc[i]);
                    accz[i] += a;
                }
            }
        }
    }
    float* pt = part_t + ((size_t)b * n_chunks + ch) * S * SC_D;
#pragma unroll
    for (int i = 0; i < SC_SMAX / 2; ++i) {
        if (s_lo + i < s_hi) {
            pt[(s_lo + i) * SC_D + f] = acc[i];
            if (f == 0) part_z[((size_t)b * n_chunks + ch) * S + s_lo + i] = accz[i];
        }
    }
}

// t[b][s] = sum_chunks part_t / sum_chunks part_z;  Z[b][s] = sum_chunks part_z      (deterministic: fixed chunk order)
__global__ void sa_core_finish_kernel(const float* __restrict__ part_t, const float* __restrict__ part_z, float* __restrict__ t,
                                      float* __restrict__ Z, int S, int n_chunks) {
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < S * SC_D; i += blockDim.x) {
        const int s = i / SC_D;
        float z = 0.f, acc = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
            z += part_z[((size_t)b * n_chunks + c) * S + s];
            acc += part_t[((size_t)b * n_chunks + c) * S * SC_D + i];
        }
        t[(size_t)b * S * SC_D + i] = acc / z;
        if ((i % SC_D) == 0) Z[b * S + s] = z;
    }
}

// backward: given dt (B,S,D) (gradient of the normalised weighted mean), A, xn, t, Z:
//   dw[s][n] = dt[s].xn[n];  r[s] = dt[s].t[s];  da = (dw - r) / Z;  dd = p (da - sum_s' p da)
//   dA[s] += dd[s][n] xn[n];   dxn[n] = sum_s ( (a[s][n]/Z[s]) dt[s] + dd[s][n] A[s] )
// dxn is ACCUMULATED (+=) when accumulate != 0 (the iterations of one forward sum into the same buffer).
__global__ void __launch_bounds__(256) sa_core_bwd_kernel(const float* __restrict__ A, const float* __restrict__ xn, const float* __restrict__ t,
                                                         const float* __restrict__ Z, const float* __restrict__ dt, float* __restrict__ dxn,
                                                         float* __restrict__ part_dA, int S, int N, int n_chunks, int tok_per_chunk,
                                                         float eps, int accumulate) {
    const int b = blockIdx.y, ch = blockIdx.x;
    extern __shared__ float sm[];
    float* s_x = sm;                               // [64][132]
    float* s_A = s_x + SC_TB * SC_XS;              // [16][128]
    float* s_dt = s_A + SC_SMAX * SC_D;            // [16][128]
    float* s_dd = s_dt + SC_SMAX * SC_D;           // [64][16] logit gradients
    float* s_w = s_dd + SC_TB * SC_SMAX;           // [64][16] normalised attention weights
    float* s_r = s_w + SC_TB * SC_SMAX;            // [16]
    float* s_iz = s_r + SC_SMAX;                   // [16]
    for (int i = threadIdx.x; i < S * SC_D; i += 256) {
        s_A[i] = A[(size_t)b * S * SC_D + i];
        s_dt[i] = dt[(size_t)b * S * SC_D + i];
    }
    __syncthreads();
    if (threadIdx.x < S) {
        float r = 0.f;
        for (int j = 0; j < SC_D; ++j) r = fmaf(s_dt[threadIdx.x * SC_D + j], t[((size_t)b * S + threadIdx.x) * SC_D + j], r);
        s_r[threadIdx.x] = r;
        s_iz[threadIdx.x] = 1.0f / Z[b * S + threadIdx.x];
    }
    const int q4 = threadIdx.x & 3, tl = threadIdx.x >> 2;
    const int f = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int s_half = (S + 1) >> 1, s_lo = half * s_half, s_hi = min(S, s_lo + s_half);
    float accA[SC_SMAX / 2];
#pragma unroll
    for (int i = 0; i < SC_SMAX / 2; ++i) accA[i] = 0.f;
    const int n0 = ch * tok_per_chunk, n1 = min(N, n0 + tok_per_chunk);
    for (int nb = n0; nb < n1; nb += SC_TB) {
        const int nv = min(SC_TB, n1 - nb);
        __syncthreads();
        sc_load_tile(s_x, xn, (size_t)b * N + nb, nv);
        __syncthreads();
        {   // phase 1: logits, softmax, logit gradient per token
            float d[SC_SMAX], dw[SC_SMAX];
            float mx = -3.0e38f;
            const float* xr = s_x + tl * SC_XS + q4 * 33;
#pragma unroll
            for (int s = 0; s < SC_SMAX; ++s) {
                if (s < S) {
                    float v = 0.f, w = 0.f;
                    const float* a = s_A + s * SC_D + q4 * 32;
                    const float* g = s_dt + s * SC_D + q4 * 32;
#pragma unroll 8
                    for (int j = 0; j < 32; ++j) { v = fmaf(a[j], xr[j], v); w = fmaf(g[j], xr[j], w); }
                    v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
                    w += __shfl_xor_sync(0xffffffffu, w, 1); w += __shfl_xor_sync(0xffffffffu, w, 2);
                    d[s] = v; dw[s] = w;
                    mx = fmaxf(mx, v);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int s = 0; s < SC_SMAX; ++s) if (s < S) { d[s] = expf(d[s] - mx); sum += d[s]; }
            const float inv = 1.0f / sum;
            float pda = 0.f;
#pragma unroll
            for (int s = 0; s < SC_SMAX; ++s) {
                if (s < S) {
                    d[s] *= inv;                                       // p
                    dw[s] = (dw[s] - s_r[s]) * s_iz[s];                // da
                    pda = fmaf(d[s], dw[s], pda);
                }
            }
            if (q4 == 0) {
#pragma unroll
                for (int s = 0; s < SC_SMAX; ++s) {
                    if (s < S) {
                        const bool ok = tl < nv;
                        s_dd[tl * SC_SMAX + s] = ok ? d[s] * (dw[s] - pda) : 0.f;
                        s_w[tl * SC_SMAX + s] = ok ? (d[s] + eps) * s_iz[s] : 0.f;
                    }
                }
            }
        }
        __syncthreads();
        // phase 2a: dA[s][f] += sum_n dd[n][s] x[n][f]
        for (int n = 0; n < nv; ++n) {
            const float xv = s_x[sc_xi(n, f)];
#pragma unroll
            for (int i = 0; i < SC_SMAX / 2; ++i)
                if (s_lo + i < s_hi) accA[i] = fmaf(s_dd[n * SC_SMAX + s_lo + i], xv, accA[i]);
        }
        // phase 2b: dxn[n][f] = sum_s w[n][s] dt[s][f] + dd[n][s] A[s][f]; thread = (feature f, token parity `half`)
        for (int n = half; n < nv; n += 2) {
            float g = 0.f;
            for (int s = 0; s < S; ++s) g = fmaf(s_w[n * SC_SMAX + s], s_dt[s * SC_D + f], fmaf(s_dd[n * SC_SMAX + s], s_A[s * SC_D + f], g));
            float* o = dxn + ((size_t)b * N + nb + n) * SC_D + f;
            *o = accumulate ? *o + g : g;
        }
    }
    float* pd = part_dA + ((size_t)b * n_chunks + ch) * S * SC_D;
#pragma unroll
    for (int i = 0; i < SC_SMAX / 2; ++i)
        if (s_lo + i < s_hi) pd[(s_lo + i) * SC_D + f] = accA[i];
}

__global__ void sum_chunks_kernel(const float* __restrict__ part, float* __restrict__ out, int per_b, int n_chunks) {
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < per_b; i += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < n_chunks; ++c) acc += part[((size_t)b * n_chunks + c) * per_b + i];
        out[(size_t)b * per_b + i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ GRU gates
// gi = W_ih u + b_ih, gh = W_hh h + b_hh (rows [r; z; n]);  r = s(gi_r + gh_r), z = s(gi_z + gh_z), n = tanh(gi_n + r gh_n),
// h' = (1 - z) n + z h                                                                    (torch.nn.GRUCell)
__global__ void gru_gates_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h,
                                     float* __restrict__ hout, float* __restrict__ gates, long long M, int D) {
    const long long total = M * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / D; const int d = (int)(i - m * D);
        const float* a = gi + m * 3 * D; const float* c = gh + m * 3 * D;
        const float r = sigmoidf_(a[d] + c[d]);
        const float z = sigmoidf_(a[D + d] + c[D + d]);
        const float n = tanhf(a[2 * D + d] + r * c[2 * D + d]);
        hout[i] = (1.0f - z) * n + z * h[i];
        gates[m * 3 * D + d] = r; gates[m * 3 * D + D + d] = z; gates[m * 3 * D + 2 * D + d] = n;
    }
}
__global__ void gru_gates_bwd_kernel(const float* __restrict__ dhout, const float* __restrict__ gates, const float* __restrict__ gh,
                                     const float* __restrict__ h, float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh,
                                     long long M, int D) {
    const long long total = M * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / D; const int d = (int)(i - m * D);
        const float r = gates[m * 3 * D + d], z = gates[m * 3 * D + D + d], n = gates[m * 3 * D + 2 * D + d];
        const float g = dhout[i];
        const float dn = g * (1.0f - z);
        const float dz = g * (h[i] - n);
        const float dpre_n = dn * (1.0f - n * n);
        const float ghn = gh[m * 3 * D + 2 * D + d];
        const float dr = dpre_n * ghn;
        const float dpre_r = dr * r * (1.0f - r);
        const float dpre_z = dz * z * (1.0f - z);
        dgi[m * 3 * D + d] = dpre_r; dgi[m * 3 * D + D + d] = dpre_z; dgi[m * 3 * D + 2 * D + d] = dpre_n;
        dgh[m * 3 * D + d] = dpre_r; dgh[m * 3 * D + D + d] = dpre_z; dgh[m * 3 * D + 2 * D + d] = dpre_n * r;
        dh[i] = g * z;
    }
}

// ------------------------------------------------------------------------------------------------ conv weight gradients
// planes: rows of 64 channels, fp16 hi (+ fp16 lo scaled by 2048), zero borders; position p of an image-major padded grid.
__device__ __forceinline__ float plane_val(const uint16_t* hi, const uint16_t* lo, size_t idx) {
    float v = __half2float(*reinterpret_cast<const __half*>(hi + idx));
    if (lo) v = fmaf(__half2float(*reinterpret_cast<const __half*>(lo + idx)), LO_INV, v);
    return v;
}

// dW[tap = ky*5+kx][co][ci] += scale * sum_p dY[p][co] * X[p + (ky-2)*Wp + (kx-2)][ci]
// grid (n_chunks, 5): one CTA = one filter row ky for a chunk of positions; thread = 4 co x 4 ci x 5 kx accumulators.
constexpr int WG_PB = 32;
__global__ void __launch_bounds__(256) conv5x5_planes_wgrad_kernel(const uint16_t* __restrict__ dy_hi, const uint16_t* __restrict__ dy_lo,
                                                                  const uint16_t* __restrict__ x_hi, const uint16_t* __restrict__ x_lo,
                                                                  float* __restrict__ dW, long long P, int Wp, long long pos_per_chunk,
                                                                  float scale) {
    __shared__ __align__(16) float s_dy[WG_PB][64];
    __shared__ __align__(16) float s_x[WG_PB + 4][64];
    const int ky = blockIdx.y;
    const long long p_begin = (long long)blockIdx.x * pos_per_chunk;
    const long long p_end = min(P, p_begin + pos_per_chunk);
    const int co0 = (threadIdx.x >> 4) * 4, ci0 = (threadIdx.x & 15) * 4;
    float acc[5][4][4];
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[k][a][b] = 0.f;
    const long long shift = (long long)(ky - 2) * Wp - 2;          // X row of kx = 0
    for (long long p0 = p_begin; p0 < p_end; p0 += WG_PB) {
        __syncthreads();
        for (int i = threadIdx.x; i < WG_PB * 64; i += 256) {
            const long long p = p0 + (i >> 6);
            s_dy[i >> 6][i & 63] = p < p_end ? plane_val(dy_hi, dy_lo, (size_t)p * 64 + (i & 63)) : 0.f;
        }
        for (int i = threadIdx.x; i < (WG_PB + 4) * 64; i += 256) {
            const long long p = p0 + shift + (i >> 6);
            s_x[i >> 6][i & 63] = (p >= 0 && p < P) ? plane_val(x_hi, x_lo, (size_t)p * 64 + (i & 63)) : 0.f;
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < WG_PB; ++j) {
            const float4 g = *reinterpret_cast<const float4*>(&s_dy[j][co0]);
            const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float4 xv4 = *reinterpret_cast<const float4*>(&s_x[j + k][ci0]);
                const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[k][a][b] = fmaf(gv[a], xv[b], acc[k][a][b]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                atomicAdd(dW + ((size_t)(ky * 5 + k) * 64 + co0 + a) * 64 + ci0 + b, acc[k][a][b] * scale);
}

// out[c] += scale * sum_p planes[p][c]
__global__ void __launch_bounds__(256) planes_colsum_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ out,
                                                           long long P, long long pos_per_chunk, float scale) {
    // thread = (position group 0..31, 8 consecutive channels): 16-byte loads of the hi and lo rows, four positions in flight
    // (the first version read one 2-byte element per thread and step and ran at 13 % of the HBM bandwidth)
    const int cg = threadIdx.x & 7, pg = threadIdx.x >> 3;
    const long long p_begin = (long long)blockIdx.x * pos_per_chunk, p_end = min(P, p_begin + pos_per_chunk);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    auto add_row = [&](long long p) {
        const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(hi + (size_t)p * 64) + cg);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
            v[2 * j] = f.x; v[2 * j + 1] = f.y;
        }
        if (lo) {
            const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(lo + (size_t)p * 64) + cg);
            const uint32_t lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
                v[2 * j] = fmaf(f.x, LO_INV, v[2 * j]); v[2 * j + 1] = fmaf(f.y, LO_INV, v[2 * j + 1]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
    };
    long long p = p_begin + pg;
    for (; p + 96 < p_end; p += 128) { add_row(p); add_row(p + 32); add_row(p + 64); add_row(p + 96); }
    for (; p < p_end; p += 32) add_row(p);
    __shared__ float s[32][65];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[pg][cg * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
        float t = 0.f;
#pragma unroll 8
        for (int g = 0; g < 32; ++g) t += s[g][threadIdx.x];
        atomicAdd(out + threadIdx.x, t * scale);
    }
}

// first encoder convolution: dW[co][ci][ky][kx] += scale * sum dH[p(img,y,x)][co] * x[img][ci][y+ky-pad][x+kx-pad]  (zero padding)
// x: fp32 NCHW (or uint8 / 255).  grid (H / 16 bands, n_img); thread = (co 0..63, tap quarter): <= 25 taps in registers over the band.
__global__ void __launch_bounds__(256) conv_first_wgrad_kernel(const uint16_t* __restrict__ dh_hi, const uint16_t* __restrict__ dh_lo,
                                                              const float* __restrict__ x, const uint8_t* __restrict__ x8,
                                                              float* __restrict__ dW, int Cin, int H, int W, int ks, int band, float scale) {
    extern __shared__ float s_rows[];                 // [Cin][ks][W + ks - 1] input rows around y (zero padded)
    const int img = blockIdx.y, pad = ks / 2, Wr = W + ks - 1;
    const int co = threadIdx.x & 63, part = threadIdx.x >> 6;        // 4 parts split the Cin*ks*ks taps
    const int n_taps = Cin * ks * ks;
    const int Wp = W + 4, Hp = H + 4;
    float acc[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) acc[i] = 0.f;
    const int y0 = blockIdx.x * band, y1 = min(H, y0 + band);
    for (int y = y0; y < y1; ++y) {
        __syncthreads();
        for (int i = threadIdx.x; i < Cin * ks * Wr; i += 256) {
            const int ci = i / (ks * Wr), r = (i / Wr) % ks, xx = i % Wr - pad, yy = y + r - pad;
            float v = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t o = (((size_t)img * Cin + ci) * H + yy) * W + xx;
                v = x8 ? __fdiv_rn((float)x8[o], 255.0f) : x[o];
            }
            s_rows[i] = v;
        }
        __syncthreads();
        const size_t row_base = (((size_t)img * Hp + y + 2) * Wp + 2) * 64;
        for (int xx = 0; xx < W; ++xx) {
            const float g = plane_val(dh_hi, dh_lo, row_base + (size_t)xx * 64 + co);
#pragma unroll
            for (int i = 0; i < 25; ++i) {
                const int tap = part + 4 * i;
                if (tap < n_taps) {
                    const int ci = tap / (ks * ks), kyx = tap - ci * ks * ks, ky = kyx / ks, kx = kyx - ky * ks;
                    acc[i] = fmaf(g, s_rows[(ci * ks + ky) * Wr + kx + xx], acc[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 25; ++i) {
        const int tap = part + 4 * i;
        if (tap < n_taps) atomicAdd(dW + (size_t)co * n_taps + tap, acc[i] * scale);
    }
}

// The same with register reuse (W % 4 == 0): a thread owns whole filter ROWS (ci, ky) -- its part takes rows part, part + 4, ... --
// and walks the image row four pixels at a time: the 4 + ks - 1 input values of a filter row come in with two LDS.128 and feed
// 4 * ks FMAs (the first version issued one shared-memory load per FMA: 14.2 ms per SAVi training step, LDS-bound).
template <int KS>
__global__ void __launch_bounds__(256) conv_first_wgrad4_kernel(const uint16_t* __restrict__ dh_hi, const uint16_t* __restrict__ dh_lo,
                                                               const float* __restrict__ x, const uint8_t* __restrict__ x8,
                                                               float* __restrict__ dW, int Cin, int H, int W, int band, float scale) {
    extern __shared__ __align__(16) float s_rows4[];  // [Cin * KS][Wq] input rows around y (zero padded), Wq = pitch rounded up to 4 floats (+ 4 slack)
    constexpr int PAD = KS / 2, RMAX = 5;             // Cin * KS <= 20 rows -> at most 5 per part
    const int img = blockIdx.y, Wr = W + KS - 1, Wq = ((Wr + 3) & ~3) + 4;
    const int co = threadIdx.x & 63, part = threadIdx.x >> 6;
    const int n_rows = Cin * KS;
    const int Wp = W + 4, Hp = H + 4;
    float acc[RMAX][KS];
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
#pragma unroll
        for (int k = 0; k < KS; ++k) acc[j][k] = 0.f;
    const int y0 = blockIdx.x * band, y1 = min(H, y0 + band);
    for (int y = y0; y < y1; ++y) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_rows * Wq; i += 256) {
            const int r = i / Wq, c = i - r * Wq;
            const int ci = r / KS, ky = r - ci * KS, xx = c - PAD, yy = y + ky - PAD;
            float v = 0.f;
            if (c < Wr && yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t o = (((size_t)img * Cin + ci) * H + yy) * W + xx;
                v = x8 ? __fdiv_rn((float)x8[o], 255.0f) : x[o];
            }
            s_rows4[i] = v;
        }
        __syncthreads();
        const size_t row_base = (((size_t)img * Hp + y + 2) * Wp + 2) * 64;
        for (int xx = 0; xx < W; xx += 4) {
            float g[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) g[c] = plane_val(dh_hi, dh_lo, row_base + (size_t)(xx + c) * 64 + co);
#pragma unroll
            for (int j = 0; j < RMAX; ++j) {
                const int r = part + 4 * j;
                if (r < n_rows) {
                    const float4 w0 = *reinterpret_cast<const float4*>(s_rows4 + r * Wq + xx);
                    const float4 w1 = *reinterpret_cast<const float4*>(s_rows4 + r * Wq + xx + 4);
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int k = 0; k < KS; ++k)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[j][k] = fmaf(g[c], w[c + k], acc[j][k]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
        const int r = part + 4 * j;
        if (r < n_rows) {
            const int ci = r / KS, ky = r - ci * KS;
#pragma unroll
            for (int k = 0; k < KS; ++k)
                atomicAdd(dW + (size_t)co * (Cin * KS * KS) + (ci * KS + ky) * KS + k, acc[j][k] * scale);
        }
    }
}

// last decoder convolution (64 -> 4, 3x3): dW[co][ci][ky][kx] += sum g[img][co][y][x] h[img][ci][y+ky-1][x+kx-1]; db[co] += sum g.
// grid (n_img); thread = (ci 0..63, tap group 0..3 handling taps {g, g+4, g+8}); rows streamed through shared memory.
__global__ void __launch_bounds__(256) conv3x3_final_wgrad_kernel(const float* __restrict__ g_o, const float* __restrict__ h,
                                                                 const uint16_t* __restrict__ h_hi, const uint16_t* __restrict__ h_lo,
                                                                 float* __restrict__ dW, float* __restrict__ db, int H, int W) {
    extern __shared__ float sm[];
    const int img = blockIdx.x, Wr = W + 2;
    float* s_h = sm;                                 // [64][3][Wr + 1]
    float* s_g = sm + 64 * 3 * (Wr + 1);             // [4][W]
    const size_t HW = (size_t)H * W;
    const int ci = threadIdx.x >> 2, tg = threadIdx.x & 3;
    float acc[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
    float accb = 0.f;
    for (int y = 0; y < H; ++y) {
        __syncthreads();
        // rolling three-row window: image row yy lives in slot (yy + 3) % 3; only row y + 1 is new (rows -1, 0, 1 at y = 0)
        for (int yy = (y == 0 ? -1 : y + 1); yy <= y + 1; ++yy) {
            const int slot = (yy + 3) % 3;
            const bool row_ok = yy >= 0 && yy < H;
            if (h) {
                for (int i = threadIdx.x; i < 64 * Wr; i += 256) {
                    const int c = i / Wr, col = i - c * Wr, xx = col - 1;
                    s_h[(c * 3 + slot) * (Wr + 1) + col] = (row_ok && xx >= 0 && xx < W) ? h[((size_t)img * 64 + c) * HW + (size_t)yy * W + xx] : 0.f;
                }
            } else {
                // saved activation as padded split planes [img][(H + 4)(W + 4)][64]: consecutive threads read consecutive channels of a
                // position (coalesced 2-byte loads of the hi and lo rows); the zero border of the planes supplies the padding columns
                const size_t prow = (((size_t)img * (H + 4) + yy + 2) * (W + 4) + 1) * 64;      // column -1 of image row yy
                for (int i = threadIdx.x; i < 64 * Wr; i += 256) {
                    const int col = i >> 6, c = i & 63;
                    s_h[(c * 3 + slot) * (Wr + 1) + col] = row_ok ? plane_val(h_hi, h_lo, prow + (size_t)col * 64 + c) : 0.f;
                }
            }
        }
        for (int i = threadIdx.x; i < 4 * W; i += 256) s_g[i] = g_o[((size_t)img * 4 + i / W) * HW + (size_t)y * W + (i % W)];
        __syncthreads();
        if ((W & 3) == 0) {
            // four pixels per step: the 16 gradient values (4 outputs x 4 pixels) come in with four LDS.128 and are shared by the
            // thread's three taps -- 48 FMAs per 16 shared-memory loads (the scalar loop below issues five loads per four FMAs)
            const float* hr[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int tap = min(tg + 4 * i, 8), ky = tap / 3, kx = tap - 3 * ky;
                hr[i] = s_h + (ci * 3 + (y + ky + 2) % 3) * (Wr + 1) + kx;          // filter row ky reads image row y + ky - 1
            }
            for (int xx = 0; xx < W; xx += 4) {
                const float4 g0 = *reinterpret_cast<const float4*>(s_g + xx), g1 = *reinterpret_cast<const float4*>(s_g + W + xx);
                const float4 g2 = *reinterpret_cast<const float4*>(s_g + 2 * W + xx), g3 = *reinterpret_cast<const float4*>(s_g + 3 * W + xx);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float h0 = hr[i][xx], h1 = hr[i][xx + 1], h2 = hr[i][xx + 2], h3 = hr[i][xx + 3];
                    acc[i][0] = fmaf(g0.x, h0, fmaf(g0.y, h1, fmaf(g0.z, h2, fmaf(g0.w, h3, acc[i][0]))));
                    acc[i][1] = fmaf(g1.x, h0, fmaf(g1.y, h1, fmaf(g1.z, h2, fmaf(g1.w, h3, acc[i][1]))));
                    acc[i][2] = fmaf(g2.x, h0, fmaf(g2.y, h1, fmaf(g2.z, h2, fmaf(g2.w, h3, acc[i][2]))));
                    acc[i][3] = fmaf(g3.x, h0, fmaf(g3.y, h1, fmaf(g3.z, h2, fmaf(g3.w, h3, acc[i][3]))));
                }
            }
        } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int tap = tg + 4 * i;
            if (tap < 9) {
                const int ky = tap / 3, kx = tap - 3 * ky;
                const float* hr = s_h + (ci * 3 + (y + ky + 2) % 3) * (Wr + 1) + kx;
                for (int xx = 0; xx < W; ++xx) {
                    const float hv = hr[xx];
                    acc[i][0] = fmaf(s_g[xx], hv, acc[i][0]); acc[i][1] = fmaf(s_g[W + xx], hv, acc[i][1]);
                    acc[i][2] = fmaf(s_g[2 * W + xx], hv, acc[i][2]); acc[i][3] = fmaf(s_g[3 * W + xx], hv, acc[i][3]);
                }
            }
        }
        }
        if (threadIdx.x < 4) for (int xx = 0; xx < W; ++xx) accb += s_g[threadIdx.x * W + xx];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int tap = tg + 4 * i;
        if (tap < 9)
#pragma unroll
            for (int c = 0; c < 4; ++c) atomicAdd(dW + ((size_t)c * 64 + ci) * 9 + tap, acc[i][c]);
    }
    if (db && threadIdx.x < 4) atomicAdd(db + threadIdx.x, accb);
}

// out[n] = scale * sum_m X[m][n]   (deterministic: one thread per column, rows in order)
__global__ void colsum_kernel(const float* __restrict__ X, float* __restrict__ out, long long M, long long N, float scale) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (long long m = 0; m < M; ++m) acc += X[m * N + n];
    out[n] = acc * scale;
}

// fp32 NHWC tokens [n_img][H*W][ld] (first 64 columns) * scale -> zero-bordered split planes
__global__ void nhwc_to_split_scaled_kernel(const float* __restrict__ tok, int ld, const float* __restrict__ scale_ptr, uint16_t* __restrict__ hi,
                                            uint16_t* __restrict__ lo, int n_img, int H, int W) {
    const float sc = scale_ptr ? *scale_ptr : 1.0f;
    const long long total = (long long)n_img * H * W * 16;           // 4 channels per thread
    const int Wp = W + 4, Hp = H + 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i & 15);
        const long long pix = i >> 4;
        const int x = (int)(pix % W), y = (int)((pix / W) % H);
        const long long img = pix / ((long long)H * W);
        const float4 v = *reinterpret_cast<const float4*>(tok + (size_t)pix * ld + c4 * 4);
        const float f[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};
        const size_t o = ((((size_t)img * Hp + y + 2) * Wp) + x + 2) * 64 + c4 * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half hh = __float2half_rn(f[j]);
            *reinterpret_cast<__half*>(hi + o + j) = hh;
            if (lo) *reinterpret_cast<__half*>(lo + o + j) = __float2half_rn((f[j] - __half2float(hh)) * 2048.0f);
        }
    }
}

static inline unsigned ew_grid(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

}  // namespace

// ------------------------------------------------------------------------------------------------ C ABI

extern "C" int ps_sa_core_chunks(int N) { int c = (N + 511) / 512; return c < 1 ? 1 : c; }

// A (B,S,128) = scale * q W_k (queries folded through the key projection), xn (B,N,128) = norm_input(inputs)  ->
// t (B,S,128) = normalised attention-weighted mean of xn, Z (B,S) = sum_n (softmax_s + eps), masks (B,S,N) (softmax_s + eps) or null.
// workspace: part_t (B, chunks, S, 128), part_z (B, chunks, S) with chunks = ps_sa_core_chunks(N).
extern "C" int ps_sa_core_fwd(const float* A, const float* xn, float* t, float* Z, float* masks, float* part_t, float* part_z,
                              int B, int S, int N, float eps, void* stream) {
    PS_REQUIRE(A && xn && t && Z && part_t && part_z, "ps_sa_core_fwd: null pointer");
    PS_REQUIRE(S >= 1 && S <= SC_SMAX, "ps_sa_core_fwd: 1 <= num_slots <= %d (got %d)", SC_SMAX, S);
    PS_REQUIRE(ps_aligned16(xn), "ps_sa_core_fwd: features must be 16-byte aligned");
    if (B == 0 || N == 0) return PS_OK;
    PS_REQUIRE(B <= 65535, "ps_sa_core_fwd: at most 65535 frames per call");
    const int chunks = ps_sa_core_chunks(N), tpc = (N + chunks - 1) / chunks;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(SC_TB * SC_XS + SC_SMAX * SC_D + SC_TB * SC_SMAX) * sizeof(float);
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(sa_core_fwd_kernel, smem, smem_cfg));
    sa_core_fwd_kernel<<<dim3(chunks, B), 256, smem, st>>>(A, xn, part_t, part_z, masks, S, N, chunks, tpc, eps);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    sa_core_finish_kernel<<<B, 256, 0, st>>>(part_t, part_z, t, Z, S, chunks);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// gradients of ps_sa_core_fwd w.r.t. A (dA (B,S,128)) and xn (dxn (B,N,128); += when accumulate).  workspace part_dA (B, chunks, S, 128).
extern "C" int ps_sa_core_bwd(const float* A, const float* xn, const float* t, const float* Z, const float* dt, float* dA, float* dxn,
                              float* part_dA, int B, int S, int N, float eps, int accumulate, void* stream) {
    PS_REQUIRE(A && xn && t && Z && dt && dA && dxn && part_dA, "ps_sa_core_bwd: null pointer");
    PS_REQUIRE(S >= 1 && S <= SC_SMAX, "ps_sa_core_bwd: 1 <= num_slots <= %d (got %d)", SC_SMAX, S);
    PS_REQUIRE(ps_aligned16(xn) && ps_aligned16(dxn), "ps_sa_core_bwd: features must be 16-byte aligned");
    if (B == 0 || N == 0) return PS_OK;
    PS_REQUIRE(B <= 65535, "ps_sa_core_bwd: at most 65535 frames per call");
    const int chunks = ps_sa_core_chunks(N), tpc = (N + chunks - 1) / chunks;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)(SC_TB * SC_XS + 2 * SC_SMAX * SC_D + 2 * SC_TB * SC_SMAX + 2 * SC_SMAX) * sizeof(float);
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(sa_core_bwd_kernel, smem, smem_cfg));
    sa_core_bwd_kernel<<<dim3(chunks, B), 256, smem, st>>>(A, xn, t, Z, dt, dxn, part_dA, S, N, chunks, tpc, eps, accumulate);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    sum_chunks_kernel<<<B, 256, 0, st>>>(part_dA, dA, S * SC_D, chunks);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_gru_gates_fwd(const float* gi, const float* gh, const float* h, float* hout, float* gates, long long M, int D, void* stream) {
    PS_REQUIRE(gi && gh && h && hout && gates && D > 0, "ps_gru_gates_fwd: bad arguments");
    if (M == 0) return PS_OK;
    gru_gates_fwd_kernel<<<ew_grid(M * D), 256, 0, (cudaStream_t)stream>>>(gi, gh, h, hout, gates, M, D);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_gru_gates_bwd(const float* dhout, const float* gates, const float* gh, const float* h, float* dgi, float* dgh, float* dh,
                                long long M, int D, void* stream) {
    PS_REQUIRE(dhout && gates && gh && h && dgi && dgh && dh && D > 0, "ps_gru_gates_bwd: bad arguments");
    if (M == 0) return PS_OK;
    gru_gates_bwd_kernel<<<ew_grid(M * D), 256, 0, (cudaStream_t)stream>>>(dhout, gates, gh, h, dgi, dgh, dh, M, D);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// dW (25, 64, 64) [tap][cout][cin] += scale * correlation of the gradient planes dy with the activation planes x (both padded NHWC
// split planes of n_img images, precision 3: hi + lo, else hi only).  dW must be zeroed by the caller (gradients accumulate).
extern "C" int ps_conv5x5_planes_wgrad(const void* dy_planes, const void* x_planes, float* dW, int n_img, int H, int W, int dy_precision,
                                       int x_precision, float scale, void* stream) {
    PS_REQUIRE(dy_planes && x_planes && dW, "ps_conv5x5_planes_wgrad: null pointer");
    if (n_img == 0) return PS_OK;
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    const uint16_t* dy_hi = reinterpret_cast<const uint16_t*>(dy_planes);
    const uint16_t* x_hi = reinterpret_cast<const uint16_t*>(x_planes);
    const uint16_t* dy_lo = dy_precision == 3 ? dy_hi + P * 64 : nullptr;
    const uint16_t* x_lo = x_precision == 3 ? x_hi + P * 64 : nullptr;
    long long chunks = (P + 4095) / 4096;
    if (chunks > 148) chunks = 148;
    const long long ppc = ((P + chunks - 1) / chunks + WG_PB - 1) / WG_PB * WG_PB;
    chunks = (P + ppc - 1) / ppc;
    conv5x5_planes_wgrad_kernel<<<dim3((unsigned)chunks, 5), 256, 0, (cudaStream_t)stream>>>(dy_hi, dy_lo, x_hi, x_lo, dW, P, W + 4, ppc, scale);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// out (64) += scale * sum over positions of the planes (bias gradient); out must be zeroed by the caller
extern "C" int ps_planes_colsum(const void* planes, float* out, int n_img, int H, int W, int precision, float scale, void* stream) {
    PS_REQUIRE(planes && out, "ps_planes_colsum: null pointer");
    if (n_img == 0) return PS_OK;
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    const uint16_t* hi = reinterpret_cast<const uint16_t*>(planes);
    const uint16_t* lo = precision == 3 ? hi + P * 64 : nullptr;
    long long chunks = (P + 8191) / 8192;
    if (chunks > 1184) chunks = 1184;
    const long long ppc = (P + chunks - 1) / chunks;
    planes_colsum_kernel<<<(unsigned)chunks, 256, 0, (cudaStream_t)stream>>>(hi, lo, out, P, ppc, scale);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// dW (64, Cin, ks, ks) += scale * correlation of the gradient planes dh (64 channels) with the input images x (fp32 NCHW) or x8 (uint8 / 255)
extern "C" int ps_conv_first_wgrad(const void* dh_planes, const float* x, const unsigned char* x8, float* dW, int n_img, int Cin, int H, int W,
                                   int ksize, int precision, float scale, void* stream) {
    PS_REQUIRE(dh_planes && (x || x8) && dW, "ps_conv_first_wgrad: null pointer");
    PS_REQUIRE(Cin >= 1 && Cin <= 4 && (ksize == 3 || ksize == 5), "ps_conv_first_wgrad: Cin <= 4, ksize 3 | 5");
    if (n_img == 0) return PS_OK;
    PS_REQUIRE(n_img <= 65535, "ps_conv_first_wgrad: at most 65535 images per call");
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    const uint16_t* hi = reinterpret_cast<const uint16_t*>(dh_planes);
    const uint16_t* lo = precision == 3 ? hi + P * 64 : nullptr;
    const int band = 16;
    if (W % 4 == 0) {
        const size_t smem4 = (size_t)Cin * ksize * ((((W + ksize - 1) + 3) & ~3) + 4) * sizeof(float);
        const dim3 grid((H + band - 1) / band, n_img);
        if (ksize == 5) conv_first_wgrad4_kernel<5><<<grid, 256, smem4, (cudaStream_t)stream>>>(hi, lo, x, x8, dW, Cin, H, W, band, scale);
        else conv_first_wgrad4_kernel<3><<<grid, 256, smem4, (cudaStream_t)stream>>>(hi, lo, x, x8, dW, Cin, H, W, band, scale);
        PS_COUNT_LAUNCH();
        PS_LAUNCH_CHECK();
        return PS_OK;
    }
    const size_t smem = (size_t)Cin * ksize * (W + ksize - 1) * sizeof(float);
    conv_first_wgrad_kernel<<<dim3((H + band - 1) / band, n_img), 256, smem, (cudaStream_t)stream>>>(hi, lo, x, x8, dW, Cin, H, W, ksize, band, scale);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// dW (4, 64, 3, 3) += sum g_o * shifted h, db (4) += sum g_o; g_o (n_img, 4, H, W), h (n_img, 64, H, W) fp32 NCHW; zeroed by the caller
static int conv3x3_final_wgrad_run(const float* g_o, const float* h, const void* h_planes, int precision, float* dW, float* db, int n_img,
                                   int H, int W, void* stream);

extern "C" int ps_conv3x3_final_wgrad(const float* g_o, const float* h, float* dW, float* db, int n_img, int H, int W, void* stream) {
    PS_REQUIRE(h, "ps_conv3x3_final_wgrad: null pointer");
    return conv3x3_final_wgrad_run(g_o, h, nullptr, 0, dW, db, n_img, H, W, stream);
}

// the same with the saved activation h given as the padded split planes of the tensor-core layers (precision 2: hi plane only)
extern "C" int ps_conv3x3_final_wgrad_planes(const float* g_o, const void* h_planes, int precision, float* dW, float* db, int n_img, int H,
                                             int W, void* stream) {
    PS_REQUIRE(h_planes && (precision == 2 || precision == 3), "ps_conv3x3_final_wgrad_planes: planes with precision 2 | 3 required");
    return conv3x3_final_wgrad_run(g_o, nullptr, h_planes, precision, dW, db, n_img, H, W, stream);
}

static int conv3x3_final_wgrad_run(const float* g_o, const float* h, const void* h_planes, int precision, float* dW, float* db, int n_img,
                                   int H, int W, void* stream) {
    PS_REQUIRE(g_o && dW, "ps_conv3x3_final_wgrad: null pointer");
    if (n_img == 0) return PS_OK;
    PS_REQUIRE(n_img <= 65535 && W <= 256, "ps_conv3x3_final_wgrad: at most 65535 images, W <= 256");
    const size_t smem = (size_t)(64 * 3 * (W + 3) + 4 * W) * sizeof(float);
    static PsSmemConfig smem_cfg = {};
    if (smem > 48 * 1024) PS_CUDA(ps_ensure_smem(conv3x3_final_wgrad_kernel, smem, smem_cfg));
    const uint16_t* hi = reinterpret_cast<const uint16_t*>(h_planes);
    const uint16_t* lo = (hi && precision == 3) ? hi + (long long)n_img * (H + 4) * (W + 4) * 64 : nullptr;
    conv3x3_final_wgrad_kernel<<<n_img, 256, smem, (cudaStream_t)stream>>>(g_o, h, hi, lo, dW, db, H, W);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_colsum(const float* X, float* out, long long M, long long N, float scale, void* stream) {
    PS_REQUIRE(X && out && N > 0, "ps_colsum: bad arguments");
    colsum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, out, M, N, scale);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// tokens (n_img, H*W, ld) fp32 (first 64 columns) * scale[0] -> padded split planes (borders must already be zero)
extern "C" int ps_nhwc_to_split_scaled(const float* tokens, int ld, const float* scale, void* planes, int n_img, int H, int W, int precision,
                                       void* stream) {
    PS_REQUIRE(tokens && planes && ld >= 64 && ld % 4 == 0 && ps_aligned16(tokens), "ps_nhwc_to_split_scaled: bad arguments");
    if (n_img == 0) return PS_OK;
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    uint16_t* hi = reinterpret_cast<uint16_t*>(planes);
    uint16_t* lo = precision == 3 ? hi + P * 64 : nullptr;
    nhwc_to_split_scaled_kernel<<<ew_grid((long long)n_img * H * W * 16), 256, 0, (cudaStream_t)stream>>>(tokens, ld, scale, hi, lo, n_img, H, W);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
