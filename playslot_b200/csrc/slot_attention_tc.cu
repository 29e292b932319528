// SlotAttention streaming pass on the tensor cores (tcgen05 / TMEM) -- same contract as ps_sa_main
// (slot_attention.cu) and the same re-association of src/models/BlocksUtils/attention.py:92-110:
//     logits[j,s] = c_j . A_s + a0_s            c_j = (x_j - mu_j) * rstd_j   (LayerNorm without its affine part)
//     attn        = softmax over slots + eps     (masks, :107-108)
//     U[s,:]      = sum_j attn[j,s] * c_j ,  rs[s] = sum_j attn[j,s]           (:109-110, finished by ps_sa_update)
// Both contractions run as tcgen05.mma with fp16 hi/lo split operands (x = hi + lo/2048, 22 mantissa bits,
// fp32 accumulation in TMEM; the lo*lo term is dropped):
//   MMA1  L[128 tokens x (2*SPAD)] = C[128 x 128 feat] * [A_hi | A_lo]^T      C K-major (features contiguous)
//   MMA2  V[128 feat x (2*SPAD)]  += C^T[128 feat x 128 tokens] * [w_hi | w_lo]  the SAME shared-memory tile read
//         as an MN-major operand (no transposed copy), w written by the softmax warps as an MN-major B operand.
// Per 128-token tile the features are read from HBM exactly once (coalesced float4), normalised, split and
// stored 128B-swizzled; nothing but the masks and a (S x 129)-float partial per CTA goes back.
//
// CTA = 288 threads: warps 0-7 load / normalise / split (warps 0-3 additionally do the per-token softmax out
// of TMEM and the epilogue), warp 8 issues the MMAs.  Synchronisation is mbarrier-only:
//   c_full (8 warps) -> MMA1 -> l_full -> softmax -> w_full (4 warps) -> MMA2 -> v_done -> tile buffer free.
// Several CTAs share an SM so one CTA's loads overlap another's softmax / MMA phases.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int SA_D = 128;
constexpr int TILE = 128;
constexpr int C_HALF = TILE * 128;                   // bytes of one [128 tokens][64 features] fp16 block
constexpr float LO_SCALE = 2048.0f;
// attention weights live in [1e-8, 1]: scaled by 2^15 before the fp16 split so that hi AND lo stay in fp16's
// normal range for every value (a "dead" slot whose weights are all ~eps must keep its relative precision:
// its update is the ratio U / rs of two tiny sums).  Undone, exactly, in the epilogue.
constexpr float W_SCALE = 32768.0f;

struct SaTcParams {
    const float* x;       // [B][N][128]
    const float* A;       // [B][S][128]
    const float* a0;      // [B][S]
    float* partial;       // [B][CH][S*128 + S]
    float* masks;         // [B][S][N] or null
    int B, N, S, CH, tpc;
    float ln_eps, attn_eps;
};

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn((v[2 * j] - f.x) * LO_SCALE, (v[2 * j + 1] - f.y) * LO_SCALE);
        hw[j] = *reinterpret_cast<const uint32_t*>(&h);
        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    hi = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

template <int SPAD>
__global__ void __launch_bounds__(288, 2) sa_main_tc_kernel(const SaTcParams p) {
    constexpr int NB = SPAD / 8;                     // 8-slot blocks per part
    constexpr int N1 = 2 * SPAD;                     // operand rows / accumulator columns: [hi | lo]
    constexpr int AOP_HALF = N1 * 128;               // folded queries, bytes per 64-feature half
    constexpr int WOP_GROUP = 2 * NB * 128;          // attention weights, bytes per 8-token group
    constexpr uint32_t TM_L0 = 0, TM_L1 = N1, TM_V0 = 2 * N1, TM_V1 = 3 * N1, TM_COLS = 4 * N1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* c_hi = smem;                            // [2 halves][128 tokens][128 B], 128B swizzle
    uint8_t* c_lo = smem + 2 * C_HALF;
    uint8_t* aop = smem + 4 * C_HALF;                // [2 halves][N1 rows][128 B], 128B swizzle
    uint8_t* wop = aop + 2 * AOP_HALF;               // [16 groups][2*NB blocks][8 tokens][8 slots]
    __shared__ uint64_t bar_cfull, bar_l, bar_wfull, bar_v;
    __shared__ uint32_t tmem_base;
    __shared__ float s_a0[SPAD];
    __shared__ float s_red[4][SPAD];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int S = p.S;
    const int tok0 = chunk * p.tpc;
    const int ntok = min(p.tpc, p.N - tok0);
    const int ntiles = (ntok + TILE - 1) / TILE;
    const float* xb = p.x + ((size_t)b * p.N + tok0) * SA_D;

    if (tid == 0) {
        tc::mbar_init(&bar_cfull, 8);
        tc::mbar_init(&bar_l, 1);
        tc::mbar_init(&bar_wfull, 4);
        tc::mbar_init(&bar_v, 1);
        tc::fence_mbar_init();
    }
    if (warp == 8) tc::tmem_alloc(&tmem_base, TM_COLS);
    // folded queries of this frame -> fp16 hi rows [0, SPAD) and lo rows [SPAD, 2*SPAD) of the B operand of MMA1
    for (int idx = tid; idx < SPAD * 16; idx += 288) {
        const int s = idx >> 4, c = idx & 15;
        float v[8];
        if (s < S) {
            const float4 v0 = *reinterpret_cast<const float4*>(p.A + ((size_t)b * S + s) * SA_D + c * 8);
            const float4 v1 = *reinterpret_cast<const float4*>(p.A + ((size_t)b * S + s) * SA_D + c * 8 + 4);
            v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        const int off = (c >> 3) * AOP_HALF + (((c & 7) ^ (s & 7)) << 4);
        *reinterpret_cast<uint4*>(aop + off + s * 128) = hi;
        *reinterpret_cast<uint4*>(aop + off + (SPAD + s) * 128) = lo;
    }
    if (tid < SPAD) s_a0[tid] = (tid < S) ? p.a0[b * S + tid] : 0.f;
    tc::fence_async_proxy();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;

    if (warp < 8) {
        const int q = lane >> 3, j = lane & 7;
        float rs[SPAD];
#pragma unroll
        for (int s = 0; s < SPAD; ++s) rs[s] = 0.f;

        // Software pipeline over "group pairs" (2 x 4 tokens per warp): the loads of the next pair -- across tile
        // boundaries too -- are in flight while the current pair is normalised, split and stored, and (warps 0-3)
        // while the softmax of the finished tile runs.
        auto load_pair = [&](float4 (&v)[2][4], int t, int gp) {
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const int r = warp * 16 + (gp * 2 + g2) * 4 + q;
                const int tok = t * TILE + r;
                const float* src = xb + (size_t)tok * SA_D + 8 * j;
                if (t < ntiles && tok < ntok) {
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        v[g2][2 * i] = __ldg(reinterpret_cast<const float4*>(src + 64 * i));
                        v[g2][2 * i + 1] = __ldg(reinterpret_cast<const float4*>(src + 64 * i + 4));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[g2][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        auto store_pair = [&](const float4 (&v)[2][4], int gp) {
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const int r = warp * 16 + (gp * 2 + g2) * 4 + q;
                float f[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) { f[4 * i] = v[g2][i].x; f[4 * i + 1] = v[g2][i].y; f[4 * i + 2] = v[g2][i].z; f[4 * i + 3] = v[g2][i].w; }
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) s += f[i];
                s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
                const float mean = s * (1.0f / SA_D);
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 16; ++i) { f[i] -= mean; ss = fmaf(f[i], f[i], ss); }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2); ss += __shfl_xor_sync(0xffffffffu, ss, 4);
                const float rstd = 1.0f / sqrtf(ss * (1.0f / SA_D) + p.ln_eps);
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] *= rstd;
                const int off = r * 128 + ((j ^ (r & 7)) << 4);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    uint4 hi, lo;
                    split8(f + 8 * i, hi, lo);
                    *reinterpret_cast<uint4*>(c_hi + i * C_HALF + off) = hi;
                    *reinterpret_cast<uint4*>(c_lo + i * C_HALF + off) = lo;
                }
            }
        };
        float4 va[2][4], vb[2][4];
        load_pair(va, 0, 0);
        for (int t = 0; t < ntiles; ++t) {
            // ---------------------------------------------------------------- load, normalise, split, store
            load_pair(vb, t, 1);
            if (t > 0) tc::mbar_wait(&bar_v, (t - 1) & 1);         // previous tile fully consumed
            store_pair(va, 0);
            load_pair(va, t + 1, 0);                                           // zeros past the last tile
            store_pair(vb, 1);
            tc::fence_async_proxy();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_cfull);

            // ---------------------------------------------------------------- softmax over slots (warps 0-3)
            if (warp < 4) {
                tc::mbar_wait(&bar_l, t & 1);
                __syncwarp();
                tc::tc_fence_after();
                const int r = warp * 32 + lane;                       // token row == TMEM lane
                const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
                float l0[N1], l1[SPAD];
                if constexpr (SPAD == 8) { tc::tmem_ld_32x16(tl + TM_L0, l0); tc::tmem_ld_32x8(tl + TM_L1, l1); }
                else { tc::tmem_ld_32x32(tl + TM_L0, l0); tc::tmem_ld_32x16(tl + TM_L1, l1); }
                float w[SPAD];
                float mx = -INFINITY;
#pragma unroll
                for (int s = 0; s < SPAD; ++s) {
                    w[s] = l0[s] + (l0[SPAD + s] + l1[s]) * (1.0f / LO_SCALE) + s_a0[s];
                    if (s < S) mx = fmaxf(mx, w[s]);
                }
                float den = 0.f;
#pragma unroll
                for (int s = 0; s < SPAD; ++s) {
                    w[s] = (s < S) ? expf(w[s] - mx) : 0.f;
                    den += w[s];
                }
                const int tok = t * TILE + r;
                const bool valid = tok < ntok;
                const float inv = 1.0f / den;
#pragma unroll
                for (int s = 0; s < SPAD; ++s) {
                    w[s] = (s < S && valid) ? w[s] * inv + p.attn_eps : 0.f;      // softmax(dim=slots) + eps
                    rs[s] += w[s];
                }
                if (p.masks && valid) {
                    float* mrow = p.masks + (size_t)b * S * p.N + tok0 + tok;
#pragma unroll
                    for (int s = 0; s < SPAD; ++s)
                        if (s < S) mrow[(size_t)s * p.N] = w[s];
                }
                uint8_t* wrow = wop + (r >> 3) * WOP_GROUP + (r & 7) * 16;
#pragma unroll
                for (int s = 0; s < SPAD; ++s) w[s] *= W_SCALE;
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) {
                    uint4 hi, lo;
                    split8(w + 8 * blk, hi, lo);
                    *reinterpret_cast<uint4*>(wrow + blk * 128) = hi;
                    *reinterpret_cast<uint4*>(wrow + (NB + blk) * 128) = lo;
                }
                tc::tc_fence_before();
                tc::fence_async_proxy();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_wfull);
            }
        }

        // -------------------------------------------------------------------- epilogue: partial sums of this CTA
        if (warp < 4) {
            tc::mbar_wait(&bar_v, (ntiles - 1) & 1);
            __syncwarp();
            tc::tc_fence_after();
            const int k = warp * 32 + lane;                            // feature == TMEM lane
            const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16);
            float v0[N1], v1[SPAD];
            if constexpr (SPAD == 8) { tc::tmem_ld_32x16(tl + TM_V0, v0); tc::tmem_ld_32x8(tl + TM_V1, v1); }
            else { tc::tmem_ld_32x32(tl + TM_V0, v0); tc::tmem_ld_32x16(tl + TM_V1, v1); }
            const int pstride = S * SA_D + S;
            float* pout = p.partial + ((size_t)b * p.CH + chunk) * pstride;
#pragma unroll
            for (int s = 0; s < SPAD; ++s)
                if (s < S) pout[s * SA_D + k] = (v0[s] + (v0[SPAD + s] + v1[s]) * (1.0f / LO_SCALE)) * (1.0f / W_SCALE);
#pragma unroll
            for (int s = 0; s < SPAD; ++s) {
                const float r = warp_sum(rs[s]);
                if (lane == 0) s_red[warp][s] = r;
            }
            asm volatile("bar.sync 1, 128;\n" ::: "memory");
            if (tid < S) pout[S * SA_D + tid] = (s_red[0][tid] + s_red[1][tid]) + (s_red[2][tid] + s_red[3][tid]);
        }
    } else {
        // ------------------------------------------------------------------------ MMA issuer (warp 8)
        if (tc::elect_one()) {
            const uint32_t idesc1 = tc::umma_idesc_f16_major(128, N1, 0, 0);
            const uint32_t idesc2 = tc::umma_idesc_f16_major(128, N1, 1, 1);
            const uint32_t chi = tc::smem_u32(c_hi), clo = tc::smem_u32(c_lo);
            const uint32_t ao = tc::smem_u32(aop), wo = tc::smem_u32(wop);
            for (int t = 0; t < ntiles; ++t) {
                tc::mbar_wait(&bar_cfull, t & 1);
                tc::tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = tc::umma_desc_sw128(ao + h * AOP_HALF + k * 32, 1024, 0);
                        const uint32_t acc = (h | k) ? 1u : 0u;
                        tc::umma_bf16(tm + TM_L0, tc::umma_desc_sw128(chi + h * C_HALF + k * 32, 1024, 0), bd, idesc1, acc);
                        tc::umma_bf16(tm + TM_L1, tc::umma_desc_sw128(clo + h * C_HALF + k * 32, 1024, 0), bd, idesc1, acc);
                    }
                tc::umma_commit(&bar_l);
                tc::mbar_wait(&bar_wfull, t & 1);
                tc::tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t bd = tc::umma_desc(wo + kk * 2 * WOP_GROUP, WOP_GROUP, 128, 0);
                    const uint32_t acc = (t | kk) ? 1u : 0u;
                    tc::umma_bf16(tm + TM_V0, tc::umma_desc(chi + kk * 2048, C_HALF, 1024, 2), bd, idesc2, acc);
                    tc::umma_bf16(tm + TM_V1, tc::umma_desc(clo + kk * 2048, C_HALF, 1024, 2), bd, idesc2, acc);
                }
                tc::umma_commit(&bar_v);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tm, TM_COLS);
}

template <int SPAD>
int launch_sa_main_tc(const SaTcParams& p, cudaStream_t st) {
    const size_t smem = 4 * C_HALF + 2 * (2 * SPAD * 128) + 16 * (2 * (SPAD / 8) * 128) + 1024;
    static bool configured = false;
    if (!configured) {
        PS_CUDA(cudaFuncSetAttribute(sa_main_tc_kernel<SPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    sa_main_tc_kernel<SPAD><<<dim3(p.CH, p.B), 288, smem, st>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

extern "C" int ps_sa_main_tc(const float* x, const float* A, const float* a0, float* partial, float* masks,
                             int B, int N, int S, int dim, int n_chunks, float ln_eps, float attn_eps, void* stream) {
    PS_REQUIRE(x && A && a0 && partial, "ps_sa_main_tc: null pointer");
    PS_REQUIRE(dim == SA_D, "ps_sa_main_tc: feature/slot dim must be %d (got %d)", SA_D, dim);
    PS_REQUIRE(S >= 1 && S <= 16, "ps_sa_main_tc: 1 <= num_slots <= 16 (got %d)", S);
    PS_REQUIRE(B >= 0 && N > 0 && n_chunks > 0, "ps_sa_main_tc: bad shape");
    if (B == 0) return PS_OK;
    PS_REQUIRE(B <= 65535, "ps_sa_main_tc: batch too large for one launch");
    PS_REQUIRE(ps_aligned16(x) && ps_aligned16(A), "ps_sa_main_tc: pointers must be 16-byte aligned");
    SaTcParams p;
    p.x = x; p.A = A; p.a0 = a0; p.partial = partial; p.masks = masks;
    p.B = B; p.N = N; p.S = S; p.CH = n_chunks; p.tpc = (N + n_chunks - 1) / n_chunks;
    p.ln_eps = ln_eps; p.attn_eps = attn_eps;
    PS_REQUIRE((size_t)(n_chunks - 1) * p.tpc < (size_t)N, "ps_sa_main_tc: n_chunks=%d leaves empty chunks for N=%d", n_chunks, N);
    cudaStream_t st = (cudaStream_t)stream;
    if (S <= 8) return launch_sa_main_tc<8>(p, st);
    return launch_sa_main_tc<16>(p, st);
}
