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
// CTA (ps_sa_main_tc) = 288 threads: warps 0-7 load / normalise / split (warps 0-3 additionally do the per-token softmax out
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
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(sa_main_tc_kernel<SPAD>, smem, smem_cfg));
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

// ================================================================================================
// Pre-normalised feature planes.  ps_sa_prepare applies norm_input's statistics (no affine part) once and stores
// c = (x - mu) * rstd as two fp16 planes [hi | lo][B*N tokens][128] (lo scaled by 2048; 512 bytes per token, the size of
// the fp32 input).  ps_sa_main_tc2 then streams those planes with TMA straight into the swizzled operand layout: the
// per-iteration pass has no CUDA-core work on the features at all (TMA -> tcgen05 -> per-token softmax out of TMEM),
// so it runs at HBM speed for every iteration of every frame.
// ================================================================================================
namespace {

__global__ void __launch_bounds__(256) sa_prepare_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi_plane,
                                                        uint16_t* __restrict__ lo_plane, long long n_tok, float ln_eps) {
    const int lane = threadIdx.x & 31, q = lane >> 3, j = lane & 7;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long g0 = warp_global * 8; g0 < n_tok; g0 += n_warps * 8) {
        float4 v[2][4];
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
            const long long tok = g0 + g2 * 4 + q;
            const float* src = x + tok * SA_D + 8 * j;
            if (tok < n_tok) {
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
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
            const long long tok = g0 + g2 * 4 + q;
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
            const float rstd = 1.0f / sqrtf(ss * (1.0f / SA_D) + ln_eps);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] *= rstd;
            if (tok < n_tok) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    uint4 hi, lo;
                    split8(f + 8 * i, hi, lo);
                    *reinterpret_cast<uint4*>(hi_plane + tok * SA_D + 64 * i + 8 * j) = hi;
                    *reinterpret_cast<uint4*>(lo_plane + tok * SA_D + 64 * i + 8 * j) = lo;
                }
            }
        }
    }
}

constexpr int TC2_STAGES = 3;
constexpr int TC2_STAGE_BYTES = 4 * C_HALF;          // c_hi half0 | c_hi half1 | c_lo half0 | c_lo half1

struct SaTc2Params {
    const float* A;       // [B][S][128]
    const float* a0;      // [B][S]
    float* partial;       // [B][CH][S*128 + S]
    float* masks;         // [B][S][N] or null
    int B, N, S, CH, tpc;
    int lo_row0;          // row offset of the lo plane relative to the hi plane
    float attn_eps;
};

template <int SPAD>
__global__ void __launch_bounds__(320, 1) sa_main_tc2_kernel(const __grid_constant__ CUtensorMap mapC, const SaTc2Params p) {
    constexpr int NB = SPAD / 8;
    constexpr int N1 = 2 * SPAD;
    constexpr int AOP_HALF = N1 * 128;
    constexpr int AOP_BYTES = 2 * AOP_HALF;
    constexpr int WOP_GROUP = 2 * NB * 128;
    constexpr int WOP_BYTES = 16 * WOP_GROUP;
    constexpr uint32_t TM_L = 0;                     // 2 buffers x [L0 (N1) | L1 (N1)]
    constexpr uint32_t TM_V = 4 * N1;                // 2 buffers x [V0 (N1) | V1 (N1)]
    constexpr uint32_t TM_COLS = 8 * N1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* stage0 = smem;
    uint8_t* aop0 = smem + TC2_STAGES * TC2_STAGE_BYTES;
    uint8_t* wop0 = aop0 + 2 * AOP_BYTES;
    __shared__ uint64_t bar_full[TC2_STAGES], bar_empty[TC2_STAGES], bar_lfull[2], bar_lfree[2], bar_wfull[2], bar_wfree[2],
        bar_afull[2], bar_vfull[2], bar_vfree[2];
    __shared__ uint32_t tmem_base;
    __shared__ float s_a0[2][SPAD];
    __shared__ float s_red[2][2][4][SPAD];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay on the uniform datapath
    const int S = p.S;
    // Work partition: the B * tpf token tiles (tpf = tiles per frame) are dealt to the CTAs as contiguous, equally long
    // ranges; a range is cut at frame boundaries into segments (the work items below).  Segment j of frame b -- counted over
    // the CTAs that share the frame -- writes partial slot j; the CTA that finishes a frame zero-fills its unused slots.
    const int tpf = (p.N + TILE - 1) / TILE;
    const long long total_tiles = (long long)p.B * tpf;
    const int g_begin = (int)((long long)blockIdx.x * total_tiles / gridDim.x);
    const int g_end = (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x);
    auto owner = [&](long long gt) { return (int)(((gt + 1) * gridDim.x - 1) / total_tiles); };     // CTA that owns global tile gt

    if (tid == 0) {
        for (int s = 0; s < TC2_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&bar_lfull[b], 1); tc::mbar_init(&bar_lfree[b], 4);
            tc::mbar_init(&bar_wfull[b], 4); tc::mbar_init(&bar_wfree[b], 1);
            tc::mbar_init(&bar_afull[b], 4); tc::mbar_init(&bar_vfull[b], 1); tc::mbar_init(&bar_vfree[b], 4);
        }
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base, TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    // PDL: the feature planes were written long before the previous kernel (the slot update that produces A / a0), so the TMA
    // producer starts filling the ring while that kernel is still running; everything that reads A / a0 or writes partial /
    // masks (the softmax warps) waits for it.
    tc::pdl_trigger();

    if (warp == 0) {
        // ------------------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&mapC);
            const int lo_row0 = p.lo_row0;                         // first row of the lo plane
            int g = 0;
            for (int gs = g_begin; gs < g_end;) {
                const int b = gs / tpf, t_begin = gs - b * tpf;
                const int ntiles = min(tpf - t_begin, g_end - gs);
                const int tok0 = t_begin * TILE;
                gs += ntiles;
                for (int t = 0; t < ntiles; ++t, ++g) {
                    // (measured and rejected: an L2 prefetch of tiles ahead of the three-stage ring, 8-14 % slower; releasing the
                    //  stage after the logits MMA and re-loading the tile from L2 for the weighted sums, 6 % slower)
                    const int s = g % TC2_STAGES;
                    if (g >= TC2_STAGES) tc::mbar_wait(&bar_empty[s], ((g / TC2_STAGES) - 1) & 1);
                    uint8_t* st = stage0 + s * TC2_STAGE_BYTES;
                    const int row = b * p.N + tok0 + t * TILE;
                    tc::mbar_expect_tx(&bar_full[s], TC2_STAGE_BYTES);
                    tc::tma_load_2d(st, &mapC, &bar_full[s], 0, row);
                    tc::tma_load_2d(st + C_HALF, &mapC, &bar_full[s], 64, row);
                    tc::tma_load_2d(st + 2 * C_HALF, &mapC, &bar_full[s], 0, lo_row0 + row);
                    tc::tma_load_2d(st + 3 * C_HALF, &mapC, &bar_full[s], 64, lo_row0 + row);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------ MMA issuer
        // (the whole warp walks the loop with uniform control flow; one elected lane issues the MMAs and the commits)
        const bool leader = tc::elect_one();
        {
            const uint32_t idesc1 = tc::umma_idesc_f16_major(128, N1, 0, 0);
            const uint32_t idesc2 = tc::umma_idesc_f16_major(128, N1, 1, 1);
            const uint32_t st_base = tc::smem_u32(stage0), ao = tc::smem_u32(aop0), wo = tc::smem_u32(wop0);
            // descriptors are built once; the issue loop only adds 16-byte-unit offsets to their start-address field
            const uint64_t dA_k = tc::umma_desc_sw128(st_base, 1024, 0);                 // feature tile, K-major (MMA1)
            const uint64_t dA_mn = tc::umma_desc(st_base, C_HALF, 1024, 2);              // same bytes, MN-major (MMA2)
            const uint64_t dQ = tc::umma_desc_sw128(ao, 1024, 0);                        // folded queries
            const uint64_t dW = tc::umma_desc(wo, WOP_GROUP, 128, 0);                    // attention weights, MN-major
            auto mma1 = [&](int g, int k_item) {                  // logits of tile g into L buffer g & 1
                const uint64_t a = dA_k + (uint64_t)((g % TC2_STAGES) * (TC2_STAGE_BYTES / 16));
                const uint64_t q = dQ + (uint64_t)((k_item & 1) * (AOP_BYTES / 16));
                const uint32_t lb = tm + TM_L + (g & 1) * 2 * N1;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t bd = q + (uint64_t)((h * AOP_HALF + k * 32) / 16);
                        const uint32_t acc = (h | k) ? 1u : 0u;
                        if (leader) tc::umma_bf16(lb, a + (uint64_t)((h * C_HALF + k * 32) / 16), bd, idesc1, acc);
                        if (leader) tc::umma_bf16(lb + N1, a + (uint64_t)(((2 + h) * C_HALF + k * 32) / 16), bd, idesc1, acc);
                    }
                if (leader) tc::umma_commit(&bar_lfull[g & 1]);
            };
            int g = 0, k_item = 0;
            // The logits of the NEXT tile in stream order are always issued before the weighted sums of the current one -- also
            // across an item boundary (the next item's folded queries were converted one item ahead), so the softmax groups are
            // never starved at the start of an item.
            auto issue_logits = [&](int gn, int item_k) {
                tc::mbar_wait(&bar_full[gn % TC2_STAGES], (gn / TC2_STAGES) & 1);
                if (gn >= 2) tc::mbar_wait(&bar_lfree[gn & 1], ((gn >> 1) - 1) & 1);
                tc::tc_fence_after();
                mma1(gn, item_k);
            };
            if (g_begin < g_end) {
                tc::mbar_wait(&bar_afull[0], 0);                                   // folded queries of the first item are in smem
                issue_logits(0, 0);
            }
            for (int gs = g_begin; gs < g_end; ++k_item) {
                const int t_begin = gs % tpf;
                const int ntiles = min(tpf - t_begin, g_end - gs);
                gs += ntiles;
                for (int t = 0; t < ntiles; ++t, ++g) {
                    if (t + 1 < ntiles) {
                        issue_logits(g + 1, k_item);
                    } else if (gs < g_end) {
                        tc::mbar_wait(&bar_afull[(k_item + 1) & 1], ((k_item + 1) >> 1) & 1);
                        issue_logits(g + 1, k_item + 1);
                    }
                    if (t == 0 && k_item >= 2) tc::mbar_wait(&bar_vfree[k_item & 1], ((k_item >> 1) - 1) & 1);
                    tc::mbar_wait(&bar_wfull[g & 1], (g >> 1) & 1);
                    tc::tc_fence_after();
                    const uint64_t a = dA_mn + (uint64_t)((g % TC2_STAGES) * (TC2_STAGE_BYTES / 16));
                    const uint64_t wd = dW + (uint64_t)((g & 1) * (WOP_BYTES / 16));
                    const uint32_t vb = tm + TM_V + (k_item & 1) * 2 * N1;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint64_t bd = wd + (uint64_t)(kk * 2 * WOP_GROUP / 16);
                        const uint32_t acc = (t | kk) ? 1u : 0u;
                        if (leader) tc::umma_bf16(vb, a + (uint64_t)(kk * 2048 / 16), bd, idesc2, acc);
                        if (leader) tc::umma_bf16(vb + N1, a + (uint64_t)((2 * C_HALF + kk * 2048) / 16), bd, idesc2, acc);
                    }
                    if (leader) tc::umma_commit(&bar_empty[g % TC2_STAGES]);                   // feature tile consumed
                    if (leader) tc::umma_commit(&bar_wfree[g & 1]);
                    if (t == ntiles - 1) if (leader) tc::umma_commit(&bar_vfull[k_item & 1]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------------ softmax / epilogue warps (2..9)
        // Two groups of four warps; group G owns the tiles with (global tile counter & 1) == G, i.e. always the L / weight
        // buffer G, so a tile's softmax has two tile periods to complete.  Group 1 converts the folded queries one item
        // ahead, group 0 writes the finished item's partial sums.
        tc::pdl_wait();
        const int G = (warp - 2) >> 2;                           // group
        const int quad = warp & 3;                               // TMEM lane quadrant of this warp
        const int r = quad * 32 + lane;                          // token row of a tile == feature row of the V accumulators
        const int gtid = (tid - 64) & 127;                       // thread index inside the group
        const uint32_t tl = tm + ((uint32_t)(quad * 32) << 16);
        float rs[SPAD];
#pragma unroll
        for (int s = 0; s < SPAD; ++s) rs[s] = 0.f;

        auto convert_queries = [&](int k_item, int b) {         // A, a0 of frame b -> operand buffer k_item & 1 (group 1)
            uint8_t* aop = aop0 + (k_item & 1) * AOP_BYTES;
            for (int idx = gtid; idx < SPAD * 16; idx += 128) {
                const int s = idx >> 4, c = idx & 15;
                float v[8];
                if (s < S) {
                    const float4 q0 = *reinterpret_cast<const float4*>(p.A + ((size_t)b * S + s) * SA_D + c * 8);
                    const float4 q1 = *reinterpret_cast<const float4*>(p.A + ((size_t)b * S + s) * SA_D + c * 8 + 4);
                    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
                } else {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) v[jj] = 0.f;
                }
                uint4 hi, lo;
                split8(v, hi, lo);
                const int off = (c >> 3) * AOP_HALF + (((c & 7) ^ (s & 7)) << 4);
                *reinterpret_cast<uint4*>(aop + off + s * 128) = hi;
                *reinterpret_cast<uint4*>(aop + off + (SPAD + s) * 128) = lo;
            }
            if (gtid < SPAD) s_a0[k_item & 1][gtid] = (gtid < S) ? p.a0[b * S + gtid] : 0.f;
            tc::fence_async_proxy();
            asm volatile("bar.sync 3, 128;\n" ::: "memory");
            if (lane == 0) tc::mbar_arrive(&bar_afull[k_item & 1]);
        };

        auto seg_len = [&](int gs) { return min(tpf - gs % tpf, g_end - gs); };
        if (G == 1) {
            if (g_begin < g_end) {
                convert_queries(0, g_begin / tpf);
                const int gs1 = g_begin + seg_len(g_begin);
                if (gs1 < g_end) convert_queries(1, gs1 / tpf);
            }
        }
        int g = 0, k_item = 0;
        for (int gs = g_begin; gs < g_end; ++k_item) {
            const int b = gs / tpf, t_begin = gs - b * tpf;
            const int ntiles = seg_len(gs);
            const int tok0 = t_begin * TILE;
            const int ntok = min(p.N, (t_begin + ntiles) * TILE) - tok0;
            const int slot = (int)blockIdx.x - owner((long long)b * tpf);          // partial slot of this segment inside frame b
            const bool frame_done = t_begin + ntiles == tpf;
            gs += ntiles;
            // queries of the NEXT item: its buffer was last read by the MMAs of item k-1, all complete by now
            if (G == 1 && k_item >= 1 && gs < g_end) convert_queries(k_item + 1, gs / tpf);

            for (int t = 0; t < ntiles; ++t, ++g) {
                if ((g & 1) != G) continue;
                const int lb = G;
                tc::mbar_wait(&bar_lfull[lb], (g >> 1) & 1);
                __syncwarp();
                tc::tc_fence_after();
                float l0[N1], l1[SPAD];
                const uint32_t lbase = tl + TM_L + lb * 2 * N1;
                if constexpr (SPAD == 8) { tc::tmem_ld_32x16(lbase, l0); tc::tmem_ld_32x8(lbase + N1, l1); }
                else { tc::tmem_ld_32x32(lbase, l0); tc::tmem_ld_32x16(lbase + N1, l1); }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_lfree[lb]);
                float w[SPAD];
                float mx = -INFINITY;
#pragma unroll
                for (int s = 0; s < SPAD; ++s) {
                    w[s] = l0[s] + (l0[SPAD + s] + l1[s]) * (1.0f / LO_SCALE) + s_a0[k_item & 1][s];
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
                if (g >= 2) tc::mbar_wait(&bar_wfree[lb], ((g >> 1) - 1) & 1);     // MMA2 of tile g-2 has read this buffer
                uint8_t* wrow = wop0 + lb * WOP_BYTES + (r >> 3) * WOP_GROUP + (r & 7) * 16;
#pragma unroll
                for (int s = 0; s < SPAD; ++s) w[s] *= W_SCALE;
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) {
                    uint4 hi, lo;
                    split8(w + 8 * blk, hi, lo);
                    *reinterpret_cast<uint4*>(wrow + blk * 128) = hi;
                    *reinterpret_cast<uint4*>(wrow + (NB + blk) * 128) = lo;
                }
                tc::fence_async_proxy();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_wfull[lb]);
            }
            // ---- item finished: column sums of the attention weights from both groups, then the accumulators
#pragma unroll
            for (int s = 0; s < SPAD; ++s) {
                const float x = warp_sum(rs[s]);
                if (lane == 0) s_red[k_item & 1][G][quad][s] = x;
                rs[s] = 0.f;
            }
            asm volatile("bar.sync 2, 256;\n" ::: "memory");
            if (G == 0) {
                float rs_tot = 0.f;
                if (gtid < S) {
#pragma unroll
                    for (int gg = 0; gg < 2; ++gg)
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) rs_tot += s_red[k_item & 1][gg][qq][gtid];
                }
                tc::mbar_wait(&bar_vfull[k_item & 1], (k_item >> 1) & 1);
                __syncwarp();
                tc::tc_fence_after();
                float v0[N1], v1[SPAD];
                const uint32_t vb = tl + TM_V + (k_item & 1) * 2 * N1;
                if constexpr (SPAD == 8) { tc::tmem_ld_32x16(vb, v0); tc::tmem_ld_32x8(vb + N1, v1); }
                else { tc::tmem_ld_32x32(vb, v0); tc::tmem_ld_32x16(vb + N1, v1); }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_vfree[k_item & 1]);
                const int pstride = S * SA_D + S;
                float* pout = p.partial + ((size_t)b * p.CH + slot) * pstride;
#pragma unroll
                for (int s = 0; s < SPAD; ++s)
                    if (s < S) pout[s * SA_D + r] = (v0[s] + (v0[SPAD + s] + v1[s]) * (1.0f / LO_SCALE)) * (1.0f / W_SCALE);
                if (gtid < S) pout[S * SA_D + gtid] = rs_tot;
                if (frame_done)                                   // the update kernel sums all CH slots of a frame
                    for (int z = slot + 1; z < p.CH; ++z)
                        for (int idx = gtid; idx < pstride; idx += 128) p.partial[((size_t)b * p.CH + z) * pstride + idx] = 0.f;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tm, TM_COLS);
}

template <int SPAD>
int launch_sa_main_tc2(const CUtensorMap& mC, const SaTc2Params& p, cudaStream_t st) {
    const size_t smem = (size_t)TC2_STAGES * TC2_STAGE_BYTES + 2 * (2 * (2 * SPAD) * 128) + 2 * (16 * 2 * (SPAD / 8) * 128) + 1024;
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(sa_main_tc2_kernel<SPAD>, smem, smem_cfg));
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long total_tiles = (long long)p.B * ((p.N + TILE - 1) / TILE);
    PS_CUDA(tc::launch_pdl(sa_main_tc2_kernel<SPAD>, dim3((unsigned)(total_tiles < sms ? total_tiles : sms)), dim3(320), smem, st, mC, p));
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

extern "C" int ps_sa_prepare(const float* x, void* planes, long long n_tokens, int dim, float ln_eps, void* stream) {
    // planes: [hi | lo][n_tokens][128] 16-bit
    PS_REQUIRE(x && planes, "ps_sa_prepare: null pointer");
    PS_REQUIRE(dim == SA_D, "ps_sa_prepare: feature dim must be %d (got %d)", SA_D, dim);
    PS_REQUIRE(ps_aligned16(x) && ps_aligned16(planes), "ps_sa_prepare: pointers must be 16-byte aligned");
    if (n_tokens == 0) return PS_OK;
    uint16_t* hi = reinterpret_cast<uint16_t*>(planes);
    long long blocks = (n_tokens + 63) / 64;                   // 8 warps x 8 tokens per pass
    if (blocks > 148 * 16) blocks = 148 * 16;
    sa_prepare_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, hi + n_tokens * SA_D, n_tokens, ln_eps);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// Partial slots per frame ps_sa_main_tc2 needs: its CTAs take equally long contiguous ranges of the B * ceil(N / 128) token
// tiles, a frame is shared by at most this many CTAs (each writes one slot; unused slots are zero-filled).
extern "C" int ps_sa_tc2_chunks(int B, int N) {
    if (B <= 0 || N <= 0) return 1;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tpf = (N + TILE - 1) / TILE, total = (long long)B * tpf;
    const long long G = total < sms ? total : sms;
    int worst = 1;
    for (long long b = 0; b < B; ++b) {
        const long long first = ((b * tpf + 1) * G - 1) / total, last = ((b * tpf + tpf) * G - 1) / total;
        if ((int)(last - first + 1) > worst) worst = (int)(last - first + 1);
    }
    return worst;
}

extern "C" int ps_sa_main_tc2(const void* planes, long long lo_offset_tokens, const float* A, const float* a0, float* partial,
                              float* masks, int B, int N, int S, int dim, int n_chunks, float attn_eps, void* stream) {
    PS_REQUIRE(planes && A && a0 && partial, "ps_sa_main_tc2: null pointer");
    PS_REQUIRE(dim == SA_D, "ps_sa_main_tc2: feature/slot dim must be %d (got %d)", SA_D, dim);
    PS_REQUIRE(S >= 1 && S <= 16, "ps_sa_main_tc2: 1 <= num_slots <= 16 (got %d)", S);
    PS_REQUIRE(B >= 0 && N > 0 && n_chunks > 0, "ps_sa_main_tc2: bad shape");
    if (B == 0) return PS_OK;
    PS_REQUIRE(lo_offset_tokens >= (long long)B * N, "ps_sa_main_tc2: the lo plane must start at or after the end of the hi rows");
    PS_REQUIRE(lo_offset_tokens + (long long)B * N < (1ll << 31), "ps_sa_main_tc2: too many tokens for 32-bit TMA coordinates");
    PS_REQUIRE(ps_aligned16(planes) && ps_aligned16(A), "ps_sa_main_tc2: pointers must be 16-byte aligned");
    SaTc2Params p;
    p.A = A; p.a0 = a0; p.partial = partial; p.masks = masks;
    p.B = B; p.N = N; p.S = S; p.CH = n_chunks; p.tpc = 0;
    p.attn_eps = attn_eps; p.lo_row0 = (int)lo_offset_tokens;
    PS_REQUIRE(n_chunks >= ps_sa_tc2_chunks(B, N), "ps_sa_main_tc2: partial needs %d slots per frame for B=%d, N=%d (got %d): "
               "size it with ps_sa_tc2_chunks", ps_sa_tc2_chunks(B, N), B, N, n_chunks);
    CUtensorMap mC;
    const int r = tc::make_tmap_2d_bf16_sw128(&mC, planes, (uint64_t)(lo_offset_tokens + (long long)B * N), SA_D, TILE, 64);
    PS_REQUIRE(r == 0, "ps_sa_main_tc2: cuTensorMapEncodeTiled failed (%d)", r);
    cudaStream_t st = (cudaStream_t)stream;
    if (S <= 8) return launch_sa_main_tc2<8>(mC, p, st);
    return launch_sa_main_tc2<16>(mC, p, st);
}
