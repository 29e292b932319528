// Fused SAVi encoder MLP on tcgen05:  tokens -> LayerNorm -> Linear(C,128) -> ReLU -> Linear(128,128) -> norm_input
// statistics -> fp16 hi/lo planes for the tensor-core SlotAttention pass.
// Replaces `self.encoder_mlp` (src/models/SAVi.py:111-116, applied at :214-216) followed by the `norm_input` LayerNorm of
// SlotAttention (BlocksUtils/attention.py:57,88; its affine part is folded into the query matrices, see
// slot_attention_tc.cu) for the whole clip: 5.5 M tokens per bench step.  The unfused chain (two ps_linear_tc launches +
// ps_sa_prepare) moves the hidden and the output activations through HBM three times (16.8 GB per step); this kernel
// reads the tokens once and writes the planes once (4.2 GB).
//
// Arithmetic: the same fp16x3 scheme as linear_tc.cu (x = hi + lo/2048, D0 += A_hi W_hi, D1 += A_hi W_lo + A_lo W_hi,
// y = D0 + D1/2048), with the SAME packed weights (ops.pack_linear_weight_tc), so results agree with the unfused chain
// to fp32 rounding.
//
// Persistent, one CTA per SM, 128-token tiles, 672 threads:
//   warps 12-19 : token producers -- read the fp32 tokens (float4, four lanes per row), LayerNorm over C in registers,
//                 fp16 split, 128B-swizzled K-major A1 tile
//   warp  20    : MMA issuer (whole warp walks the loop, one elected lane issues), TMEM owner, loads both weight
//                 matrices by TMA once (they stay resident in shared memory: 96 KB)
//   warps 0-3   : epilogue 1 -- hidden accumulators (TMEM) -> +b1, ReLU -> fp16 split -> A2 tiles (operand of GEMM 2)
//   warps 4-11  : epilogue 2 -- output accumulators -> +b2 -> norm_input statistics (two threads per token row exchange
//                 their partial sums through shared memory) -> fp16 split -> hi / lo planes in HBM
// The MMA warp issues GEMM 1 of tile i before GEMM 2 of tile i-1, so epilogue 1 of tile i overlaps GEMM 2 of tile i-1
// and epilogue 2 of tile i-1 overlaps both GEMMs of the next tile.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int EM_BM = 128;                 // tokens per tile
constexpr int EM_D = 128;                  // hidden == output features
constexpr int EM_TILE = EM_BM * 128;       // bytes of one 16-bit [128 rows][64 K] operand tile
constexpr int EM_P1_WARPS = 8;              // token-producer warps (4 lanes per row -> 8 rows per warp and pass)
constexpr int EM_P1_PASSES = EM_BM / (EM_P1_WARPS * 8);
constexpr int EM_MMA_WARP = 12 + EM_P1_WARPS;
constexpr int EM_THREADS = (EM_MMA_WARP + 1) * 32;
constexpr float EM_LO = 2048.0f;

// shared-memory map (bytes, from the 1024-aligned base)
constexpr int EM_W1 = 0;                            // [W1_hi 128 rows ; W1_lo 128 rows] x 128 B          32 KB
constexpr int EM_W2 = EM_W1 + 2 * EM_TILE;          // 2 K-blocks x [W2_hi ; W2_lo]                        64 KB
constexpr int EM_A1 = EM_W2 + 4 * EM_TILE;          // [A1_hi | A1_lo]                                     32 KB
constexpr int EM_A2 = EM_A1 + 2 * EM_TILE;          // 2 K-blocks x [A2_hi | A2_lo]                        64 KB
constexpr int EM_SMEM = EM_A2 + 4 * EM_TILE;        // 192 KB

struct EncMlpParams {
    const float* X; int ldx; int K1;
    const float* ln_g; const float* ln_b; float ln_eps;
    const float* b1; const float* b2;
    float s1, s2;                 // 1 / (power-of-two scale folded into the packed weights)
    float out_eps;
    uint16_t* out_hi; uint16_t* out_lo;
    long long M;
    int n_tiles;
    int pair_hw;                  // > 0: input rows are (pair, position, image-in-pair); output row = (2 * pair + image) * pair_hw + position
};

__device__ __forceinline__ void em_split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn((v[2 * j] - f.x) * EM_LO, (v[2 * j + 1] - f.y) * EM_LO);
        hw[j] = *reinterpret_cast<const uint32_t*>(&h);
        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    hi = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    lo = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// one 256-bit store: a lane writes a whole 32-byte sector (the rows of a warp are 256 bytes apart, so 16-byte stores would
// touch every sector twice)
__device__ __forceinline__ void em_store32(void* dst, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                 ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

__global__ void __launch_bounds__(EM_THREADS, 1)
enc_mlp_tc_kernel(const __grid_constant__ CUtensorMap mapW1, const __grid_constant__ CUtensorMap mapW2, const EncMlpParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_w, bar_a1full, bar_a1empty, bar_d1full, bar_d1empty, bar_a2full, bar_a2empty, bar_d2full, bar_d2empty;
    __shared__ uint32_t tmem_base;
    __shared__ float s_b1[EM_D], s_b2[EM_D];
    __shared__ __align__(16) float s_g[64], s_be[64];              // LayerNorm affine parameters (zero beyond K1)
    __shared__ float s_sum[2][2][EM_BM], s_sq[2][2][EM_BM];     // [tile parity][column half][row]

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    if (tid == 0) {
        tc::mbar_init(&bar_w, 1);
        tc::mbar_init(&bar_a1full, EM_P1_WARPS);  tc::mbar_init(&bar_a1empty, 1);
        tc::mbar_init(&bar_d1full, 1);  tc::mbar_init(&bar_d1empty, 4);
        tc::mbar_init(&bar_a2full, 4);  tc::mbar_init(&bar_a2empty, 1);
        tc::mbar_init(&bar_d2full, 1);  tc::mbar_init(&bar_d2empty, 8);
        tc::fence_mbar_init();
    }
    if (tid < EM_D) { s_b1[tid] = p.b1 ? p.b1[tid] : 0.f; s_b2[tid] = p.b2 ? p.b2[tid] : 0.f; }
    if (tid < 64) { s_g[tid] = tid < p.K1 ? p.ln_g[tid] : 0.f; s_be[tid] = tid < p.K1 ? p.ln_b[tid] : 0.f; }
    if (warp == EM_MMA_WARP) tc::tmem_alloc(&tmem_base, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    constexpr uint32_t TM_D1 = 0, TM_D2 = 256;

    if (warp >= 12 && warp < EM_MMA_WARP) {
        // ------------------------------------------------------------------ token producers (8 warps)
        // Four lanes per token row, each owning the float4 columns {lane4*4 + 16*j}: the 64-byte row segments of a warp-wide
        // load are contiguous, and the LayerNorm reductions need two shuffle steps (instead of four with 16 lanes per row).
        const int ptid = tid - 12 * 32;             // 0..255
        const int lane4 = ptid & 3;
        const int r_sub = ptid >> 2;                // row inside a pass
        const int nj = p.K1 >> 4;                   // float4 columns per lane (K1 is a multiple of 16)
        float4 g4[4], b4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            g4[j] = *reinterpret_cast<const float4*>(s_g + lane4 * 4 + 16 * j);
            b4[j] = *reinterpret_cast<const float4*>(s_be + lane4 * 4 + 16 * j);
        }
        const float inv_k = 1.0f / (float)p.K1;
        uint8_t* a_hi = smem + EM_A1;
        uint8_t* a_lo = a_hi + EM_TILE;
        int i = 0;
        for (int it = blockIdx.x; it < p.n_tiles; it += gridDim.x, ++i) {
            const long long m0 = (long long)it * EM_BM;
            float4 v[EM_P1_PASSES][4];
#pragma unroll
            for (int r = 0; r < EM_P1_PASSES; ++r) {
                const long long m = m0 + r * (EM_P1_WARPS * 8) + r_sub;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[r][j] = (m < p.M && j < nj) ? __ldg(reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + lane4 * 4 + 16 * j))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < EM_P1_PASSES; ++r) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) s += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                const float mean = s * inv_k;
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < nj) {
                        v[r][j].x -= mean; v[r][j].y -= mean; v[r][j].z -= mean; v[r][j].w -= mean;
                        ss = fmaf(v[r][j].x, v[r][j].x, ss); ss = fmaf(v[r][j].y, v[r][j].y, ss);
                        ss = fmaf(v[r][j].z, v[r][j].z, ss); ss = fmaf(v[r][j].w, v[r][j].w, ss);
                    }
                }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1);
                ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                const float rstd = 1.0f / sqrtf(ss * inv_k + p.ln_eps);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[r][j].x = fmaf(v[r][j].x * rstd, g4[j].x, b4[j].x); v[r][j].y = fmaf(v[r][j].y * rstd, g4[j].y, b4[j].y);
                    v[r][j].z = fmaf(v[r][j].z * rstd, g4[j].z, b4[j].z); v[r][j].w = fmaf(v[r][j].w * rstd, g4[j].w, b4[j].w);
                }
            }
            if (i > 0) tc::mbar_wait(&bar_a1empty, (i - 1) & 1);      // GEMM 1 of the previous tile has read the A1 tile
#pragma unroll
            for (int r = 0; r < EM_P1_PASSES; ++r) {
                const int row = r * (EM_P1_WARPS * 8) + r_sub;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // columns lane4*4 + 16*j .. +3 = bytes (lane4*8 + 32*j) of the 128-byte fp16 row: 16-byte chunk 2*j + (lane4 >> 1)
                    const int off = row * 128 + (((2 * j + (lane4 >> 1)) ^ (row & 7)) << 4) + ((lane4 & 1) << 3);
                    const float4 x = (j < nj) ? v[r][j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    const __half2 l0 = __floats2half2_rn((x.x - f0.x) * EM_LO, (x.y - f0.y) * EM_LO);
                    const __half2 l1 = __floats2half2_rn((x.z - f1.x) * EM_LO, (x.w - f1.y) * EM_LO);
                    *reinterpret_cast<uint2*>(a_hi + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                    *reinterpret_cast<uint2*>(a_lo + off) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
                }
            }
            tc::fence_async_proxy();               // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_a1full);
        }
    } else if (warp == EM_MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        const bool leader = tc::elect_one();
        if (leader) {
            tc::tma_prefetch_desc(&mapW1);
            tc::tma_prefetch_desc(&mapW2);
            tc::mbar_expect_tx(&bar_w, 6 * EM_TILE);
            tc::tma_load_2d(smem + EM_W1, &mapW1, &bar_w, 0, 0);
            tc::tma_load_2d(smem + EM_W2, &mapW2, &bar_w, 0, 0);
            tc::tma_load_2d(smem + EM_W2 + 2 * EM_TILE, &mapW2, &bar_w, 64, 0);
        }
        const uint32_t idesc256 = tc::umma_idesc_f16(EM_BM, 256);
        const uint32_t idesc128 = tc::umma_idesc_f16(EM_BM, 128);
        const uint32_t base = tc::smem_u32(smem);
        const uint64_t dW1 = tc::umma_desc_sw128(base + EM_W1, 1024, 0);
        const uint64_t dW2 = tc::umma_desc_sw128(base + EM_W2, 1024, 0);
        const uint64_t dA1 = tc::umma_desc_sw128(base + EM_A1, 1024, 0);
        const uint64_t dA2 = tc::umma_desc_sw128(base + EM_A2, 1024, 0);
        constexpr uint64_t TILE_U = EM_TILE / 16;                     // descriptor units (16 bytes) per operand tile
        const int nk1 = (p.K1 + 15) >> 4;
        tc::mbar_wait(&bar_w, 0);
        auto gemm2 = [&](int j) {                                     // output accumulators of tile-local index j
            tc::mbar_wait(&bar_a2full, j & 1);
            if (j > 0) tc::mbar_wait(&bar_d2empty, (j - 1) & 1);
            tc::tc_fence_after();
            if (leader) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t w = dW2 + (uint64_t)(kb * 2 * TILE_U + k * 2);
                        tc::umma_bf16(tm + TM_D2, dA2 + (uint64_t)(kb * 2 * TILE_U + k * 2), w, idesc256, (kb | k) ? 1u : 0u);
                        tc::umma_bf16(tm + TM_D2 + 128, dA2 + (uint64_t)(kb * 2 * TILE_U + TILE_U + k * 2), w, idesc128, 1u);
                    }
                tc::umma_commit(&bar_a2empty);
                tc::umma_commit(&bar_d2full);
            }
        };
        int i = 0;
        for (int it = blockIdx.x; it < p.n_tiles; it += gridDim.x, ++i) {
            tc::mbar_wait(&bar_a1full, i & 1);
            if (i > 0) tc::mbar_wait(&bar_d1empty, (i - 1) & 1);
            tc::tc_fence_after();
            if (leader) {
                for (int k = 0; k < nk1; ++k) {
                    // A_hi x [W1_hi ; W1_lo] (N = 256: D0 | D1 adjacent in TMEM), then D1 += A_lo x W1_hi
                    tc::umma_bf16(tm + TM_D1, dA1 + (uint64_t)(k * 2), dW1 + (uint64_t)(k * 2), idesc256, k ? 1u : 0u);
                    tc::umma_bf16(tm + TM_D1 + 128, dA1 + (uint64_t)(TILE_U + k * 2), dW1 + (uint64_t)(k * 2), idesc128, 1u);
                }
                tc::umma_commit(&bar_a1empty);
                tc::umma_commit(&bar_d1full);
            }
            if (i > 0) gemm2(i - 1);
        }
        if (i > 0) gemm2(i - 1);
    } else if (warp < 4) {
        // ------------------------------------------------------------------ epilogue 1: hidden = relu(D + b1) -> A2 operand tiles
        const int row = warp * 32 + lane;
        const uint32_t tl = tm + ((uint32_t)(warp * 32) << 16) + TM_D1;
        int i = 0;
        for (int it = blockIdx.x; it < p.n_tiles; it += gridDim.x, ++i) {
            tc::mbar_wait(&bar_d1full, i & 1);
            __syncwarp();
            tc::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < EM_D; c0 += 32) {
                float a[32], b[32];
                tc::tmem_ld_32x32(tl + c0, a);
                tc::tmem_ld_32x32(tl + 128 + c0, b);
                if (c0 == EM_D - 32) {                              // every hidden accumulator column is in registers
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_d1empty);
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) a[c] = fmaxf(fmaf(fmaf(b[c], 1.0f / EM_LO, a[c]), p.s1, s_b1[c0 + c]), 0.f);
                if (c0 == 0 && i > 0) tc::mbar_wait(&bar_a2empty, (i - 1) & 1);     // GEMM 2 of the previous tile has read A2
                uint8_t* dst = smem + EM_A2 + (c0 >> 6) * 2 * EM_TILE + row * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 hi, lo;
                    em_split8(a + 8 * j, hi, lo);
                    const int chunk = ((c0 & 63) >> 3) + j;
                    const int off = (chunk ^ (row & 7)) << 4;
                    *reinterpret_cast<uint4*>(dst + off) = hi;
                    *reinterpret_cast<uint4*>(dst + EM_TILE + off) = lo;
                }
            }
            tc::fence_async_proxy();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_a2full);
        }
    } else {
        // ------------------------------------------------------------------ epilogue 2 (warps 4..11): norm_input + split -> planes
        const int quad = warp & 3;
        const int half = (warp - 4) >> 2;                          // column half [64 * half, +64)
        const int row = quad * 32 + lane;
        const uint32_t tl = tm + ((uint32_t)(quad * 32) << 16) + TM_D2 + half * 64;
        int i = 0;
        for (int it = blockIdx.x; it < p.n_tiles; it += gridDim.x, ++i) {
            const int par = i & 1;
            tc::mbar_wait(&bar_d2full, par);
            __syncwarp();
            tc::tc_fence_after();
            float y[64];
            {
                float lo[32];
                tc::tmem_ld_32x32(tl, y);
                tc::tmem_ld_32x32(tl + 128, lo);
#pragma unroll
                for (int c = 0; c < 32; ++c) y[c] = fmaf(fmaf(lo[c], 1.0f / EM_LO, y[c]), p.s2, s_b2[half * 64 + c]);
                tc::tmem_ld_32x32(tl + 32, y + 32);
                tc::tmem_ld_32x32(tl + 128 + 32, lo);
#pragma unroll
                for (int c = 0; c < 32; ++c) y[32 + c] = fmaf(fmaf(lo[c], 1.0f / EM_LO, y[32 + c]), p.s2, s_b2[half * 64 + 32 + c]);
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_d2empty);           // accumulators are in registers
            // norm_input statistics of the 128-feature row: this thread's half gives (sum, M2 about its own mean); the two halves are
            // combined with the exact pairwise update (Chan et al.): one exchange and one barrier instead of a mean pass and a variance pass
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) s += y[c];
            const float mh = s * (1.0f / 64.0f);
            float m2 = 0.f;
#pragma unroll
            for (int c = 0; c < 64; ++c) { const float d = y[c] - mh; m2 = fmaf(d, d, m2); }
            s_sum[par][half][row] = s;
            s_sq[par][half][row] = m2;
            asm volatile("bar.sync 1, 256;\n" ::: "memory");
            const float s0 = s_sum[par][0][row], s1 = s_sum[par][1][row];
            const float mean = (s0 + s1) * (1.0f / EM_D);
            const float dm = (s0 - s1) * (1.0f / 64.0f);
            const float var = (s_sq[par][0][row] + s_sq[par][1][row] + dm * dm * 32.0f) * (1.0f / EM_D);
            const float rstd = 1.0f / sqrtf(var + p.out_eps);
            const long long m = (long long)it * EM_BM + row;
            if (m < p.M) {
                long long mo = m;
                if (p.pair_hw > 0) {                               // un-interleave the two images of a channel pair (see ps_enc_mlp_planes_paired)
                    const long long q = m >> 1, pr = q / p.pair_hw;
                    mo = (2 * pr + (m & 1)) * p.pair_hw + (q - pr * p.pair_hw);
                }
                uint16_t* oh = p.out_hi + (size_t)mo * EM_D + half * 64;
                uint16_t* ol = p.out_lo + (size_t)mo * EM_D + half * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {                      // 16 features -> one full 32-byte sector per plane
                    float f[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) f[c] = (y[16 * j + c] - mean) * rstd;
                    uint4 h0, l0, h1, l1;
                    em_split8(f, h0, l0);
                    em_split8(f + 8, h1, l1);
                    em_store32(oh + 16 * j, h0, h1);
                    em_store32(ol + 16 * j, l0, l1);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == EM_MMA_WARP) tc::tmem_dealloc(tm, 512);
}

}  // namespace

// tokens X [n_tokens][ldx] fp32 (K1 features used) -> planes [2][n_tokens][128] (fp16 bit patterns, [hi | lo]):
//   h = relu(LN(x; ln_g, ln_b, ln_eps) W1^T + b1);  y = h W2^T + b2;  planes = split((y - mean(y)) * rstd(y; out_eps))
// w1_packed / w2_packed: ops.pack_linear_weight_tc layout [2][128][64] and [2][128][128] (fp16 hi/lo, scaled by w*_scale).
extern "C" int ps_enc_mlp_planes(const float* X, int ldx, int K1, const float* ln_g, const float* ln_b, float ln_eps,
                                 const void* w1_packed, float w1_scale, const float* b1, const void* w2_packed, float w2_scale,
                                 const float* b2, float out_eps, void* planes, long long n_tokens, void* stream) {
    return ps_enc_mlp_planes_paired(X, ldx, K1, ln_g, ln_b, ln_eps, w1_packed, w1_scale, b1, w2_packed, w2_scale, b2, out_eps, planes,
                                    n_tokens, 0, stream);
}

// The same for the tokens of a 32-channel encoder whose images were convolved in channel PAIRS (two images side by side in the
// 64 channels of the tensor-core convolution, SAVi._encode_convs_tc): input row r = (pair * pair_hw + position) * 2 + image,
// i.e. the two images' tokens alternate; the planes are written in image order, row (2 * pair + image) * pair_hw + position.
extern "C" int ps_enc_mlp_planes_paired(const float* X, int ldx, int K1, const float* ln_g, const float* ln_b, float ln_eps,
                                        const void* w1_packed, float w1_scale, const float* b1, const void* w2_packed, float w2_scale,
                                        const float* b2, float out_eps, void* planes, long long n_tokens, int pair_hw, void* stream) {
    PS_REQUIRE(pair_hw >= 0 && (pair_hw == 0 || n_tokens % (2ll * pair_hw) == 0),
               "ps_enc_mlp_planes_paired: n_tokens must be a multiple of 2 * pair_hw (got %lld, %d)", n_tokens, pair_hw);
    PS_REQUIRE(X && ln_g && ln_b && w1_packed && w2_packed && planes, "ps_enc_mlp_planes: null pointer");
    PS_REQUIRE(K1 >= 16 && K1 <= 64 && K1 % 16 == 0, "ps_enc_mlp_planes: K1 must be 16, 32, 48 or 64 (got %d)", K1);
    PS_REQUIRE(ldx >= K1 && ldx % 4 == 0 && ps_aligned16(X) && ps_aligned16(ln_g) && ps_aligned16(ln_b) && ps_aligned16(planes),
               "ps_enc_mlp_planes: tokens / LayerNorm parameters / planes must be 16-byte aligned (ldx multiple of 4)");
    PS_REQUIRE(w1_scale > 0.f && w2_scale > 0.f, "ps_enc_mlp_planes: weight scales must be positive");
    if (n_tokens <= 0) return PS_OK;
    PS_REQUIRE((n_tokens + EM_BM - 1) / EM_BM < (1ll << 31), "ps_enc_mlp_planes: too many tokens");
    EncMlpParams p;
    p.X = X; p.ldx = ldx; p.K1 = K1; p.ln_g = ln_g; p.ln_b = ln_b; p.ln_eps = ln_eps;
    p.b1 = b1; p.b2 = b2; p.s1 = 1.0f / w1_scale; p.s2 = 1.0f / w2_scale; p.out_eps = out_eps;
    p.out_hi = reinterpret_cast<uint16_t*>(planes); p.out_lo = p.out_hi + n_tokens * EM_D;
    p.M = n_tokens; p.n_tiles = (int)((n_tokens + EM_BM - 1) / EM_BM); p.pair_hw = pair_hw;
    CUtensorMap mW1, mW2;
    int r = tc::make_tmap_2d_bf16_sw128(&mW1, w1_packed, 256, 64, 256, 64);
    PS_REQUIRE(r == 0, "ps_enc_mlp_planes: cuTensorMapEncodeTiled(W1) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mW2, w2_packed, 256, 128, 256, 64);
    PS_REQUIRE(r == 0, "ps_enc_mlp_planes: cuTensorMapEncodeTiled(W2) failed (%d)", r);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = EM_SMEM + 1024;
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(enc_mlp_tc_kernel, smem, smem_cfg));
    enc_mlp_tc_kernel<<<p.n_tiles < sms ? p.n_tiles : sms, EM_THREADS, smem, (cudaStream_t)stream>>>(mW1, mW2, p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
