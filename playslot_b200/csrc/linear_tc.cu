// tcgen05 linear layer:  Y[M,N] = act( LN?(X)[M,K] * W[N,K]^T + bias ) + res   with fp32-grade accuracy.
// Replaces the nn.LayerNorm + nn.Linear (+ReLU / GELU / residual) chains of the reference on the tensor
// cores: encoder MLP (src/models/SAVi.py:111-116,214-216), transformer q/k/v/out projections and MLPs
// (src/models/BlocksUtils/attention.py:296,462-475,539-563; model_blocks.py:293-309), the cOCVP embeddings and
// regression head (Predictors/DynamicsModels.py:119-148,487-510), the InvDyn encoders (LatentAction.py).
//
// Arithmetic.  Same "fp16x3" scheme as conv_tc.cu: x = hi + lo/2048 with hi = fp16(x), lo = fp16((x-hi)*2048);
// D0 += A_hi*W_hi and D1 += A_hi*W_lo + A_lo*W_hi in two TMEM accumulators, Y = D0 + D1/2048 (22 mantissa bits).
// precision 2 / 1: a single fp16 / bf16 product.
//
// Structure.  One CTA per (128-row, BN-column) output tile, 320 threads:
//   warps 0-7 : A producers -- read the fp32 activations (coalesced float4, 8 loads in flight per thread), apply
//               the LayerNorm (two-pass statistics held in registers), split to fp16 hi/lo and store them
//               128B-swizzled K-major into the stage ring; afterwards the epilogue (TMEM -> registers ->
//               shared-memory transpose -> bias / activation / residual -> coalesced 128-byte row stores)
//   warp 8    : MMA issuer (one elected thread) and TMEM owner
//   warp 9    : TMA producer of the packed weights [hi | lo][Npad][Kp] (128B swizzle)
// K is consumed in blocks of 64 (one 128-byte swizzle row of fp16) through an mbarrier ring.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128;
constexpr int KB = 64;                     // K elements per stage
constexpr int A_TILE = BM * 128;           // bytes of one 16-bit A tile
constexpr int LT_MAX_STAGES = 4;
constexpr int LT_THREADS = 320;
constexpr float LT_LO_SCALE = 2048.0f;

struct LinTcParams {
    const float* X; int ldx;
    const float* bias;
    const float* ln_g; const float* ln_b; float ln_eps;
    const float* res; int ldres; int res_div; int res_mod;
    float* Y; int ldy;
    int M, N, K, nkb, npad;
    int act, stages, bf16, vec_out;
    float out_scale;
    int a_planes;            // 1: the activations arrive as fp16 [hi | lo] planes and are loaded by TMA (no conversion work)
    int a_lo_row0;           // first row of the lo plane inside the plane tensor map
    uint16_t* y_hi;          // planes output instead of Y (fp16 [hi | lo][M][ldy], lo scaled by 2048), else null
    uint16_t* y_lo;
};

// 4 consecutive K elements -> 8 bytes of the hi tile (and of the lo tile)
template <int NSPLIT>
__device__ __forceinline__ void split_store4(const float4 v, uint8_t* hi, uint8_t* lo, bool bf16) {
    if (bf16) {
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
        return;
    }
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    if (NSPLIT == 2) {
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        const __half2 l0 = __floats2half2_rn((v.x - f0.x) * LT_LO_SCALE, (v.y - f0.y) * LT_LO_SCALE);
        const __half2 l1 = __floats2half2_rn((v.z - f1.x) * LT_LO_SCALE, (v.w - f1.y) * LT_LO_SCALE);
        *reinterpret_cast<uint2*>(lo) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
}

// ring slot / round of K block kb for a run-time ring depth (1..4) without an integer division
__device__ __forceinline__ void ring_pos(int kb, int stages, int& s, int& q) {
    switch (stages) {
        case 4: s = kb & 3; q = kb >> 2; break;
        case 3: q = kb / 3; s = kb - 3 * q; break;
        case 2: s = kb & 1; q = kb >> 1; break;
        default: s = 0; q = kb; break;
    }
}

template <int BN, int NSPLIT, bool PIPE>
__global__ void __launch_bounds__(LT_THREADS, PIPE ? 1 : 2) linear_tc_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapA, const LinTcParams p) {
    constexpr int W_TILE = BN * 128;
    constexpr int STAGE_BYTES = NSPLIT * (A_TILE + W_TILE);
    constexpr int TM_COLS = (NSPLIT * BN < 32) ? 32 : NSPLIT * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_afull[LT_MAX_STAGES], bar_wfull[LT_MAX_STAGES], bar_empty[LT_MAX_STAGES], bar_acc;
    __shared__ uint32_t tmem_base;
    __shared__ float s_mean[BM], s_rstd[BM];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay on the uniform datapath
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int STAGES = p.stages;

    if (tid == 0) {
        for (int s = 0; s < LT_MAX_STAGES; ++s) {
            tc::mbar_init(&bar_afull[s], 8);      // one arrive per producer warp
            tc::mbar_init(&bar_wfull[s], 1);
            tc::mbar_init(&bar_empty[s], 1);
        }
        tc::mbar_init(&bar_acc, 1);
        tc::fence_mbar_init();
    }
    if (warp == 8) tc::tmem_alloc(&tmem_base, TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    // PDL: the prologue above and the weight TMA loads (parameters, never written inside a step) overlap the tail of the
    // previous kernel; everything that reads activations or writes Y waits for it here.
    tc::pdl_trigger();

    if (warp < 8) {
        // ------------------------------------------------------------------ A producers
        tc::pdl_wait();
        if (!p.a_planes) {
        const int r_sub = tid >> 4;               // 0..15
        const int c4 = tid & 15;                  // float4 index inside the 64-wide K block
        const bool ln = p.ln_g != nullptr;
        const bool ln_fused = ln && p.K <= 128 && !PIPE;     // K <= 128: one load round serves statistics AND both K blocks
        if (ln && !ln_fused) {
            // two-pass LayerNorm statistics, 16 lanes per row, 2 rows per round.  K <= 256: the row slice stays in
            // registers between the passes; larger K re-reads (L1-resident).
            const int nk = (p.K + 63) >> 6;
#pragma unroll 1
            for (int i0 = 0; i0 < 8; i0 += 2) {
                float4 v[2][4];
                float s[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int m = m0 + (i0 + i) * 16 + r_sub;
                    s[i] = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int k = c * 64 + c4 * 4;
                        v[i][c] = (m < p.M && k < p.K) ? *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (nk > 4 && m < p.M)
                        for (int k = 256 + c4 * 4; k < p.K; k += 64) {
                            const float4 t = *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k);
                            s[i] += (t.x + t.y) + (t.z + t.w);
                        }
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) s[i] += (v[i][c].x + v[i][c].y) + (v[i][c].z + v[i][c].w);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int row = (i0 + i) * 16 + r_sub;
                    const int m = m0 + row;
                    const float mean = s[i] / (float)p.K;
                    float ss = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c * 64 + c4 * 4 < p.K) {
                            const float a = v[i][c].x - mean, b = v[i][c].y - mean, cc = v[i][c].z - mean, d = v[i][c].w - mean;
                            ss += (a * a + b * b) + (cc * cc + d * d);
                        }
                    }
                    if (nk > 4 && m < p.M)
                        for (int k = 256 + c4 * 4; k < p.K; k += 64) {
                            const float4 t = *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k);
                            const float a = t.x - mean, b = t.y - mean, cc = t.z - mean, d = t.w - mean;
                            ss += (a * a + b * b) + (cc * cc + d * d);
                        }
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                    if (c4 == 0) { s_mean[row] = mean; s_rstd[row] = 1.0f / sqrtf(ss / (float)p.K + p.ln_eps); }
                }
            }
            __syncwarp();                          // the rows a warp normalises are the rows it measured
        }
        auto load_kb = [&](float4 (&v)[8], int kb) {
            const int k = kb * KB + c4 * 4;
            const bool k_ok = kb < p.nkb && k < p.K;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + i * 16 + r_sub;
                v[i] = (m < p.M && k_ok) ? *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto store_kb = [&](const float4 (&v)[8], int kb) {
            int s, rq;
            ring_pos(kb, STAGES, s, rq);
            uint8_t* a_hi = smem + s * STAGE_BYTES;
            uint8_t* a_lo = a_hi + A_TILE;
            const int k = kb * KB + c4 * 4;
            const bool k_ok = k < p.K;
            float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ln && k_ok) {
                g4 = *reinterpret_cast<const float4*>(p.ln_g + k);
                b4 = *reinterpret_cast<const float4*>(p.ln_b + k);
            }
            if (kb >= STAGES) tc::mbar_wait(&bar_empty[s], (rq - 1) & 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = i * 16 + r_sub;
                float4 x = v[i];
                if (ln) {
                    const float mu = s_mean[row], rs = s_rstd[row];
                    x.x = (x.x - mu) * rs * g4.x + b4.x; x.y = (x.y - mu) * rs * g4.y + b4.y;
                    x.z = (x.z - mu) * rs * g4.z + b4.z; x.w = (x.w - mu) * rs * g4.w + b4.w;
                    if (!(m0 + row < p.M && k_ok)) x = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int off = row * 128 + ((((c4 >> 1) ^ (row & 7))) << 4) + ((c4 & 1) << 3);
                split_store4<NSPLIT>(x, a_hi + off, a_lo + off, p.bf16 != 0);
            }
            tc::fence_async_proxy();               // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_afull[s]);
        };
        if (ln_fused) {
            // LayerNorm over K <= 128 (every transformer / encoder LayerNorm of the shipped configs): the 8 rows x 2 K-blocks
            // a thread owns are loaded once, reduced over the 16 lanes of a row (two passes, in registers), normalised,
            // split and stored -- no statistics pre-pass, no second read
            float4 v[8][2];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + i * 16 + r_sub;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int k = c * 64 + c4 * 4;
                    v[i][c] = (m < p.M && k < p.K) ? *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float mean[8], rstd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float sm_ = ((v[i][0].x + v[i][0].y) + (v[i][0].z + v[i][0].w)) + ((v[i][1].x + v[i][1].y) + (v[i][1].z + v[i][1].w));
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) sm_ += __shfl_xor_sync(0xffffffffu, sm_, o);
                mean[i] = sm_ / (float)p.K;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c * 64 + c4 * 4 < p.K) {
                        const float a = v[i][c].x - mean[i], b = v[i][c].y - mean[i], cc = v[i][c].z - mean[i], d = v[i][c].w - mean[i];
                        ss += (a * a + b * b) + (cc * cc + d * d);
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                rstd[i] = 1.0f / sqrtf(ss / (float)p.K + p.ln_eps);
            }
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                if (kb >= p.nkb) break;
                int s, rq;
                ring_pos(kb, STAGES, s, rq);
                uint8_t* a_hi = smem + s * STAGE_BYTES;
                uint8_t* a_lo = a_hi + A_TILE;
                const int k = kb * KB + c4 * 4;
                const bool k_ok = k < p.K;
                float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k_ok) {
                    g4 = *reinterpret_cast<const float4*>(p.ln_g + k);
                    b4 = *reinterpret_cast<const float4*>(p.ln_b + k);
                }
                if (kb >= STAGES) tc::mbar_wait(&bar_empty[s], (rq - 1) & 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = i * 16 + r_sub;
                    float4 x = v[i][kb];
                    x.x = (x.x - mean[i]) * rstd[i] * g4.x + b4.x; x.y = (x.y - mean[i]) * rstd[i] * g4.y + b4.y;
                    x.z = (x.z - mean[i]) * rstd[i] * g4.z + b4.z; x.w = (x.w - mean[i]) * rstd[i] * g4.w + b4.w;
                    if (!(m0 + row < p.M && k_ok)) x = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int off = row * 128 + ((((c4 >> 1) ^ (row & 7))) << 4) + ((c4 & 1) << 3);
                    split_store4<NSPLIT>(x, a_hi + off, a_lo + off, p.bf16 != 0);
                }
                tc::fence_async_proxy();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bar_afull[s]);
            }
        } else if (PIPE) {
            // long K (small-M rollout GEMMs, one CTA per SM): the loads of the next K block are in flight while this one
            // is normalised, split and stored
            // (register ring of four K blocks: three loads are always in flight behind the block being converted)
            float4 v0[8], v1[8], v2[8], v3[8];
            load_kb(v0, 0); load_kb(v1, 1); load_kb(v2, 2);
            for (int kb = 0; kb < p.nkb; kb += 4) {
                load_kb(v3, kb + 3);
                store_kb(v0, kb);
                load_kb(v0, kb + 4);
                if (kb + 1 < p.nkb) store_kb(v1, kb + 1);
                load_kb(v1, kb + 5);
                if (kb + 2 < p.nkb) store_kb(v2, kb + 2);
                load_kb(v2, kb + 6);
                if (kb + 3 < p.nkb) store_kb(v3, kb + 3);
            }
        } else {
            for (int kb = 0; kb < p.nkb; ++kb) {
                float4 v[8];
                load_kb(v, kb);
                store_kb(v, kb);
            }
        }
        }   // !a_planes
        // ------------------------------------------------------------------ epilogue
        // warp w owns TMEM lanes (w % 4) * 32 .. +31 (rows) and the column half w / 4.  Each 32 x CW chunk is
        // transposed through shared memory so that every global store instruction writes one contiguous row piece.
        constexpr int CW = (BN >= 64) ? 32 : 16;                     // chunk width
        constexpr int CHUNKS = BN / (2 * CW);                        // chunks per warp
        tc::mbar_wait(&bar_acc, 0);
        tc::tc_fence_after();
        const int q = warp & 3, ch = warp >> 2;
        float* scratch = reinterpret_cast<float*>(smem) + warp * (32 * 33);   // stage ring is dead now
        const uint32_t tl = tm + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < CHUNKS; ++c) {
            const int col0 = (ch * CHUNKS + c) * CW;
            float v[CW];
            if constexpr (CW == 32) tc::tmem_ld_32x32(tl + col0, v); else tc::tmem_ld_32x16(tl + col0, v);
            if (NSPLIT == 2) {
                float u[CW];
                if constexpr (CW == 32) tc::tmem_ld_32x32(tl + BN + col0, u); else tc::tmem_ld_32x16(tl + BN + col0, u);
#pragma unroll
                for (int j = 0; j < CW; ++j) v[j] = fmaf(u[j], 1.0f / LT_LO_SCALE, v[j]);
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) scratch[lane * 33 + j] = v[j] * p.out_scale;
            __syncwarp();
            if (p.vec_out && n0 + col0 + CW <= p.N) {
                // lanes walk along rows in float4 pieces: one store instruction writes RPV rows x CW floats
                constexpr int LPR = CW / 4, RPV = 32 / LPR;          // lanes per row, rows per store instruction
                const int rsub4 = lane / LPR, c4o = (lane % LPR) * 4;
                const int n = n0 + col0 + c4o;
                float4 bv4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias) bv4 = *reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll 2
                for (int r = 0; r < 32; r += RPV) {
                    const int row = q * 32 + r + rsub4;
                    const int m = m0 + row;
                    const float* sp = scratch + (r + rsub4) * 33 + c4o;
                    float t[4] = {sp[0] + bv4.x, sp[1] + bv4.y, sp[2] + bv4.z, sp[3] + bv4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (p.act == 1) t[j] = fmaxf(t[j], 0.f);
                        else if (p.act == 2) t[j] = gelu_erf(t[j]);
                        else if (p.act == 3) t[j] = fabsf(t[j]);
                    }
                    if (m < p.M) {
                        if (p.res) {
                            const float4 rv = *reinterpret_cast<const float4*>(p.res + (size_t)((m / p.res_div) % p.res_mod) * p.ldres + n);
                            t[0] += rv.x; t[1] += rv.y; t[2] += rv.z; t[3] += rv.w;
                        }
                        if (p.y_hi) {          // the consumer is another tensor-core linear: hand it the split operand directly
                            const size_t o = (size_t)m * p.ldy + n;
                            split_store4<2>(make_float4(t[0], t[1], t[2], t[3]), reinterpret_cast<uint8_t*>(p.y_hi + o),
                                            reinterpret_cast<uint8_t*>(p.y_lo + o), false);
                        } else
                        *reinterpret_cast<float4*>(p.Y + (size_t)m * p.ldy + n) = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
                __syncwarp();
                continue;
            }
            // lanes now walk along a row: lane = column (two rows per instruction when CW == 16)
            constexpr int RPI = 32 / CW;                             // rows per instruction
            const int jj = lane % CW, rsub = lane / CW;
            const int n = n0 + col0 + jj;
            const bool n_ok = n < p.N;
            const float bv = (p.bias && n_ok) ? __ldg(p.bias + n) : 0.f;
#pragma unroll 4
            for (int r = 0; r < 32; r += RPI) {
                const int row = q * 32 + r + rsub;
                const int m = m0 + row;
                float t = scratch[(r + rsub) * 33 + jj] + bv;
                if (p.act == 1) t = fmaxf(t, 0.f);
                else if (p.act == 2) t = gelu_erf(t);
                else if (p.act == 3) t = fabsf(t);
                if (m < p.M && n_ok) {
                    if (p.res) t += __ldg(p.res + (size_t)((m / p.res_div) % p.res_mod) * p.ldres + n);
                    p.Y[(size_t)m * p.ldy + n] = t;
                }
            }
            __syncwarp();
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        // the whole warp walks the loop (uniform control flow and descriptor arithmetic, ring slot / parity advanced without
        // divisions); one elected lane issues the MMAs and the commits
        const bool leader = tc::elect_one();
        {
            const uint32_t idesc = p.bf16 ? tc::umma_idesc_bf16(BM, BN) : tc::umma_idesc_f16(BM, BN);
            const uint32_t idesc2 = tc::umma_idesc_f16(BM, 2 * BN <= 256 ? 2 * BN : BN);
            const uint64_t dA = tc::umma_desc_sw128(tc::smem_u32(smem), 1024, 0);
            constexpr uint64_t STAGE_U = STAGE_BYTES / 16, A_U = A_TILE / 16, W_U = W_TILE / 16;
            int st = 0;
            uint32_t ph = 0;
            for (int kb = 0; kb < p.nkb; ++kb) {
                tc::mbar_wait(&bar_wfull[st], ph);
                if (!p.a_planes) tc::mbar_wait(&bar_afull[st], ph);
                tc::tc_fence_after();
                const uint64_t a_hi0 = dA + (uint64_t)st * STAGE_U;
                const uint64_t w_hi0 = a_hi0 + NSPLIT * A_U;
#pragma unroll
                for (int k = 0; k < KB / 16; ++k) {
                    const uint32_t first = (uint32_t)(kb | k) ? 1u : 0u;
                    const uint64_t a_hi = a_hi0 + (uint64_t)(k * 2);
                    const uint64_t b_hi = w_hi0 + (uint64_t)(k * 2);
                    if (NSPLIT == 2 && 2 * BN <= 256) {
                        // A_hi x [W_hi ; W_lo] in one N = 2*BN MMA (the two weight tiles are adjacent in the stage, D0 | D1
                        // adjacent in TMEM), then D1 += A_lo x W_hi
                        if (leader) tc::umma_bf16(tm, a_hi, b_hi, idesc2, first);
                        if (leader) tc::umma_bf16(tm + BN, a_hi + A_U, b_hi, idesc, 1u);
                    } else if (NSPLIT == 2) {                      // BN = 256: an MMA cannot be wider, three N = BN products
                        if (leader) tc::umma_bf16(tm, a_hi, b_hi, idesc, first);
                        if (leader) tc::umma_bf16(tm + BN, a_hi, b_hi + W_U, idesc, first);
                        if (leader) tc::umma_bf16(tm + BN, a_hi + A_U, b_hi, idesc, 1u);
                    } else {
                        if (leader) tc::umma_bf16(tm, a_hi, b_hi, idesc, first);
                    }
                }
                if (leader) tc::umma_commit(&bar_empty[st]);
                if (++st == STAGES) { st = 0; ph ^= 1u; }
            }
            if (leader) tc::umma_commit(&bar_acc);
        }
    } else {
        // ------------------------------------------------------------------ TMA producer of the weights
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&mapW);
            if (p.a_planes) tc::tma_prefetch_desc(&mapA);
            for (int kb = 0; kb < p.nkb; ++kb) {
                int s, rq;
                ring_pos(kb, STAGES, s, rq);
                if (kb >= STAGES) tc::mbar_wait(&bar_empty[s], (rq - 1) & 1);
                uint8_t* a_hi = smem + s * STAGE_BYTES;
                uint8_t* w_hi = a_hi + NSPLIT * A_TILE;
                tc::mbar_expect_tx(&bar_wfull[s], NSPLIT * W_TILE + (p.a_planes ? NSPLIT * A_TILE : 0));
                tc::tma_load_2d(w_hi, &mapW, &bar_wfull[s], kb * KB, n0);
                if (NSPLIT == 2) tc::tma_load_2d(w_hi + W_TILE, &mapW, &bar_wfull[s], kb * KB, p.npad + n0);
                if (p.a_planes) {
                    // activation planes written by the previous kernel: the weight loads above did not have to wait for it
                    if (kb == 0) tc::pdl_wait();
                    tc::tma_load_2d(a_hi, &mapA, &bar_wfull[s], kb * KB, m0);
                    if (NSPLIT == 2) tc::tma_load_2d(a_hi + A_TILE, &mapA, &bar_wfull[s], kb * KB, p.a_lo_row0 + m0);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tm, TM_COLS);
}

template <int BN, int NSPLIT, bool PIPE>
int launch_linear_tc_impl(const CUtensorMap& mW, const CUtensorMap& mA, const LinTcParams& p, cudaStream_t st) {
    constexpr int STAGE_BYTES = NSPLIT * (A_TILE + BN * 128);
    size_t smem = (size_t)p.stages * STAGE_BYTES;
    if (smem < 8 * 32 * 33 * sizeof(float)) smem = 8 * 32 * 33 * sizeof(float);      // epilogue transpose scratch
    smem += 1024;
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(linear_tc_kernel<BN, NSPLIT, PIPE>, smem, smem_cfg));
    dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN);
    PS_CUDA(tc::launch_pdl(linear_tc_kernel<BN, NSPLIT, PIPE>, grid, dim3(LT_THREADS), smem, st, mW, mA, p));
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

template <int BN, int NSPLIT>
int launch_linear_tc(const CUtensorMap& mW, const CUtensorMap& mA, const LinTcParams& p, cudaStream_t st) {
    // software-pipelined producers when K is long and the grid leaves at most one CTA per SM anyway
    const long long ctas = (long long)((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    if (!p.a_planes && p.nkb >= 3 && ctas <= 2 * 148) return launch_linear_tc_impl<BN, NSPLIT, true>(mW, mA, p, st);
    return launch_linear_tc_impl<BN, NSPLIT, false>(mW, mA, p, st);
}

}  // namespace

static int linear_tc_run(const float* X, int ldx, const void* x_planes, long long x_plane_rows, const void* w_tc, float w_scale,
                         int Kp, int Npad, const float* bias, const float* ln_gamma, const float* ln_beta, float ln_eps,
                         const float* res, int ldres, int res_div, int res_mod, float* Y, void* y_planes, long long y_plane_rows,
                         int ldy, int M, int N, int K, int act, int precision, void* stream) {
    PS_REQUIRE((X || x_planes) && w_tc && (Y || y_planes), "ps_linear_tc: null pointer");
    if (x_planes) {
        PS_REQUIRE(precision == 3 && !ln_gamma, "ps_linear_tc_planes: plane input needs the fp16x3 mode and no LayerNorm prologue");
        PS_REQUIRE(ldx % 8 == 0 && ldx >= K && ps_aligned16(x_planes) && x_plane_rows >= M && 2 * x_plane_rows < (1ll << 31),
                   "ps_linear_tc_planes: plane input needs a row pitch (%d) that is a multiple of 8 elements and 16-byte alignment", ldx);
        X = reinterpret_cast<const float*>(x_planes);      // only for the alignment checks below
    }
    if (y_planes) {
        PS_REQUIRE(precision == 3 && ldy % 4 == 0 && N % 32 == 0 && ps_aligned16(y_planes) && y_plane_rows >= M,
                   "ps_linear_tc_planes: plane output needs the fp16x3 mode, N (%d) multiple of 32 and ldy (%d) multiple of 4", N, ldy);
    }
    PS_REQUIRE(M >= 0 && N > 0 && K > 0, "ps_linear_tc: bad shape M=%d N=%d K=%d", M, N, K);
    if (M == 0) return PS_OK;
    PS_REQUIRE(precision >= 1 && precision <= 3, "ps_linear_tc: precision must be 1 (bf16), 2 (fp16) or 3 (fp16x3)");
    PS_REQUIRE(K % 4 == 0 && ldx % 4 == 0 && ps_aligned16(X), "ps_linear_tc: K (%d), ldx (%d) must be multiples of 4 and X 16-byte aligned", K, ldx);
    PS_REQUIRE(Kp % KB == 0 && Kp >= K && Npad % 128 == 0 && Npad >= N, "ps_linear_tc: packed weight geometry Kp=%d Npad=%d does not cover K=%d N=%d", Kp, Npad, K, N);
    PS_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "ps_linear_tc: LayerNorm needs both gamma and beta");
    PS_REQUIRE(!ln_gamma || (ps_aligned16(ln_gamma) && ps_aligned16(ln_beta)), "ps_linear_tc: LN params must be 16-byte aligned");
    PS_REQUIRE(act >= 0 && act <= 3, "ps_linear_tc: unknown activation %d", act);
    PS_REQUIRE(w_scale > 0.f, "ps_linear_tc: w_scale must be positive");
    PS_REQUIRE((N + 31) / 32 <= 65535, "ps_linear_tc: N too large");
    const int nsplit = precision == 3 ? 2 : 1;
    LinTcParams p;
    p.X = X; p.ldx = ldx; p.bias = bias; p.ln_g = ln_gamma; p.ln_b = ln_beta; p.ln_eps = ln_eps;
    p.res = res; p.ldres = ldres; p.res_div = res_div > 0 ? res_div : 1; p.res_mod = res_mod > 0 ? res_mod : M;
    p.Y = Y; p.ldy = ldy; p.M = M; p.N = N; p.K = K; p.nkb = Kp / KB; p.npad = Npad;
    p.act = act; p.bf16 = precision == 1 ? 1 : 0; p.out_scale = 1.0f / w_scale;
    p.vec_out = (ldy % 4 == 0 && (y_planes || ps_aligned16(Y)) && (!bias || ps_aligned16(bias)) && (!res || (ldres % 4 == 0 && ps_aligned16(res)))) ? 1 : 0;
    p.a_planes = x_planes ? 1 : 0;
    p.a_lo_row0 = (int)x_plane_rows;
    p.y_hi = reinterpret_cast<uint16_t*>(y_planes);
    p.y_lo = p.y_hi ? p.y_hi + (size_t)y_plane_rows * ldy : nullptr;
    PS_REQUIRE(!y_planes || p.vec_out, "ps_linear_tc_planes: plane output needs 16-byte aligned bias / residual");
    // Column tile.  Every CTA normalises / splits its own copy of the 128-row activation tile, so narrow tiles multiply that
    // CUDA-core work: take the widest tile (<= 256) that still leaves about one CTA per two SMs.
    const long long mt = (M + BM - 1) / BM;
    const int n32 = ((N + 31) / 32) * 32;
    int bn = 256;
    while (bn > 32 && (bn / 2 >= n32 || mt * ((N + bn - 1) / bn) < 74)) bn >>= 1;
    if (bn == 256 && nsplit == 1) bn = 128;            // (the single-product instantiation of the widest tile is not built)
    const int stage_bytes = nsplit * (A_TILE + bn * 128);
    int stages = (100 * 1024) / stage_bytes;         // at most ~100 KB per CTA: two CTAs per SM overlap each other's phases
    if (bn == 256) stages = 2;                       // 96 KB stages, one CTA per SM (its 512 TMEM columns fill the SM anyway)
    if (mt * ((N + bn - 1) / bn) <= 148) {           // one CTA per SM anyway: use the shared memory for a deeper ring
        stages = (200 * 1024) / stage_bytes;
        if (stages < 1) stages = 1;
        if (stages > LT_MAX_STAGES) stages = LT_MAX_STAGES;
    }
    if (stages < 1) stages = 1;
    if (stages > LT_MAX_STAGES) stages = LT_MAX_STAGES;
    if (stages > p.nkb) stages = p.nkb;
    p.stages = stages;
    CUtensorMap mW, mA;
    int r = tc::make_tmap_2d_bf16_sw128(&mW, w_tc, 2ull * Npad, (uint64_t)Kp, (uint32_t)bn, KB);
    PS_REQUIRE(r == 0, "ps_linear_tc: cuTensorMapEncodeTiled(W) failed (%d)", r);
    mA = mW;                                          // unused unless the activations arrive as planes
    if (x_planes) {
        // [hi rows 0..rows-1 | lo rows rows..2*rows-1] x ldx fp16; columns >= K of a box are weighted by zero-padded weight columns
        r = tc::make_tmap_2d_bf16_sw128(&mA, x_planes, 2ull * (uint64_t)x_plane_rows, (uint64_t)ldx, BM, KB);
        PS_REQUIRE(r == 0, "ps_linear_tc_planes: cuTensorMapEncodeTiled(A) failed (%d)", r);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (bn == 256) return launch_linear_tc<256, 2>(mW, mA, p, st);
    if (bn == 128) return nsplit == 2 ? launch_linear_tc<128, 2>(mW, mA, p, st) : launch_linear_tc<128, 1>(mW, mA, p, st);
    if (bn == 64) return nsplit == 2 ? launch_linear_tc<64, 2>(mW, mA, p, st) : launch_linear_tc<64, 1>(mW, mA, p, st);
    return nsplit == 2 ? launch_linear_tc<32, 2>(mW, mA, p, st) : launch_linear_tc<32, 1>(mW, mA, p, st);
}

extern "C" int ps_linear_tc(const float* X, int ldx, const void* w_tc, float w_scale, int Kp, int Npad, const float* bias,
                            const float* ln_gamma, const float* ln_beta, float ln_eps,
                            const float* res, int ldres, int res_div, int res_mod,
                            float* Y, int ldy, int M, int N, int K, int act, int precision, void* stream) {
    PS_REQUIRE(X && Y, "ps_linear_tc: null pointer");
    return linear_tc_run(X, ldx, nullptr, 0, w_tc, w_scale, Kp, Npad, bias, ln_gamma, ln_beta, ln_eps, res, ldres, res_div, res_mod,
                         Y, nullptr, 0, ldy, M, N, K, act, precision, stream);
}

// The same linear layer with fp16 [hi | lo] plane operands on either side, for chains of tensor-core linears (MLP fc1 -> fc2,
// attention -> out projection): x_planes [2][x_plane_rows][ldx] replaces X (loaded by TMA: no conversion work in the kernel),
// y_planes [2][y_plane_rows][ldy] replaces Y (the epilogue writes the split of the fp32 result).  Either may be NULL.
extern "C" int ps_linear_tc_planes(const float* X, int ldx, const void* x_planes, long long x_plane_rows, const void* w_tc,
                                   float w_scale, int Kp, int Npad, const float* bias, const float* ln_gamma, const float* ln_beta,
                                   float ln_eps, const float* res, int ldres, int res_div, int res_mod, float* Y, void* y_planes,
                                   long long y_plane_rows, int ldy, int M, int N, int K, int act, void* stream) {
    PS_REQUIRE((X != nullptr) != (x_planes != nullptr), "ps_linear_tc_planes: exactly one of X / x_planes");
    PS_REQUIRE((Y != nullptr) != (y_planes != nullptr), "ps_linear_tc_planes: exactly one of Y / y_planes");
    return linear_tc_run(X, ldx, x_planes, x_plane_rows, w_tc, w_scale, Kp, Npad, bias, ln_gamma, ln_beta, ln_eps, res, ldres, res_div,
                         res_mod, Y, y_planes, y_plane_rows, ldy, M, N, K, act, 3, stream);
}
