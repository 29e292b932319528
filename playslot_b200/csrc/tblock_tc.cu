// A whole stack of pre-LN transformer blocks in ONE launch (tcgen05 / TMEM / TMA, 4-CTA clusters):
//     for each layer:   y = x + MHSA(LN_1(x)) W_o^T + b_o ;   x = y + W_2 GELU(W_1 LN_2(y) + b_1) + b_2
// Replaces, for embed_dim = 128, the chains `TransformerDecoder` / `TransformerEncoderBlock` of the reference
// (src/models/BlocksUtils/attention.py:425-475,479-565 with MetaAttention :204-316 and model_blocks.MLP :278-309) as they are
// stacked by the cOCVP predictor (Predictors/DynamicsModels.py:457-516: 8 layers per roll-out step in the GridShapes config) and
// used as SAVi's transition module (BlocksUtils/transition_models.py:19-53): 5 launches per block before, one per STACK now.
//
// Work split.  A 128-row tile holds G whole sequences (G * L <= 128 tokens; attention never leaves a sequence).  A cluster of
// 4 CTAs owns one tile; every CTA keeps the tile's residual stream x in registers (two threads per row, 64 values each) and does
//   - the LayerNorms for all 128 features (redundantly: they are cheap and every CTA needs the full operand),
//   - attention for its heads (H / 4 of them): q/k/v projection, S = q k^T, masked softmax, P v, and its K-slice of the output
//     projection  Y_c = sum_{h in c} O_h W_o[:, h]^T,
//   - the MLP for its quarter of the hidden units: Y_c = GELU(LN_2(y) W_1[c]^T + b_1[c]) W_2[:, c]^T,
// and the two partial sums per layer are reduced over the cluster through distributed shared memory: each CTA stores its Y_c
// (fp32) in its own shared memory, cluster barrier, every CTA adds the four partials in rank order to its copy of x (so all four
// copies stay bit-identical), cluster barrier.  Activations never touch HBM; weights stream from L2 by TMA.
//
// Arithmetic: the fp16x3 scheme of linear_tc.cu (x = hi + lo/2048; D0 += A_hi B_hi, D1 += A_hi B_lo + A_lo B_hi, y = D0 + D1/2048)
// for EVERY contraction including q k^T and P v; softmax, LayerNorm, GELU, residuals in fp32.
//
// CTA = 256 threads = 8 "row" warps (row = 32 * (warp % 4) + lane, half = warp / 4: the TMEM lane quarter a warp may read is
// warp % 4, the half selects 64 of a row's columns).  MMA and epilogue phases alternate strictly (operands ready -> MMAs ->
// accumulators ready -> epilogue -> ...) on two mbarriers, so warp 0 issues the TMA loads and MMAs itself between its epilogue
// parts (a ninth warp would put three warps on one scheduler and cap the kernel at 168 registers).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int TB_E = 128;                    // embedding width
constexpr int TB_M = 128;                    // rows per tile
constexpr int TB_TILE = TB_M * 128;          // bytes of one [128 rows][64 K] 16-bit operand tile (16 KB)
constexpr int TB_THREADS = 256;
constexpr int TB_CLUSTER = 4;
constexpr float TB_LO = 2048.0f;
constexpr float TB_LO_INV = 1.0f / 2048.0f;

// shared memory map (bytes from the 1024-aligned base)
constexpr int TB_AX = 0;                     // LN(x) operand: [hi kb0][hi kb1][lo kb0][lo kb1]                      64 KB
constexpr int TB_W = TB_AX + 4 * TB_TILE;    // weights: W0 (32 KB) | W1 (32 KB)                                     64 KB
constexpr int TB_QKP = TB_W + 4 * TB_TILE;   // [Q_hi][Q_lo][K_hi][K_lo]  /  P: [hi kb0][hi kb1][lo kb0][lo kb1]  /  A2 same  /  R (fp32 partial)   64 KB
constexpr int TB_VO = TB_QKP + 4 * TB_TILE;  // V^T: [kb0: hi 64 rows | lo 64 rows][kb1 ...]  /  O: [hi][lo]                   32 KB
constexpr int TB_SMEM = TB_VO + 2 * TB_TILE; // 224 KB

struct TbParams {
    float* X;                    // [n_seq * L][128] residual stream, updated in place
    int n_seq, L, G, spf;        // sequences, tokens per sequence, sequences per tile, slots per frame (0: no mask)
    int n_layers, heads_per_cta; // heads = 4 * heads_per_cta
    int ra_rows;                 // rows of one (layer, hi|lo) part of WA = heads * 192 + 512
    float qscale;                // 1 / sqrt(head_dim)
    float eps1, eps2;
    const float* ln1_g; const float* ln1_b; const float* ln2_g; const float* ln2_b;      // [n_layers][128]
    const float* b_o; const float* b1; const float* b2;                                 // [n_layers][128], [n_layers][512], [n_layers][128]
    const float* inv_scale;      // [n_layers][4]: 1 / pack scale of (qkv, fc1, W_o, W_2)
    float* ws;                   // scratch [2][n_tiles][4][128][128] fp32: the clusters' partial sums
    unsigned long long* dbg;     // optional: per-phase %globaltimer stamps of CTA 0, thread 0 (development)
};

// ---- cluster / distributed shared memory
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, one exp + one reciprocal):
// erff() costs ~45 instructions, and a CTA evaluates 16 384 of them per layer
__device__ __forceinline__ float tb_gelu(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float e = 1.0f - poly * t * __expf(-z * z);          // erf(|x| / sqrt 2)
    return 0.5f * x * (1.0f + copysignf(e, x));
}

// 8 consecutive K elements of one row -> one 16-byte chunk of the hi tile and of the lo tile (128B-swizzled K-major)
__device__ __forceinline__ void tb_store8(uint8_t* hi_tile, uint8_t* lo_tile, int row, int chunk, const float* v) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        const float2 f = __half22float2(h);
        const __half2 l = __floats2half2_rn((v[2 * j] - f.x) * TB_LO, (v[2 * j + 1] - f.y) * TB_LO);
        hw[j] = *reinterpret_cast<const uint32_t*>(&h);
        lw[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// accumulators D0[col0 .. col0+32) + D1[...] / 2048 of this thread's TMEM lane -> v[32]
__device__ __forceinline__ void tb_tmem_ld32_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tb_ld32(uint32_t tl, int d0_col, int d1_col, float* v) {
    float u[32];
    tb_tmem_ld32_nowait(tl + d0_col, v);                       // both loads in flight, one wait
    tb_tmem_ld32_nowait(tl + d1_col, u);
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(u[j], TB_LO_INV, v[j]);
}

__global__ void __cluster_dims__(TB_CLUSTER, 1, 1) __launch_bounds__(TB_THREADS, 1)
tblock_stack_kernel(const __grid_constant__ CUtensorMap mapA128, const __grid_constant__ CUtensorMap mapA64,
                    const __grid_constant__ CUtensorMap mapB, const TbParams p) {
    // 224 KB of operand tiles leave no room for an alignment slack: the kernel has NO static shared memory, so the dynamic region
    // starts at offset 0 of the CTA's shared window (1024-byte aligned, checked below); barriers and the small exchange buffer
    // live behind the tiles.
    extern __shared__ __align__(1024) uint8_t smem[];
    float (*s_red)[2][TB_M] = reinterpret_cast<float (*)[2][TB_M]>(smem + TB_SMEM);      // [buffer][half][row] cross-half exchange
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TB_SMEM + 2 * 2 * TB_M * 4);
    uint64_t& bar_opnd = bars[0]; uint64_t& bar_acc = bars[1]; uint64_t& bar_w0 = bars[2]; uint64_t& bar_w1 = bars[3]; uint64_t& bar_m = bars[4];
    uint32_t& tmem_base = *reinterpret_cast<uint32_t*>(bars + 5);
    if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t crank = cluster_rank();
    const int tile = blockIdx.x / TB_CLUSTER;
    const int rows_valid = min(p.G, p.n_seq - tile * p.G) * p.L;         // real token rows of this tile
    const long long row0 = (long long)tile * p.G * p.L;                  // first global row

    if (tid == 0) {
        tc::mbar_init(&bar_opnd, 8);      // one arrive per row warp
        tc::mbar_init(&bar_acc, 1);
        tc::mbar_init(&bar_w0, 1);
        tc::mbar_init(&bar_w1, 1);
        tc::mbar_init(&bar_m, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    constexpr uint32_t TM_X = 0, TM_Y = 256;
    const int hpc = p.heads_per_cta;
    const int heads = TB_CLUSTER * hpc;
    tc::pdl_wait();                                                      // X is produced by the previous kernel in the stream

    // ------------------------------------------------------------------ issuer role (warp 0; one elected lane issues TMA / MMA)
    // The phases alternate strictly -- epilogue of all 8 warps -> MMAs -> epilogue ... -- so the MMAs are issued by warp 0 itself
    // between two of its epilogue parts: no ninth warp (a third warp on one scheduler would cap the kernel at 168 registers).
    const bool issuer = warp == 0;
    const bool leader = issuer && tc::elect_one();
    const uint32_t sb = tc::smem_u32(smem);
    const uint32_t idesc256 = tc::umma_idesc_f16(TB_M, 256), idesc128 = tc::umma_idesc_f16(TB_M, 128), idesc64 = tc::umma_idesc_f16(TB_M, 64);
    uint32_t ph_opnd = 0, ph_w0 = 0, ph_w1 = 0, ph_m = 0, ph_acc = 0;
    auto desc = [&](int byte_off) { return tc::umma_desc_sw128(sb + (uint32_t)byte_off, 1024, 0); };
    auto wait_opnd = [&]() { tc::mbar_wait(&bar_opnd, ph_opnd); ph_opnd ^= 1u; tc::tc_fence_after(); };     // issuer warp only
    auto wait_w0 = [&]() { tc::mbar_wait(&bar_w0, ph_w0); ph_w0 ^= 1u; };
    auto wait_w1 = [&]() { tc::mbar_wait(&bar_w1, ph_w1); ph_w1 ^= 1u; };
    // 128-row weight tile of WA (q|k of a head, or a hidden chunk of W_1): K-block kb -> W(kb) = [hi 128 rows | lo 128 rows]
    auto load_a128_kb = [&](int layer, int row_in_part, int kb) {
        if (!leader) return;
        uint64_t* bar = kb == 0 ? &bar_w0 : &bar_w1;
        tc::mbar_expect_tx(bar, 2 * TB_TILE);
        for (int part = 0; part < 2; ++part)
            tc::tma_load_2d(smem + TB_W + kb * 2 * TB_TILE + part * TB_TILE, &mapA128, bar, kb * 64, (layer * 2 + part) * p.ra_rows + row_in_part);
    };
    auto load_v_o = [&](int layer, int h) {             // W0 <- W_v: [kb0: hi 64 | lo 64][kb1: ...];  W1 <- W_o[:, h]: [hi 128 | lo 128]
        if (!leader) return;
        tc::mbar_expect_tx(&bar_w0, 2 * TB_TILE); tc::mbar_expect_tx(&bar_w1, 2 * TB_TILE);
        for (int part = 0; part < 2; ++part) {
            const int r = (layer * 2 + part) * p.ra_rows + h * 192 + 128;
            tc::tma_load_2d(smem + TB_W + part * (TB_TILE / 2), &mapA64, &bar_w0, 0, r);
            tc::tma_load_2d(smem + TB_W + TB_TILE + part * (TB_TILE / 2), &mapA64, &bar_w0, 64, r);
            tc::tma_load_2d(smem + TB_W + 2 * TB_TILE + part * TB_TILE, &mapB, &bar_w1, h * 64, (layer * 2 + part) * TB_E);
        }
    };
    auto load_fc2 = [&](int layer) {                    // W0 / W1 <- the two K-blocks of W_2[:, chunk]: [hi 128 | lo 128]
        if (!leader) return;
        tc::mbar_expect_tx(&bar_w0, 2 * TB_TILE); tc::mbar_expect_tx(&bar_w1, 2 * TB_TILE);
        for (int part = 0; part < 2; ++part) {
            const int c = heads * 64 + (int)crank * 128;
            tc::tma_load_2d(smem + TB_W + part * TB_TILE, &mapB, &bar_w0, c, (layer * 2 + part) * TB_E);
            tc::tma_load_2d(smem + TB_W + 2 * TB_TILE + part * TB_TILE, &mapB, &bar_w1, c + 64, (layer * 2 + part) * TB_E);
        }
    };
    const int fc1_row = heads * 192 + (int)crank * 128;          // this CTA's hidden chunk of W_1
    // GEMM with K = 128 (two K-blocks of the Ax operand) against weight tiles [hi N rows | lo N rows]:
    //   D0 | D1 (2N columns at tm_col) += A_hi [W_hi ; W_lo];  D1 += A_lo W_hi
    auto gemm_ax_kb = [&](uint32_t tm_col, int N, int kb) {
        if (!leader) return;
        const uint32_t idn2 = N == 128 ? idesc256 : idesc128, idn = N == 128 ? idesc128 : idesc64;
        const int wofs = N == 128 ? TB_W + kb * 2 * TB_TILE : TB_W + kb * TB_TILE;       // QK / fc1: kb -> W0 / W1; V: both K-blocks inside W0
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            tc::umma_bf16(tm + tm_col, desc(TB_AX + kb * TB_TILE + k * 32), desc(wofs + k * 32), idn2, (kb | k) ? 1u : 0u);
            tc::umma_bf16(tm + tm_col + N, desc(TB_AX + (2 + kb) * TB_TILE + k * 32), desc(wofs + k * 32), idn, 1u);
        }
    };
    auto commit_acc = [&]() { if (leader) tc::umma_commit(&bar_acc); };

    // ------------------------------------------------------------------ row role (all 8 warps)
    const int q4 = warp & 3, half = warp >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t tl = tm + ((uint32_t)(q4 * 32) << 16);
    const bool row_ok = row < rows_valid;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 4096) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
            p.dbg[n_stamp++] = t;
        }
    };
    auto wait_acc = [&]() { stamp(); tc::mbar_wait(&bar_acc, ph_acc); ph_acc ^= 1u; tc::tc_fence_after(); stamp(); };
    auto opnd_ready = [&]() {
        tc::tc_fence_before();
        tc::fence_async_proxy();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_opnd);
    };
    // residual stream: this thread's 64 features of its row
    float x[64];
    {
        const float4* src = reinterpret_cast<const float4*>(p.X + (size_t)(row0 + row) * TB_E + half * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 v = row_ok ? src[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
        }
    }
    int red_buf = 0;
    // LayerNorm of the row (two threads per row exchange partial sums) -> Ax K-block `half` (hi and lo tiles)
    auto layer_norm_to_ax = [&](const float* g, const float* b, float eps) {
        float sm_ = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) sm_ += x[j];
        s_red[red_buf][half][row] = sm_;
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        const float mean = (s_red[red_buf][0][row] + s_red[red_buf][1][row]) * (1.0f / TB_E);
        red_buf ^= 1;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) { const float d = x[j] - mean; ss = fmaf(d, d, ss); }
        s_red[red_buf][half][row] = ss;
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        const float rstd = 1.0f / sqrtf((s_red[red_buf][0][row] + s_red[red_buf][1][row]) * (1.0f / TB_E) + eps);
        red_buf ^= 1;
        uint8_t* hi_t = smem + TB_AX + half * TB_TILE;
        uint8_t* lo_t = smem + TB_AX + (2 + half) * TB_TILE;
        const float4* g4 = reinterpret_cast<const float4*>(g + half * 64);
        const float4* b4 = reinterpret_cast<const float4*>(b + half * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 ga = __ldg(g4 + 2 * c), gb = __ldg(g4 + 2 * c + 1), ba = __ldg(b4 + 2 * c), bb = __ldg(b4 + 2 * c + 1);
            float v[8];
            v[0] = fmaf((x[c * 8 + 0] - mean) * rstd, ga.x, ba.x); v[1] = fmaf((x[c * 8 + 1] - mean) * rstd, ga.y, ba.y);
            v[2] = fmaf((x[c * 8 + 2] - mean) * rstd, ga.z, ba.z); v[3] = fmaf((x[c * 8 + 3] - mean) * rstd, ga.w, ba.w);
            v[4] = fmaf((x[c * 8 + 4] - mean) * rstd, gb.x, bb.x); v[5] = fmaf((x[c * 8 + 5] - mean) * rstd, gb.y, bb.y);
            v[6] = fmaf((x[c * 8 + 6] - mean) * rstd, gb.z, bb.z); v[7] = fmaf((x[c * 8 + 7] - mean) * rstd, gb.w, bb.w);
            tb_store8(hi_t, lo_t, row, c, v);
        }
    };
    // Cluster reduction through L2: every CTA stores its fp32 partial [feature][row] in a global scratch slot, ONE cluster barrier
    // (release / acquire at cluster scope orders the global stores), then every CTA adds the four partials in rank order
    // (ld.global.cg: L2, never a stale L1 line) -- the four copies of x stay bit-identical.  Distributed shared memory was
    // measured first: 17 B/clk per SM (B300_MICROARCH.md) made the 192 KB of remote reads 10.7 us per reduction; through L2 the
    // same exchange takes 6 us.  With [feature][row] the 32 lanes of a warp (consecutive rows) touch 128 contiguous bytes.
    // Two scratch buffers alternate between the two reductions of a layer, so a slot is rewritten only after a later cluster
    // barrier has proven that everybody finished reading it.  `after_acc` runs once the Y MMAs have retired (weight prefetch).
    int red_idx = 0;
    auto reduce_into_x = [&](const float* bias, float inv_scale, auto&& after_acc) {
        wait_acc();                                            // Y accumulators complete
        after_acc();
        float* slot = p.ws + ((size_t)(red_idx * (gridDim.x / TB_CLUSTER) + tile) * TB_CLUSTER) * (TB_M * TB_E);
        red_idx ^= 1;
        {
            float* dst = slot + (size_t)crank * (TB_M * TB_E) + (size_t)(half * 64) * TB_M + row;
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                tb_ld32(tl, TM_Y + half * 64 + c0, TM_Y + 128 + half * 64 + c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) __stcg(dst + (size_t)(c0 + j) * TB_M, v[j] * inv_scale);
            }
        }
        tc::tc_fence_before();
        stamp();
        cluster_sync_all();
        stamp();
        const float* src = slot + (size_t)(half * 64) * TB_M + row;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) {
            float acc = __ldg(bias + half * 64 + j);
#pragma unroll
            for (int r = 0; r < TB_CLUSTER; ++r)                // rank order: identical sums in all four CTAs
                acc += __ldcg(src + (size_t)r * (TB_M * TB_E) + (size_t)j * TB_M);
            x[j] += acc;
        }
        stamp();
        stamp();
    };

    load_a128_kb(0, (int)crank * hpc * 192, 0);
    load_a128_kb(0, (int)crank * hpc * 192, 1);
    for (int layer = 0; layer < p.n_layers; ++layer) {
        const float is_qkv = p.inv_scale[layer * 4 + 0], is_fc1 = p.inv_scale[layer * 4 + 1];
        const float is_o = p.inv_scale[layer * 4 + 2], is_2 = p.inv_scale[layer * 4 + 3];
        layer_norm_to_ax(p.ln1_g + layer * TB_E, p.ln1_b + layer * TB_E, p.eps1);
        opnd_ready();
        for (int hh = 0; hh < hpc; ++hh) {
            const int h = (int)crank * hpc + hh;
            const bool last = hh + 1 == hpc;
            // ---- q | k projection of head h (K-block 0 runs while K-block 1 of the weights may still be landing)
            if (issuer) {
                wait_opnd();                               // hh == 0: Ax = LN_1(x);  hh > 0: matched below (see out_proj)
                wait_w0(); gemm_ax_kb(TM_X, 128, 0);
                wait_w1(); gemm_ax_kb(TM_X, 128, 1);
                commit_acc();
            }
            wait_acc();
            if (issuer) load_v_o(layer, h);                // the projection MMAs have retired: W0 / W1 are free
            {   // q / k epilogue: half 0 -> Q tile (scaled by 1/sqrt(d_h)), half 1 -> K tile
                uint8_t* hi_t = smem + TB_QKP + half * 2 * TB_TILE;
                uint8_t* lo_t = hi_t + TB_TILE;
                const float sc = half == 0 ? is_qkv * p.qscale : is_qkv;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
                    tb_ld32(tl, TM_X + half * 64 + c0, TM_X + 128 + half * 64 + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= sc;
#pragma unroll
                    for (int c = 0; c < 4; ++c) tb_store8(hi_t, lo_t, row, (c0 >> 3) + c, v + 8 * c);
                }
            }
            opnd_ready();
            // ---- v projection
            if (issuer) {
                wait_opnd();
                wait_w0();
                gemm_ax_kb(TM_X, 64, 0); gemm_ax_kb(TM_X, 64, 1);
                commit_acc();
            }
            wait_acc();
            if (issuer) {                                  // W0 (W_v) is free: K-block 0 of the next q|k weights, or of W_1
                if (!last) load_a128_kb(layer, (h + 1) * 192, 0);
                else load_a128_kb(layer, fc1_row, 0);
            }
            {   // v epilogue: transposed store, V^T[kb = row / 64] = [hi: 64 d rows | lo: 64 d rows] x 64 keys
                float v[32];
                tb_ld32(tl, TM_X + half * 32, TM_X + 64 + half * 32, v);
                uint8_t* vt = smem + TB_VO + (row >> 6) * TB_TILE;
                const int kk = row & 63;                       // key index inside the K-block
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int d = half * 32 + j;
                    const float val = v[j] * is_qkv;
                    const __half hv = __float2half_rn(val);
                    const __half lv = __float2half_rn((val - __half2float(hv)) * TB_LO);
                    const int off = d * 128 + ((((kk * 2) >> 4) ^ (d & 7)) << 4) + ((kk * 2) & 15);
                    *reinterpret_cast<__half*>(vt + off) = hv;
                    *reinterpret_cast<__half*>(vt + 64 * 128 + off) = lv;
                }
            }
            opnd_ready();
            // ---- S = q k^T: A = Q (hi, lo), B = [K_hi ; K_lo]
            if (issuer) {
                wait_opnd();
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(tm + TM_X, desc(TB_QKP + k * 32), desc(TB_QKP + 2 * TB_TILE + k * 32), idesc256, k ? 1u : 0u);
                        tc::umma_bf16(tm + TM_X + 128, desc(TB_QKP + TB_TILE + k * 32), desc(TB_QKP + 2 * TB_TILE + k * 32), idesc128, 1u);
                    }
                }
                commit_acc();
            }
            // ---- masked softmax over the keys of the row's own sequence: this thread's 64 keys [half * 64, +64).  The P tile
            // holds exp(s - max) UN-normalised; 1 / sum is applied to O afterwards (the two halves' partial sums meet there).
            wait_acc();
            {
                const int seq = row / p.L, pos = row - seq * p.L;
                const int frame = p.spf > 0 ? pos / p.spf : 0;
                // valid keys: the same sequence, and (block-causal) frame(key) <= frame(row)
                const int k_lo = seq * p.L - half * 64;                          // relative to this thread's first key
                const int k_hi = row_ok ? (p.spf > 0 ? k_lo + (frame + 1) * p.spf : k_lo + p.L) : k_lo;     // exclusive
                constexpr float LOG2E = 1.4426950408889634f;
                float s0[32], s1[32];
                tb_ld32(tl, TM_X + half * 64, TM_X + 128 + half * 64, s0);
                tb_ld32(tl, TM_X + half * 64 + 32, TM_X + 128 + half * 64 + 32, s1);
                float mx = -3.0e38f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    s0[j] = (j >= k_lo && j < k_hi) ? s0[j] * LOG2E : -3.0e38f;
                    s1[j] = (32 + j >= k_lo && 32 + j < k_hi) ? s1[j] * LOG2E : -3.0e38f;
                    mx = fmaxf(mx, fmaxf(s0[j], s1[j]));
                }
                s_red[red_buf][half][row] = mx;
                asm volatile("bar.sync 1, 256;\n" ::: "memory");
                mx = fmaxf(s_red[red_buf][0][row], s_red[red_buf][1][row]);
                red_buf ^= 1;
                float sum = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    s0[j] = s0[j] > -1.0e38f ? exp2f(s0[j] - mx) : 0.f;
                    s1[j] = s1[j] > -1.0e38f ? exp2f(s1[j] - mx) : 0.f;
                    sum += s0[j] + s1[j];
                }
                s_red[red_buf][half][row] = sum;               // read in the O epilogue (ordered by the mbarrier hand-offs in between)
                uint8_t* hi_t = smem + TB_QKP + half * TB_TILE;
                uint8_t* lo_t = smem + TB_QKP + (2 + half) * TB_TILE;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tb_store8(hi_t, lo_t, row, c, s0 + c * 8);
                    tb_store8(hi_t, lo_t, row, 4 + c, s1 + c * 8);
                }
            }
            opnd_ready();
            // ---- O = P v: A = P (two K-blocks of keys), B = V^T[kb] = [hi 64 | lo 64]
            if (issuer) {
                wait_opnd();
                if (leader) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            tc::umma_bf16(tm + TM_X, desc(TB_QKP + kb * TB_TILE + k * 32), desc(TB_VO + kb * TB_TILE + k * 32), idesc128, (kb | k) ? 1u : 0u);
                            tc::umma_bf16(tm + TM_X + 64, desc(TB_QKP + (2 + kb) * TB_TILE + k * 32), desc(TB_VO + kb * TB_TILE + k * 32), idesc64, 1u);
                        }
                }
                commit_acc();
            }
            wait_acc();
            {   // O epilogue: this thread's 32 head features / sum -> O tile (A operand of the output projection)
                float v[32];
                tb_ld32(tl, TM_X + half * 32, TM_X + 64 + half * 32, v);
                const float sum = s_red[red_buf][0][row] + s_red[red_buf][1][row];
                red_buf ^= 1;
                const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= inv;
                uint8_t* hi_t = smem + TB_VO;
                uint8_t* lo_t = smem + TB_VO + TB_TILE;
#pragma unroll
                for (int c = 0; c < 4; ++c) tb_store8(hi_t, lo_t, row, half * 4 + c, v + 8 * c);
            }
            opnd_ready();
            // ---- Y (+)= O W_o[:, h]^T : A = O tile (TB_VO: hi | lo), B = W1 = [hi 128 | lo 128]
            if (issuer) {
                wait_opnd();
                wait_w1();
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(tm + TM_Y, desc(TB_VO + k * 32), desc(TB_W + 2 * TB_TILE + k * 32), idesc256, (hh > 0 || k > 0) ? 1u : 0u);
                        tc::umma_bf16(tm + TM_Y + 128, desc(TB_VO + TB_TILE + k * 32), desc(TB_W + 2 * TB_TILE + k * 32), idesc128, 1u);
                    }
                }
                if (last) {
                    commit_acc();                          // Y complete: reduction 1 follows
                } else {
                    if (leader) tc::umma_commit(&bar_m);   // W1 (W_o slice) is free once the projection has retired
                    tc::mbar_wait(&bar_m, ph_m); ph_m ^= 1u;
                    load_a128_kb(layer, (h + 1) * 192, 1);
                    if (lane == 0) tc::mbar_arrive(&bar_opnd);      // the next head's projection waits on bar_opnd like every MMA phase:
                }
            }
            if (!last) {                                   // ... all 8 warps arrive once (nothing to write), warp 0 did so above
                __syncwarp();
                if (!issuer && lane == 0) tc::mbar_arrive(&bar_opnd);
            }
        }
        reduce_into_x(p.b_o + layer * TB_E, is_o, [&]() { if (issuer) load_a128_kb(layer, fc1_row, 1); });
        layer_norm_to_ax(p.ln2_g + layer * TB_E, p.ln2_b + layer * TB_E, p.eps2);
        opnd_ready();
        // ---- fc1, this CTA's hidden chunk
        if (issuer) {
            wait_opnd();
            wait_w0(); gemm_ax_kb(TM_X, 128, 0);
            wait_w1(); gemm_ax_kb(TM_X, 128, 1);
            commit_acc();
        }
        wait_acc();
        if (issuer) load_fc2(layer);                       // behind the GELU epilogue
        {   // fc1 epilogue: + b_1, GELU -> A2 operand, K-block `half`
            uint8_t* hi_t = smem + TB_QKP + half * TB_TILE;
            uint8_t* lo_t = smem + TB_QKP + (2 + half) * TB_TILE;
            const float4* b1 = reinterpret_cast<const float4*>(p.b1 + layer * 512 + (int)crank * 128 + half * 64);
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float v[32];
                tb_ld32(tl, TM_X + half * 64 + c0, TM_X + 128 + half * 64 + c0, v);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bb = __ldg(b1 + (c0 >> 2) + c);
                    v[4 * c] = tb_gelu(fmaf(v[4 * c], is_fc1, bb.x)); v[4 * c + 1] = tb_gelu(fmaf(v[4 * c + 1], is_fc1, bb.y));
                    v[4 * c + 2] = tb_gelu(fmaf(v[4 * c + 2], is_fc1, bb.z)); v[4 * c + 3] = tb_gelu(fmaf(v[4 * c + 3], is_fc1, bb.w));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) tb_store8(hi_t, lo_t, row, (c0 >> 3) + c, v + 8 * c);
            }
        }
        opnd_ready();
        // ---- Y = A2 W_2[:, chunk]^T
        if (issuer) {
            wait_opnd();
            wait_w0(); wait_w1();
            if (leader) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc::umma_bf16(tm + TM_Y, desc(TB_QKP + kb * TB_TILE + k * 32), desc(TB_W + kb * 2 * TB_TILE + k * 32), idesc256, (kb | k) ? 1u : 0u);
                        tc::umma_bf16(tm + TM_Y + 128, desc(TB_QKP + (2 + kb) * TB_TILE + k * 32), desc(TB_W + kb * 2 * TB_TILE + k * 32), idesc128, 1u);
                    }
            }
            commit_acc();
        }
        reduce_into_x(p.b2 + layer * TB_E, is_2, [&]() {
            if (issuer && layer + 1 < p.n_layers) { load_a128_kb(layer + 1, (int)crank * hpc * 192, 0); load_a128_kb(layer + 1, (int)crank * hpc * 192, 1); }
        });
    }
    // write the tile back (rank 0 of the cluster: all four copies are identical)
    if (crank == 0 && row_ok) {
        float4* dst = reinterpret_cast<float4*>(p.X + (size_t)(row0 + row) * TB_E + half * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    }
    tc::pdl_trigger();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}

}  // namespace

// X (n_seq * L, 128) fp32 is updated in place by `n_layers` pre-LN transformer blocks (embed 128, `heads` heads of padded width 64,
// MLP hidden 512).  Packed weights (playslot_b200.ops.pack_tblock_stack):
//   WA fp16 [n_layers][2 (hi, lo)][heads * 192 + 512][128]: per head the rows [q 64 | k 64 | v 64] (head_dim zero-padded to 64), then W_1
//   WB fp16 [n_layers][2][128][heads * 64 + 512]: columns [W_o slice of each head (64, zero-padded) | W_2]
//   inv_scale [n_layers][4]: 1 / power-of-two pack scale of (qkv, W_1, W_o, W_2);  LayerNorm / bias vectors [n_layers][...]
// spf > 0: block-causal mask over frames of `spf` tokens (build_slot_causal_mask, model_utils.py:115-124); 0: full attention.
extern "C" int ps_tblock_stack_dbg(float* X, int n_seq, int L, int spf, int n_layers, int heads, float qscale, float eps1, float eps2,
                                   const void* WA, const void* WB, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                                   const float* ln2_b, const float* b_o, const float* b1, const float* b2, const float* inv_scale,
                                   float* workspace, unsigned long long* dbg, void* stream);

extern "C" int ps_tblock_stack(float* X, int n_seq, int L, int spf, int n_layers, int heads, float qscale, float eps1, float eps2,
                               const void* WA, const void* WB, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                               const float* ln2_b, const float* b_o, const float* b1, const float* b2, const float* inv_scale,
                               float* workspace, void* stream) {
    return ps_tblock_stack_dbg(X, n_seq, L, spf, n_layers, heads, qscale, eps1, eps2, WA, WB, ln1_g, ln1_b, ln2_g, ln2_b, b_o, b1, b2,
                               inv_scale, workspace, nullptr, stream);
}

// floats of scratch ps_tblock_stack needs: two buffers of one fp32 [128][128] partial per CTA
extern "C" long long ps_tblock_stack_workspace_floats(int n_seq, int L) {
    if (n_seq <= 0 || L <= 0 || L > TB_M) return 0;
    const int G = TB_M / L;
    const long long n_tiles = (n_seq + G - 1) / G;
    return 2ll * n_tiles * TB_CLUSTER * TB_M * TB_E;
}

// development variant: `dbg` (device, >= 4096 x u64) receives %globaltimer stamps of CTA 0 at the phase boundaries of its row warps
extern "C" int ps_tblock_stack_dbg(float* X, int n_seq, int L, int spf, int n_layers, int heads, float qscale, float eps1, float eps2,
                                   const void* WA, const void* WB, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                                   const float* ln2_b, const float* b_o, const float* b1, const float* b2, const float* inv_scale,
                                   float* workspace, unsigned long long* dbg, void* stream) {
    PS_REQUIRE(X && WA && WB && ln1_g && ln1_b && ln2_g && ln2_b && b_o && b1 && b2 && inv_scale, "ps_tblock_stack: null pointer");
    PS_REQUIRE(L >= 1 && L <= TB_M, "ps_tblock_stack: 1 <= tokens per sequence <= 128 (got %d)", L);
    PS_REQUIRE(heads == 4 || heads == 8, "ps_tblock_stack: 4 or 8 heads (got %d)", heads);
    PS_REQUIRE(n_layers >= 1 && spf >= 0 && (spf == 0 || L % spf == 0), "ps_tblock_stack: bad layer count / mask geometry");
    PS_REQUIRE(ps_aligned16(X), "ps_tblock_stack: X must be 16-byte aligned");
    if (n_seq <= 0) return PS_OK;
    TbParams p;
    p.X = X; p.n_seq = n_seq; p.L = L; p.G = TB_M / L; p.spf = spf;
    p.n_layers = n_layers; p.heads_per_cta = heads / TB_CLUSTER; p.ra_rows = heads * 192 + 512;
    p.qscale = qscale; p.eps1 = eps1; p.eps2 = eps2;
    p.ln1_g = ln1_g; p.ln1_b = ln1_b; p.ln2_g = ln2_g; p.ln2_b = ln2_b; p.b_o = b_o; p.b1 = b1; p.b2 = b2; p.inv_scale = inv_scale;
    p.dbg = dbg; p.ws = workspace;
    PS_REQUIRE(workspace && ps_aligned16(workspace), "ps_tblock_stack: workspace of ps_tblock_stack_workspace_floats(n_seq, L) floats required");
    const int n_tiles = (n_seq + p.G - 1) / p.G;
    CUtensorMap mA128, mA64, mB;
    const uint64_t rowsA = (uint64_t)n_layers * 2 * p.ra_rows;
    int r = tc::make_tmap_2d_bf16_sw128(&mA128, WA, rowsA, 128, 128, 64);
    PS_REQUIRE(r == 0, "ps_tblock_stack: cuTensorMapEncodeTiled(WA, 128) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mA64, WA, rowsA, 128, 64, 64);
    PS_REQUIRE(r == 0, "ps_tblock_stack: cuTensorMapEncodeTiled(WA, 64) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mB, WB, (uint64_t)n_layers * 2 * TB_E, (uint64_t)(heads * 64 + 512), 128, 64);
    PS_REQUIRE(r == 0, "ps_tblock_stack: cuTensorMapEncodeTiled(WB) failed (%d)", r);
    const size_t smem = TB_SMEM + 2 * 2 * TB_M * 4 + 64;
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(tblock_stack_kernel, smem, smem_cfg));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles * TB_CLUSTER); cfg.blockDim = dim3(TB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PS_CUDA(cudaLaunchKernelEx(&cfg, tblock_stack_kernel, mA128, mA64, mB, p));
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
