// tcgen05 implicit-GEMM 5x5 convolution, 64 -> 64 channels (decoder conv2..4 of every config, encoder conv2..4; 32-channel
// encoders run zero-padded to 64 channels).  Reference op: nn.Conv2d(C, C, 5, padding=2) (+ReLU) in
// src/models/BlocksUtils/model_blocks.py:48-79 as stacked by EncodersDecoders/{encoders,decoders}.py.
//
// Layout.  Activations live in HBM as zero-bordered NHWC 16-bit planes: position p = (y+2)*Wp + (x+2),
// Wp = W+4, Hp = H+4, one 128-byte row of 64 channels per position; an activation tensor is two such planes
// [hi | lo] (fp16 split of the fp32 value, see below; a single bf16 plane in the bf16 mode).  On this "padded-linear" grid a
// filter tap (ky,kx) is a constant row shift (ky-2)*Wp + (kx-2), so for a tile of 128 consecutive positions the A operand
// of every tap is a 128-row window of ONE shared-memory resident region (tile + halo), addressed purely by the
// UMMA descriptor start address (128B swizzle is applied on absolute smem address bits, verified by
// tools/tc_probe.cu; tools/tc_rate.cu: windows that start off an 8-row swizzle atom cost nothing).  Outputs computed at
// border columns are junk by construction and are written as zeros, which maintains the zero border of the next layer's
// input.
//
// GEMM per tile: D[128 pos x 64 cout] += A_tap[128 x 64 cin] * W_tap[64 cout x 64 cin]^T over 25 taps, 16-bit
// operands, fp32 accumulation in TMEM.  Precision modes:
//   3 = "fp16x3": x = hi + lo/2048 with hi = fp16(x), lo = fp16((x - hi) * 2048)  (22 mantissa bits, the scaled
//       low part never goes subnormal); D0 += A_hi*W_hi and D1 += A_hi*W_lo + A_lo*W_hi in a second TMEM
//       accumulator, result = D0 + D1/2048.  The dropped lo*lo term is 2^-22 relative: fp32-grade.
//   2 = "fp16" : A_hi*W_hi only with fp16 operands;   1 = "bf16": the same with bf16 operands.
// (A bf16 hi/lo split was measured first and rejected: 16 mantissa bits put the decoder's slot-softmax masks
//  at 1.6e-4..3.7e-4 max error, above the 1e-4 parity bar.)
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 epilogue.  Persistent (one CTA per
// SM), double-buffered TMEM accumulators, 4-stage weight ring, staged region refilled chunk by chunk (see the kernel).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int C = 64;              // channels in == channels out
constexpr int ROWB = 128;          // bytes per position row (64 bf16)
constexpr int TILE_M = 128;
constexpr int MAX_STAGES = 4;
constexpr int ABOX = 32;           // rows per TMA box of the activation region
constexpr int WTAP_BYTES = 64 * ROWB;   // one [cout][cin] bf16 tile

constexpr float LO_SCALE = 2048.0f;      // low parts are stored multiplied by 2^11

// 8 consecutive channels -> 16 bytes of the hi plane (and of the lo plane when lo != nullptr).
// bf16 == true: single bf16 plane (precision 1);  otherwise fp16 hi (+ scaled fp16 lo).
__device__ __forceinline__ void split_store8(const float* v, uint16_t* hi, uint16_t* lo, bool bf16) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        if (bf16) {
            hw[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
            lw[j] = 0;
        } else {
            const __half ah = __float2half_rn(a), bh = __float2half_rn(b);
            const __half al = __float2half_rn((a - __half2float(ah)) * LO_SCALE);
            const __half bl = __float2half_rn((b - __half2float(bh)) * LO_SCALE);
            hw[j] = (uint32_t)__half_as_ushort(ah) | ((uint32_t)__half_as_ushort(bh) << 16);
            lw[j] = (uint32_t)__half_as_ushort(al) | ((uint32_t)__half_as_ushort(bl) << 16);
        }
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (lo) *reinterpret_cast<uint4*>(lo) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// 16 consecutive channels -> one 32-byte (256-bit) store per plane: the rows a warp writes are 128 bytes apart, so 16-byte
// stores would touch every 32-byte sector twice
__device__ __forceinline__ void split_store16(const float* v, uint16_t* hi, uint16_t* lo, bool bf16) {
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float a = v[2 * j], b = v[2 * j + 1];
        if (bf16) {
            hw[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
            lw[j] = 0;
        } else {
            const __half2 h = __floats2half2_rn(a, b);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn((a - f.x) * LO_SCALE, (b - f.y) * LO_SCALE);
            hw[j] = *reinterpret_cast<const uint32_t*>(&h);
            lw[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
    }
    tc::st_global_v8(hi, hw[0], hw[1], hw[2], hw[3], hw[4], hw[5], hw[6], hw[7]);
    if (lo) tc::st_global_v8(lo, lw[0], lw[1], lw[2], lw[3], lw[4], lw[5], lw[6], lw[7]);
}

struct ConvTcParams {
    const float* bias;
    const float* addend;     // mode 2: [H*W][64]
    const uint16_t* mask_hi; // optional: hi plane of a padded split activation; outputs are zeroed where it is <= 0
    uint16_t* out_hi;        // mode 0
    uint16_t* out_lo;        // mode 0 (null in the single-product modes)
    float* out_f32;          // mode 1 (NCHW) / mode 2 (NHWC)
    int n_img, H, W, Wp, Hp;
    int items_per_img;       // work items (NPOS positions each) per image
    int halo, a_rows;        // halo rows on each side, rows reserved per staged plane (multiple of 64: 1024-byte aligned regions)
    int chunk_b[6];          // the staged region is (re)filled in five row chunks [chunk_b[c], chunk_b[c+1]): chunk c < 4 is dead once
                             // the taps of filter row ky = c have been issued, chunk 4 at the end of the item
    int n_early;             // chunks the first filter row needs (4, or 5 when the image is too narrow to split)
    int stages;              // weight ring depth (2..MAX_STAGES)
    int n_ctas;              // persistent grid size (one CTA per SM)
    int relu, out_mode;
    int bf16;                // operand format: 1 = bf16, 0 = fp16
    float out_scale;         // 1 / (power-of-two scale folded into the packed weights)
};

// Persistent: every CTA walks over work items (NPOS consecutive positions of one image) with stride gridDim.x.  The
// accumulators are double-buffered in TMEM, so the epilogue of item i (TMEM -> registers -> global) overlaps the TMA
// loads and the MMAs of item i+1; the weight ring keeps streaming across items.
template <int TILES, int NSPLIT>
__global__ void __launch_bounds__(192, 1)
conv5x5_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                  const __grid_constant__ CUtensorMap mapA_hi8, const __grid_constant__ CUtensorMap mapA_lo8,
                  const __grid_constant__ CUtensorMap mapW, const ConvTcParams p) {
    constexpr int NPOS = TILE_M * TILES;
    constexpr uint32_t ACC_COLS = 64 * TILES * NSPLIT;      // TMEM columns of one accumulator buffer
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = p.a_rows * ROWB;
    uint8_t* sA[2] = {smem, smem + a_bytes};
    uint8_t* sW = smem + NSPLIT * a_bytes;                 // stages x NSPLIT x 8 KB
    __shared__ uint64_t bar_afull[2], bar_aempty[5], bar_wfull[MAX_STAGES], bar_wempty[MAX_STAGES], bar_accfull[2], bar_accempty[2];
    const int STAGES = p.stages;
    __shared__ uint32_t tmem_base;
    __shared__ float s_bias[C];

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform: role branches stay on the uniform datapath
    const int n_items = p.n_img * p.items_per_img;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) tc::mbar_init(&bar_afull[b], 1);
        for (int c = 0; c < 5; ++c) tc::mbar_init(&bar_aempty[c], 1);
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&bar_accfull[b], 1); tc::mbar_init(&bar_accempty[b], 4); }
        for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(&bar_wfull[s], 1); tc::mbar_init(&bar_wempty[s], 1); }
        tc::fence_mbar_init();
    }
    if (tid < C) s_bias[tid] = p.bias ? p.bias[tid] : 0.f;
    if (warp == 1) tc::tmem_alloc(&tmem_base, 2 * ACC_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&mapA_hi);
            tc::tma_prefetch_desc(&mapW);
            int ws = 0;                    // weight ring slot and the parity of its "empty" barrier, advanced per tap
            uint32_t wph = 1;              // (a fresh barrier passes a wait on parity 1: the first STAGES loads do not block)
            auto region_row0 = [&](int it) {                        // first global row of the staged region of work item `it`
                const int img = it / p.items_per_img;
                const int item = it - img * p.items_per_img;
                return (int)((long long)img * p.Hp * p.Wp + 2 * p.Wp + item * NPOS - p.halo);
            };
            // chunk c of the region of an item: 32-row boxes, then 8-row boxes up to the chunk end (both multiples of 8 rows)
            auto load_chunk = [&](int g0, int c) {
                const int fb = c < p.n_early ? 0 : 1;
                if (c == 0 || c == p.n_early) {
                    const int rows = c == 0 ? p.chunk_b[p.n_early] : p.chunk_b[5] - p.chunk_b[p.n_early];
                    tc::mbar_expect_tx(&bar_afull[fb], (uint32_t)(NSPLIT * rows * ROWB));
                }
                int r = p.chunk_b[c];
                const int r1 = p.chunk_b[c + 1];
                for (; r + ABOX <= r1; r += ABOX) {
                    tc::tma_load_2d(sA[0] + r * ROWB, &mapA_hi, &bar_afull[fb], 0, g0 + r);
                    if (NSPLIT == 2) tc::tma_load_2d(sA[1] + r * ROWB, &mapA_lo, &bar_afull[fb], 0, g0 + r);
                }
                for (; r < r1; r += 8) {
                    tc::tma_load_2d(sA[0] + r * ROWB, &mapA_hi8, &bar_afull[fb], 0, g0 + r);
                    if (NSPLIT == 2) tc::tma_load_2d(sA[1] + r * ROWB, &mapA_lo8, &bar_afull[fb], 0, g0 + r);
                }
            };
            int i = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
                const int g0 = region_row0(it);
                const bool has_next = it + (int)gridDim.x < n_items;
                const int g0_next = has_next ? region_row0(it + gridDim.x) : 0;
                if (i == 0)
                    for (int c = 0; c < 5; ++c) load_chunk(g0, c);
                for (int tap = 0; tap < 25; ++tap) {
                    tc::mbar_wait(&bar_wempty[ws], wph);
                    tc::mbar_expect_tx(&bar_wfull[ws], NSPLIT * WTAP_BYTES);
                    tc::tma_load_2d(sW + ws * NSPLIT * WTAP_BYTES, &mapW, &bar_wfull[ws], 0, tap * 128);
                    if (++ws == STAGES) { ws = 0; wph ^= 1u; }
                    // Region refill, interleaved with the weight stream at the points where the MMA warp is known to be past the
                    // rows in question (it has consumed tap - STAGES when the wait above returns), so these waits do not stall
                    // the weight ring: chunk 4 of THIS item once the previous item is finished, chunks 0..3 of the NEXT item as
                    // the filter rows of this item retire.  The next item's first filter row only reads chunks 0..3.
                    if (i > 0 && tap == STAGES - 1) {
                        tc::mbar_wait(&bar_aempty[4], (i - 1) & 1);
                        load_chunk(g0, 4);
                    }
                    if (has_next) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (tap == 5 * c + 4 + STAGES) {                     // <= 23: STAGES <= MAX_STAGES = 4
                                tc::mbar_wait(&bar_aempty[c], i & 1);
                                load_chunk(g0_next, c);
                            }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the loop (uniform control flow and uniform-register address arithmetic); one elected lane
        // issues the MMAs and the commits.
        const bool leader = tc::elect_one();
        {
            const uint32_t idesc = p.bf16 ? tc::umma_idesc_bf16(TILE_M, C) : tc::umma_idesc_f16(TILE_M, C);
            const uint32_t idesc2 = tc::umma_idesc_f16(TILE_M, 2 * C);
            // The single issuing thread must not become the bottleneck (the MMA queue is shallow: scalar work between two
            // MMAs that takes longer than the queued MMAs idles the tensor pipe).  So: descriptors are built once and per
            // MMA only a compile-time constant is added to the 14-bit start-address field (16-byte units); ring slots,
            // parities and tap shifts are advanced incrementally (no division by the run-time stage count); and the wait
            // for the NEXT tap's weights is issued in the middle of the current tap's MMAs.
            const uint64_t a_hi0 = tc::umma_desc_sw128(tc::smem_u32(sA[0]) + (uint32_t)p.halo * ROWB, 1024, 0);
            const uint64_t a_lo0 = tc::umma_desc_sw128(tc::smem_u32(sA[1]) + (uint32_t)p.halo * ROWB, 1024, 0);
            const uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(sW), 1024, 0);
            constexpr uint64_t TILE_UNITS = (uint64_t)TILE_M * ROWB / 16;             // descriptor units between the tiles
            constexpr uint64_t STAGE_UNITS = (uint64_t)NSPLIT * WTAP_BYTES / 16;
            const int row_step = (p.Wp - 5) * (ROWB / 16);                            // ky -> ky+1 after five kx steps
            int ws = 0;
            uint32_t wph = 0;
            int i = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
                const int buf = i & 1;
                if (i >= 2) tc::mbar_wait(&bar_accempty[buf], ((i >> 1) - 1) & 1);   // epilogue drained this accumulator buffer
                tc::mbar_wait(&bar_afull[0], i & 1);                                 // the rows the first filter row reads
                tc::mbar_wait(&bar_wfull[ws], wph);                                  // weights of tap 0
                tc::tc_fence_after();
                int a_off = (-2 * p.Wp - 2) * (ROWB / 16);                           // tap (0, 0) shift, 16-byte units
                const uint32_t acc_buf = tm + buf * ACC_COLS;
#pragma unroll 1
                for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) {
                        const uint64_t a_hi = a_hi0 + (uint64_t)(int64_t)a_off;
                        const uint64_t a_lo = a_lo0 + (uint64_t)(int64_t)a_off;
                        const uint64_t b_st = bdesc0 + (uint64_t)ws * STAGE_UNITS;
                        uint64_t* const done_bar = &bar_wempty[ws];
                        if (++ws == STAGES) { ws = 0; wph ^= 1u; }
#pragma unroll
                        for (int tile = 0; tile < TILES; ++tile) {
                            const uint32_t acc0 = acc_buf + tile * NSPLIT * C;        // D0 | D1 are adjacent: 2 x 64 columns
#pragma unroll
                            for (int k = 0; k < C / 16; ++k) {
                                const uint32_t first = (ky | kx | k) ? 1u : 0u;
                                if (NSPLIT == 2) {
                                    // one N=128 MMA against [W_hi ; W_lo] (the two weight tiles are adjacent in the stage):
                                    // D0 += A_hi W_hi and D1 += A_hi W_lo share a single read of A_hi; then D1 += A_lo W_hi.
                                    // 14 KB of shared-memory operand traffic per k-step instead of 18 KB for three N=64 MMAs.
                                    if (leader) tc::umma_bf16(acc0, a_hi + (tile * TILE_UNITS + k * 2), b_st + k * 2, idesc2, first);
                                    if (leader) tc::umma_bf16(acc0 + C, a_lo + (tile * TILE_UNITS + k * 2), b_st + k * 2, idesc, 1u);
                                } else {
                                    if (leader) tc::umma_bf16(acc0, a_hi + (tile * TILE_UNITS + k * 2), b_st + k * 2, idesc, first);
                                }
                            }
                            if (tile == 0 && !(ky == 4 && kx == 4)) {
                                // the next tap's weights: normally long there (ring of STAGES taps); checked here, with
                                // this tap's MMAs still queued, so that the check costs no tensor-pipe time
                                tc::mbar_wait(&bar_wfull[ws], wph);
                                tc::tc_fence_after();
                            }
                        }
                        if (leader) tc::umma_commit(done_bar);             // frees the weight stage once these MMAs retire
                        a_off += ROWB / 16;
                    }
                    a_off += row_step;
                    if (leader) tc::umma_commit(&bar_aempty[ky]);      // region chunk ky is dead (chunk 4: the rest of the region)
                    if (ky == 0 && p.n_early < 5) {
                        tc::mbar_wait(&bar_afull[1], i & 1);           // the remaining rows (loaded while filter row 0 was issued)
                        tc::tc_fence_after();
                    }
                }
                if (leader) tc::umma_commit(&bar_accfull[buf]);        // the accumulators of this item are complete
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int quad = warp & 3;                          // TMEM lane quadrant this warp may access
        const int HW = p.H * p.W;
        int i = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
            const int buf = i & 1;
            const int img = it / p.items_per_img;
            const int item = it - img * p.items_per_img;
            const int p0 = 2 * p.Wp + item * NPOS;
            const long long img_base = (long long)img * p.Hp * p.Wp;
            tc::mbar_wait(&bar_accfull[buf], (i >> 1) & 1);
            __syncwarp();
            tc::tc_fence_after();
#pragma unroll 1
            for (int tile = 0; tile < TILES; ++tile) {
                const int r = quad * 32 + lane;
                const int pp = p0 + tile * TILE_M + r;          // padded-linear position inside the image
                const int yp = pp / p.Wp, xp = pp - yp * p.Wp;
                const bool row_ok = yp < p.H + 2;               // rows >= 2 by construction
                const bool interior = row_ok && xp >= 2 && xp < p.W + 2;
                float v[64];
                const uint32_t tbase = tm + ((uint32_t)(quad * 32) << 16) + buf * ACC_COLS + tile * NSPLIT * C;
                tc::tmem_ld_32x32(tbase, v);
                tc::tmem_ld_32x32(tbase + 32, v + 32);
                if (NSPLIT == 2) {
                    float u[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        tc::tmem_ld_32x32(tbase + C + h * 32, u);
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[h * 32 + c] = fmaf(u[c], 1.0f / LO_SCALE, v[h * 32 + c]);
                    }
                }
                if (tile == TILES - 1) {
                    // every accumulator column of this buffer is in registers: hand the buffer back to the MMA warp
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&bar_accempty[buf]);
                }
#pragma unroll
                for (int c = 0; c < 64; ++c) {
                    float t = fmaf(v[c], p.out_scale, s_bias[c]);
                    if (p.relu) t = fmaxf(t, 0.f);
                    v[c] = interior ? t : 0.f;
                }
                if (p.mask_hi && interior) {                    // ReLU data gradient: pass where the saved activation was positive
                    const uint4* mk = reinterpret_cast<const uint4*>(p.mask_hi + ((size_t)img_base + pp) * C);
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const uint4 m4 = __ldg(mk + g);
                        const uint32_t mw[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if ((mw[j] & 0x7FFFu) == 0u || (mw[j] & 0x8000u)) v[g * 8 + 2 * j] = 0.f;
                            if ((mw[j] & 0x7FFF0000u) == 0u || (mw[j] & 0x80000000u)) v[g * 8 + 2 * j + 1] = 0.f;
                        }
                    }
                }
                if (p.out_mode == 0) {
                    if (row_ok) {
                        const size_t o = ((size_t)img_base + pp) * C;
#pragma unroll
                        for (int g = 0; g < 4; ++g)          // 16 channels = one full 32-byte sector per plane and store
                            split_store16(v + g * 16, p.out_hi + o + g * 16, p.out_lo ? p.out_lo + o + g * 16 : nullptr, p.bf16 != 0);
                    }
                } else if (interior) {
                    const int y = yp - 2, x = xp - 2;
                    if (p.out_mode == 1) {
                        float* dst = p.out_f32 + (size_t)img * C * HW + (size_t)y * p.W + x;
#pragma unroll
                        for (int c = 0; c < 64; ++c) dst[(size_t)c * HW] = v[c];
                    } else {
                        const size_t pix = (size_t)y * p.W + x;
                        float* dst = p.out_f32 + ((size_t)img * HW + pix) * C;
                        const float4* ad = p.addend ? reinterpret_cast<const float4*>(p.addend + pix * C) : nullptr;
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            float4 o0 = make_float4(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3]);
                            float4 o1 = make_float4(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]);
                            if (ad) {
                                const float4 a = __ldg(ad + 2 * g), b = __ldg(ad + 2 * g + 1);
                                o0.x += a.x; o0.y += a.y; o0.z += a.z; o0.w += a.w;
                                o1.x += b.x; o1.y += b.y; o1.z += b.z; o1.w += b.w;
                            }
                            tc::st_global_v8(dst + 8 * g, __float_as_uint(o0.x), __float_as_uint(o0.y), __float_as_uint(o0.z), __float_as_uint(o0.w),
                                             __float_as_uint(o1.x), __float_as_uint(o1.y), __float_as_uint(o1.z), __float_as_uint(o1.w));
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tm, 2 * ACC_COLS);
}

template <int TILES, int NSPLIT>
int launch_conv_tc(const CUtensorMap& mA_hi, const CUtensorMap& mA_lo, const CUtensorMap& mA_hi8, const CUtensorMap& mA_lo8,
                   const CUtensorMap& mW, const ConvTcParams& p, cudaStream_t st) {
    const size_t smem = (size_t)NSPLIT * p.a_rows * ROWB + (size_t)p.stages * NSPLIT * WTAP_BYTES + 1024;
    PS_REQUIRE(smem <= 227 * 1024, "ps_conv5x5_tc: image width %d needs %zu bytes of shared memory", p.W, smem);
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(conv5x5_tc_kernel<TILES, NSPLIT>, smem, smem_cfg));
    const int n_items = p.n_img * p.items_per_img;
    conv5x5_tc_kernel<TILES, NSPLIT><<<n_items < p.n_ctas ? n_items : p.n_ctas, 192, smem, st>>>(mA_hi, mA_lo, mA_hi8, mA_lo8, mW, p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

extern "C" long long ps_conv_tc_plane_elems(int n_img, int H, int W) {
    return (long long)n_img * (H + 4) * (W + 4) * C;
}

static int conv5x5_tc_run(const void* act_in, const void* w_tc, const float* bias, float w_scale, void* act_out,
                          float* out_f32, const float* addend, const void* mask_act, int n_img, int H, int W, int relu,
                          int out_mode, int precision, void* stream) {
    PS_REQUIRE(act_in && w_tc, "ps_conv5x5_tc: null pointer");
    PS_REQUIRE(precision >= 1 && precision <= 3, "ps_conv5x5_tc: precision must be 1 (bf16), 2 (fp16) or 3 (fp16x3)");
    PS_REQUIRE(w_scale > 0.f, "ps_conv5x5_tc: w_scale must be positive");
    PS_REQUIRE(out_mode >= 0 && out_mode <= 2, "ps_conv5x5_tc: out_mode must be 0, 1 or 2");
    PS_REQUIRE(out_mode == 0 ? act_out != nullptr : out_f32 != nullptr, "ps_conv5x5_tc: missing output buffer");
    PS_REQUIRE(H >= 8 && W >= 8 && W <= 252, "ps_conv5x5_tc: unsupported image size %dx%d", H, W);
    if (n_img == 0) return PS_OK;
    const int Wp = W + 4, Hp = H + 4;
    const long long P = (long long)n_img * Hp * Wp;
    PS_REQUIRE(P < (1ll << 31) - 4096, "ps_conv5x5_tc: too many positions for 32-bit TMA coordinates");
    const int nsplit = precision == 3 ? 2 : 1;
    const int halo = 2 * Wp + 2;
    // two 128-position tiles per CTA when the staged region fits, else one
    int tiles = 2;
    auto a_rows_for = [&](int t) { return ((TILE_M * t + 2 * halo) + 63) / 64 * 64; };   // multiple of 64 rows = 8 KB: keeps every region 1024-byte aligned
    auto stages_for = [&](int t) {
        const long long room = 227ll * 1024 - 1024 - (long long)nsplit * a_rows_for(t) * ROWB;
        long long st = room / ((long long)nsplit * WTAP_BYTES);
        return (int)(st > MAX_STAGES ? MAX_STAGES : st);
    };
    if (stages_for(2) < 3) tiles = 1;
    PS_REQUIRE(stages_for(tiles) >= 2, "ps_conv5x5_tc: image width %d does not fit the shared-memory staging", W);
    ConvTcParams p;
    p.bias = bias; p.addend = addend; p.out_f32 = out_f32;
    p.mask_hi = reinterpret_cast<const uint16_t*>(mask_act);
    p.out_hi = reinterpret_cast<uint16_t*>(act_out);
    p.out_lo = (act_out && precision == 3) ? p.out_hi + P * C : nullptr;
    p.bf16 = precision == 1 ? 1 : 0;
    p.out_scale = 1.0f / w_scale;
    p.n_img = n_img; p.H = H; p.W = W; p.Wp = Wp; p.Hp = Hp;
    p.items_per_img = (H * Wp + TILE_M * tiles - 1) / (TILE_M * tiles);
    p.halo = halo; p.a_rows = a_rows_for(tiles); p.stages = stages_for(tiles);
    // refill chunks: rows below Wp * (c + 1) are dead after filter row c (rounded down to the 8-row swizzle atom)
    p.chunk_b[0] = 0;
    for (int c = 0; c < 4; ++c) p.chunk_b[c + 1] = Wp * (c + 1) / 8 * 8;
    p.chunk_b[5] = (TILE_M * tiles + 2 * halo + 7) / 8 * 8;
    p.n_early = (TILE_M * tiles + 4 <= p.chunk_b[4]) ? 4 : 5;
    p.relu = relu; p.out_mode = out_mode;
    {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        p.n_ctas = ps_persistent_ctas(sms);
    }
    const uint16_t* in_hi = reinterpret_cast<const uint16_t*>(act_in);
    CUtensorMap mA_hi, mA_lo, mA_hi8, mA_lo8, mW;
    const uint16_t* in_lo = precision == 3 ? in_hi + P * C : in_hi;
    int r = tc::make_tmap_2d_bf16_sw128(&mA_hi, in_hi, (uint64_t)P, C, ABOX, C);
    PS_REQUIRE(r == 0, "ps_conv5x5_tc: cuTensorMapEncodeTiled(A_hi) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mA_lo, in_lo, (uint64_t)P, C, ABOX, C);
    PS_REQUIRE(r == 0, "ps_conv5x5_tc: cuTensorMapEncodeTiled(A_lo) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mA_hi8, in_hi, (uint64_t)P, C, 8, C);
    PS_REQUIRE(r == 0, "ps_conv5x5_tc: cuTensorMapEncodeTiled(A_hi, 8 rows) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mA_lo8, in_lo, (uint64_t)P, C, 8, C);
    PS_REQUIRE(r == 0, "ps_conv5x5_tc: cuTensorMapEncodeTiled(A_lo, 8 rows) failed (%d)", r);
    // weights: [25 taps][2 (hi, lo)][64 cout][64 cin]; the single-product modes load the hi half of each tap only
    r = tc::make_tmap_2d_bf16_sw128(&mW, w_tc, 25ull * 2 * 64, C, 64 * nsplit, C);
    PS_REQUIRE(r == 0, "ps_conv5x5_tc: cuTensorMapEncodeTiled(W) failed (%d)", r);
    cudaStream_t st = (cudaStream_t)stream;
    if (tiles == 2) return nsplit == 2 ? launch_conv_tc<2, 2>(mA_hi, mA_lo, mA_hi8, mA_lo8, mW, p, st) : launch_conv_tc<2, 1>(mA_hi, mA_lo, mA_hi8, mA_lo8, mW, p, st);
    return nsplit == 2 ? launch_conv_tc<1, 2>(mA_hi, mA_lo, mA_hi8, mA_lo8, mW, p, st) : launch_conv_tc<1, 1>(mA_hi, mA_lo, mA_hi8, mA_lo8, mW, p, st);
}

extern "C" int ps_conv5x5_tc(const void* act_in, const void* w_tc, const float* bias, float w_scale, void* act_out,
                             float* out_f32, const float* addend, int n_img, int H, int W, int relu, int out_mode,
                             int precision, void* stream) {
    return conv5x5_tc_run(act_in, w_tc, bias, w_scale, act_out, out_f32, addend, nullptr, n_img, H, W, relu, out_mode, precision, stream);
}

// Data gradient of ConvBlock(64, 64, 5) = Conv2d + ReLU w.r.t. its input, given the gradient w.r.t. the conv output
// (already multiplied by relu' of THIS block by the caller chain) : conv with the flipped / transposed weights,
// then masked by relu' of the PREVIOUS block (mask_act = that block's saved output = this block's input).
extern "C" int ps_conv5x5_tc_dgrad(const void* g_in, const void* w_tc_flipped, float w_scale, const void* mask_act,
                                   void* act_out, float* out_tokens, int n_img, int H, int W, int out_mode, int precision,
                                   void* stream) {
    PS_REQUIRE(out_mode == 0 || out_mode == 2, "ps_conv5x5_tc_dgrad: out_mode must be 0 (padded split) or 2 (fp32 NHWC tokens)");
    PS_REQUIRE(precision == 2 || precision == 3, "ps_conv5x5_tc_dgrad: precision must be 2 (fp16) or 3 (fp16x3)");
    return conv5x5_tc_run(g_in, w_tc_flipped, nullptr, w_scale, act_out, out_tokens, nullptr, mask_act, n_img, H, W, 0, out_mode, precision, stream);
}

// ------------------------------------------------------------------------------------------------
// producers of the zero-bordered NHWC bf16 [hi | lo] activation format
// ------------------------------------------------------------------------------------------------
namespace {


__device__ __forceinline__ int border_class_tc(int v, int n, int pad) {
    if (v < pad) return v;
    if (v > n - 1 - pad) return 2 * pad - (n - 1 - v);
    return pad;
}

// h1[img, y, x, :] = relu(G[img, cls(y), cls(x), :] + P[y, x, :]) written as padded NHWC split;
// one CTA per (img, image row y) -- the row decomposition happens once per block, not per element: the kernel is
// instruction-bound (80 % issue utilisation) and integer divisions were half of its instructions -- one thread per
// (pixel x, group of 8 channels)
__global__ void __launch_bounds__(256) dec_expand_split_kernel(const float* __restrict__ G, const float* __restrict__ Phwc,
                                                                uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo,
                                                                int n_img, int H, int W, int pad, long long n_rows, int bf16) {
    const int ncls = 2 * pad + 1;
    const int Wp = W + 4, Hp = H + 4;
    for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const int img = (int)(row / H), y = (int)(row - (long long)img * H);
        const float* grow = G + ((size_t)img * ncls + border_class_tc(y, H, pad)) * ncls * C;
        const float* prow = Phwc + (size_t)y * W * C;
        const size_t orow = (((size_t)img * Hp + y + 2) * Wp + 2) * C;
        for (int i = threadIdx.x; i < W * 8; i += 256) {
            const int g = i & 7, x = i >> 3;
            const float* gp = grow + border_class_tc(x, W, pad) * C + g * 8;
            const float* pp = prow + x * C + g * 8;
            const float4 g0 = *reinterpret_cast<const float4*>(gp), g1 = *reinterpret_cast<const float4*>(gp + 4);
            const float4 q0 = *reinterpret_cast<const float4*>(pp), q1 = *reinterpret_cast<const float4*>(pp + 4);
            float v[8] = {fmaxf(g0.x + q0.x, 0.f), fmaxf(g0.y + q0.y, 0.f), fmaxf(g0.z + q0.z, 0.f), fmaxf(g0.w + q0.w, 0.f),
                          fmaxf(g1.x + q1.x, 0.f), fmaxf(g1.y + q1.y, 0.f), fmaxf(g1.z + q1.z, 0.f), fmaxf(g1.w + q1.w, 0.f)};
            const size_t o = orow + (size_t)x * C + g * 8;
            split_store8(v, out_hi + o, out_lo ? out_lo + o : nullptr, bf16 != 0);
        }
    }
}

// fp32 NCHW [n][64][H][W] -> padded NHWC split; one CTA per (img, row y), transpose through shared memory
__global__ void __launch_bounds__(256) nchw_to_split_kernel(const float* __restrict__ x, uint16_t* __restrict__ out_hi,
                                                            uint16_t* __restrict__ out_lo, int H, int W, int bf16) {
    extern __shared__ float s_t[];                    // [64][W+1]
    const int y = blockIdx.x, img = blockIdx.y;
    const int Wp = W + 4, Hp = H + 4;
    const float* src = x + (size_t)img * C * H * W + (size_t)y * W;
    for (int i = threadIdx.x; i < C * W; i += 256) {
        const int c = i / W, xx = i - c * W;
        s_t[c * (W + 1) + xx] = src[(size_t)c * H * W + xx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < W * 8; i += 256) {
        const int xx = i >> 3, g = i & 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = s_t[(g * 8 + j) * (W + 1) + xx];
        const size_t o = (((size_t)img * Hp + y + 2) * Wp + xx + 2) * C + g * 8;
        split_store8(v, out_hi + o, out_lo ? out_lo + o : nullptr, bf16 != 0);
    }
}

}  // namespace

extern "C" int ps_dec_expand_split(const float* G, const float* P_hwc, void* act_out, int n_img, int H, int W,
                                   int ksize, int precision, void* stream) {
    PS_REQUIRE(G && P_hwc && act_out, "ps_dec_expand_split: null pointer");
    PS_REQUIRE(precision >= 1 && precision <= 3, "ps_dec_expand_split: precision must be 1, 2 or 3");
    PS_REQUIRE(ksize % 2 == 1 && H > ksize && W > ksize, "ps_dec_expand_split: unsupported geometry");
    if (n_img == 0) return PS_OK;
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    uint16_t* hi = reinterpret_cast<uint16_t*>(act_out);
    uint16_t* lo = precision == 3 ? hi + P * C : nullptr;
    const long long n_rows = (long long)n_img * H;
    long long blocks = n_rows;
    if (blocks > 148 * 64) blocks = 148 * 64;
    dec_expand_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(G, P_hwc, hi, lo, n_img, H, W, ksize / 2, n_rows, precision == 1);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_nchw_to_split(const float* x, void* act_out, int n_img, int H, int W, int precision, void* stream) {
    PS_REQUIRE(x && act_out, "ps_nchw_to_split: null pointer");
    PS_REQUIRE(precision >= 1 && precision <= 3, "ps_nchw_to_split: precision must be 1, 2 or 3");
    PS_REQUIRE(W <= 160 && H <= 65535, "ps_nchw_to_split: unsupported image size");
    if (n_img == 0) return PS_OK;
    PS_REQUIRE(n_img <= 65535, "ps_nchw_to_split: at most 65535 images per launch");
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    uint16_t* hi = reinterpret_cast<uint16_t*>(act_out);
    uint16_t* lo = precision == 3 ? hi + P * C : nullptr;
    const size_t smem = (size_t)C * (W + 1) * sizeof(float);
    nchw_to_split_kernel<<<dim3(H, n_img), 256, smem, (cudaStream_t)stream>>>(x, hi, lo, H, W, precision == 1);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ================================================================================================
// Final decoder stage on the tensor cores: Conv2d(64, 4, 3, padding=1) of every slot image + softmax of the alpha channel
// over the slot axis + compositing (src/models/SAVi.py:226-234, EncodersDecoders/decoders.py:103-111).
// Padded-linear implicit GEMM like above, but an MMA with 4 output channels reads a whole 4 KB activation operand for
// 128 x 8 x 16 products: nine such taps made the kernel shared-memory-operand bound (72 MMAs = 288 KB of operand reads per slot
// and tile, 2.4 ms per step).  The three taps of a filter ROW therefore share one MMA: their weights sit side by side in the B
// operand (N = 32: [hi: kx0 o0-3 | kx1 | kx2 | pad ; lo: same]) and the A window is shifted by whole image rows only,
//     Z_kx[q] = sum_ky W[ky][kx] . A[q + (ky - 1) Wp]                      (24 MMAs per slot and tile)
//     out[p]  = Z_0[p - 1] + Z_1[p] + Z_2[p + 1]                           (epilogue: neighbouring TMEM lanes = neighbouring positions)
// The horizontal shift happens between the epilogue threads: warp shuffles inside a warp, a 4 KB shared-memory exchange across
// the three warp boundaries; rows 0 and 127 of a tile have no complete neighbourhood, so tiles advance by 126 positions.
// One work item = one tile of one frame: the S slot images stream through a double-buffered activation region, their
// accumulators (32 columns per slot) sit side by side in TMEM, and the epilogue thread of a position sees all S slots: bias,
// slot softmax, recons / masks / recon written once.
// ================================================================================================
namespace {

struct CompTcParams {
    const float* bias;       // [4]
    float* recon;            // [M][3][H][W]
    float* recons;           // [M][S][3][H][W]
    float* masks;            // [M][S][1][H][W]
    int M, S, H, W, Wp, Hp;
    int tiles_per_img, halo, a_rows, nbuf, n_ctas;
    float out_scale;
};

constexpr int CT_WBYTES = 3 * 32 * ROWB;     // 3 filter rows x (16 hi + 16 lo rows) x 64 cin fp16
constexpr int CT_STRIDE = TILE_M - 2;        // useful positions per tile

template <int SPAD>      // slots padded to 8 or 16 (TMEM columns = 32 per slot)
__global__ void __launch_bounds__(192, 1)
dec_composite_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                        const __grid_constant__ CUtensorMap mapW, const CompTcParams p) {
    constexpr uint32_t ACC_COLS = 32 * SPAD;
    constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;    // 16 slots fill the 512 TMEM columns once: no accumulator double buffer
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = p.a_rows * ROWB;
    uint8_t* sW = smem;                                    // 12 KB
    uint8_t* sA0 = smem + ((CT_WBYTES + 1023) & ~1023);    // nbuf x [hi | lo] regions
    __shared__ uint64_t bar_w, bar_afull[2], bar_aempty[2], bar_accfull[2], bar_accempty[2];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(16) float s_edge[2][4][2][SPAD][4];             // [item parity][lane quarter][0: Z_0 of lane 31 | 1: Z_2 of lane 0][slot][o]
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // provably warp-uniform (see conv5x5_tc_kernel)
    const int n_items = p.M * p.tiles_per_img;
    const int S = p.S, NBUF = p.nbuf;

    if (tid == 0) {
        tc::mbar_init(&bar_w, 1);
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&bar_afull[b], 1); tc::mbar_init(&bar_aempty[b], 1);
            tc::mbar_init(&bar_accfull[b], 1); tc::mbar_init(&bar_accempty[b], 4);
        }
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base, NACC * ACC_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;

    if (warp == 0) {
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&mapA_hi);
            tc::mbar_expect_tx(&bar_w, CT_WBYTES);
            tc::tma_load_2d(sW, &mapW, &bar_w, 0, 0);
            int q = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int m = it / p.tiles_per_img, tile = it - m * p.tiles_per_img;
                const int p0 = 2 * p.Wp + tile * CT_STRIDE - 1;               // position of the tile's row 0
                for (int s = 0; s < S; ++s, ++q) {
                    const int buf = q % NBUF;
                    if (q >= NBUF) tc::mbar_wait(&bar_aempty[buf], ((q / NBUF) - 1) & 1);
                    const int g0 = (int)((long long)(m * S + s) * p.Hp * p.Wp + p0 - p.halo);
                    uint8_t* a = sA0 + buf * 2 * a_bytes;
                    tc::mbar_expect_tx(&bar_afull[buf], (uint32_t)(2 * a_bytes));
                    for (int r = 0; r < p.a_rows; r += ABOX) {
                        tc::tma_load_2d(a + r * ROWB, &mapA_hi, &bar_afull[buf], 0, g0 + r);
                        tc::tma_load_2d(a + a_bytes + r * ROWB, &mapA_lo, &bar_afull[buf], 0, g0 + r);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // the whole warp walks the loop (uniform control flow, uniform-register descriptor arithmetic, no division by the
        // run-time buffer count); one elected lane issues -- same reasoning as in conv5x5_tc_kernel
        const bool leader = tc::elect_one();
        {
            const uint32_t idesc32 = tc::umma_idesc_f16(TILE_M, 32), idesc16 = tc::umma_idesc_f16(TILE_M, 16);
            const uint64_t dW = tc::umma_desc_sw128(tc::smem_u32(sW), 1024, 0);
            const uint64_t dA = tc::umma_desc_sw128(tc::smem_u32(sA0) + (uint32_t)p.halo * ROWB, 1024, 0);
            const uint64_t buf_units = (uint64_t)((2 * a_bytes) / 16), lo_units = (uint64_t)(a_bytes / 16);
            const int row_step = p.Wp * (ROWB / 16);
            tc::mbar_wait(&bar_w, 0);
            int buf = 0;
            uint32_t bph = 0;
            int i = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
                const int ab = NACC == 2 ? (i & 1) : 0;
                if (i >= NACC) tc::mbar_wait(&bar_accempty[ab], (NACC == 2 ? ((i >> 1) - 1) : (i - 1)) & 1);
                for (int s = 0; s < S; ++s) {
                    tc::mbar_wait(&bar_afull[buf], bph);
                    tc::tc_fence_after();
                    const uint64_t a_hi0 = dA + (uint64_t)buf * buf_units;
                    const uint32_t acc = tm + ab * ACC_COLS + s * 32;
                    int a_off = -row_step;                                       // filter row 0: one image row up, 16-byte units
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint64_t a_hi = a_hi0 + (uint64_t)(int64_t)a_off;
#pragma unroll
                        for (int k = 0; k < C / 16; ++k) {
                            const uint64_t bd = dW + (uint64_t)((ky * 32 * ROWB + k * 32) / 16);
                            if (leader) tc::umma_bf16(acc, a_hi + (uint64_t)(k * 2), bd, idesc32, (ky | k) ? 1u : 0u);      // D0 | D1 += A_hi [w_hi | w_lo]
                            if (leader) tc::umma_bf16(acc + 16, a_hi + lo_units + (uint64_t)(k * 2), bd, idesc16, 1u);       // D1 += A_lo w_hi
                        }
                        a_off += row_step;
                    }
                    if (leader) tc::umma_commit(&bar_aempty[buf]);
                    if (++buf == NBUF) { buf = 0; bph ^= 1u; }
                }
                if (leader) tc::umma_commit(&bar_accfull[ab]);
            }
        }
    } else {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;                    // row of the tile == TMEM lane
        const int HW = p.H * p.W;
        const float b0 = p.bias[0], b1 = p.bias[1], b2 = p.bias[2], b3 = p.bias[3];
        int i = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++i) {
            const int ab = NACC == 2 ? (i & 1) : 0;
            const int par = i & 1;
            const int m = it / p.tiles_per_img, tile = it - m * p.tiles_per_img;
            const int pp = 2 * p.Wp + tile * CT_STRIDE - 1 + r;
            const int yp = pp / p.Wp, xp = pp - yp * p.Wp;
            const bool interior = r >= 1 && r <= CT_STRIDE && yp < p.H + 2 && xp >= 2 && xp < p.W + 2;
            tc::mbar_wait(&bar_accfull[ab], (NACC == 2 ? (i >> 1) : i) & 1);
            __syncwarp();
            tc::tc_fence_after();
            float o[SPAD][4];
            const uint32_t tb = tm + ((uint32_t)(quad * 32) << 16) + ab * ACC_COLS;
#pragma unroll
            for (int s = 0; s < SPAD; ++s) {
                if (s < S) {
                    float v[32];
                    tc::tmem_ld_32x32(tb + s * 32, v);
                    float z[12];                               // z[kx * 4 + o] = Z_kx[this position][o]
#pragma unroll
                    for (int t = 0; t < 12; ++t) z[t] = fmaf(v[16 + t], 1.0f / LO_SCALE, v[t]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float up = __shfl_up_sync(0xffffffffu, z[c], 1);          // Z_0 of the position to the left
                        const float dn = __shfl_down_sync(0xffffffffu, z[8 + c], 1);    // Z_2 of the position to the right
                        o[s][c] = z[4 + c] + (lane > 0 ? up : 0.f) + (lane < 31 ? dn : 0.f);
                    }
                    if (lane == 31) *reinterpret_cast<float4*>(s_edge[par][quad][0][s]) = make_float4(z[0], z[1], z[2], z[3]);
                    if (lane == 0) *reinterpret_cast<float4*>(s_edge[par][quad][1][s]) = make_float4(z[8], z[9], z[10], z[11]);
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bar_accempty[ab]);
            asm volatile("bar.sync 1, 128;\n" ::: "memory");       // the four epilogue warps: edge values are in shared memory
            if (lane == 0 && quad > 0) {
#pragma unroll
                for (int s = 0; s < SPAD; ++s)
                    if (s < S) {
                        const float4 e = *reinterpret_cast<const float4*>(s_edge[par][quad - 1][0][s]);
                        o[s][0] += e.x; o[s][1] += e.y; o[s][2] += e.z; o[s][3] += e.w;
                    }
            }
            if (lane == 31 && quad < 3) {
#pragma unroll
                for (int s = 0; s < SPAD; ++s)
                    if (s < S) {
                        const float4 e = *reinterpret_cast<const float4*>(s_edge[par][quad + 1][1][s]);
                        o[s][0] += e.x; o[s][1] += e.y; o[s][2] += e.z; o[s][3] += e.w;
                    }
            }
            if (!interior) continue;
#pragma unroll
            for (int s = 0; s < SPAD; ++s)
                if (s < S) {
                    o[s][0] = fmaf(o[s][0], p.out_scale, b0); o[s][1] = fmaf(o[s][1], p.out_scale, b1);
                    o[s][2] = fmaf(o[s][2], p.out_scale, b2); o[s][3] = fmaf(o[s][3], p.out_scale, b3);
                }
            float mx = -INFINITY;
#pragma unroll
            for (int s = 0; s < SPAD; ++s)
                if (s < S) mx = fmaxf(mx, o[s][3]);
            float den = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int s = 0; s < SPAD; ++s)
                if (s < S) { const float e = expf(o[s][3] - mx); o[s][3] = e; den += e; }
            const float inv = 1.0f / den;
            const size_t pix = (size_t)(yp - 2) * p.W + (xp - 2);
#pragma unroll
            for (int s = 0; s < SPAD; ++s) {
                if (s < S) {
                    const float mk = o[s][3] * inv;
                    const size_t ms = (size_t)m * S + s;
                    p.recons[(ms * 3 + 0) * HW + pix] = o[s][0];
                    p.recons[(ms * 3 + 1) * HW + pix] = o[s][1];
                    p.recons[(ms * 3 + 2) * HW + pix] = o[s][2];
                    p.masks[ms * HW + pix] = mk;
                    r0 = fmaf(o[s][0], mk, r0); r1 = fmaf(o[s][1], mk, r1); r2 = fmaf(o[s][2], mk, r2);
                }
            }
            p.recon[((size_t)m * 3 + 0) * HW + pix] = r0;
            p.recon[((size_t)m * 3 + 1) * HW + pix] = r1;
            p.recon[((size_t)m * 3 + 2) * HW + pix] = r2;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tm, NACC * ACC_COLS);
}

template <int SPAD>
int launch_comp_tc(const CUtensorMap& mA_hi, const CUtensorMap& mA_lo, const CUtensorMap& mW, const CompTcParams& p, cudaStream_t st) {
    const size_t smem = ((CT_WBYTES + 1023) & ~1023) + (size_t)p.nbuf * 2 * p.a_rows * ROWB + 1024;
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(dec_composite_tc_kernel<SPAD>, smem, smem_cfg));
    const int n_items = p.M * p.tiles_per_img;
    dec_composite_tc_kernel<SPAD><<<n_items < p.n_ctas ? n_items : p.n_ctas, 192, smem, st>>>(mA_hi, mA_lo, mW, p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

}  // namespace

extern "C" int ps_dec_composite_tc(const void* act_in, const void* w_tc, float w_scale, const float* bias, float* recon,
                                   float* recons, float* masks, int M, int S, int H, int W, void* stream) {
    PS_REQUIRE(act_in && w_tc && bias && recon && recons && masks, "ps_dec_composite_tc: null pointer");
    PS_REQUIRE(S >= 1 && S <= 16, "ps_dec_composite_tc: 1 <= num_slots <= 16 (got %d)", S);
    PS_REQUIRE(H >= 8 && W >= 8 && W <= 252, "ps_dec_composite_tc: unsupported image size %dx%d", H, W);
    PS_REQUIRE(w_scale > 0.f, "ps_dec_composite_tc: w_scale must be positive");
    if (M == 0) return PS_OK;
    const int Wp = W + 4, Hp = H + 4;
    const long long P = (long long)M * S * Hp * Wp;
    PS_REQUIRE(P < (1ll << 31) - 4096, "ps_dec_composite_tc: too many positions for 32-bit TMA coordinates");
    CompTcParams p;
    p.bias = bias; p.recon = recon; p.recons = recons; p.masks = masks;
    p.M = M; p.S = S; p.H = H; p.W = W; p.Wp = Wp; p.Hp = Hp;
    p.tiles_per_img = (H * Wp + CT_STRIDE - 1) / CT_STRIDE;
    p.halo = Wp;                                   // whole image rows only: the +-1 column shift happens in the epilogue
    p.a_rows = ((TILE_M + 2 * p.halo) + ABOX - 1) / ABOX * ABOX;
    const long long room = 227ll * 1024 - 1024 - ((CT_WBYTES + 1023) & ~1023);
    p.nbuf = (2ll * 2 * p.a_rows * ROWB <= room) ? 2 : 1;
    PS_REQUIRE(2ll * p.a_rows * ROWB <= room, "ps_dec_composite_tc: image width %d does not fit the shared-memory staging", W);
    p.out_scale = 1.0f / w_scale;
    {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        p.n_ctas = ps_persistent_ctas(sms);
    }
    const uint16_t* in_hi = reinterpret_cast<const uint16_t*>(act_in);
    CUtensorMap mA_hi, mA_lo, mW;
    int r = tc::make_tmap_2d_bf16_sw128(&mA_hi, in_hi, (uint64_t)P, C, ABOX, C);
    PS_REQUIRE(r == 0, "ps_dec_composite_tc: cuTensorMapEncodeTiled(A_hi) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mA_lo, in_hi + P * C, (uint64_t)P, C, ABOX, C);
    PS_REQUIRE(r == 0, "ps_dec_composite_tc: cuTensorMapEncodeTiled(A_lo) failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mW, w_tc, 3ull * 32, C, 3 * 32, C);
    PS_REQUIRE(r == 0, "ps_dec_composite_tc: cuTensorMapEncodeTiled(W) failed (%d)", r);
    cudaStream_t st = (cudaStream_t)stream;
    return S <= 8 ? launch_comp_tc<8>(mA_hi, mA_lo, mW, p, st) : launch_comp_tc<16>(mA_hi, mA_lo, mW, p, st);
}
