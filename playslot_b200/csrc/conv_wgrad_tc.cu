// Weight gradient of the 64 -> 64 channel 5x5 convolutions on tcgen05 (SAVi training, src/02_train_savi.py:21-70:
// what loss.backward() computes for the ConvBlock weights of encoders.py:46-104 / decoders.py:45-113):
//     dW[tap][co][ci] = scale * sum_p dY[p][co] * X[p + off(tap)][ci],      off(ky,kx) = (ky-2) * Wp + (kx-2)
// over the positions p of the zero-bordered ("padded-linear") NHWC planes the forward / data-gradient kernels read and write
// (conv_tc.cu): the tap shift is a constant row offset, borders are zero, so the sum may run over ALL padded positions.
//
// GEMM view: M = co, N = ci, K = positions.  Both operands are stored position-major (rows of 64 channels = 128 bytes), i.e. they
// are MN-MAJOR operands of the MMA (the contraction index runs over rows): A = [dY_hi | dY_lo] (M = 128, two 64-element blocks),
// B = [X_hi | X_lo] (N = 128) read from a window of the X planes at the tap's row shift.  One M=128 x N=128 x K=16 MMA per tap and
// 16 positions yields  rows 0-63:  [dY_hi X_hi | dY_hi X_lo]   rows 64-127: [dY_lo X_hi | (dY_lo X_lo, dropped)]  and
//     dW = scale * (D[co][ci] + (D[co][64+ci] + D[64+co][ci]) / 2048)                     (the fp16x3 scheme of conv_tc.cu).
// TMEM holds 4 taps x 128 fp32 columns, so the 25 taps are dealt to 7 groups -- {(ky, kx 0..3)} for ky = 0..4, {(0..2, 4)},
// {(3..4, 4)} -- and a CTA owns (tap group, chunk of positions): it streams 128-position blocks of both tensors through a
// TMA / mbarrier ring, accumulates in TMEM over its whole chunk and adds its result to dW once (atomics).  CTAs of the 7 groups
// of one chunk are adjacent in the grid, so they read the same blocks at about the same time (L2 hits).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_fp16.h>

namespace {

constexpr int WG_KB = 128;                       // positions per block (8 MMA k-steps)
constexpr int WG_ROWB = 128;                     // bytes per position row (64 channels x 2 bytes)
constexpr int WG_DY_TILE = WG_KB * WG_ROWB;      // 16 KB per plane
constexpr int WG_BOX = 136;                      // rows per X TMA box (17 KB, a multiple of 1024 bytes)
constexpr int WG_X_BOXB = WG_BOX * WG_ROWB;
constexpr int WG_MAX_STAGES = 3;
constexpr int WG_THREADS = 192;                  // warp 0: TMA, warp 1: MMA + TMEM, warps 2-5: epilogue

struct WgParams {
    float* dW;                                   // [25][64][64]
    long long P;                                 // padded positions of all images
    long long pos_per_chunk;                     // multiple of WG_KB
    int Wp;
    int lo;                                      // 1: hi/lo planes (fp16x3), 0: hi planes only
    float scale;
};

// tap group g -> (number of taps, (ky, kx) of each tap, X boxes per plane)
__device__ __forceinline__ void wg_group(int g, int& ntaps, int* ky, int* kx, int& nbox) {
    if (g < 5) { ntaps = 4; nbox = 1; for (int t = 0; t < 4; ++t) { ky[t] = g; kx[t] = t; } }
    else if (g == 5) { ntaps = 3; nbox = 2; for (int t = 0; t < 3; ++t) { ky[t] = t; kx[t] = 4; } }
    else { ntaps = 2; nbox = 2; for (int t = 0; t < 2; ++t) { ky[t] = 3 + t; kx[t] = 4; } }
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv5x5_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mDyHi, const __grid_constant__ CUtensorMap mDyLo,
                        const __grid_constant__ CUtensorMap mXHi, const __grid_constant__ CUtensorMap mXLo, const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[WG_MAX_STAGES], bar_empty[WG_MAX_STAGES], bar_done;
    __shared__ uint32_t tmem_base;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = blockIdx.x % 7, chunk = blockIdx.x / 7;
    int ntaps, ky[4], kx[4], nbox;
    wg_group(g, ntaps, ky, kx, nbox);
    const int nplanes = p.lo ? 2 : 1;
    const int x_plane_bytes = nbox * WG_X_BOXB;
    const int stage_bytes = nplanes * (WG_DY_TILE + x_plane_bytes);
    const int stages = min(WG_MAX_STAGES, (220 * 1024) / stage_bytes);
    const long long off0 = (long long)(ky[0] - 2) * p.Wp + (kx[0] - 2);          // shift of the group's first tap = window start
    const long long p_begin = (long long)chunk * p.pos_per_chunk;
    const long long p_end = min(p.P, p_begin + p.pos_per_chunk);
    const int nblocks = p_begin < p_end ? (int)((p_end - p_begin + WG_KB - 1) / WG_KB) : 0;

    if (tid == 0) {
        for (int s = 0; s < WG_MAX_STAGES; ++s) { tc::mbar_init(&bar_full[s], 1); tc::mbar_init(&bar_empty[s], 1); }
        tc::mbar_init(&bar_done, 1);
        tc::fence_mbar_init();
    }
    if (warp == 1) tc::tmem_alloc(&tmem_base, 512);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (tc::elect_one()) {
            tc::tma_prefetch_desc(&mDyHi); tc::tma_prefetch_desc(&mXHi);
            for (int b = 0; b < nblocks; ++b) {
                const int s = b % stages, r = b / stages;
                if (r > 0) tc::mbar_wait(&bar_empty[s], (r - 1) & 1);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                const long long p0 = p_begin + (long long)b * WG_KB;
                tc::mbar_expect_tx(&bar_full[s], (uint32_t)stage_bytes);
                // positions beyond the tensors (negative window rows, the tail of the last block) are zero-filled by TMA; the
                // tail of a chunk's last block belongs to the next chunk's positions: mask it with a row count instead --
                // chunks are multiples of WG_KB, so only the global tail (p >= P) exists and that is zero-filled
                tc::tma_load_2d(st, &mDyHi, &bar_full[s], 0, (int)p0);
                if (p.lo) tc::tma_load_2d(st + WG_DY_TILE, &mDyLo, &bar_full[s], 0, (int)p0);
                uint8_t* xs = st + nplanes * WG_DY_TILE;
                for (int bx = 0; bx < nbox; ++bx) {
                    tc::tma_load_2d(xs + bx * WG_X_BOXB, &mXHi, &bar_full[s], 0, (int)(p0 + off0) + bx * WG_BOX);
                    if (p.lo) tc::tma_load_2d(xs + x_plane_bytes + bx * WG_X_BOXB, &mXLo, &bar_full[s], 0, (int)(p0 + off0) + bx * WG_BOX);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (tc::elect_one()) {
            // A: M = 128 = [hi 64 | lo 64] (LBO = plane distance), K = positions (8-row groups 1024 bytes apart), MN-major, 128B swizzle
            // B: N = 128 = [X_hi 64 | X_lo 64] likewise; in the single-plane modes M = N = 64
            const int MN = p.lo ? 128 : 64;
            const uint32_t idesc = tc::umma_idesc_f16_major(MN == 128 ? 128 : 64, MN, 1, 1);
            int row_shift[4];
            for (int t = 0; t < ntaps; ++t) row_shift[t] = (int)(((long long)(ky[t] - 2) * p.Wp + (kx[t] - 2)) - off0);
            for (int b = 0; b < nblocks; ++b) {
                const int s = b % stages, r = b / stages;
                tc::mbar_wait(&bar_full[s], r & 1);
                tc::tc_fence_after();
                const uint32_t st = tc::smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t xs = st + nplanes * WG_DY_TILE;
                for (int t = 0; t < ntaps; ++t) {
                    const uint32_t bt = xs + (uint32_t)row_shift[t] * WG_ROWB;
#pragma unroll
                    for (int k = 0; k < WG_KB / 16; ++k) {
                        const uint64_t ad = tc::umma_desc(st + k * 16 * WG_ROWB, WG_DY_TILE, 1024, 2);
                        const uint64_t bd = tc::umma_desc(bt + k * 16 * WG_ROWB, (uint32_t)x_plane_bytes, 1024, 2);
                        tc::umma_bf16(tm + t * 128, ad, bd, idesc, (b | k) ? 1u : 0u);
                    }
                }
                tc::umma_commit(&bar_empty[s]);
            }
            tc::umma_commit(&bar_done);
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5 = TMEM lane quarters 2,3,0,1)
        if (nblocks > 0) {
            tc::mbar_wait(&bar_done, 0);
            tc::tc_fence_after();
            const int q = warp & 3;                              // a warp may only read the 32 TMEM lanes of its quarter (warp % 4)
            const int row = q * 32 + lane;                       // 0..127
            const uint32_t tl = tm + ((uint32_t)(q * 32) << 16);
            const int co = row & 63;
            const bool lo_row = row >= 64;
            if (!(lo_row && !p.lo)) {
                for (int t = 0; t < ntaps; ++t) {
                    float* dst = p.dW + ((size_t)(ky[t] * 5 + kx[t]) * 64 + co) * 64;
#pragma unroll 1
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        float v[32];
                        tc::tmem_ld_32x32(tl + t * 128 + c0, v);
                        if (!lo_row && p.lo) {
                            float u[32];
                            tc::tmem_ld_32x32(tl + t * 128 + 64 + c0, u);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fmaf(u[j], 1.0f / 2048.0f, v[j]);
                        }
                        const float sc = lo_row ? p.scale * (1.0f / 2048.0f) : p.scale;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + c0 + j, v[j] * sc);
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tm, 512);
}

}  // namespace

// dW (25, 64, 64) [tap][cout][cin] += scale * correlation of the gradient planes dy with the activation planes x, on the tensor
// cores; same contract as ps_conv5x5_planes_wgrad (both tensors in the same precision: 3 = hi + lo planes, 2 = fp16 hi only).
extern "C" int ps_conv5x5_planes_wgrad_tc(const void* dy_planes, const void* x_planes, float* dW, int n_img, int H, int W, int precision,
                                          float scale, void* stream) {
    PS_REQUIRE(dy_planes && x_planes && dW, "ps_conv5x5_planes_wgrad_tc: null pointer");
    PS_REQUIRE(precision == 3, "ps_conv5x5_planes_wgrad_tc: precision must be 3 (fp16x3 hi / lo planes)");
    if (n_img == 0) return PS_OK;
    const int Wp = W + 4, Hp = H + 4;
    const long long P = (long long)n_img * Hp * Wp;
    PS_REQUIRE(P < (1ll << 31) - 4096, "ps_conv5x5_planes_wgrad_tc: too many positions for 32-bit TMA coordinates");
    PS_REQUIRE(2 * Wp + WG_KB <= 2 * WG_BOX, "ps_conv5x5_planes_wgrad_tc: image width %d exceeds the staged window", W);
    const uint16_t* dy_hi = reinterpret_cast<const uint16_t*>(dy_planes);
    const uint16_t* x_hi = reinterpret_cast<const uint16_t*>(x_planes);
    const uint16_t* dy_lo = precision == 3 ? dy_hi + P * 64 : dy_hi;
    const uint16_t* x_lo = precision == 3 ? x_hi + P * 64 : x_hi;
    CUtensorMap mDyHi, mDyLo, mXHi, mXLo;
    int r = tc::make_tmap_2d_bf16_sw128(&mDyHi, dy_hi, (uint64_t)P, 64, WG_KB, 64);
    PS_REQUIRE(r == 0, "ps_conv5x5_planes_wgrad_tc: cuTensorMapEncodeTiled failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mDyLo, dy_lo, (uint64_t)P, 64, WG_KB, 64);
    PS_REQUIRE(r == 0, "ps_conv5x5_planes_wgrad_tc: cuTensorMapEncodeTiled failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mXHi, x_hi, (uint64_t)P, 64, WG_BOX, 64);
    PS_REQUIRE(r == 0, "ps_conv5x5_planes_wgrad_tc: cuTensorMapEncodeTiled failed (%d)", r);
    r = tc::make_tmap_2d_bf16_sw128(&mXLo, x_lo, (uint64_t)P, 64, WG_BOX, 64);
    PS_REQUIRE(r == 0, "ps_conv5x5_planes_wgrad_tc: cuTensorMapEncodeTiled failed (%d)", r);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    WgParams p;
    p.dW = dW; p.P = P; p.Wp = Wp; p.lo = precision == 3 ? 1 : 0; p.scale = scale;
    long long chunks = sms / 7;
    const long long blocks = (P + WG_KB - 1) / WG_KB;
    if (chunks > blocks) chunks = blocks;
    if (chunks < 1) chunks = 1;
    p.pos_per_chunk = (blocks + chunks - 1) / chunks * WG_KB;
    chunks = (P + p.pos_per_chunk - 1) / p.pos_per_chunk;
    const size_t smem = 222 * 1024;
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(conv5x5_wgrad_tc_kernel, smem, smem_cfg));
    conv5x5_wgrad_tc_kernel<<<(unsigned)(chunks * 7), WG_THREADS, smem, (cudaStream_t)stream>>>(mDyHi, mDyLo, mXHi, mXLo, p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
