// Development tool: HBM read bandwidth of a TMA-only streaming kernel (no compute) for the two candidate layouts of the
// SlotAttention feature planes: (A) [tokens][128 feats] 16-bit, a 128-token tile = 2 boxes of 128 rows x 128 B at a 256 B row
// pitch per plane; (B) the two 64-feature halves stored as separate sub-planes, every box 16 KB contiguous.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tma_stream tools/tma_stream.cu -lcuda
#include <cstdio>
#include <cstdint>
#include "../playslot_b200/csrc/tc_common.cuh"

constexpr int STAGES = 3, STAGE_BYTES = 65536, TILE = 128;

__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int layout, long long n_tok, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[STAGES], empty[STAGES];
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); } tc::fence_mbar_init(); }
    __syncthreads();
    const int t0 = (int)((long long)blockIdx.x * n_tiles / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * n_tiles / gridDim.x);
    if (warp == 0 && tc::elect_one()) {
        int g = 0;
        for (int t = t0; t < t1; ++t, ++g) {
            const int s = g % STAGES;
            if (g >= STAGES) tc::mbar_wait(&empty[s], ((g / STAGES) - 1) & 1);
            uint8_t* st = smem + s * STAGE_BYTES;
            const int row = t * TILE;
            tc::mbar_expect_tx(&full[s], STAGE_BYTES);
            if (layout == 0) {              // planes [hi | lo][n_tok][128]: map rows = 2 * n_tok, cols = 128
                tc::tma_load_2d(st, &map, &full[s], 0, row);
                tc::tma_load_2d(st + 16384, &map, &full[s], 64, row);
                tc::tma_load_2d(st + 32768, &map, &full[s], 0, (int)n_tok + row);
                tc::tma_load_2d(st + 49152, &map, &full[s], 64, (int)n_tok + row);
            } else {                        // sub-planes [hi0 | hi1 | lo0 | lo1][n_tok][64]: map rows = 4 * n_tok, cols = 64
                for (int q = 0; q < 4; ++q) tc::tma_load_2d(st + q * 16384, &map, &full[s], 0, (int)(q * n_tok) + row);
            }
        }
    } else if (warp == 1 && tc::elect_one()) {
        int g = 0;
        for (int t = t0; t < t1; ++t, ++g) {
            const int s = g % STAGES;
            tc::mbar_wait(&full[s], (g / STAGES) & 1);
            tc::mbar_arrive(&empty[s]);
        }
    }
}

int main() {
    const long long n_tok = 64ll * 4096;
    void* buf; cudaMalloc(&buf, n_tok * 512); cudaMemset(buf, 1, n_tok * 512);
    const int smem = STAGES * STAGE_BYTES + 1024;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int layout = 0; layout < 2; ++layout) {
        CUtensorMap map;
        int r = layout == 0 ? tc::make_tmap_2d_bf16_sw128(&map, buf, 2 * n_tok, 128, TILE, 64)
                            : tc::make_tmap_2d_bf16_sw128(&map, buf, 4 * n_tok, 64, TILE, 64);
        if (r) { printf("tensor map failed %d\n", r); return 1; }
        for (int grid : {128, 148}) {
            const int n_tiles = (int)(n_tok / TILE);
            for (int i = 0; i < 3; ++i) stream_kernel<<<grid, 64, smem>>>(map, layout, n_tok, n_tiles);
            cudaEventRecord(e0);
            for (int i = 0; i < 20; ++i) stream_kernel<<<grid, 64, smem>>>(map, layout, n_tok, n_tiles);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaError_t e = cudaGetLastError();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("layout %s, %3d CTAs: %.1f us per pass, %.0f GB/s (%s)\n", layout ? "B (contiguous 16 KB boxes)" : "A (128 B of 256 B rows)    ",
                   grid, ms / 20 * 1e3, n_tok * 512.0 / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(e));
        }
    }
    return 0;
}
