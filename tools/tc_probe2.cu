// Generic tcgen05 descriptor probe (development tool, not part of the library).
// The host supplies raw shared-memory images of the A and B operand regions, a list of (A, B) shared-memory
// descriptors (address field = byte offset inside the region >> 4; the kernel adds the region base), one
// instruction descriptor and the number of accumulator columns; the kernel issues the MMAs into one TMEM
// accumulator and returns its 128 lanes x ncols.  tools/tc_probe2.py drives it with layout hypotheses.
#include <cstdio>
#include <cstdint>
#include "../playslot_b200/csrc/tc_common.cuh"

constexpr int REGION = 64 * 1024;

__global__ void __launch_bounds__(128) probe2_kernel(const uint4* imgA, int bytesA, const uint4* imgB, int bytesB,
                                                     const uint64_t* adesc, const uint64_t* bdesc, int n_mma,
                                                     uint32_t idesc, float* out, int ncols) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + REGION;
    __shared__ uint64_t bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < bytesA / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = imgA[i];
    for (int i = tid; i < bytesB / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = imgB[i];
    if (tid == 0) { tc::mbar_init(&bar_mma, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
    tc::fence_async_proxy();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    if (warp == 0 && tc::elect_one()) {
        const uint64_t abase = (uint64_t)((tc::smem_u32(sA) & 0x3FFFF) >> 4);
        const uint64_t bbase = (uint64_t)((tc::smem_u32(sB) & 0x3FFFF) >> 4);
        for (int i = 0; i < n_mma; ++i)
            tc::umma_bf16(tm, adesc[i] + abase, bdesc[i] + bbase, idesc, i > 0);
        tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    float v[32];
    for (int c = 0; c < ncols; c += 32) {
        tc::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int i = 0; i < 32 && c + i < ncols; ++i) out[(size_t)tid * ncols + c + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 256);
}

extern "C" int probe_mma(const void* imgA, int bytesA, const void* imgB, int bytesB, const void* adesc,
                         const void* bdesc, int n_mma, unsigned idesc, float* out, int ncols) {
    const int smem = 2 * REGION + 1024;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    probe2_kernel<<<1, 128, smem>>>((const uint4*)imgA, bytesA, (const uint4*)imgB, bytesB, (const uint64_t*)adesc,
                                    (const uint64_t*)bdesc, n_mma, idesc, out, ncols);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe_mma: %s\n", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
