// Stand-alone tcgen05 probe (development tool, not part of the library):
//   D[128x64] = A[off : off+128, 0:64] * B[64x64]^T  with bf16 operands loaded by TMA (128B swizzle),
//   for row offsets `off` that are not multiples of the 8-row swizzle atom and two base_offset conventions.
// Prints the max abs error against a host reference for each (off, mode).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../playslot_b200/csrc/tc_common.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int AROWS = 192;   // rows of A staged in shared memory
constexpr int KCH = 64;      // channels (K)

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    float* out, int off, int mode) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // 192 rows x 128 B
    uint8_t* sB = smem + AROWS * 128;         // 64 rows x 128 B
    __shared__ uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { tc::mbar_init(&bar_load, 1); tc::mbar_init(&bar_mma, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    if (warp == 0 && tc::elect_one()) {
        tc::mbar_expect_tx(&bar_load, (AROWS + 64) * 128);
        for (int i = 0; i < AROWS / 64; ++i) tc::tma_load_2d(sA + i * 64 * 128, &mapA, &bar_load, 0, i * 64);
        tc::tma_load_2d(sB, &mapB, &bar_load, 0, 0);
        tc::mbar_wait(&bar_load, 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::umma_idesc_bf16(128, 64);
        for (int k = 0; k < KCH / 16; ++k) {
            const uint32_t a_addr = tc::smem_u32(sA) + off * 128 + k * 32;
            const uint32_t b_addr = tc::smem_u32(sB) + k * 32;
            const uint32_t bo = (mode == 1) ? ((a_addr >> 7) & 7) : 0;
            tc::umma_bf16(tm, tc::umma_desc_sw128(a_addr, 1024, bo), tc::umma_desc_sw128(b_addr, 1024, 0), idesc, k > 0);
        }
        tc::umma_commit(&bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(&bar_mma, 0);
    tc::tc_fence_after();
    float v[32];
    for (int c = 0; c < 2; ++c) {
        tc::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c * 32, v);
        for (int i = 0; i < 32; ++i) out[(size_t)tid * 64 + c * 32 + i] = v[i];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 64);
}

int main() {
    const int RA = 256;
    std::vector<__nv_bfloat16> hA(RA * KCH), hB(64 * KCH);
    std::vector<float> fA(RA * KCH), fB(64 * KCH);
    srand(1);
    for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
    __nv_bfloat16 *dA, *dB; float* dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 128 * 64 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap mapA, mapB;
    int r1 = tc::make_tmap_2d_bf16_sw128(&mapA, dA, RA, KCH, 64, 64);
    int r2 = tc::make_tmap_2d_bf16_sw128(&mapB, dB, 64, KCH, 64, 64);
    printf("tensor maps: %d %d\n", r1, r2);
    if (r1 || r2) return 1;
    const int smem = (AROWS + 64) * 128 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> hO(128 * 64);
    for (int mode = 0; mode < 2; ++mode)
        for (int off = 0; off <= 20; ++off) {
            CK(cudaMemset(dO, 0, 128 * 64 * 4));
            probe_kernel<<<1, 128, smem>>>(mapA, mapB, dO, off, mode);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d off %d: kernel failed: %s\n", mode, off, cudaGetErrorString(e)); return 2; }
            CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    double acc = 0;
                    for (int k = 0; k < KCH; ++k) acc += (double)fA[(m + off) * KCH + k] * fB[n * KCH + k];
                    maxerr = fmax(maxerr, fabs(acc - hO[m * 64 + n]));
                    maxref = fmax(maxref, fabs(acc));
                }
            printf("mode %d off %2d: max|err| = %.3e (max|ref| %.2f) %s\n", mode, off, maxerr, maxref, maxerr < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
