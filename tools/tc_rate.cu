// Development tool (not part of the library): tcgen05.mma issue-rate probe.  One CTA per SM issues a long stream of
// SS-mode kind::f16 MMAs (M = 128) against shared-memory operands and reports cycles per MMA, for several N and
// several row offsets of the A window inside its 128B-swizzled region (the implicit-GEMM convolution addresses its
// taps as row shifts, so its A windows start on arbitrary rows, not on 8-row swizzle-atom boundaries).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tc_rate tools/tc_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <vector>
#include "../playslot_b200/csrc/tc_common.cuh"

struct Case { int n1, n2, a_off1, a_off2, b_off, m1, m2; };   // m1 / m2: M of the two MMAs (0 = 128)     // per k-step: MMA(N = n1, A rows from a_off1) [+ MMA(N = n2, A from plane 2 at a_off2)]

__global__ void __launch_bounds__(128, 1) rate_kernel(Case c, int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA0 = smem;                  // 600 rows x 128 B
    uint8_t* sA1 = smem + 600 * 128;
    uint8_t* sB = smem + 1200 * 128;      // 2 stages x 32 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (1200 * 128 + 4 * 16384) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
    tc::fence_async_proxy();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    if (warp == 0 && tc::elect_one()) {
        const uint32_t id1 = tc::umma_idesc_f16(c.m1 ? c.m1 : 128, c.n1);
        const uint32_t id2 = c.n2 ? tc::umma_idesc_f16(c.m2 ? c.m2 : 128, c.n2) : 0;
        const uint64_t a0 = tc::umma_desc_sw128(tc::smem_u32(sA0), 1024, 0);
        const uint64_t a1 = tc::umma_desc_sw128(tc::smem_u32(sA1), 1024, 0);
        const uint64_t b0 = tc::umma_desc_sw128(tc::smem_u32(sB), 1024, 0);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int tap = it % 25;
            const int ky = tap / 5, kx = tap % 5;
            const uint64_t bst = b0 + (uint64_t)((it & 1) * 2048) + (uint64_t)(c.b_off * 8);
#pragma unroll
            for (int tile = 0; tile < 2; ++tile) {
                // conv-like: the window moves with the tap; a_off selects the residue of the start row modulo 8
                const int r1 = tile * 128 + ky * 64 + kx * 8 + c.a_off1 * (kx + ky + 1);
                const int r2 = tile * 128 + ky * 64 + kx * 8 + c.a_off2 * (kx + ky + 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    tc::umma_bf16(tm + tile * 256, a0 + (uint64_t)(r1 * 8 + k * 2), bst + (uint64_t)(k * 2), id1, 1u);
                    if (c.n2) tc::umma_bf16(tm + tile * 256 + (c.n2 > 192 ? 0 : 64), a1 + (uint64_t)(r2 * 8 + k * 2), bst + (uint64_t)(k * 2), id2, 1u);
                }
            }
        }
        tc::umma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}

int main() {
    const int smem = 1200 * 128 + 4 * 16384 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const Case cases[] = {
        {64, 0, 0, 0, 0}, {64, 0, 1, 0, 0}, {128, 0, 0, 0, 0}, {128, 0, 1, 0, 0}, {128, 0, 4, 0, 0}, {256, 0, 0, 0, 0}, {256, 0, 1, 0, 0},
        {128, 64, 0, 0, 0}, {128, 64, 1, 1, 0}, {128, 64, 3, 3, 0}, {128, 128, 0, 0, 0}, {128, 128, 1, 1, 0}, {192, 0, 0, 0, 0}, {192, 0, 1, 0, 0},
        // weights-stationary costing: what an M = 64 product costs next to an M = 128 one at N = 256 / 128
        {256, 0, 0, 0, 0, 64, 0}, {128, 0, 0, 0, 0, 64, 0}, {256, 256, 0, 0, 0, 128, 64}, {128, 128, 0, 0, 0, 128, 64},
    };
    for (int grid : {1, 148}) {
        for (const Case& c : cases) {
            const int iters = 2000;
            for (int rep = 0; rep < 2; ++rep) {
                rate_kernel<<<grid, 128, smem>>>(c, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            std::vector<long long> h(grid);
            cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
            const double ksteps = (double)iters * 8;
            printf("grid %3d  M1 %3d N1 %3d  M2 %3d N2 %3d  a_off %d/%d : %.1f clk per k-step (%.1f per MMA)\n", grid, c.m1 ? c.m1 : 128, c.n1,
                   c.n2 ? (c.m2 ? c.m2 : 128) : 0, c.n2, c.a_off1, c.a_off2, mx / ksteps, mx / ksteps / (c.n2 ? 2 : 1));
        }
    }
    return 0;
}
