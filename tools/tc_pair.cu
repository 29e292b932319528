// Development tool (not part of the library): semantics + rate probe of 2-CTA tcgen05.mma (cta_group::2, M = 256).
// Each CTA of a pair holds 128 rows of A and N/2 rows of B (K-major, 128B swizzle) in its own shared memory; the leader
// CTA issues the MMAs; D rows 0-127 land in CTA 0's TMEM, rows 128-255 in CTA 1's.  Checks the result against the host
// and times the conv-like mix (N=128 + N=64 per k-step).
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tc_pair tools/tc_pair.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_fp16.h>
#include "../playslot_b200/csrc/tc_common.cuh"

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(tc::smem_u32(smem_result)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::);
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}

constexpr int A_ROWS = 600;

// mode 0: correctness (one K=64 GEMM, N given); mode 1: timing of the conv-like mix
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(int mode, int N, int iters, const __half* Aall, const __half* Ball, float* Dout, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA0 = smem;
    uint8_t* sA1 = smem + A_ROWS * 128;
    uint8_t* sB = smem + 2 * A_ROWS * 128;          // 2 stages x 16 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = tid; i < (2 * A_ROWS * 128 + 2 * 16384) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (mode == 0) {
        // A: this CTA's 128 rows x 64 K; B: this CTA's N/2 rows x 64 K   (fp16, swizzled K-major)
        for (int idx = tid; idx < 128 * 64; idx += 128) {
            const int r = idx >> 6, k = idx & 63;
            const int off = r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2;
            *reinterpret_cast<__half*>(sA0 + off) = Aall[(rank * 128 + r) * 64 + k];
        }
        for (int idx = tid; idx < (N / 2) * 64; idx += 128) {
            const int r = idx >> 6, k = idx & 63;
            const int off = r * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2;
            *reinterpret_cast<__half*>(sB + off) = Ball[(rank * (N / 2) + r) * 64 + k];
        }
    }
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    if (warp == 0) tmem_alloc2(&tmem_base, 512);
    tc::fence_async_proxy();
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc::tc_fence_after();
    const uint32_t tm = tmem_base;
    if (rank == 0 && warp == 0 && tc::elect_one()) {
        const uint64_t a0 = tc::umma_desc_sw128(tc::smem_u32(sA0), 1024, 0);
        const uint64_t a1 = tc::umma_desc_sw128(tc::smem_u32(sA1), 1024, 0);
        const uint64_t b0 = tc::umma_desc_sw128(tc::smem_u32(sB), 1024, 0);
        if (mode == 0) {
            const uint32_t id = tc::umma_idesc_f16(256, N);
            for (int k = 0; k < 4; ++k) umma2(tm, a0 + (uint64_t)(k * 2), b0 + (uint64_t)(k * 2), id, k ? 1u : 0u);
        } else {
            const uint32_t id1 = tc::umma_idesc_f16(256, 128), id2 = tc::umma_idesc_f16(256, 64);
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const int tap = it % 25, ky = tap / 5, kx = tap % 5;
                const uint64_t bst = b0 + (uint64_t)((it & 1) * 1024);
#pragma unroll
                for (int tile = 0; tile < 2; ++tile) {
                    const int r = tile * 128 + ky * 68 + kx;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma2(tm + tile * 256, a0 + (uint64_t)(r * 8 + k * 2), bst + (uint64_t)(k * 2), id1, 1u);
                        umma2(tm + tile * 256 + 64, a1 + (uint64_t)(r * 8 + k * 2), bst + (uint64_t)(512 + k * 2), id2, 1u);
                    }
                }
            }
            umma_commit2(&bar, 3);
            tc::mbar_wait(&bar, 0);
            cycles[blockIdx.x >> 1] = clock64() - t0;
        }
        if (mode == 0) umma_commit2(&bar, 3);
    }
    __syncwarp();
    if (mode == 0) {
        tc::mbar_wait(&bar, 0);                    // multicast commit: both CTAs' barriers
        tc::tc_fence_after();
        float v[32];
        for (int c = 0; c < N; c += 32) {
            tc::tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
            for (int i = 0; i < 32; ++i) Dout[(size_t)(rank * 128 + tid) * N + c + i] = v[i];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) tmem_dealloc2(tm, 512);
}

int main() {
    const int smem = 2 * A_ROWS * 128 + 2 * 16384 + 1024;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {128, 64, 256}) {
        std::vector<__half> A(256 * 64), B(N * 64);
        for (int i = 0; i < 256 * 64; ++i) A[i] = __float2half((float)((i * 7 + (i >> 6)) % 13 - 6));
        for (int i = 0; i < N * 64; ++i) B[i] = __float2half((float)((i * 5 + (i >> 6) * 3) % 11 - 5));
        __half *dA, *dB; float* dD; long long* dC;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, 256 * N * 4); cudaMalloc(&dC, 1024);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, 256 * N * 4);
        pair_kernel<<<2, 128, smem>>>(0, N, 0, dA, dB, dD, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: error %s\n", N, cudaGetErrorString(e)); return 1; }
        std::vector<float> D(256 * N);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0; double maxerr = 0;
        for (int r = 0; r < 256; ++r)
            for (int n = 0; n < N; ++n) {
                float ref = 0;
                for (int k = 0; k < 64; ++k) ref += __half2float(A[r * 64 + k]) * __half2float(B[n * 64 + k]);
                const double err = fabs(ref - D[r * N + n]);
                if (err > maxerr) maxerr = err;
                if (err > 1e-3 && bad++ < 5) printf("  mismatch r=%d n=%d got %f want %f\n", r, n, D[r * N + n], ref);
            }
        printf("N=%3d: %s (max err %.3g, %d bad)\n", N, bad ? "MISMATCH" : "OK: D[256 x N] = [A_cta0 ; A_cta1] x [B_cta0 ; B_cta1]^T", maxerr, bad);
    }
    long long* dC; cudaMalloc(&dC, 1024 * 8);
    for (int grid : {2, 148}) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
            pair_kernel<<<grid, 128, smem>>>(1, 128, iters, nullptr, nullptr, nullptr, dC);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("timing: error %s\n", cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> h(grid / 2);
        cudaMemcpy(h.data(), dC, h.size() * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (long long v : h) mx = v > mx ? v : mx;
        printf("grid %3d (pairs): conv mix at M=256: %.1f clk per k-step per pair (single-CTA M=128: 112.0)\n", grid, mx / (iters * 8.0));
    }
    return 0;
}
