// Shared helpers for the playslot_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define PS_OK 0
#define PS_EINVAL (-1)
#define PS_EUNSUPPORTED (-2)

// thread-local last-error message, readable through ps_last_error()
void ps_set_error(const char* fmt, ...);

#define PS_REQUIRE(cond, ...)                       \
    do {                                            \
        if (!(cond)) {                              \
            ps_set_error(__VA_ARGS__);              \
            return PS_EINVAL;                       \
        }                                           \
    } while (0)

#define PS_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            ps_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                         __FILE__, __LINE__);                                      \
            return (int)e__;                                                       \
        }                                                                          \
    } while (0)

#define PS_LAUNCH_CHECK()                                                          \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            ps_set_error("kernel launch failed: %s (%s:%d)",                       \
                         cudaGetErrorString(e__), __FILE__, __LINE__);             \
            return (int)e__;                                                       \
        }                                                                          \
    } while (0)

// every launch made through the C-ABI bumps this counter (bench.py reports it as gpu_launches)
extern unsigned long long g_ps_launch_count;
#define PS_COUNT_LAUNCH() (++g_ps_launch_count)
// SMs the persistent kernels leave free (ps_set_sm_reserve)
extern int g_ps_sm_reserve;
static inline int ps_persistent_ctas(int sms) { const int n = sms - g_ps_sm_reserve; return n < 1 ? 1 : n; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE setting of a kernel: remember, per device, the largest size that
// has been configured (a process may drive several GPUs).  `set_bytes` (default: smem) is the value passed to the driver.
struct PsSmemConfig { size_t bytes[64]; };
template <typename Kernel>
static inline cudaError_t ps_ensure_smem(Kernel kernel, size_t smem, PsSmemConfig& cfg, size_t set_bytes = 0) {
    if (set_bytes == 0) set_bytes = smem;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)set_bytes);
    if (smem > cfg.bytes[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)set_bytes);
        if (e != cudaSuccess) return e;
        cfg.bytes[dev] = set_bytes;
    }
    return cudaSuccess;
}

static inline bool ps_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
#endif
