// Fused loss / metric tail of the slot pipeline.
//   ps_frame_sq_err : per-frame sum of squared differences between predicted and target frames (optionally clamped to [0, 1]) in ONE
//                     pass over both tensors -- the MSE terms of the training loss (PredImgMSE / PredSlotMSE, src/lib/loss.py:201-238:
//                     nn.MSELoss = mean over all elements) and the per-frame PSNR of the evaluation
//                     (src/05_evaluate_PlaySlot.py:98-108 -> lib/metrics.py:225-235 -> piqa.psnr: 10 log10(1 / (mse + 1e-8)) per
//                     image) are both functions of these sums, so neither (pred - target)^2 tensors nor the frames themselves
//                     leave the device.
//   ps_mse_bwd      : d/d pred of scale * sum (pred - target)^2  =  2 scale (pred - target) g, one pass.
// Deterministic: one CTA per frame, fixed reduction tree (no atomics).
#include "common.cuh"
#include "../../include/playslot_b200.h"

namespace {

__global__ void __launch_bounds__(256) frame_sq_err_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          float* __restrict__ out, long long elems, int frames_per_video,
                                                          long long pred_video_stride, long long pred_frame_stride,
                                                          long long tgt_video_stride, long long tgt_frame_stride, int clamp01) {
    const int f = blockIdx.x;
    const int v = f / frames_per_video, t = f - v * frames_per_video;
    const float* p = pred + v * pred_video_stride + t * pred_frame_stride;
    const float* q = target + v * tgt_video_stride + t * tgt_frame_stride;
    float acc = 0.f;
    const bool vec = (elems % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q)) % 16 == 0);
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const float4* q4 = reinterpret_cast<const float4*>(q);
        for (long long i = threadIdx.x; i < elems / 4; i += 256) {
            float4 a = p4[i], b = q4[i];
            if (clamp01) {
                a.x = fminf(fmaxf(a.x, 0.f), 1.f); a.y = fminf(fmaxf(a.y, 0.f), 1.f); a.z = fminf(fmaxf(a.z, 0.f), 1.f); a.w = fminf(fmaxf(a.w, 0.f), 1.f);
                b.x = fminf(fmaxf(b.x, 0.f), 1.f); b.y = fminf(fmaxf(b.y, 0.f), 1.f); b.z = fminf(fmaxf(b.z, 0.f), 1.f); b.w = fminf(fmaxf(b.w, 0.f), 1.f);
            }
            const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
            acc += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
    } else {
        for (long long i = threadIdx.x; i < elems; i += 256) {
            float a = p[i], b = q[i];
            if (clamp01) { a = fminf(fmaxf(a, 0.f), 1.f); b = fminf(fmaxf(b, 0.f), 1.f); }
            const float d = a - b;
            acc = fmaf(d, d, acc);
        }
    }
    __shared__ float s[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += s[w];
        out[f] = r;
    }
}

__global__ void __launch_bounds__(256) mse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                     const float* __restrict__ g, float scale, float* __restrict__ d_pred,
                                                     long long elems, int frames_per_video, long long n_frames,
                                                     long long tgt_video_stride, long long tgt_frame_stride) {
    const float gs = 2.0f * scale * (g ? *g : 1.0f);
    const long long total = n_frames * elems;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long f = i / elems, e = i - f * elems;
        const long long v = f / frames_per_video, t = f - v * frames_per_video;
        d_pred[i] = gs * (pred[i] - target[v * tgt_video_stride + t * tgt_frame_stride + e]);
    }
}

}  // namespace

// out[f] = sum over the `elems` values of frame f of (pred - target)^2, f = video * frames_per_video + frame.  Frame f of `pred`
// starts at pred + video * pred_video_stride + frame * pred_frame_stride (elements); the same for `target`, so a time slice of the
// input clip (videos[:, ns:ns+np]) is read in place.
extern "C" int ps_frame_sq_err(const float* pred, const float* target, float* out, long long n_videos, int frames_per_video,
                               long long elems, long long pred_video_stride, long long pred_frame_stride,
                               long long tgt_video_stride, long long tgt_frame_stride, int clamp01, void* stream) {
    PS_REQUIRE(pred && target && out, "ps_frame_sq_err: null pointer");
    PS_REQUIRE(n_videos >= 0 && frames_per_video > 0 && elems > 0, "ps_frame_sq_err: bad shape");
    const long long n_frames = n_videos * frames_per_video;
    if (n_frames == 0) return PS_OK;
    PS_REQUIRE(n_frames < (1ll << 31), "ps_frame_sq_err: too many frames");
    frame_sq_err_kernel<<<(unsigned)n_frames, 256, 0, (cudaStream_t)stream>>>(pred, target, out, elems, frames_per_video, pred_video_stride,
                                                                             pred_frame_stride, tgt_video_stride, tgt_frame_stride, clamp01);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// d_pred = 2 * scale * g[0] * (pred - target); pred / d_pred contiguous (n_videos * frames_per_video frames of `elems` values),
// target addressed like in ps_frame_sq_err.  g may be null (= 1).
extern "C" int ps_mse_bwd(const float* pred, const float* target, const float* g, float scale, float* d_pred, long long n_videos,
                          int frames_per_video, long long elems, long long tgt_video_stride, long long tgt_frame_stride, void* stream) {
    PS_REQUIRE(pred && target && d_pred, "ps_mse_bwd: null pointer");
    const long long n_frames = n_videos * frames_per_video;
    if (n_frames * elems == 0) return PS_OK;
    long long blocks = (n_frames * elems + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    mse_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, g, scale, d_pred, elems, frames_per_video, n_frames,
                                                                      tgt_video_stride, tgt_frame_stride);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
