// Spatial-broadcast decoder pieces (reference: src/models/SAVi.py:220-248, EncodersDecoders/decoders.py:89-113)
//   ps_dec_expand    : first decoder conv on the broadcast input, collapsed to a per-slot border-class table
//                      G[img][class][co] (a GEMV done with ps_linear) + a weight-only map P[co][y][x]:
//                      h1[img][co][y][x] = relu(G[img][cls(y,x)][co] + P[co][y][x])
//   ps_dec_composite : final 3x3 conv (C -> 4), softmax over the slot axis of the alpha channel, compositing
#include "common.cuh"
#include "../../include/playslot_b200.h"

__device__ __forceinline__ int border_class(int v, int n, int pad) {
    // class index of coordinate v along an axis of length n for a (2*pad+1)-tap "same" convolution
    if (v < pad) return v;
    if (v > n - 1 - pad) return 2 * pad - (n - 1 - v);
    return pad;
}

__global__ void dec_expand_kernel(const float* __restrict__ G, const float* __restrict__ P, float* __restrict__ out,
                                  int n_img, int C, int H, int W, int pad, size_t total4) {
    const int ncls = 2 * pad + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e = i * 4;
        const int x = (int)(e % W);
        const int y = (int)((e / W) % H);
        const int co = (int)((e / ((size_t)W * H)) % C);
        const int img = (int)(e / ((size_t)W * H * C));
        const int cy = border_class(y, H, pad);
        const float* g = G + ((size_t)img * ncls * ncls + (size_t)cy * ncls) * C + co;
        const float4 pv = *reinterpret_cast<const float4*>(P + ((size_t)co * H + y) * W + x);
        float4 o;
        o.x = fmaxf(g[(size_t)border_class(x + 0, W, pad) * C] + pv.x, 0.f);
        o.y = fmaxf(g[(size_t)border_class(x + 1, W, pad) * C] + pv.y, 0.f);
        o.z = fmaxf(g[(size_t)border_class(x + 2, W, pad) * C] + pv.z, 0.f);
        o.w = fmaxf(g[(size_t)border_class(x + 3, W, pad) * C] + pv.w, 0.f);
        *reinterpret_cast<float4*>(out + e) = o;
    }
}

extern "C" int ps_dec_expand(const float* G, const float* P, float* out, int n_img, int C, int H, int W,
                             int ksize, void* stream) {
    PS_REQUIRE(G && P && out, "ps_dec_expand: null pointer");
    PS_REQUIRE(W % 4 == 0 && ksize % 2 == 1 && H > ksize && W > ksize, "ps_dec_expand: unsupported geometry");
    PS_REQUIRE(ps_aligned16(P) && ps_aligned16(out), "ps_dec_expand: pointers must be 16-byte aligned");
    if (n_img == 0) return PS_OK;
    const size_t total4 = (size_t)n_img * C * H * W / 4;
    const int threads = 256;
    size_t blocks = (total4 + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    dec_expand_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(G, P, out, n_img, C, H, W, ksize / 2, total4);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// final conv (3x3, Cin -> 4) + slot softmax + compositing
// ------------------------------------------------------------------------------------------------
struct CompParams {
    const float* h;       // [M*S][Cin][H][W]
    const float* wp;      // packed [Cin][9][4]
    const float* bias;    // [4]
    float* recon;         // [M][3][H][W]
    float* recons;        // [M][S][3][H][W]
    float* masks;         // [M][S][1][H][W]
    int M, S, Cin, H, W;
};

#define CP_TW 64          // tile width  (pixels)
#define CP_TH 16          // tile height
#define CP_CI 8           // input channels staged per step
#define CP_HS 68          // row stride of the staged halo tile (floats): 66 used, 16-byte aligned rows

// 256 threads, 4 horizontally adjacent pixels per thread: per (channel, kernel row) one LDS.128 + one LDS.64 of inputs and
// three broadcast LDS.128 of weights feed 48 FMAs.  The slot softmax is evaluated online while looping over the slots.
__global__ __launch_bounds__(256) void dec_composite_kernel(CompParams p) {
    extern __shared__ __align__(16) float sm[];
    float* s_w = sm;                              // [Cin][9][4]
    float* s_in = sm + p.Cin * 36;                // [CP_CI][CP_TH + 2][CP_HS]
    const int tid = threadIdx.x;
    const int tiles_x = (p.W + CP_TW - 1) / CP_TW;
    const int ty0 = (blockIdx.x / tiles_x) * CP_TH, tx0 = (blockIdx.x % tiles_x) * CP_TW;
    const int m = blockIdx.y;
    const int py = tid >> 4, px = (tid & 15) * 4;
    const int oy = ty0 + py, ox = tx0 + px;
    const bool active = ox < p.W;                 // W % 4 == 0: a thread's four pixels are all inside or all outside
    const size_t HW = (size_t)p.H * p.W;
    for (int i = tid; i < p.Cin * 36; i += 256) s_w[i] = p.wp[i];
    const float bias[4] = {p.bias[0], p.bias[1], p.bias[2], p.bias[3]};

    float mrun[4], drun[4], rr[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) { mrun[i] = -INFINITY; drun[i] = 0.f; rr[i][0] = rr[i][1] = rr[i][2] = 0.f; }
    for (int s = 0; s < p.S; ++s) {
        const float* hin = p.h + ((size_t)m * p.S + s) * p.Cin * HW;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = bias[c];
        for (int c0 = 0; c0 < p.Cin; c0 += CP_CI) {
            __syncthreads();
            for (int idx = tid; idx < CP_CI * (CP_TH + 2) * (CP_TW + 2); idx += 256) {
                const int ci = idx / ((CP_TH + 2) * (CP_TW + 2)), r = idx - ci * ((CP_TH + 2) * (CP_TW + 2));
                const int hy = r / (CP_TW + 2), hx = r - hy * (CP_TW + 2);
                const int gy = ty0 + hy - 1, gx = tx0 + hx - 1;
                float v = 0.f;
                if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = hin[(size_t)(c0 + ci) * HW + (size_t)gy * p.W + gx];
                s_in[(ci * (CP_TH + 2) + hy) * CP_HS + hx] = v;
            }
            __syncthreads();
#pragma unroll 2
            for (int ci = 0; ci < CP_CI; ++ci) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float* row = s_in + (ci * (CP_TH + 2) + py + ky) * CP_HS + px;
                    const float4 v0 = *reinterpret_cast<const float4*>(row);
                    const float2 v1 = *reinterpret_cast<const float2*>(row + 4);
                    const float v[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4 w = *reinterpret_cast<const float4*>(s_w + ((c0 + ci) * 9 + ky * 3 + kx) * 4);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc[i][0] = fmaf(v[i + kx], w.x, acc[i][0]); acc[i][1] = fmaf(v[i + kx], w.y, acc[i][1]);
                            acc[i][2] = fmaf(v[i + kx], w.z, acc[i][2]); acc[i][3] = fmaf(v[i + kx], w.w, acc[i][3]);
                        }
                    }
                }
            }
        }
        if (active && oy < p.H) {
            const size_t pix = (size_t)oy * p.W + ox;
            float* rc = p.recons + (((size_t)m * p.S + s) * 3) * HW + pix;
#pragma unroll
            for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(rc + c * HW) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
            *reinterpret_cast<float4*>(p.masks + ((size_t)m * p.S + s) * HW + pix) = make_float4(acc[0][3], acc[1][3], acc[2][3], acc[3][3]);   // alpha logits, normalised below
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a3 = acc[i][3];
            const float mnew = fmaxf(mrun[i], a3);
            const float corr = expf(mrun[i] - mnew);                // exp(-inf) = 0 on the first slot
            const float e = expf(a3 - mnew);
            drun[i] = drun[i] * corr + e;
            rr[i][0] = rr[i][0] * corr + e * acc[i][0]; rr[i][1] = rr[i][1] * corr + e * acc[i][1]; rr[i][2] = rr[i][2] * corr + e * acc[i][2];
            mrun[i] = mnew;
        }
    }
    if (!(active && oy < p.H)) return;
    const size_t pix = (size_t)oy * p.W + ox;
    float inv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) inv[i] = 1.0f / drun[i];
    for (int s = 0; s < p.S; ++s) {
        float4* mk = reinterpret_cast<float4*>(p.masks + ((size_t)m * p.S + s) * HW + pix);
        const float4 a = *mk;
        *mk = make_float4(expf(a.x - mrun[0]) * inv[0], expf(a.y - mrun[1]) * inv[1], expf(a.z - mrun[2]) * inv[2], expf(a.w - mrun[3]) * inv[3]);
    }
    float* ro = p.recon + (size_t)m * 3 * HW + pix;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float4*>(ro + c * HW) = make_float4(rr[0][c] * inv[0], rr[1][c] * inv[1], rr[2][c] * inv[2], rr[3][c] * inv[3]);
}

extern "C" int ps_dec_composite(const float* h, const float* w_packed, const float* bias, float* recon,
                                float* recons, float* masks, int M, int S, int Cin, int H, int W, void* stream) {
    PS_REQUIRE(h && w_packed && bias && recon && recons && masks, "ps_dec_composite: null pointer");
    PS_REQUIRE(Cin % CP_CI == 0 && Cin <= 128, "ps_dec_composite: Cin must be a multiple of %d and <= 128", CP_CI);
    PS_REQUIRE(H % CP_TH == 0 && W % 4 == 0, "ps_dec_composite: H must be a multiple of %d and W of 4", CP_TH);
    PS_REQUIRE(ps_aligned16(recon) && ps_aligned16(recons) && ps_aligned16(masks), "ps_dec_composite: outputs must be 16-byte aligned");
    PS_REQUIRE(S >= 1, "ps_dec_composite: bad slot count");
    if (M == 0) return PS_OK;
    PS_REQUIRE(M <= 65535, "ps_dec_composite: at most 65535 frames per launch");
    CompParams p;
    p.h = h; p.wp = w_packed; p.bias = bias; p.recon = recon; p.recons = recons; p.masks = masks;
    p.M = M; p.S = S; p.Cin = Cin; p.H = H; p.W = W;
    const size_t smem = (size_t)(Cin * 36 + CP_CI * (CP_TH + 2) * CP_HS) * sizeof(float);
    dec_composite_kernel<<<dim3((H / CP_TH) * ((W + CP_TW - 1) / CP_TW), M), 256, smem, (cudaStream_t)stream>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}
