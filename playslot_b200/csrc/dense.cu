// fp32 CUDA-core building blocks shared by the encoder MLP, the transition block, InvDyn and cOCVP:
//   ps_linear  : Y = act(LN?(X) W^T + b) (+ residual), tiled SGEMM with fused LayerNorm prologue
//   ps_mha     : multi-head self attention over short token sequences (<= a few hundred tokens),
//                optional block-causal (time x slot staircase) mask
//   ps_conv2d  : direct KxK stride-1 "same" convolution, NCHW, fused bias/ReLU, optional NHWC(+addend) output
// These are the exact-fp32 reference kernels of the library; the tcgen05 kernels are checked against them.
#include "common.cuh"
#include "../../include/playslot_b200.h"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

// ------------------------------------------------------------------------------------------------
// linear
// ------------------------------------------------------------------------------------------------
struct LinearParams {
    const float* X; int ldx;
    const float* W;
    const float* bias;
    const float* ln_g; const float* ln_b; float ln_eps;
    const float* res; int ldres; int res_div; int res_mod;
    float* Y; int ldy;
    int M, N, K;
    int act;
    int vec_store;
    int ksplit;      // > 1: the K range is split over gridDim.z CTAs that atomically accumulate into a zeroed Y (no LayerNorm,
                     // activation or residual in that mode; the bias is added by the first slice)
};

#define LBM 64
#define LBN 64
#define LBK 16

template <bool LN>
__global__ __launch_bounds__(256) void linear_kernel(LinearParams p) {
    __shared__ __align__(16) float As[LBK][LBM + 4];
    __shared__ __align__(16) float Bs[LBK][LBN + 4];
    __shared__ float s_mean[LBM], s_rstd[LBM];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * LBM, n0 = blockIdx.y * LBN;
    const int lr = tid >> 2;          // tile row handled by this thread when loading (0..63)
    const int lk = (tid & 3) * 4;     // k offset inside the k-tile (0,4,8,12)

    float mean = 0.f, rstd = 1.f;
    if (LN) {
        const int m = m0 + lr;
        float s = 0.f;
        if (m < p.M)
            for (int k = lk; k < p.K; k += 16) {
                float4 v = *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k);
                s += (v.x + v.y) + (v.z + v.w);
            }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        mean = s / (float)p.K;
        float ss = 0.f;
        if (m < p.M)
            for (int k = lk; k < p.K; k += 16) {
                float4 v = *reinterpret_cast<const float4*>(p.X + (size_t)m * p.ldx + k);
                float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
                ss += (a * a + b * b) + (c * c + d * d);
            }
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        rstd = 1.0f / sqrtf(ss / (float)p.K + p.ln_eps);
        if ((tid & 3) == 0) { s_mean[lr] = mean; s_rstd[lr] = rstd; }
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int tx = tid & 15, ty = tid >> 4;
    // the global loads of k-tile i+1 are issued before the FMAs of k-tile i (register prefetch): their latency hides behind
    // the 256 FMAs per thread instead of being paid once per tile
    auto load_tile = [&](int k0, float4& av, float4& bv) {
        av = make_float4(0.f, 0.f, 0.f, 0.f); bv = av;
        const int k = k0 + lk;
        if (m0 + lr < p.M && k < p.K) {
            av = *reinterpret_cast<const float4*>(p.X + (size_t)(m0 + lr) * p.ldx + k);
            if (LN) {
                float4 g = *reinterpret_cast<const float4*>(p.ln_g + k);
                float4 b = *reinterpret_cast<const float4*>(p.ln_b + k);
                av.x = (av.x - mean) * rstd * g.x + b.x;
                av.y = (av.y - mean) * rstd * g.y + b.y;
                av.z = (av.z - mean) * rstd * g.z + b.z;
                av.w = (av.w - mean) * rstd * g.w + b.w;
            }
        }
        if (n0 + lr < p.N && k < p.K)
            bv = *reinterpret_cast<const float4*>(p.W + (size_t)(n0 + lr) * p.K + k);
    };
    // split-K: this CTA's K range (multiples of LBK)
    int kbeg = 0, kend = p.K;
    if (p.ksplit > 1) {
        const int per = ((p.K + p.ksplit - 1) / p.ksplit + LBK - 1) / LBK * LBK;
        kbeg = blockIdx.z * per;
        kend = min(p.K, kbeg + per);
    }
    float4 av, bv;
    load_tile(kbeg, av, bv);
    for (int k0 = kbeg; k0 < kend; k0 += LBK) {
        As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
        Bs[lk + 0][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
        __syncthreads();
        if (k0 + LBK < kend) load_tile(k0 + LBK, av, bv);
#pragma unroll
        for (int kk = 0; kk < LBK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        const int n = n0 + tx * 4;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = acc[i][j];
            if (n + j < p.N) {
                if (p.bias) v += p.bias[n + j];
                if (p.act == 1) v = fmaxf(v, 0.f);
                else if (p.act == 2) v = gelu_erf(v);
                else if (p.act == 3) v = fabsf(v);
                if (p.res) {
                    const int rr = (m / p.res_div) % p.res_mod;
                    v += p.res[(size_t)rr * p.ldres + n + j];
                }
            }
            o[j] = v;
        }
        float* dst = p.Y + (size_t)m * p.ldy + n;
        if (p.ksplit > 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) atomicAdd(dst + j, (blockIdx.z == 0 || !p.bias) ? o[j] : o[j] - p.bias[n + j]);
        } else if (p.vec_store && n + 3 < p.N) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n + j < p.N) dst[j] = o[j];
        }
    }
}

extern "C" int ps_linear(const float* X, int ldx, const float* W, const float* bias,
                         const float* ln_gamma, const float* ln_beta, float ln_eps,
                         const float* res, int ldres, int res_div, int res_mod,
                         float* Y, int ldy, int M, int N, int K, int act, void* stream) {
    PS_REQUIRE(X && W && Y, "ps_linear: null pointer");
    PS_REQUIRE(M >= 0 && N > 0 && K > 0, "ps_linear: bad shape M=%d N=%d K=%d", M, N, K);
    if (M == 0) return PS_OK;
    PS_REQUIRE(K % 4 == 0 && ldx % 4 == 0, "ps_linear: K (%d) and ldx (%d) must be multiples of 4", K, ldx);
    PS_REQUIRE(ps_aligned16(X) && ps_aligned16(W), "ps_linear: X and W must be 16-byte aligned");
    PS_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "ps_linear: LayerNorm needs both gamma and beta");
    PS_REQUIRE(!ln_gamma || (ps_aligned16(ln_gamma) && ps_aligned16(ln_beta)), "ps_linear: LN params must be 16-byte aligned");
    PS_REQUIRE(act >= 0 && act <= 3, "ps_linear: unknown activation %d", act);
    LinearParams p;
    p.X = X; p.ldx = ldx; p.W = W; p.bias = bias; p.ln_g = ln_gamma; p.ln_b = ln_beta; p.ln_eps = ln_eps;
    p.res = res; p.ldres = ldres; p.res_div = res_div > 0 ? res_div : 1; p.res_mod = res_mod > 0 ? res_mod : M;
    p.Y = Y; p.ldy = ldy; p.M = M; p.N = N; p.K = K; p.act = act;
    p.vec_store = (ldy % 4 == 0 && ps_aligned16(Y)) ? 1 : 0;
    dim3 grid((M + LBM - 1) / LBM, (N + LBN - 1) / LBN);
    PS_REQUIRE(grid.y <= 65535, "ps_linear: N=%d too large for one launch", N);
    cudaStream_t st = (cudaStream_t)stream;
    // Few output tiles and a long K (the training-path GEMMs at the reference's batch size: 256 rows): split K over up to 8
    // CTAs per tile so the launch fills the GPU; partial sums meet in Y through atomics (summation order not fixed).
    p.ksplit = 1;
    if (!ln_gamma && act == 0 && !res && ldy == N && grid.x * grid.y < 74 && K >= 256) {
        int ks = 148 / (int)(grid.x * grid.y);
        if (ks > 8) ks = 8;
        if (ks > K / 128) ks = K / 128;
        if (ks > 1) {
            p.ksplit = ks;
            grid.z = ks;
            PS_CUDA(cudaMemsetAsync(Y, 0, (size_t)M * N * sizeof(float), st));
        }
    }
    if (ln_gamma) linear_kernel<true><<<grid, 256, 0, st>>>(p);
    else linear_kernel<false><<<grid, 256, 0, st>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-head self attention for short sequences
// ------------------------------------------------------------------------------------------------
struct MhaParams {
    const float* qkv; int ld;
    const float* res;      // optional residual added to the output, same layout as out
    float* out; int ldo;
    uint16_t* out_hi;      // tiled kernel only: write the result as fp16 [hi | lo] planes (operand of a tensor-core linear)
    uint16_t* out_lo;      // instead of fp32 `out`
    int n_seq, L, H, dh, spf;
    float scale;
    int rows_per_chunk;    // row-streaming kernel: query rows per CTA (blockIdx.z selects the chunk; L when the grid has one chunk)
};

// One CTA per (sequence, head, chunk of query rows).  K and V are staged in shared memory; every warp processes FOUR query rows at a time so that
// each shared-memory read of a key / value element feeds four FMAs (the kernel is LDS-bound otherwise): the four queries are
// stored interleaved [d][4] and the four probability rows [j][4], both read as one broadcast LDS.128.
__global__ __launch_bounds__(128) void mha_kernel(MhaParams p) {
    extern __shared__ __align__(16) float sm[];
    const int seq = blockIdx.x, h = blockIdx.y;
    const int L = p.L, dh = p.dh, inner = p.H * dh;
    float* Ks = sm;                      // [L][dh+1]
    float* Vs = Ks + L * (dh + 1);       // [L][dh]
    float* Ps = sm + ((L * (dh + 1) + L * dh + 3) & ~3);    // [4 warps][L][4], 16-byte aligned
    float* Qs = Ps + 4 * L * 4;          // [4 warps][dh][4]
    const float* base = p.qkv + (size_t)seq * L * p.ld;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // Long sequences on few (sequence, head) pairs -- the 160-token windows of the 16-slot configuration at 8 videos per GPU --
    // split the query rows over blockIdx.z so that the launch fills the GPU.  With the block-causal mask the rows of a chunk only
    // see the keys up to the chunk's last frame: the keys behind it would be filled with -1e9 and contribute exp(.) = 0 exactly,
    // so they are neither staged nor visited (bit-identical results).
    const int r_begin = blockIdx.z * p.rows_per_chunk, r_end = min(L, r_begin + p.rows_per_chunk);
    const int Lk = p.spf > 0 ? min(L, ((r_end - 1) / p.spf + 1) * p.spf) : L;      // keys any row of this chunk attends to
    for (int idx = tid; idx < Lk * dh; idx += 128) {
        const int j = idx / dh, d = idx - j * dh;
        Ks[j * (dh + 1) + d] = base[(size_t)j * p.ld + inner + h * dh + d];
        Vs[j * dh + d] = base[(size_t)j * p.ld + 2 * inner + h * dh + d];
    }
    __syncthreads();
    float* Pw = Ps + w * L * 4;
    float* Qw = Qs + w * dh * 4;
    for (int i0 = r_begin + w * 4; i0 < r_end; i0 += 16) {
        // the four query rows of this pass, interleaved; rows past the end are zero (their results are not stored)
        for (int idx = lane; idx < dh * 4; idx += 32) {
            const int d = idx >> 2, r = idx & 3;
            Qw[idx] = (i0 + r < r_end) ? base[(size_t)(i0 + r) * p.ld + h * dh + d] : 0.f;
        }
        __syncwarp();
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int j = lane; j < Lk; j += 32) {
            const float* kr = Ks + j * (dh + 1);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
            for (int d = 0; d < dh; ++d) {
                const float k = kr[d];
                const float4 q = *reinterpret_cast<const float4*>(Qw + d * 4);
                a0 = fmaf(q.x, k, a0); a1 = fmaf(q.y, k, a1); a2 = fmaf(q.z, k, a2); a3 = fmaf(q.w, k, a3);
            }
            float dot[4] = {a0 * p.scale, a1 * p.scale, a2 * p.scale, a3 * p.scale};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (p.spf > 0 && (j / p.spf) > ((i0 + r) / p.spf)) dot[r] = -1e9f;   // masked_fill(mask == 0, -1e9)
                mx[r] = fmaxf(mx[r], dot[r]);
            }
            *reinterpret_cast<float4*>(Pw + j * 4) = make_float4(dot[0], dot[1], dot[2], dot[3]);
        }
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 4; ++r) mx[r] = warp_max(mx[r]);
        for (int j = lane; j < Lk; j += 32) {
            float4 e = *reinterpret_cast<const float4*>(Pw + j * 4);
            e.x = expf(e.x - mx[0]); e.y = expf(e.y - mx[1]); e.z = expf(e.z - mx[2]); e.w = expf(e.w - mx[3]);
            sum[0] += e.x; sum[1] += e.y; sum[2] += e.z; sum[3] += e.w;
            *reinterpret_cast<float4*>(Pw + j * 4) = e;
        }
        float inv[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) inv[r] = 1.0f / warp_sum(sum[r]);
        __syncwarp();
        for (int d = lane; d < dh; d += 32) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
            for (int j = 0; j < Lk; ++j) {
                const float v = Vs[j * dh + d];
                const float4 pr = *reinterpret_cast<const float4*>(Pw + j * 4);
                a0 = fmaf(pr.x * inv[0], v, a0); a1 = fmaf(pr.y * inv[1], v, a1);
                a2 = fmaf(pr.z * inv[2], v, a2); a3 = fmaf(pr.w * inv[3], v, a3);
            }
            const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (i0 + r < r_end) {
                    const size_t oi = ((size_t)seq * L + i0 + r) * p.ldo + h * dh + d;
                    p.out[oi] = p.res ? acc[r] + p.res[oi] : acc[r];
                }
            }
        }
        __syncwarp();
    }
}

// Register-tiled variant for head_dim % 4 == 0 (every shipped config): one CTA per (sequence, head), 128 threads.
//   S = scale * Q K^T : thread tile 3 query rows x 6 keys, operands read as LDS.128 along the head dimension (9 loads per 72 FMAs);
//                       rows / keys of a thread are strided (i = ty + nTy*r, j = tx + nTx*c) so that the 128-bit loads of a warp
//                       fall into distinct banks with the (dh + 4)-float row pitch
//   P = softmax(S)    : one warp per row, written transposed (Pt[j][i]) for the next stage
//   O = P V           : thread tile 6 consecutive rows x 4 consecutive head columns (3 LDS.64 + 1 LDS.128 per 24 FMAs)
// For the cOCVP shape (L = 48, dh = 64) both stages use all 128 threads exactly once.
#define MHA_RT 3
#define MHA_CT 6
__global__ __launch_bounds__(128) void mha_tiled_kernel(MhaParams p) {
    extern __shared__ __align__(16) float sm[];
    const int seq = blockIdx.x, h = blockIdx.y;
    const int L = p.L, dh = p.dh, inner = p.H * dh;
    const int L6 = (L + 5) / 6 * 6;            // padded row count of Q / K (zero rows) and of the transposed probabilities
    const int pitch = dh + 4;
    const int Lt = (L6 + 3) & ~3;              // pitch of Pt rows (16-byte aligned)
    float* Qs = sm;                            // [L6][pitch]; reused as Pt [L][Lt] after the first stage
    float* Ks = Qs + max(L6 * pitch, L * Lt);  // [L6][pitch]
    float* Vs = Ks + L6 * pitch;               // [L][dh]
    float* Ss = Vs + L * dh;                   // [L][L + 1]
    int* frame = reinterpret_cast<int*>(Ss + L * (L + 1));     // [L6]: frame index of a token (block-causal mask), 0 when unmasked
    const float* base = p.qkv + (size_t)seq * L * p.ld + h * dh;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int dq = dh >> 2;
    for (int j = tid; j < L6; j += 128) frame[j] = p.spf > 0 ? j / p.spf : 0;
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");     // PDL (see tc_common.cuh): the next kernel's prologue may start
    asm volatile("griddepcontrol.wait;\n" ::: "memory");                  // qkv comes from the previous kernel
    for (int idx0 = tid; idx0 < L6 * dq; idx0 += 128 * 4) {
        float4 q[4], k[4], v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                          // 12 independent 16-byte loads in flight per thread
            const int idx = idx0 + u * 128;
            const int j = idx / dq, d4 = (idx - j * dq) * 4;
            q[u] = k[u] = v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < L6 * dq && j < L) {
                const float* row = base + (size_t)j * p.ld + d4;
                q[u] = __ldg(reinterpret_cast<const float4*>(row));
                k[u] = __ldg(reinterpret_cast<const float4*>(row + inner));
                v[u] = __ldg(reinterpret_cast<const float4*>(row + 2 * inner));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * 128;
            const int j = idx / dq, d4 = (idx - j * dq) * 4;
            if (idx < L6 * dq) {
                *reinterpret_cast<float4*>(Qs + j * pitch + d4) = q[u];
                *reinterpret_cast<float4*>(Ks + j * pitch + d4) = k[u];
                if (j < L) *reinterpret_cast<float4*>(Vs + j * dh + d4) = v[u];
            }
        }
    }
    __syncthreads();
    {
        const int nTy = L6 / MHA_RT, nTx = L6 / MHA_CT;
        for (int t = tid; t < nTy * nTx; t += 128) {
            const int ty = t / nTx, tx = t - ty * nTx;
            // packed fp32x2 FMAs (sm_100 FFMA2): every accumulator is a pair of partial sums over even / odd feature pairs
            float2 acc2[MHA_RT][MHA_CT];
#pragma unroll
            for (int r = 0; r < MHA_RT; ++r)
#pragma unroll
                for (int c = 0; c < MHA_CT; ++c) acc2[r][c] = make_float2(0.f, 0.f);
            const float* qr = Qs + ty * pitch;
            const float* kr = Ks + tx * pitch;
            const int qstep = nTy * pitch, kstep = nTx * pitch;
#pragma unroll 2
            for (int d = 0; d < dh; d += 4) {
                float4 q[MHA_RT], k[MHA_CT];
#pragma unroll
                for (int r = 0; r < MHA_RT; ++r) q[r] = *reinterpret_cast<const float4*>(qr + r * qstep + d);
#pragma unroll
                for (int c = 0; c < MHA_CT; ++c) k[c] = *reinterpret_cast<const float4*>(kr + c * kstep + d);
#pragma unroll
                for (int r = 0; r < MHA_RT; ++r)
#pragma unroll
                    for (int c = 0; c < MHA_CT; ++c) {
                        acc2[r][c] = __ffma2_rn(make_float2(q[r].x, q[r].y), make_float2(k[c].x, k[c].y), acc2[r][c]);
                        acc2[r][c] = __ffma2_rn(make_float2(q[r].z, q[r].w), make_float2(k[c].z, k[c].w), acc2[r][c]);
                    }
            }
            float acc[MHA_RT][MHA_CT];
#pragma unroll
            for (int r = 0; r < MHA_RT; ++r)
#pragma unroll
                for (int c = 0; c < MHA_CT; ++c) acc[r][c] = acc2[r][c].x + acc2[r][c].y;
#pragma unroll
            for (int r = 0; r < MHA_RT; ++r) {
                const int i = ty + nTy * r;
                const int fi = frame[i];
#pragma unroll
                for (int c = 0; c < MHA_CT; ++c) {
                    const int j = tx + nTx * c;
                    if (i < L && j < L)
                        Ss[i * (L + 1) + j] = frame[j] > fi ? -1e9f : acc[r][c] * p.scale;   // masked_fill(mask == 0, -1e9)
                }
            }
        }
    }
    __syncthreads();
    float* Pt = Qs;                            // [L][Lt]: Pt[j][i], columns i >= L zero
    // softmax over the keys of every query row: four lanes per row (two shuffle steps per reduction), 32 rows per pass; the
    // probabilities are written transposed, consecutive rows to consecutive addresses
    {
        const int l4 = tid & 3;
        for (int i0 = 0; i0 < Lt; i0 += 32) {               // uniform trip count: every lane takes part in the shuffles
            const int i = i0 + (tid >> 2);
            const bool valid = i < L;
            const float* srow = Ss + (valid ? i : 0) * (L + 1);
            float mx = -INFINITY;
            for (int j = l4; j < L; j += 4) mx = fmaxf(mx, srow[j]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float sum = 0.f;
            for (int j = l4; j < L; j += 4) sum += expf(srow[j] - mx);
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float inv = 1.0f / sum;
            if (i < Lt)
                for (int j = l4; j < L; j += 4) Pt[j * Lt + i] = valid ? expf(srow[j] - mx) * inv : 0.f;   // rows >= L: zero padding
        }
    }
    __syncthreads();
    {
        const int nTy = L6 / 6;
        for (int t = tid; t < nTy * dq; t += 128) {
            const int ty = t / dq, tx = t - ty * dq;
            // packed fp32x2 FMAs over row pairs: accp[rp][c] = (rows 2rp, 2rp+1) of output column c
            float2 accp[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) accp[r][c] = make_float2(0.f, 0.f);
            const float* pr = Pt + ty * 6;
            const float* vr = Vs + tx * 4;
#pragma unroll 4
            for (int j = 0; j < L; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(vr + j * dh);
                const float2 pp[3] = {*reinterpret_cast<const float2*>(pr + j * Lt), *reinterpret_cast<const float2*>(pr + j * Lt + 2),
                                      *reinterpret_cast<const float2*>(pr + j * Lt + 4)};
                const float2 vx = make_float2(v.x, v.x), vy = make_float2(v.y, v.y), vz = make_float2(v.z, v.z), vw = make_float2(v.w, v.w);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    accp[r][0] = __ffma2_rn(pp[r], vx, accp[r][0]); accp[r][1] = __ffma2_rn(pp[r], vy, accp[r][1]);
                    accp[r][2] = __ffma2_rn(pp[r], vz, accp[r][2]); accp[r][3] = __ffma2_rn(pp[r], vw, accp[r][3]);
                }
            }
            float acc[6][4];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) { acc[2 * r][c] = accp[r][c].x; acc[2 * r + 1][c] = accp[r][c].y; }
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                const int i = ty * 6 + r;
                if (i < L) {
                    const size_t oi = ((size_t)seq * L + i) * p.ldo + h * dh + tx * 4;
                    float4 o = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                    if (p.res) { const float4 a = *reinterpret_cast<const float4*>(p.res + oi); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
                    if (p.out_hi) {
                        const __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
                        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                        const __half2 l0 = __floats2half2_rn((o.x - f0.x) * 2048.0f, (o.y - f0.y) * 2048.0f);
                        const __half2 l1 = __floats2half2_rn((o.z - f1.x) * 2048.0f, (o.w - f1.y) * 2048.0f);
                        *reinterpret_cast<uint2*>(p.out_hi + oi) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                        *reinterpret_cast<uint2*>(p.out_lo + oi) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
                    } else
                    *reinterpret_cast<float4*>(p.out + oi) = o;
                }
            }
        }
    }
}

static int mha_run(const float* qkv, int ld, const float* res, float* out, void* out_planes, long long plane_rows, int ldo,
                   int n_seq, int L, int n_heads, int head_dim, int slots_per_frame, void* stream) {
    PS_REQUIRE(qkv && (out || out_planes), "ps_mha: null pointer");
    PS_REQUIRE(n_seq >= 0 && L > 0 && n_heads > 0 && head_dim > 0, "ps_mha: bad shape");
    if (n_seq == 0) return PS_OK;
    PS_REQUIRE(ld >= 3 * n_heads * head_dim && ldo >= n_heads * head_dim, "ps_mha: leading dimensions too small");
    MhaParams p;
    p.qkv = qkv; p.ld = ld; p.res = res; p.out = out; p.ldo = ldo;
    p.out_hi = reinterpret_cast<uint16_t*>(out_planes);
    p.out_lo = p.out_hi ? p.out_hi + (size_t)plane_rows * ldo : nullptr; p.n_seq = n_seq; p.L = L; p.H = n_heads; p.dh = head_dim;
    p.spf = slots_per_frame; p.scale = 1.0f / sqrtf((float)head_dim); p.rows_per_chunk = L;
    PS_REQUIRE(n_heads <= 65535, "ps_mha: too many heads");
    const bool vec = head_dim % 4 == 0 && ld % 4 == 0 && ldo % 4 == 0 && ps_aligned16(qkv) && (out_planes ? ps_aligned16(out_planes) : ps_aligned16(out)) &&
                     (res == nullptr || ps_aligned16(res));
    const int L6 = (L + 5) / 6 * 6, Lt = (L6 + 3) & ~3, pitch = head_dim + 4;
    const int qreg = L6 * pitch > L * Lt ? L6 * pitch : L * Lt;
    const size_t smem_tiled = (size_t)(qreg + L6 * pitch + L * head_dim + L * (L + 1) + L6) * sizeof(float);
    // long sequences (the L x L score tile does not fit) and very short ones (a handful of tiles) take the row-streaming kernel
    PS_REQUIRE(!out_planes || (vec && L >= 12 && smem_tiled <= 200 * 1024 && plane_rows >= (long long)n_seq * L),
               "ps_mha_planes: plane output needs head_dim %% 4 == 0, 12 <= L and a sequence that fits the tiled kernel");
    if (vec && L >= 12 && smem_tiled <= 200 * 1024) {
        const size_t smem = smem_tiled;
        static PsSmemConfig smem_cfg = {};
        if (smem > 48 * 1024) PS_CUDA(ps_ensure_smem(mha_tiled_kernel, smem, smem_cfg, 200 * 1024));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_seq, n_heads); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        PS_CUDA(cudaLaunchKernelEx(&cfg, mha_tiled_kernel, p));
    } else {
        size_t smem = (size_t)(L * (head_dim + 1) + L * head_dim + 4 + 16 * L + 16 * head_dim) * sizeof(float);
        PS_REQUIRE(smem <= 200 * 1024, "ps_mha: sequence of %d tokens x head_dim %d does not fit in shared memory", L, head_dim);
        static PsSmemConfig smem_cfg = {};
        if (smem > 48 * 1024) PS_CUDA(ps_ensure_smem(mha_kernel, smem, smem_cfg, 200 * 1024));
        // few (sequence, head) pairs with many rows each: split the query rows so that about two CTAs per SM exist
        int chunks = 1;
        const long long pairs = (long long)n_seq * n_heads;
        if (L >= 32 && pairs < 148) {
            chunks = (int)((2 * 148 + pairs - 1) / pairs);
            if (chunks > (L + 15) / 16) chunks = (L + 15) / 16;
            if (chunks < 1) chunks = 1;
        }
        p.rows_per_chunk = ((L + chunks - 1) / chunks + 3) / 4 * 4;
        chunks = (L + p.rows_per_chunk - 1) / p.rows_per_chunk;
        mha_kernel<<<dim3(n_seq, n_heads, chunks), 128, smem, (cudaStream_t)stream>>>(p);
    }
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_mha(const float* qkv, int ld, const float* res, float* out, int ldo, int n_seq, int L, int n_heads,
                      int head_dim, int slots_per_frame, void* stream) {
    PS_REQUIRE(out, "ps_mha: null pointer");
    return mha_run(qkv, ld, res, out, nullptr, 0, ldo, n_seq, L, n_heads, head_dim, slots_per_frame, stream);
}

// Same attention, result written as fp16 [hi | lo] planes [2][plane_rows][ldo] (lo scaled by 2048): the operand format
// ps_linear_tc_planes loads by TMA, so the output projection does no conversion work.
extern "C" int ps_mha_planes(const float* qkv, int ld, const float* res, void* out_planes, long long plane_rows, int ldo, int n_seq,
                             int L, int n_heads, int head_dim, int slots_per_frame, void* stream) {
    PS_REQUIRE(out_planes, "ps_mha_planes: null pointer");
    return mha_run(qkv, ld, res, nullptr, out_planes, plane_rows, ldo, n_seq, L, n_heads, head_dim, slots_per_frame, stream);
}

// ------------------------------------------------------------------------------------------------
// direct convolution (stride 1, zero "same" padding), fp32
// ------------------------------------------------------------------------------------------------
struct ConvParams {
    const float* x;        // [n][Cin][H][W]
    const uint8_t* x8;     // uint8 images [n][Cin][H][W] instead of x: value / 255 (torchvision ToTensor, data/load_data.py) on the fly
    const float* wp;       // packed [Cin][KS*KS][Cout]
    const float* bias;     // [Cout] or null
    const float* addend;   // NHWC addend [H*W][Cout] (only with out_nhwc) or null
    float* y;              // [n][Cout][H][W] or [n][H][W][Cout]
    int n, Cin, Cout, H, W;
    int relu, out_nhwc;
    uint16_t* split_hi;    // out_nhwc == 2: zero-bordered NHWC 16-bit planes of the tensor-core convolutions (Cout == 64)
    uint16_t* split_lo;    // null in the single-product modes
    int split_bf16;
};

#define CV_T 16        // output tile is CV_T x CV_T pixels
#define CV_CI 8        // input channels staged per step
#define CV_HS 20       // shared-memory row stride of the halo tile (floats)

template <int KS, int COUT_T>
__global__ __launch_bounds__(256) void conv2d_kernel(ConvParams p) {
    constexpr int PAD = KS / 2;
    constexpr int HH = CV_T + KS - 1;
    constexpr int KK = KS * KS;
    constexpr int COUT = 8 * COUT_T;
    extern __shared__ __align__(16) float sm[];
    float* s_in = sm;                               // [CV_CI][HH][CV_HS]
    float* s_w = sm + CV_CI * HH * CV_HS;           // [CV_CI][KK][COUT]
    const int tid = threadIdx.x, lane = tid & 31, cg = tid >> 5;
    const int tiles_x = p.W / CV_T;
    const int ty0 = (blockIdx.x / tiles_x) * CV_T, tx0 = (blockIdx.x % tiles_x) * CV_T;
    const int img = blockIdx.y;
    const int py = lane >> 1, px0 = (lane & 1) * 8;
    const float* xin = p.x ? p.x + (size_t)img * p.Cin * p.H * p.W : nullptr;
    const uint8_t* xin8 = p.x8 ? p.x8 + (size_t)img * p.Cin * p.H * p.W : nullptr;

    float acc[8][COUT_T];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < COUT_T; ++j) acc[i][j] = 0.f;

    for (int c0 = 0; c0 < p.Cin; c0 += CV_CI) {
        const int cc = min(CV_CI, p.Cin - c0);
        // stage the input halo tile (zero padded)
        for (int idx = tid; idx < cc * HH * CV_HS; idx += 256) {
            const int ci = idx / (HH * CV_HS);
            const int r = idx - ci * (HH * CV_HS);
            const int hy = r / CV_HS, hx = r - hy * CV_HS;
            const int gy = ty0 + hy - PAD, gx = tx0 + hx - PAD;
            float v = 0.f;
            if (hx < HH && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
                const size_t o = ((size_t)(c0 + ci) * p.H + gy) * p.W + gx;
                v = xin8 ? __fdiv_rn((float)xin8[o], 255.0f) : xin[o];       // IEEE division: bit-equal to ToTensor's x.div(255)
            }
            s_in[idx] = v;
        }
        // stage the weights of these channels: contiguous in the packed layout
        {
            const float4* src = reinterpret_cast<const float4*>(p.wp + (size_t)c0 * KK * COUT);
            float4* dst = reinterpret_cast<float4*>(s_w);
            for (int idx = tid; idx < cc * KK * COUT / 4; idx += 256) dst[idx] = src[idx];
        }
        __syncthreads();
        for (int ci = 0; ci < cc; ++ci) {
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const float* row = s_in + (ci * HH + py + ky) * CV_HS + px0;
                float in[12];
                *reinterpret_cast<float4*>(&in[0]) = *reinterpret_cast<const float4*>(row);
                *reinterpret_cast<float4*>(&in[4]) = *reinterpret_cast<const float4*>(row + 4);
                *reinterpret_cast<float4*>(&in[8]) = *reinterpret_cast<const float4*>(row + 8);
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const float* wr = s_w + (ci * KK + ky * KS + kx) * COUT + cg * COUT_T;
                    float wv[COUT_T];
#pragma unroll
                    for (int j = 0; j < COUT_T; j += 4)
                        *reinterpret_cast<float4*>(&wv[j]) = *reinterpret_cast<const float4*>(wr + j);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < COUT_T; ++j) acc[i][j] = fmaf(in[i + kx], wv[j], acc[i][j]);
                }
            }
        }
        __syncthreads();
    }

    const int oy = ty0 + py, ox = tx0 + px0;
#pragma unroll
    for (int j = 0; j < COUT_T; ++j) {
        const int co = cg * COUT_T + j;
        const float b = p.bias ? p.bias[co] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[i][j] + b;
            if (p.relu) v = fmaxf(v, 0.f);
            acc[i][j] = v;
        }
    }
    if (p.out_nhwc == 2) {
        if constexpr (COUT_T == 8) {
            const int Wp = p.W + 4, Hp = p.H + 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const size_t o = (((size_t)img * Hp + oy + 2) * Wp + ox + i + 2) * 64 + cg * 8;
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = acc[i][2 * j], b = acc[i][2 * j + 1];
                    if (p.split_bf16) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
                        hw[j] = *reinterpret_cast<const uint32_t*>(&h2); lw[j] = 0;
                    } else {
                        const __half2 h2 = __floats2half2_rn(a, b);
                        const float2 f = __half22float2(h2);
                        const __half2 l2 = __floats2half2_rn((a - f.x) * 2048.0f, (b - f.y) * 2048.0f);
                        hw[j] = *reinterpret_cast<const uint32_t*>(&h2); lw[j] = *reinterpret_cast<const uint32_t*>(&l2);
                    }
                }
                *reinterpret_cast<uint4*>(p.split_hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                if (p.split_lo) *reinterpret_cast<uint4*>(p.split_lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
        }
    } else if (!p.out_nhwc) {
#pragma unroll
        for (int j = 0; j < COUT_T; ++j) {
            const int co = cg * COUT_T + j;
            float* dst = p.y + (((size_t)img * COUT + co) * p.H + oy) * p.W + ox;
            *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const size_t pix = (size_t)oy * p.W + ox + i;
            float* dst = p.y + ((size_t)img * p.H * p.W + pix) * COUT + cg * COUT_T;
            const float* ad = p.addend ? p.addend + pix * COUT + cg * COUT_T : nullptr;
#pragma unroll
            for (int j = 0; j < COUT_T; j += 4) {
                float4 v = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
                if (ad) {
                    float4 a = *reinterpret_cast<const float4*>(ad + j);
                    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                }
                *reinterpret_cast<float4*>(dst + j) = v;
            }
        }
    }
}

template <int KS, int COUT_T>
static int launch_conv(const ConvParams& p, cudaStream_t st) {
    constexpr int HH = CV_T + KS - 1;
    const size_t smem = (size_t)(CV_CI * HH * CV_HS + CV_CI * KS * KS * 8 * COUT_T) * sizeof(float);
    static PsSmemConfig smem_cfg = {};
    PS_CUDA(ps_ensure_smem(conv2d_kernel<KS, COUT_T>, smem, smem_cfg));
    dim3 grid((p.H / CV_T) * (p.W / CV_T), p.n);
    conv2d_kernel<KS, COUT_T><<<grid, 256, smem, st>>>(p);
    PS_COUNT_LAUNCH();
    PS_LAUNCH_CHECK();
    return PS_OK;
}

extern "C" int ps_conv2d(const float* x, const float* w_packed, const float* bias, const float* addend_nhwc,
                         float* y, int n_img, int Cin, int Cout, int H, int W, int ksize, int relu,
                         int out_nhwc, void* stream) {
    PS_REQUIRE(x && w_packed && y, "ps_conv2d: null pointer");
    PS_REQUIRE(n_img >= 0 && Cin > 0, "ps_conv2d: bad shape");
    if (n_img == 0) return PS_OK;
    PS_REQUIRE(H % CV_T == 0 && W % CV_T == 0, "ps_conv2d: H and W must be multiples of %d (got %dx%d)", CV_T, H, W);
    PS_REQUIRE(n_img <= 65535, "ps_conv2d: at most 65535 images per launch (got %d)", n_img);
    PS_REQUIRE(ps_aligned16(x) && ps_aligned16(w_packed) && ps_aligned16(y), "ps_conv2d: pointers must be 16-byte aligned");
    PS_REQUIRE(!addend_nhwc || out_nhwc, "ps_conv2d: the addend is only supported with NHWC output");
    ConvParams p;
    p.x = x; p.x8 = nullptr; p.wp = w_packed; p.bias = bias; p.addend = addend_nhwc; p.y = y;
    p.n = n_img; p.Cin = Cin; p.Cout = Cout; p.H = H; p.W = W; p.relu = relu; p.out_nhwc = out_nhwc ? 1 : 0;
    p.split_hi = nullptr; p.split_lo = nullptr; p.split_bf16 = 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (ksize == 5 && Cout == 64) return launch_conv<5, 8>(p, st);
    if (ksize == 5 && Cout == 32) return launch_conv<5, 4>(p, st);
    if (ksize == 3 && Cout == 64) return launch_conv<3, 8>(p, st);
    if (ksize == 3 && Cout == 32) return launch_conv<3, 4>(p, st);
    ps_set_error("ps_conv2d: unsupported ksize=%d Cout=%d (supported: ksize 3|5, Cout 32|64)", ksize, Cout);
    return PS_EUNSUPPORTED;
}

// ps_conv2d writing the zero-bordered NHWC split planes consumed by ps_conv5x5_tc (Cout == 64): the first encoder
// convolution feeds the tensor-core layers without an fp32 NCHW round trip.
static int conv2d_split_run(const float* x, const uint8_t* x8, const float* w_packed, const float* bias, void* act_out, int n_img, int Cin,
                            int H, int W, int ksize, int relu, int precision, void* stream);

extern "C" int ps_conv2d_split(const float* x, const float* w_packed, const float* bias, void* act_out, int n_img, int Cin,
                               int H, int W, int ksize, int relu, int precision, void* stream) {
    PS_REQUIRE(x, "ps_conv2d_split: null pointer");
    return conv2d_split_run(x, nullptr, w_packed, bias, act_out, n_img, Cin, H, W, ksize, relu, precision, stream);
}

// the same with uint8 input images (0..255): the ToTensor normalisation x / 255 of the reference's data pipeline
// (src/data/load_data.py:86-91 DataLoader -> .to(device), 05_evaluate_PlaySlot.py:75) is applied while the halo tile is staged,
// so a clip travels over PCIe / through HBM as 1 byte per value instead of 4
extern "C" int ps_conv2d_split_u8(const unsigned char* x8, const float* w_packed, const float* bias, void* act_out, int n_img, int Cin,
                                  int H, int W, int ksize, int relu, int precision, void* stream) {
    PS_REQUIRE(x8, "ps_conv2d_split_u8: null pointer");
    return conv2d_split_run(nullptr, x8, w_packed, bias, act_out, n_img, Cin, H, W, ksize, relu, precision, stream);
}

static int conv2d_split_run(const float* x, const uint8_t* x8, const float* w_packed, const float* bias, void* act_out, int n_img, int Cin,
                            int H, int W, int ksize, int relu, int precision, void* stream) {
    PS_REQUIRE(w_packed && act_out, "ps_conv2d_split: null pointer");
    PS_REQUIRE(precision >= 1 && precision <= 3, "ps_conv2d_split: precision must be 1, 2 or 3");
    PS_REQUIRE(n_img >= 0 && Cin > 0, "ps_conv2d_split: bad shape");
    if (n_img == 0) return PS_OK;
    PS_REQUIRE(H % CV_T == 0 && W % CV_T == 0, "ps_conv2d_split: H and W must be multiples of %d (got %dx%d)", CV_T, H, W);
    PS_REQUIRE(n_img <= 65535, "ps_conv2d_split: at most 65535 images per launch (got %d)", n_img);
    PS_REQUIRE((!x || ps_aligned16(x)) && ps_aligned16(w_packed) && ps_aligned16(act_out), "ps_conv2d_split: pointers must be 16-byte aligned");
    ConvParams p;
    p.x = x; p.x8 = x8; p.wp = w_packed; p.bias = bias; p.addend = nullptr; p.y = nullptr;
    p.n = n_img; p.Cin = Cin; p.Cout = 64; p.H = H; p.W = W; p.relu = relu; p.out_nhwc = 2;
    const long long P = (long long)n_img * (H + 4) * (W + 4);
    p.split_hi = reinterpret_cast<uint16_t*>(act_out);
    p.split_lo = precision == 3 ? p.split_hi + P * 64 : nullptr;
    p.split_bf16 = precision == 1 ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (ksize == 5) return launch_conv<5, 8>(p, st);
    if (ksize == 3) return launch_conv<3, 8>(p, st);
    ps_set_error("ps_conv2d_split: unsupported ksize=%d (supported: 3, 5)", ksize);
    return PS_EUNSUPPORTED;
}
