/*
 * playslot_b200 -- C ABI of the B200 (sm_100a) kernels behind PlaySlot's per-frame slot pipeline.
 *
 * The reference (chrono-motz/PlaySlot) is pure PyTorch and has no FFI of its own; the drop-in boundary is
 * its `src/models` nn.Module surface.  The functions below are what the Python
 * mirror of that surface (playslot_b200/models) binds through ctypes: one entry point per fused device
 * op, each replacing the chain of ATen calls cited next to it (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers to contiguous fp32 (`float`) data unless stated otherwise;
 *     tensors are row-major with the last listed axis contiguous;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); every call is
 *     stream-ordered, asynchronous, allocates nothing and keeps no pointer after it returns;
 *   - return value: 0 = ok, negative = invalid argument / unsupported shape, positive = cudaError_t;
 *     `ps_last_error()` returns a thread-local human readable message; nothing throws;
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef PLAYSLOT_B200_H
#define PLAYSLOT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PLAYSLOT_B200_VERSION 100   /* 0.1.0 */

/* ---- plumbing ---------------------------------------------------------------------------------------- */
int ps_version(void);
const char* ps_last_error(void);
unsigned long long ps_launch_count(void);      /* kernels launched through this library by this process */
void ps_reset_launch_count(void);
int ps_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* SMs the persistent kernels (ps_conv5x5_tc, ps_dec_composite_tc) leave free for kernels of a concurrent stream: the decoder of
 * predicted frame i (side stream) overlaps the cOCVP step of frame i+1 (src/05_evaluate_PlaySlot.py:112-139 runs them back to back) */
int ps_set_sm_reserve(int n_sms);
int ps_get_sm_reserve(void);

/* ---- dense building blocks ---------------------------------------------------------------------------- */

/* Y[M,N] = act( LN_eps(X)[M,K] * W[N,K]^T + bias ) + res[(m / res_div) % res_mod, :]
 * Replaces nn.LayerNorm + nn.Linear + activation + residual chains, e.g. the encoder MLP
 * (src/models/SAVi.py:111-116,214-216), transformer-block MLPs and q/k/v/out projections
 * (src/models/BlocksUtils/attention.py:296,462-475,539-563; model_blocks.py:293-309), the cOCVP token /
 * action / variability embedding with its sinusoidal time encoding (Predictors/DynamicsModels.py:119-148,
 * 487-497) and regression head (:85-88,510).
 * ln_gamma/ln_beta may be NULL (no LayerNorm); bias, res may be NULL.
 * act: 0 none, 1 ReLU, 2 GELU (exact erf), 3 |x|.  K, ldx multiples of 4; X, W 16-byte aligned. */
int ps_linear(const float* X, int ldx, const float* W, const float* bias,
              const float* ln_gamma, const float* ln_beta, float ln_eps,
              const float* res, int ldres, int res_div, int res_mod,
              float* Y, int ldy, int M, int N, int K, int act, void* stream);

/* ps_linear on the tensor cores (tcgen05 / TMEM, weights by TMA, activations split to fp16 hi/lo in the
 * producer warps).  w_tc: 16-bit [2 (hi, lo)][Npad][Kp] holding W * w_scale (Kp = K rounded up to 64, Npad = N
 * rounded up to 128, zero padded; w_scale a power of two chosen by the packer, undone in the epilogue).
 * precision 3 ("fp16x3"): three products, fp32-grade result; 2 ("fp16") / 1 ("bf16"): one product.
 * Same reference call sites as ps_linear. */
int ps_linear_tc(const float* X, int ldx, const void* w_tc, float w_scale, int Kp, int Npad, const float* bias,
                 const float* ln_gamma, const float* ln_beta, float ln_eps,
                 const float* res, int ldres, int res_div, int res_mod,
                 float* Y, int ldy, int M, int N, int K, int act, int precision, void* stream);

/* ps_linear_tc (fp16x3) with fp16 [hi | lo] plane operands on either side, for chains of tensor-core linears (MLP fc1 ->
 * fc2, attention -> output projection): x_planes [2][x_plane_rows][ldx] (lo scaled by 2048) replaces X and is loaded by TMA
 * straight into the operand tiles (no conversion work in the kernel, no LayerNorm prologue); y_planes
 * [2][y_plane_rows][ldy] replaces Y (the epilogue writes the split of the fp32 result; N multiple of 32).  Exactly one of
 * X / x_planes and one of Y / y_planes is non-NULL.  The values are those ps_linear_tc would produce. */
int ps_linear_tc_planes(const float* X, int ldx, const void* x_planes, long long x_plane_rows, const void* w_tc, float w_scale,
                        int Kp, int Npad, const float* bias, const float* ln_gamma, const float* ln_beta, float ln_eps,
                        const float* res, int ldres, int res_div, int res_mod, float* Y, void* y_planes, long long y_plane_rows,
                        int ldy, int M, int N, int K, int act, void* stream);

/* Multi-head self attention over n_seq independent sequences of L tokens.
 * qkv: [n_seq*L, ld] rows holding q | k | v (each n_heads*head_dim wide, head-major); out: [n_seq*L, ldo].
 * slots_per_frame > 0 applies the block-causal (time x slot staircase) mask: token j is visible to token i
 * iff j / slots_per_frame <= i / slots_per_frame; masked logits are set to -1e9 as the reference does.
 * Replaces MetaAttention.attention / split_into_heads / merge_heads
 * (src/models/BlocksUtils/attention.py:204-263,288-316) and build_slot_causal_mask
 * (BlocksUtils/model_utils.py:115-124). */
int ps_mha(const float* qkv, int ld, const float* res /* optional residual, laid out like out */, float* out,
           int ldo, int n_seq, int L, int n_heads, int head_dim, int slots_per_frame, void* stream);
/* same, result as fp16 [hi | lo] planes [2][plane_rows][ldo] for ps_linear_tc_planes (head_dim % 4 == 0, L >= 12) */
int ps_mha_planes(const float* qkv, int ld, const float* res, void* out_planes, long long plane_rows, int ldo, int n_seq,
                  int L, int n_heads, int head_dim, int slots_per_frame, void* stream);

/* KxK stride-1 zero-padded ("same") convolution + bias (+ReLU).  x: [n,Cin,H,W]; w_packed: [Cin][K*K][Cout]
 * (a repack of nn.Conv2d's [Cout,Cin,K,K]); y: [n,Cout,H,W], or with out_nhwc [n,H,W,Cout] (+ addend_nhwc
 * [H*W,Cout], used to add the soft position embedding).  ksize in {3,5}, Cout in {32,64}, H,W % 16 == 0.
 * Replaces ConvBlock / SimpleConvEncoder / Decoder convolutions (src/models/BlocksUtils/model_blocks.py:48-79,
 * EncodersDecoders/encoders.py:77-104, decoders.py:78-113) and the NHWC permute + SoftPositionEmbed add
 * (SAVi.py:210-212, model_blocks.py:110-119). */
int ps_conv2d(const float* x, const float* w_packed, const float* bias, const float* addend_nhwc, float* y,
              int n_img, int Cin, int Cout, int H, int W, int ksize, int relu, int out_nhwc, void* stream);

/* ps_conv2d with Cout == 64 writing the padded NHWC split planes of ps_conv5x5_tc directly (first encoder convolution). */
int ps_conv2d_split(const float* x, const float* w_packed, const float* bias, void* act_out, int n_img, int Cin, int H, int W,
                    int ksize, int relu, int precision, void* stream);
/* ps_conv2d_split on uint8 images (0..255): x / 255 -- the ToTensor normalisation of the reference's data pipeline
 * (src/data/load_data.py:86-91, consumed at src/05_evaluate_PlaySlot.py:75) -- fused into the first encoder convolution */
int ps_conv2d_split_u8(const unsigned char* x8, const float* w_packed, const float* bias, void* act_out, int n_img, int Cin,
                       int H, int W, int ksize, int relu, int precision, void* stream);

/* ---- SlotAttention (src/models/BlocksUtils/attention.py:76-119) ---------------------------------------- */

/* floats per (frame, chunk) of the partial buffer used between ps_sa_main and ps_sa_update */
int ps_sa_partial_floats(int num_slots);

/* One streaming pass over the feature tokens of every frame: LayerNorm statistics of the inputs
 * (norm_input, :92), logits against the folded queries A/a0/a1 (= scale * to_q(norm_slot(slots)) . to_k(.),
 * :100-106), softmax over the SLOT axis + eps (:107), optional store of that tensor as the attention masks
 * [B,S,N] (:108), and the un-normalised weighted sums needed for `updates` (:109-110), written as
 * per-chunk partials [B][n_chunks][ps_sa_partial_floats(S)]. */
int ps_sa_main(const float* x, const float* A, const float* a0, const float* a1, float* partial,
               float* masks /* may be NULL */, int B, int N, int S, int dim, int n_chunks, float ln_eps,
               float attn_eps, void* stream);

/* ps_sa_main on the tensor cores: both contractions as tcgen05.mma over fp16 hi/lo split operands (fp32-grade),
 * the token tile read once from HBM, normalised and kept in shared memory as the K-major A operand of the
 * logits GEMM and, untransposed, as the MN-major A operand of the weighted-sum GEMM.  Same outputs. */
int ps_sa_main_tc(const float* x, const float* A, const float* a0, float* partial, float* masks /* may be NULL */,
                  int B, int N, int S, int dim, int n_chunks, float ln_eps, float attn_eps, void* stream);

/* Pre-normalised feature planes: c = (x - mean) * rstd of norm_input (no affine part) as two fp16 planes
 * [hi | lo][n_tokens][128] (lo scaled by 2048; 512 bytes per token, like the fp32 input).  Computed once per frame;
 * every SlotAttention iteration then streams the planes with TMA (ps_sa_main_tc2) and does no CUDA-core work on the
 * features: TMA -> tcgen05 (logits) -> per-token softmax out of TMEM -> tcgen05 (weighted sums).  Persistent over
 * (frame, chunk) items, 3-stage TMA ring, double-buffered logits / weights / accumulators in TMEM.
 * lo_offset_tokens: distance (in tokens) from the first hi row of this batch to its first lo row (>= B*N), so that a
 * batch may be a slice of planes prepared for many frames at once. */
int ps_sa_prepare(const float* x, void* planes, long long n_tokens, int dim, float ln_eps, void* stream);
/* partial slots per frame ps_sa_main_tc2 needs for this shape (its n_chunks argument / the CH extent of `partial`): the
 * persistent CTAs take equally long contiguous ranges of the B * ceil(N/128) token tiles, so a frame is shared by a few CTAs */
int ps_sa_tc2_chunks(int B, int N);
int ps_sa_main_tc2(const void* planes, long long lo_offset_tokens, const float* A, const float* a0, float* partial,
                   float* masks /* may be NULL */, int B, int N, int S, int dim, int n_chunks, float attn_eps, void* stream);

/* Fused SAVi encoder MLP + norm_input for the tensor-core SlotAttention pass (src/models/SAVi.py:111-116,214-216 and
 * the statistics part of BlocksUtils/attention.py:57,88):
 *     h = relu(LayerNorm(x; ln_g, ln_b, ln_eps) W1^T + b1);  y = h W2^T + b2;  planes = split((y - mean y) * rstd(y; out_eps))
 * One persistent tcgen05 kernel: tokens [n_tokens][ldx] fp32 (K1 <= 64 features used) are read once, the planes
 * [hi | lo][n_tokens][128] (same format as ps_sa_prepare) are written once; hidden and output activations never touch HBM.
 * w1_packed / w2_packed: fp16 hi/lo copies [2][128][64] / [2][128][128] of W1 (128 x K1, K zero-padded) and W2 (128 x 128),
 * each multiplied by the power of two w*_scale (the layout ps_linear_tc takes). */
int ps_enc_mlp_planes(const float* x, int ldx, int K1, const float* ln_g, const float* ln_b, float ln_eps,
                      const void* w1_packed, float w1_scale, const float* b1, const void* w2_packed, float w2_scale,
                      const float* b2, float out_eps, void* planes, long long n_tokens, void* stream);
/* The same for a 32-channel encoder convolved in channel pairs (two images share the 64 channels of ps_conv5x5_tc): input row
 * (pair * pair_hw + position) * 2 + image, output planes in image order.  pair_hw = H * W of one image (0: plain). */
int ps_enc_mlp_planes_paired(const float* x, int ldx, int K1, const float* ln_g, const float* ln_b, float ln_eps,
                             const void* w1_packed, float w1_scale, const float* b1, const void* w2_packed, float w2_scale,
                             const float* b2, float out_eps, void* planes, long long n_tokens, int pair_hw, void* stream);

/* Per-slot update: reduce the partials, renormalise over locations (:109), to_v (:94), nn.GRUCell (:112-115),
 * residual MLP (:117), then (do_prep) the folded queries of the next iteration.  With do_update = 0 it only
 * prepares A/a0/a1 from slots_in.  Weight matrices suffixed T are transposed to [in][out]. */
int ps_sa_update(const float* partial, int n_chunks, const float* slots_in, float* slots_out,
                 const float* ln_in_g, const float* ln_in_b, const float* WvT,
                 const float* WihT, const float* WhhT, const float* b_ih, const float* b_hh,
                 const float* ln_mlp_g, const float* ln_mlp_b,
                 const float* W1T, const float* b1, const float* W2T, const float* b2,
                 const float* ln_slot_g, const float* ln_slot_b, const float* WqT, const float* Wk,
                 float* A, float* a0, float* a1,
                 int B, int S, int dim, int mlp_hidden, float scale, float ln_eps,
                 int do_update, int do_prep, void* stream);

/* ps_sa_update with two weight-only re-associations prepared by the caller (W_iv = W_ih W_v; M_A = scale *
 * diag(gamma_in) W_k^T W_q and v0 = scale * beta_in^T W_k^T W_q, stored transposed [in][out]) and every mat-vec
 * split over K so that a stage costs about one L2 round trip.  mlp_hidden in {128, 256, 384}. */
int ps_sa_update2(const float* partial, int n_chunks, const float* slots_in, float* slots_out,
                  const float* ln_in_g, const float* ln_in_b,
                  const float* WivT, const float* WhhT, const float* b_ih, const float* b_hh,
                  const float* ln_mlp_g, const float* ln_mlp_b,
                  const float* W1T, const float* b1, const float* W2T, const float* b2,
                  const float* ln_slot_g, const float* ln_slot_b, const float* MAT, const float* v0,
                  float* A, float* a0, int B, int S, int dim, int mlp_hidden, float ln_eps,
                  int do_update, int do_prep, void* stream);

/* ---- spatial-broadcast decoder (src/models/SAVi.py:220-248, EncodersDecoders/decoders.py:89-113) -------- */

/* h1[img,co,y,x] = relu(G[img, cls(y), cls(x), co] + P[co,y,x]): the first decoder convolution applied to
 * broadcast(slot) + position embedding, collapsed into (2*pad+1)^2 border classes (exact re-association). */
int ps_dec_expand(const float* G, const float* P, float* out, int n_img, int C, int H, int W, int ksize,
                  void* stream);

/* Final 3x3 conv (Cin -> RGB + alpha), softmax of alpha over the slot axis, recon = sum_s rgb_s * mask_s.
 * h: [M*S,Cin,H,W]; w_packed: [Cin][9][4]; recon [M,3,H,W]; recons [M,S,3,H,W] (un-masked); masks [M,S,1,H,W]. */
int ps_dec_composite(const float* h, const float* w_packed, const float* bias, float* recon, float* recons,
                     float* masks, int M, int S, int Cin, int H, int W, void* stream);

/* ---- InvDyn tail (src/models/Predictors/LatentAction.py:72-87,228-249,340-359; quantization.py:256-372) -- */

/* mv: [B,T,R,2A] = mean_fc | variance_fc outputs (R = 1 for InvDynS, num_slots for InvDynM); noise [B,T-1,R,A]
 * is drawn by the host exactly as the reference draws it.  Outputs: action_dist [B,T-1,2,R,A], sampled latent
 * z, prototype z + (z_q - z), variability, int64 code indices, per-vector mean_A (z_q - z)^2. */
int ps_invdyn_tail(const float* mv, const float* noise, const float* codebook, float* dist_out, float* z_out,
                   float* proto_out, float* var_out, long long* idx_out, float* loss_out,
                   int B, int T, int R, int A, int K, int per_slot, float eps, void* stream);

/* ---- tcgen05 (tensor-core) 5x5 convolution, 64 -> 64 channels ------------------------------------------- */

/* Activations for the tensor-core convolutions are zero-bordered NHWC 16-bit planes [hi | lo]:
 * plane[img][y+2][x+2][64] with (H+4) x (W+4) positions per image; ps_conv_tc_plane_elems gives the element count
 * of ONE plane (a buffer holds two planes; only the first is used by the single-product modes).  Buffers must
 * be zero-initialised once: the kernels keep the borders zero.
 * precision 3 ("fp16x3"): hi = fp16(x), lo = fp16((x - hi) * 2048), three products, fp32-grade result;
 * precision 2 ("fp16") / 1 ("bf16"): hi plane only, one product. */
long long ps_conv_tc_plane_elems(int n_img, int H, int W);

/* nn.Conv2d(64, 64, 5, padding=2) (+bias, +ReLU) as an implicit GEMM on tcgen05 / TMEM fed by TMA.
 * w_tc: 16-bit [25 taps][2 (hi, lo)][64 cout][64 cin] in the format of `precision`, holding w * w_scale
 * (w_scale: a power of two chosen by the packer so that the weights sit mid-range; undone in the epilogue).
 * out_mode 0: padded NHWC split into act_out;  1: fp32 NCHW [n,64,H,W] into out_f32;
 * 2: fp32 NHWC tokens [n,H*W,64] + addend [H*W,64] into out_f32 (encoder output + position embedding).
 * Replaces the 64-channel ConvBlocks of EncodersDecoders/decoders.py:89-102 and encoders.py:86-96. */
int ps_conv5x5_tc(const void* act_in, const void* w_tc, const float* bias, float w_scale, void* act_out,
                  float* out_f32, const float* addend, int n_img, int H, int W, int relu, int out_mode, int precision,
                  void* stream);

/* ps_dec_composite on the tensor cores: the 3x3 convolution as a 9-tap implicit GEMM over padded NHWC split planes
 * (act_in: the last 5x5 block's output for M*S slot images, precision 3), all S slots of a 128-position tile accumulated
 * side by side in TMEM, slot softmax and compositing in the epilogue.  The three taps of a filter row share one MMA (their
 * weights side by side in N); the epilogue adds the column-shifted partial sums of neighbouring positions.  w_tc: 16-bit
 * [3 filter rows][32 rows: hi (kx0 o0-3 | kx1 o0-3 | kx2 o0-3 | 4 zero), lo (same)][64 cin] holding w * w_scale. */
int ps_dec_composite_tc(const void* act_in, const void* w_tc, float w_scale, const float* bias, float* recon, float* recons,
                        float* masks, int M, int S, int H, int W, void* stream);

/* ps_dec_expand writing the padded NHWC split format (P_hwc: [H*W][64] = conv1(position embedding) + bias). */
int ps_dec_expand_split(const float* G, const float* P_hwc, void* act_out, int n_img, int H, int W, int ksize,
                        int precision, void* stream);

/* fp32 NCHW [n,64,H,W] -> padded NHWC split (feeds the first tensor-core layer after a CUDA-core layer). */
int ps_nchw_to_split(const float* x, void* act_out, int n_img, int H, int W, int precision, void* stream);

/* ---- training step: backward kernels (src/04_train_PlaySlot.py:32-122, base/basePredictorTrainer.py:320-333) ---- */

/* Un-fused LayerNorm of the training path (keeps mean / rstd [M]) and its backward (dgamma / dbeta [K] are overwritten). */
int ps_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y, float* mean, float* rstd,
                     int M, int K, void* stream);
int ps_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx,
                     float* dgamma, float* dbeta, int M, int K, void* stream);

/* Weight gradient of nn.Linear: dW[N,K] = dY[M,N]^T X[M,K], db[N] = column sums of dY (db may be NULL); outputs overwritten.
 * The data gradient dX = dY W is ps_linear with the transposed weight. */
int ps_linear_wgrad(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int N, int K, void* stream);
/* dW += dY^T X, db += colsum(dY): accumulating variant (outputs are NOT cleared) for weights used several times in one backward */
int ps_linear_wgrad_acc(const float* dY, int lddy, const float* X, int ldx, float* dW, float* db, int M, int N, int K, void* stream);

/* activation forward on saved pre-activations / backward; act: 1 ReLU, 2 GELU (exact erf), 3 |x| */
int ps_act_fwd(const float* z, float* y, long long n, int act, void* stream);
int ps_act_bwd(const float* dy, const float* z, float* dz, long long n, int act, void* stream);

/* Backward of ps_mha (no residual): dqkv laid out like qkv; sequences of at most 96 tokens x head_dim 64. */
int ps_mha_bwd(const float* qkv, int ld, const float* dout, int ldo, float* dqkv, int n_seq, int L, int n_heads, int head_dim,
               int slots_per_frame, void* stream);

/* Backward of ps_invdyn_tail w.r.t. mv: incoming gradients (any may be NULL) for action_dist, the sampled latent, the
 * straight-through prototype, the variability and the per-vector quantisation loss mean_A (sg(z_q) - z)^2. */
int ps_invdyn_tail_bwd(const float* mv, const float* noise, const float* z, const long long* idx, const float* codebook,
                       const float* g_dist, const float* g_z, const float* g_proto, const float* g_var, const float* g_quant,
                       float* d_mv, int B, int T, int R, int A, int K, int per_slot, float eps, void* stream);

/* Decoder data gradient (the decoder is frozen in 04_train_PlaySlot.py: no weight gradients).
 * ps_dec_composite_bwd: gradients of recon / recons / masks (any may be NULL) -> g_o [M*S,4,H,W], the gradient of the
 *   last convolution's RGB+alpha output (slot softmax + compositing, SAVi.py:226-234).
 * ps_dec_final_dgrad : data gradient of that 3x3 convolution (w_packed [64][9][4]) and of the ReLU in front of it
 *   (h = its fp32 NCHW input), written as padded NHWC split planes for ps_conv5x5_tc_dgrad.
 * ps_conv5x5_tc_dgrad: data gradient of Conv2d(64,64,5)+ReLU on tcgen05: w_tc_flipped packs W'[ci][co][ky][kx] =
 *   W[co][ci][4-ky][4-kx]; the result is masked by (mask_act > 0), mask_act = the block's saved input (padded split).
 * ps_dec_expand_bwd  : gradient of the collapsed first convolution: sums of g [n,H*W,64] over the (2*pad+1)^2 border
 *   classes -> dG [n, ksize*ksize*64] (then dSlots = dG * Wsum via ps_linear).
 * g_scale (device scalar, may be NULL): a power of two the gradients are multiplied by on entry (ps_dec_final_dgrad) and
 *   divided by on exit (ps_dec_expand_bwd) so that loss gradients of any magnitude stay inside the fp16 split's range. */
int ps_dec_composite_bwd(const float* recons, const float* masks, const float* g_recon, const float* g_recons,
                         const float* g_masks, float* g_o, int M, int S, int H, int W, void* stream);
int ps_dec_final_dgrad(const float* g_o, const float* w_packed, const float* h, void* act_out, const float* g_scale,
                       int n_img, int H, int W, int precision, void* stream);
/* ps_dec_final_dgrad with the saved activation as padded split planes (the hi plane carries the sign of the ReLU input) */
int ps_dec_final_dgrad_planes(const float* g_o, const float* w_packed, const void* h_planes, void* act_out, const float* g_scale,
                              int n_img, int H, int W, int precision, void* stream);
int ps_conv5x5_tc_dgrad(const void* g_in, const void* w_tc_flipped, float w_scale, const void* mask_act, void* act_out,
                        float* out_tokens, int n_img, int H, int W, int out_mode, int precision, void* stream);
int ps_dec_expand_bwd(const float* g_tokens, float* dG, const float* g_scale, int n_img, int H, int W, int ksize, void* stream);

/* Stand-alone nearest-codeword lookup (VectorQuantizer.forward, quantization.py:294-322): z [n,A] ->
 * proto = z + (z_q - z), exact codewords z_q, int64 indices, per-vector mean_A (z_q - z)^2. */
int ps_vq_lookup(const float* z, const float* codebook, float* proto_out, float* zq_out, long long* idx_out,
                 float* loss_out, int n, int A, int K, void* stream);

/* ---- fused transformer stack --------------------------------------------------------------------------- */

/* X (n_seq * L, 128) fp32, in place: n_layers pre-LN blocks  y = x + MHSA(LN(x)) W_o^T + b_o;  x = y + W_2 GELU(W_1 LN(y) + b_1) + b_2
 * in ONE launch (4-CTA clusters, tcgen05, activations on chip).  Replaces the per-block chains of TransformerDecoder /
 * TransformerEncoderBlock (src/models/BlocksUtils/attention.py:425-565, MetaAttention :204-316, model_blocks.MLP :278-309) as
 * stacked by the cOCVP predictor (Predictors/DynamicsModels.py:457-516) and SAVi's transition module (transition_models.py:19-53),
 * for embed_dim 128, 4 or 8 heads of width <= 64, MLP hidden 512, <= 128 tokens per sequence.  spf > 0: block-causal mask over
 * frames of spf tokens (model_utils.py:115-124).  Weight packing: see csrc/tblock_tc.cu / ops.pack_tblock_stack. */
int ps_tblock_stack(float* X, int n_seq, int L, int spf, int n_layers, int heads, float qscale, float eps1, float eps2,
                    const void* WA, const void* WB, const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b,
                    const float* b_o, const float* b1, const float* b2, const float* inv_scale, float* workspace, void* stream);
/* floats of scratch (the clusters' partial sums travel through L2) ps_tblock_stack needs for n_seq sequences of L tokens */
long long ps_tblock_stack_workspace_floats(int n_seq, int L);

/* development variant: dbg (device, >= 4096 u64) receives %globaltimer stamps of CTA 0 at the phase boundaries of its row warps */
int ps_tblock_stack_dbg(float* X, int n_seq, int L, int spf, int n_layers, int heads, float qscale, float eps1, float eps2,
                        const void* WA, const void* WB, const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b,
                        const float* b_o, const float* b1, const float* b2, const float* inv_scale, float* workspace,
                        unsigned long long* dbg, void* stream);

/* ---- fused loss / metric tail ------------------------------------------------------------------------- */

/* out[f] = sum_e (pred[f][e] - target[f][e])^2 per frame (optionally both clamped to [0, 1]), one pass, deterministic.
 * Replaces nn.MSELoss of PredImgMSELoss / PredSlotMSELoss (src/lib/loss.py:201-238) in the training step and the per-frame
 * PSNR input of the evaluation (src/05_evaluate_PlaySlot.py:98-108 -> lib/metrics.py:225-235 -> piqa.psnr), which the
 * reference computes on the CPU after copying every predicted frame to the host.  Strides in elements. */
int ps_frame_sq_err(const float* pred, const float* target, float* out, long long n_videos, int frames_per_video, long long elems,
                    long long pred_video_stride, long long pred_frame_stride, long long tgt_video_stride, long long tgt_frame_stride,
                    int clamp01, void* stream);
/* d_pred = 2 * scale * g[0] * (pred - target): backward of scale * sum (pred - target)^2 (loss.backward() through nn.MSELoss) */
int ps_mse_bwd(const float* pred, const float* target, const float* g, float scale, float* d_pred, long long n_videos,
               int frames_per_video, long long elems, long long tgt_video_stride, long long tgt_frame_stride, void* stream);

/* ---- SAVi training (src/02_train_savi.py:21-70): backward pieces of SlotAttention, encoder and decoder --------------------- */

/* Attention core of one SlotAttention iteration (src/models/BlocksUtils/attention.py:97-110) on folded operands:
 * A (B,S,128) = D^-1/2 * q W_k, xn (B,N,128) = norm_input(inputs)  ->  t (B,S,128) = sum_n w[s][n] xn[n] with
 * a = softmax_s(A xn^T) + eps, w = a / sum_n a;  Z (B,S) = sum_n a;  masks (B,S,N) = a (or null).
 * Workspaces: part_t (B, chunks, S, 128), part_z (B, chunks, S), part_dA (B, chunks, S, 128), chunks = ps_sa_core_chunks(N). */
int ps_sa_core_chunks(int N);
int ps_sa_core_fwd(const float* A, const float* xn, float* t, float* Z, float* masks, float* part_t, float* part_z,
                   int B, int S, int N, float eps, void* stream);
int ps_sa_core_bwd(const float* A, const float* xn, const float* t, const float* Z, const float* dt, float* dA, float* dxn,
                   float* part_dA, int B, int S, int N, float eps, int accumulate, void* stream);
/* element-wise part of nn.GRUCell (attention.py:112-115): gi = W_ih u + b_ih, gh = W_hh h + b_hh (M, 3D) -> h' (M, D); gates (M, 3D) saved */
int ps_gru_gates_fwd(const float* gi, const float* gh, const float* h, float* hout, float* gates, long long M, int D, void* stream);
int ps_gru_gates_bwd(const float* dhout, const float* gates, const float* gh, const float* h, float* dgi, float* dgh, float* dh,
                     long long M, int D, void* stream);
/* weight / bias gradients of the convolutions (ConvBlocks of encoders.py:46-104 / decoders.py:45-113) from the padded NHWC split
 * planes the tcgen05 convolutions read and write; outputs accumulate (+=) and must be zeroed by the caller */
int ps_conv5x5_planes_wgrad(const void* dy_planes, const void* x_planes, float* dW, int n_img, int H, int W, int dy_precision,
                            int x_precision, float scale, void* stream);
/* the same on tcgen05 (both tensors precision 3 = fp16 hi + lo planes; images up to 68 pixels wide): MN-major operands straight
 * from the position-major planes, TMEM accumulation over a chunk of positions, 7 tap groups x position chunks */
int ps_conv5x5_planes_wgrad_tc(const void* dy_planes, const void* x_planes, float* dW, int n_img, int H, int W, int precision,
                               float scale, void* stream);
int ps_planes_colsum(const void* planes, float* out, int n_img, int H, int W, int precision, float scale, void* stream);
int ps_conv_first_wgrad(const void* dh_planes, const float* x, const unsigned char* x8, float* dW, int n_img, int Cin, int H, int W,
                        int ksize, int precision, float scale, void* stream);
int ps_conv3x3_final_wgrad(const float* g_o, const float* h, float* dW, float* db, int n_img, int H, int W, void* stream);
/* ... with h as the padded split planes the tensor-core layers write (precision 2: hi plane only, 3: hi + lo) */
int ps_conv3x3_final_wgrad_planes(const float* g_o, const void* h_planes, int precision, float* dW, float* db, int n_img, int H, int W,
                                  void* stream);
int ps_colsum(const float* X, float* out, long long M, long long N, float scale, void* stream);
int ps_nhwc_to_split_scaled(const float* tokens, int ld, const float* scale, void* planes, int n_img, int H, int W, int precision,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PLAYSLOT_B200_H */
