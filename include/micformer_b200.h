/*
 * micformer_b200 -- C ABI of the B200 (sm_100a) kernels behind the MicFormer dual-stream hot path.
 *
 * The reference (fxxJuses/MICFormer) is pure PyTorch: it has NO FFI / plugin interface for this path
 * (SURVEY.md §8b).  The boundary a maintainer binds is therefore the set of ATen calls its Python makes;
 * every entry point below names the reference lines (relative to /root/reference/MicFormer) it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - all tensors are contiguous fp32 device pointers owned by the caller (torch's allocator); the library
 *     never allocates, never retains pointers and never synchronises; everything is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so calls are CUDA-graph capturable.
 *   - activations are channels-last token grids (B, D, H, W, C); "rows" means the flattened (B*D*H*W) axis.
 *   - return value 0 = success, negative = error (mic_last_error_string() gives the reason, thread-local).
 *     Nothing throws or exits across the ABI.
 *   - "+=" outputs are accumulated atomically into caller-initialised buffers.
 */
#ifndef MICFORMER_B200_H
#define MICFORMER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIC_OK 0
#define MIC_ERR_INVALID (-1)
#define MIC_ERR_UNSUPPORTED (-2)
#define MIC_ERR_CUDA (-3)

int mic_version(void);
const char* mic_last_error_string(void);
/* number of kernel launches issued by this library in the calling process since load / last reset */
int64_t mic_launch_count(void);
void mic_reset_launch_count(void);
/* 0 = fp32 CUDA-core GEMMs / convs / attention everywhere (exact parity path);
 * 1 = tcgen05 tensor-core kernels where shapes allow (fp32 in/out, fp32 accumulate in TMEM): forward GEMMs run the
 *     3xTF32 split (fp32-faithful), backward GEMMs and the 3x3x3 convs single-pass TF32 with operands rounded to nearest;
 *     GEMM shapes the tensor-core kernel declines run the fp32 CUDA-core kernel of mode 0 and are reported once per
 *     entry point on stderr (MICFORMER_WARN_FALLBACK=0 silences it);
 * 2 = accepted and treated as 0 (kept for ABI stability). */
int mic_set_gemm_mode(int mode);
int mic_get_gemm_mode(void);

/* ---- LayerNorm over the channel axis (nn.LayerNorm eps=1e-5; models/MICFormer_self.py:343,477,404,559,578,
 *      1011-1012,1034 and LayerNormProxy :263-273).  The input may be the channel-concatenation of two
 *      tensors (x0:C0 | x1:C1) -- this is `torch.cat([moving, fixed], -1)` + norm2 (:1033-1034) without the
 *      cat.  The output may be zero-padded on the trailing side of D/H/W to (Dp,Hp,Wp) -- this is F.pad after
 *      norm1 (:343-350, :477-483).  mean/rstd are saved per un-padded row for the backward. */
int mic_layernorm_fwd(const float* x0, int C0, const float* x1, int C1, const float* gamma, const float* beta,
                      float* y, float* mean, float* rstd, int B, int D, int H, int W, int Dp, int Hp, int Wp,
                      float eps, void* stream);
/* dx0 = (dres0 ? dres0 : 0) + LN'(dy) (same for dx1); dgamma/dbeta += column sums. dy is in the padded layout. */
int mic_layernorm_bwd(const float* dy, const float* x0, int C0, const float* x1, int C1, const float* gamma,
                      const float* mean, const float* rstd, const float* dres0, const float* dres1, float* dx0,
                      float* dx1, float* dgamma, float* dbeta, int B, int D, int H, int W, int Dp, int Hp, int Wp,
                      void* stream);

/* ---- Linear layers (F.linear: q/kv/proj :188-201,:246-259; Mlp fc1/fc2 :28-34; concat_back_dim :1029-1030;
 *      stride==kernel convs as GEMMs :557,:576,:871,:1037).
 *      Y[M,N] = epi(X[M,K] @ Wm + bias[N]);  w_is_kn=0: Wm = W[N,K]^T (nn.Linear / Conv weights),
 *      w_is_kn=1: Wm = W[K,N] (ConvTranspose weights).  epi: act=1 -> pre (if non-null) receives the
 *      pre-activation and Y = gelu_erf(pre) (:30);  res != null -> Y = res + rowscale[m / rows_per_sample] * val
 *      (residual add :419,:424 with timm DropPath per-sample scale; rowscale null -> 1);
 *      accumulate=1 -> Y += val (two-source linear == Linear on cat[a,b] :1027-1030). */
int mic_linear_fwd(const float* X, int ldx, const float* W, int ldw, int w_is_kn, const float* bias, float* Y, int ldy,
                   int M, int N, int K, int act, float* pre, int ldpre, const float* res, int ldres,
                   const float* rowscale, int rows_per_sample, int accumulate, void* stream);
/* dX[M,K] = (rowscale * dY[M,N]) @ Wm^T, optionally * gelu'(gelu_pre[M,K]); accumulate=1 -> dX += */
int mic_linear_bwd_data(const float* dY, int lddy, const float* W, int ldw, int w_is_kn, float* dX, int lddx, int M,
                        int N, int K, const float* gelu_pre, int ldpre, const float* rowscale, int rows_per_sample,
                        int accumulate, void* stream);
/* dW += (rowscale * dY)^T @ X  (shape [N,K], or [K,N] if w_is_kn);  db[N] += column sums of rowscale*dY (nullable) */
int mic_linear_bwd_weight(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, int w_is_kn,
                          float* db, int M, int N, int K, const float* rowscale, int rows_per_sample, void* stream);

/* Decoder tail without the 4^3 block permutation (reverse_patch_embedding, reference models/MICFormer_self.py:1033-1037):
 * the NEXT mic_linear_fwd / mic_linear_bwd_data / mic_linear_bwd_weight call of this host thread addresses its Y / dY
 * matrix -- logically (B*dc*hc*wc cells) x (64*ch) ConvTranspose3d(k4,s4) rows, ld = 64*ch -- directly in the
 * (B, 4dc, 4hc, 4wc, ch) channels-last grid it is the block permutation of (5-D TMA tensor maps; no intermediate, no
 * mic_block_permute pass).  Tensor-core mode only, wc == 32, hc % 4 == 0, ch % 8 == 0; a call that cannot honour the view
 * fails with MIC_ERR_UNSUPPORTED (it never silently treats the buffer as a plain matrix).  In mic_linear_bwd_weight
 * (w_is_kn = 1 only) db receives the column sums of the buffer as laid out in memory: exact after summing the 64 block
 * positions of a channel, i.e. for a bias tiled over the block positions.  (0,0,0,0) cancels a pending view. */
int mic_linear_unpatch_view(int ch, int dc, int hc, int wc);

/* ---- Windowed multi-head attention core: window_partition + softmax(q k^T * scale) v + window_reverse
 *      (:37-50,:117-132,:193-200,:251-258) without materialising windows or scores.  q/k/v/out are token-grid
 *      tensors (B,Dp,Hp,Wp,heads*hd) with row strides ldq/ldkv/ldo (k and v usually point into one kv buffer);
 *      lse (rows, heads) is saved for the backward.  No mask, no relative-position bias.
 *      Dispatch: tensor-core mode + head_dim 32 + 128..352 tokens per window + q/k/v adjacent in one (P,3C)
 *      buffer -> the tcgen05/TMA FlashAttention-style kernels (forward: S/P in TMEM, two softmax pipelines above 192
 *      tokens; backward: scores recomputed in both orientations, P / dS fed to the dQ / dK / dV products as TMEM A
 *      operands, csrc/window_attn_tc_bwd.cu); otherwise the exact-fp32 CUDA-core kernels (8-token windows of the
 *      train config).  MICFORMER_ATTN_BWD_SIMT=1 keeps the CUDA-core backward for large windows too. */
int mic_window_attn_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* out, int ldo,
                        float* lse, int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww,
                        float scale, void* stream);
int mic_window_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* out,
                        const float* dout, int ldo, const float* lse, float* dq, int lddq, float* dk, float* dv,
                        int lddkv, int B, int Dp, int Hp, int Wp, int heads, int hd, int wd, int wh, int ww,
                        float scale, void* stream);

/* ---- 3x3x3 convolution, stride 1, zero padding 1, channels-last, up to two concatenated input tensors
 *      (conv_offset[0] on cat[LN(x), xa] :313-314,:354-356; out_conv :1046,:1053).  Inputs live on the
 *      (D,H,W) grid and are implicitly zero outside it; outputs are produced on (Dp,Hp,Wp) >= (D,H,W).
 *      Wt is the weight permuted to [27][C0+C1][Co]; Co <= 16.  out_ncdhw=1 writes (B,Co,Dp,Hp,Wp). */
int mic_conv3_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wt, const float* bias, float* y,
                  int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co, int out_ncdhw, void* stream);
int mic_conv3_bwd_data(const float* dy, const float* Wt, float* dx0, int C0, int acc0, float* dx1, int C1, int acc1,
                       int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co, int dy_ncdhw, void* stream);
/* dWt[27][C0+C1][Co] += ..., dbias[Co] += ... */
int mic_conv3_bwd_weight(const float* dy, const float* x0, int C0, const float* x1, int C1, float* dWt, float* dbias,
                         int B, int D, int H, int W, int Dp, int Hp, int Wp, int Co, int dy_ncdhw, void* stream);

/* tcgen05 implicit-GEMM forward of the same convolution (TF32 tensor cores, fp32 accumulate in TMEM): a CTA owns an
 * 8x16 x/y footprint, marches over z and reads all 27 taps as shifted views of z-planes staged once in shared memory.
 * Wk is the weight permuted to [27][Co][C0+C1]; input grid == output grid, W % 8 == 0 and H % 16 == 0, else
 * MIC_ERR_UNSUPPORTED is returned and the caller uses mic_conv3_fwd. */
int mic_conv3_tc_fwd(const float* x0, int C0, const float* x1, int C1, const float* Wk, const float* bias, float* y,
                     int B, int D, int H, int W, int Co, int out_ncdhw, void* stream);
/* tcgen05 backward-data of the same convolution: the forward kernel run on dy with mirrored taps, N tile 32 over the
 * C0+C1 input channels.  Same arguments as mic_conv3_bwd_data on an unpadded grid; Co in {8,16}, same geometry rule. */
int mic_conv3_tc_bwd_data(const float* dy, const float* Wt, float* dx0, int C0, int acc0, float* dx1, int C1, int acc1,
                          int B, int D, int H, int W, int Co, int dy_ncdhw, void* stream);
/* Tensor-core backward-weight (warp-level mma.sync m16n8k8 TF32; the reduction runs over positions, so the 27 taps
 * are index-shifted shared-memory fragment loads): same result as mic_conv3_bwd_weight on an unpadded grid, Co in {8,16}.
 * native_layout != 0: dWt is the nn.Conv3d parameter's own (Co, Cin, 3, 3, 3) gradient (accumulated in place) instead
 * of the permuted [27][Cin][Co] buffer. */
int mic_conv3_mma_bwd_weight(const float* dy, const float* x0, int C0, const float* x1, int C1, float* dWt, float* dbias,
                             int B, int D, int H, int W, int Co, int dy_ncdhw, int native_layout, void* stream);
/* The two operand layouts the conv kernels read, [27][Cin][Co] and [27][Co][Cin], of n_jobs nn.Conv3d(k=3) weights
 * (Co, Cin, 3, 3, 3) in ONE launch (once per model forward, like mic_weight_images).  jobs: device array of n_jobs x 5
 * int64 {src, dst_tap_ci_co, dst_tap_co_ci, Cin, Co}; max_elems = largest 27*Cin over the jobs, max_co = largest Co (<= 64). */
int mic_conv_weight_layouts(const void* jobs, int n_jobs, int64_t max_elems, int max_co, void* stream);

/* ---- Offset head: LayerNormProxy(16) -> GELU -> Conv3d(16->3,k1,no bias) -> + reference points
 *      (:315-317, :326-337, :360-364).  h (P,HC) -> pos (P,3) with P = B*Dp*Hp*Wp. */
int mic_offset_head_fwd(const float* h, const float* gamma, const float* beta, const float* w3, float* pos, int B,
                        int Dp, int Hp, int Wp, int HC, float eps, void* stream);
int mic_offset_head_bwd(const float* dpos, const float* h, const float* gamma, const float* beta, const float* w3,
                        float* dh, float* dgamma, float* dbeta, float* dw3, int B, int Dp, int Hp, int Wp, int HC,
                        float eps, void* stream);

/* ---- Deformable trilinear resampling == SpatialTransformer.forward (models/STN.py:9-32) on the zero-padded
 *      other-modality map (:350,:379): out[p] = trilinear(src, c) with c_i = (idx_i + pos_i) * S_i/(S_i-1) - 0.5,
 *      S = (Dp,Hp,Wp), zeros outside.  src lives on (D,H,W) (implicit zero pad up to (Dp,Hp,Wp)). */
int mic_deform_sample_fwd(const float* src, const float* pos, float* out, int B, int D, int H, int W, int Dp, int Hp,
                          int Wp, int C, void* stream);
/* dsrc += scatter (atomic), dpos = gradient w.r.t. pos (written) */
int mic_deform_sample_bwd(const float* dout, const float* src, const float* pos, float* dsrc, float* dpos, int B,
                          int D, int H, int W, int Dp, int Hp, int Wp, int C, void* stream);

/* ---- k^3 block <-> row permutation for the stride==kernel (transposed) convolutions: PatchEmbed3D :871,
 *      PatchMerging :557, PatchExpand :576, reverse_patch_embedding :1037.  grid (B, k*D', k*H', k*W', C)
 *      channels-last <-> rows (B*D'*H'*W', k^3*C) with column order (kz,ky,kx,c).  to_rows=1 packs, 0 unpacks.
 *      grid_batch_stride (floats) lets a single-channel NCDHW volume be read in place (C=1). */
int mic_block_permute(const float* src, float* dst, int B, int Dq, int Hq, int Wq, int k, int C,
                      int64_t grid_batch_stride, int to_rows, void* stream);

/* ---- MDiceLoss (loss/dice.py:130-166): per channel sigmoid -> sum p*t, sum p^2, sum t^2, sum BCE (log clamped
 *      at -100) in one pass; finalize -> loss = (0.7*sum_c dice_c + 0.3*sum_c bce_c)/C and the per-channel
 *      coefficients the backward needs.  sums: double[C*4] zeroed by the caller (and, for a global-batch loss
 *      under data parallelism, all-reduced between the two calls).  n_per_channel = B*S (global). */
int mic_dice_bce_partial(const float* logits, const float* target, double* sums, int B, int C, int64_t S, void* stream);
int mic_dice_bce_finalize(const double* sums, float* loss, float* coef /*[C*3]*/, int C, double n_per_channel,
                          void* stream);
int mic_dice_bce_bwd(const float* logits, const float* target, const float* coef, const float* dloss, float* dlogits,
                     int B, int C, int64_t S, double n_per_channel, void* stream);
/* the same with uint8 / bool one-hot labels (what dataset/MMWHS.py:392,414-425 produces before the script's .float()):
 * the loss reads 1 byte per label instead of 4 */
int mic_dice_bce_partial_u8(const float* logits, const uint8_t* target, double* sums, int B, int C, int64_t S, void* stream);
int mic_dice_bce_bwd_u8(const float* logits, const uint8_t* target, const float* coef, const float* dloss, float* dlogits,
                        int B, int C, int64_t S, void* stream);
/* loss = (w_dice * sum_c dice_c + w_bce * sum_c bce_c) / C.  (0.7, 0.3) is MDiceLoss (loss/dice.py:158-166) ==
 * mic_dice_bce_finalize; (1, 0) is MDiceLoss_Val (loss/dice.py:216-221). */
int mic_dice_bce_finalize_weighted(const double* sums, float* loss, float* coef /*[C*3]*/, int C, double n_per_channel,
                                   double w_dice, double w_bce, void* stream);

/* ---- Multi-tensor Adam (torch.optim.Adam(lr, betas, eps, weight_decay), train_mmwhs_noPad.py:114,201) over all
 *      parameter tensors in one launch.  params/grads/exp_avg/exp_avg_sq: device arrays of n_tensors pointers (a null
 *      grad skips the tensor); sizes: device int64[n_tensors]; chunk_tensor/chunk_index: device int[n_chunks] mapping
 *      each CTA to (tensor, chunk of mic_adam_chunk_elems() elements); steps: device float[n_tensors] (one counter
 *      per parameter, incremented on the device for tensors that have a gradient); lr: device float scalar
 *      (CUDA-graph capturable). */
int mic_adam_chunk_elems(void);
int mic_adam_step(void* params, void* grads, void* exp_avg, void* exp_avg_sq, const int64_t* sizes,
                  const int* chunk_tensor, const int* chunk_index, int n_chunks, int n_tensors, float* steps,
                  const float* lr, float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ---- Fused block kernels for small channel counts (the train config's stage 0: C = 48, 8-token windows).  GEMM operands
 *      are "split bf16": x = hi + lo (two bf16 roundings), products hi*hi + lo*hi + hi*lo on tcgen05 kind::f16 with fp32
 *      accumulation in TMEM (~2^-17 relative, tighter than TF32).  Weights are consumed as pre-swizzled shared-memory
 *      IMAGES produced by mic_weight_images: B operand of N rows x K reduction elements, K-major bf16, SWIZZLE_128B,
 *      panels of 64 k (n_pad rows x 128 B each), rows >= N / k >= K zero.
 *      jobs: device array of n_jobs records of 9 x int64 {src fp32 ptr, hi ptr, lo ptr, ld, N, K, transpose, n_pad,
 *      k_panels} with B[n][k] = transpose ? src[k*ld + n] : src[n*ld + k]; max_chunks = max over jobs of
 *      k_panels*n_pad*8 (grid sizing). */
int mic_weight_images(const void* jobs, int n_jobs, int64_t max_chunks, void* stream);
/* y = x + rowscale[row / rows_per_sample] * (fc2(GELU(fc1(LayerNorm(x)))))  -- Mlp.forward inside the residual (reference
 * M:28-34, 403-404 / 516-524, DropPath scale M:419-424); x, y (T, C).  w1 image: N = 4C (n_pad 4C), K = C; w2 image:
 * N = C (n_pad = C rounded up to 16), K = 4C.  C in {24, 48}; other sizes return MIC_ERR_UNSUPPORTED. */
int mic_mlp_block_fwd(const float* x, float* y, const float* gamma, const float* beta, const float* b1, const float* b2,
                      const void* w1_hi, const void* w1_lo, const void* w2_hi, const void* w2_lo, const float* rowscale,
                      int rows_per_sample, int T, int C, float eps, void* stream);
/* Backward of the same from dy and x alone (LayerNorm, fc1, GELU are recomputed on chip): dx (T, C) is written,
 * dW1 (4C, C), db1 (4C), dW2 (C, 4C), db2 (C), dgamma (C), dbeta (C) are accumulated atomically (caller-zeroed or running
 * sums).  Images: w1nk = fc1 as N = 4C rows (n_pad: 4C rounded up to 64) x K = C; w2kn = fc2 transposed, same shape;
 * w1kn = fc1 transposed as N = C rows (n_pad: C rounded up to 16) x K = 4C. */
int mic_mlp_block_bwd(const float* dy, const float* x, float* dx, const float* gamma, const float* beta, const float* b1,
                      const void* w1nk_hi, const void* w1nk_lo, const void* w2kn_hi, const void* w2kn_lo,
                      const void* w1kn_hi, const void* w1kn_lo, const float* rowscale, int rows_per_sample, float* dW1,
                      float* db1, float* dW2, float* db2, float* dgamma, float* dbeta, int T, int C, float eps, void* stream);
/* Attention half of a block on 2x2x2 windows (TransformerBlock3D M:473-499; CrossTransformerBlock3D M:339-401 with
 * CrossWindowAttention3D M:179-203): x1 = x + rowscale * proj(softmax(q k^T * scale) v) with q = Wq LN(x) + bq and
 * [k|v] = Wkv src + bkv, src = LN(x) when kvsrc is null (self block) else kvsrc (the resampled other modality, not
 * normalised).  x, kvsrc, y (B, D, H, W, C), even D/H/W; rowscale per sample.  Images: wq / wp N = C (n_pad: C up to 16) x
 * K = C, wkv N = 2C x K = C.  Built for (C, head_dim) in {(48,16), (48,24)}; other shapes return MIC_ERR_UNSUPPORTED. */
int mic_attn_block_fwd(const float* x, const float* kvsrc, float* y, const float* gamma, const float* beta, const float* bq,
                       const float* bkv, const float* bp, const void* wq_hi, const void* wq_lo, const void* wkv_hi,
                       const void* wkv_lo, const void* wp_hi, const void* wp_lo, const float* rowscale, int B, int D, int H,
                       int W, int C, int heads, float scale, float eps, void* stream);
/* Backward of the same from dy (gradient w.r.t. x1), x and kvsrc alone (LayerNorm, q / kv and the attention probabilities
 * are recomputed on chip): dx = dy + LN'(...) is written; cross blocks also write dkvsrc (T, C) (pass both kvsrc and dkvsrc
 * or neither); dgamma, dbeta, dWq (C,C), dbq, dWkv (2C,C), dbkv, dWp (C,C), dbp are accumulated atomically.
 * imgs: host array of 10 device pointers {wq hi, lo, wkv hi, lo (forward images), wpT hi, lo, wqT hi, lo (transposed,
 * N = C (n_pad: C up to 16) x K = C), wkvT hi, lo (transposed, N = C x K = 2C: two 64-k panels)}. */
int mic_attn_block_bwd(const float* x, const float* kvsrc, const float* dy, float* dx, float* dkvsrc, const float* gamma,
                       const float* beta, const float* bq, const float* bkv, const void* const* imgs, const float* rowscale,
                       float* dgamma, float* dbeta, float* dWq, float* dbq, float* dWkv, float* dbkv, float* dWp, float* dbp,
                       int B, int D, int H, int W, int C, int heads, float scale, float eps, void* stream);
/* The same MLP half-block for the deep stages (64 <= C <= 384, C % 8 == 0, HID = 4C): the work is split over 128-token
 * tiles AND 64-unit hidden chunks (grid = tiles x HID/64), operands are streamed through shared-memory rings, and every CTA
 * adds its partial product into y, which MUST BE ZERO on entry (chunk 0 adds the residual and the bias).  Images as for
 * mic_mlp_block_fwd. */
int mic_mlp_split_fwd(const float* x, float* y_zeroed, const float* gamma, const float* beta, const float* b1,
                      const float* b2, const void* w1_hi, const void* w1_lo, const void* w2_hi, const void* w2_lo,
                      const float* rowscale, int rows_per_sample, int T, int C, int HID, float eps, void* stream);
/* Backward of mic_mlp_split_fwd up to the LayerNorm: recomputes LN / fc1 / GELU per (tile, hidden chunk); dxn (T, C) -- the
 * gradient w.r.t. LN(x), ZERO on entry -- and dW1, db1, dW2, db2 are accumulated atomically; mean / rstd (T) of the LayerNorm
 * are written for mic_layernorm_bwd, which finishes the half-block: dx = dy + LN'(dxn), dgamma, dbeta.  Images: w1nk (fc1,
 * N = 4C x K = C), w2kn (fc2 transposed, N = 4C x K = C), w1kn (fc1 transposed, N = C x K = 4C). */
int mic_mlp_split_bwd(const float* dy, const float* x, float* dxn_zeroed, float* mean, float* rstd, const float* gamma,
                      const float* beta, const float* b1, const void* w1nk_hi, const void* w1nk_lo, const void* w2kn_hi,
                      const void* w2kn_lo, const void* w1kn_hi, const void* w1kn_lo, const float* rowscale, int rows_per_sample,
                      float* dW1, float* db1, float* dW2, float* db2, int T, int C, int HID, float eps, void* stream);
/* profiling hook: register (or clear with NULL) a device buffer of 4 x 8 x 32 uint64 that the fused kernels stamp with
 * %globaltimer at their phase boundaries (scripts/trace_fused.py) */
int mic_debug_t5_trace(void* buf);
/* dynamic shared memory the fused MLP kernels need for this C (-1: not built) */
int mic_mlp_block_smem(int C);

/* ---- small utilities: y = a + rowscale*b with crop from a padded grid (residual after window_reverse + crop
 *      :397-400,:419), fill, axpy ---- */
int mic_crop_residual(const float* res, const float* branch, const float* rowscale, float* y, int B, int D, int H,
                      int W, int Dp, int Hp, int Wp, int C, void* stream);
/* dbranch (padded grid) = rowscale * dy inside, 0 in the pad */
int mic_crop_residual_bwd(const float* dy, const float* rowscale, float* dbranch, int B, int D, int H, int W, int Dp,
                          int Hp, int Wp, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MICFORMER_B200_H */
