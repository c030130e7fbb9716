/* objgan_b200 -- C ABI of libobjgan_b200.so (sm_100a kernels for the Obj-GAN image_generation hot path).
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise; every call is
 * asynchronous on `stream` and re-entrant (no global mutable state), which is what the reference's one FFI
 * crossing requires (models/roi_align/src/roi_align_cuda.c:5,31: global THCState, current-stream launch).
 *
 * Return convention: og_* functions return 0 on success, else the cudaError_t of the failed call.  The two
 * ROIAlign*Laucher symbols keep the reference convention (1 = ok, 0 = error) because they are the literal
 * drop-ins for models/roi_align/src/roi_align_kernel.h:13-27.  Nothing in this library calls exit().
 *
 * Activations are NHWC fp32 ("rows" = pixels, channels contiguous), channel counts padded to a multiple of
 * 8 with zero lanes; the module boundary (NCHW, unpadded -- the reference's public layout) is crossed with
 * og_nchw_to_nhwc / og_nhwc_to_nchw.  All "ref:" paths are relative to /root/reference/image_generation/.
 */
#ifndef OBJGAN_B200_H
#define OBJGAN_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- codes ---- */
enum { OG_ACT_NONE = 0, OG_ACT_LRELU = 1, OG_ACT_TANH = 2, OG_ACT_SIGMOID = 3 };  /* conv epilogue activations  */
enum { OG_NA_NONE = 0, OG_NA_LRELU = 1, OG_NA_GLU = 2 };                          /* norm-apply fusions         */
enum { OG_PAD_ZERO = 0, OG_PAD_REFLECT = 1, OG_UPSAMPLE2X = 2, OG_TRANSPOSED = 3 }; /* conv source addressing   */

/* ------------------------------------------------------------------------------------------------------
 * ROIAlign -- replaces ref: models/roi_align/src/roi_align_kernel.h:13-27 (same names, argument order,
 * NCHW fp32 features, rois = [batch_idx, x1, y1, x2, y2] rows, caller zero-fills outputs;
 * callers: models/roi_align/src/roi_align_cuda.c:33-38, 70-73).
 * ---------------------------------------------------------------------------------------------------- */
int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois, const int height,
                           const int width, const int channels, const int aligned_height, const int aligned_width,
                           const float* bottom_rois, float* top_data, cudaStream_t stream);
int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size, const int num_rois,
                            const int height, const int width, const int channels, const int aligned_height,
                            const int aligned_width, const float* bottom_rois, float* bottom_diff, cudaStream_t stream);
/* fused RoIAlignAvg(AH, AW, scale) = align to (AH+1)x(AW+1) then avg_pool2d(2, 1)
 * -- replaces ref: models/roi_align/modules/roi_align.py:18-29 (used at model.py:1241, 1307). */
int og_roi_align_avg_fwd(const float* features, int height, int width, int channels, const float* rois, int num_rois,
                         int AH, int AW, float spatial_scale, float* out, cudaStream_t stream);
int og_roi_align_avg_bwd(const float* grad_out, int height, int width, int channels, const float* rois, int num_rois,
                         int AH, int AW, float spatial_scale, float* grad_features, cudaStream_t stream);

/* channels-last variant (features NHWC [B][H][W][C], out NHWC [R][AH][AW][C], C % 4 == 0): identical values, coalesced
 * 16-byte loads / stores / vector atomics; what the object discriminators use internally (their feature maps are
 * produced channels-last), the NCHW entry points above stay the drop-in for the reference's callers. */
int og_roi_align_avg_nhwc_fwd(const float* features, int height, int width, int channels, const float* rois,
                              int num_rois, int AH, int AW, float spatial_scale, float* out, cudaStream_t stream);
int og_roi_align_avg_nhwc_bwd(const float* grad_out, int height, int width, int channels, const float* rois,
                              int num_rois, int AH, int AW, float spatial_scale, float* grad_features,
                              cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Convolutions / linear layers as implicit GEMM -- replaces the cuDNN/cuBLAS calls behind nn.Conv2d /
 * nn.Linear in ref: model.py:36-39 (conv3x3), 43-49 (upBlock), 52-60 (downBlock_G), 63-81 (HmapResBlock),
 * 455-518 (CA_NET / INIT_STAGE_G fc), 589-617 (G_HMAP), 708-719 (GET_IMAGE_G), 999-1048 (D stacks).
 * x: source NHWC [N,H,W,C] with element strides (xsn,xsh,xsw); y likewise [N,OH,OW,K]; C % 8 == 0, K % 4 == 0.
 * mode selects how a (row pixel, tap) pair maps to a source pixel: zero padding, reflection padding
 * (nn.ReflectionPad2d), nearest-2x upsampling followed by zero padding (nn.Upsample + conv), or the
 * transposed map used for the input gradient.  wpacked is produced by og_pack_weights.
 * ---------------------------------------------------------------------------------------------------- */
int og_conv2d_simt(const float* x, int N, int H, int W, int C, long long xsn, long long xsh, long long xsw,
                   const float* wpacked, float* y, int OH, int OW, int K, long long ysn, long long ysh, long long ysw,
                   int KH, int KW, int stride, int pad, int mode, const float* bias, int act, float slope,
                   int allow_splitk, cudaStream_t stream);
int og_conv2d_wgrad_simt(const float* x, int N, int H, int W, int C, long long xsn, long long xsh, long long xsw,
                         const float* g, int OH, int OW, int K, long long gsn, long long gsh, long long gsw,
                         float* dw_packed, int KH, int KW, int stride, int pad, int mode, cudaStream_t stream);
/* <= 8 output channels with a long reduction (the k4 s2 p0 logit heads, ref: model.py:1031-1033): dot-product
 * kernels instead of GEMM tiles.  Contiguous NHWC, zero padding, output channels padded to 8, wpacked [(kh,kw,ci)][8]. */
int og_conv2d_narrow_fwd(const float* x, int N, int H, int W, int C, const float* wpacked, float* y, int OH, int OW,
                         int KH, int KW, int stride, int pad, const float* bias, int act, float slope,
                         cudaStream_t stream);
int og_conv2d_narrow_dgrad(const float* g, int N, int H, int W, int C, const float* wpacked, float* gx, int OH, int OW,
                           int KH, int KW, int stride, int pad, cudaStream_t stream);
int og_conv2d_narrow_wgrad(const float* x, int N, int H, int W, int C, const float* g, float* dw_packed, int OH, int OW,
                           int KH, int KW, int stride, int pad, cudaStream_t stream);
/* OIHW parameter (state_dict layout) <-> kernel-native matrices.  split/splitp: GLU halves of the
 * output channels are each padded to splitp.  transposed=1 gives the dgrad operand ([tap][co][ci] instead of
 * [tap][ci][co]).  og_pack_weights_f16 writes the same matrices as the fp16 hi/lo operand pair of the tensor-core
 * path, scaled by the power of two derived from *amax (og_amax over the parameter; see og_prep_split). */
int og_pack_weights(const float* w_oihw, int Co, int Ci, int KH, int KW, int Cip, int Kp, int split, int splitp,
                    int transposed, float* out, cudaStream_t stream);
int og_pack_weights_f16(const float* w_oihw, int Co, int Ci, int KH, int KW, int Cip, int Kp, int split, int splitp,
                        int transposed, const unsigned* amax, void* hi, void* lo, cudaStream_t stream);
int og_unpack_wgrad(const float* dw_packed, int Co, int Ci, int KH, int KW, int Cip, int Kp, int split, int splitp,
                    float* grad_oihw, int accumulate, int transposed, cudaStream_t stream);

/* upBlock (nearest 2x upsample + conv3x3, ref: model.py:43-49) as four 2x2 phase convolutions with pre-summed
 * weights: Wp[16][..] (tap t = ((p*2+q)*2+a)*2+b) from the OIHW parameter as fp16 hi/lo operands (scale word
 * *amax_up = 4 * *amax, the bound of the pre-sums), and the adjoint map for the gradient. */
int og_pack_upsample_weights(const float* w_oihw, int Co, int Ci, int Cip, int Kp, int split, int splitp,
                             int transposed, const unsigned* amax, unsigned* amax_up, void* hi, void* lo,
                             cudaStream_t stream);
int og_unpack_upsample_wgrad(const float* dwp, int Co, int Ci, int Cip, int Kp, int split, int splitp,
                             float* grad_oihw, cudaStream_t stream);

/* Tensor-core path (tcgen05.mma kind::f16 with fp32 accumulation in TMEM, TMA operand staging) for the
 * contractions that dominate the step (HmapResBlock / upBlock / jointConv / discriminator convs: forward, input
 * and weight gradients).  Operands are fp16 hi/lo pairs (22 mantissa bits; "3xFP16": a*b ~ ah*bh + al*bh + ah*bl)
 * of x * 2^k, k = og_scale_exp(max|x|) per tensor; the word *amax holds max|x| as float bits and travels with the
 * operand.  og_prep_split computes *amax and writes the pair for an activation tensor (pad=1 also materialises
 * the nn.ReflectionPad2d(1) halo, ref: model.py:67; s2d=1 writes the four space-to-depth phase blocks used by
 * the stride-2 convs of the discriminators and by the adjoint of the 2x upsampling); og_conv2d_tc computes
 *     y[n, osy*h+opy, osx*w+opx, :] = act(bias + sum_t x[n+dn_t, h+dh_t, w+dw_t, :] * W[widx_t])   (x OOB = 0)
 * with taps = ntaps host quadruples (dh, dw, dn, widx); nsplit = 3 is the error-compensated product (fp32-level
 * accuracy), nsplit = 1 a single fp16 product (xl / wl unused). */
int og_amax(const float* x, long long n, unsigned* amax, cudaStream_t stream);
int og_prep_split(const float* x, int N, int H, int W, int C, int pad, int s2d, unsigned* amax, int amax_ready,
                  void* xh, void* xl, cudaStream_t stream);
int og_conv2d_tc(const void* xh, const void* xl, const unsigned* amax_x, int N, int SN, int SH, int SW, int C,
                 const void* wh, const void* wl, const unsigned* amax_w, int ntaps_w, int Kw, float* y, int OH, int OW,
                 int K, long long ysn, long long ysh, long long ysw, int OHf, int OWf, int osy, int osx, int opy,
                 int opx, const int* taps_host, int ntaps, int tap_layout, int nsplit, const float* bias, int act,
                 float slope, cudaStream_t stream);
/* Weight gradient: dw[tap][co][ci] = sum_{n,h,w} G[n+gdn, h, w, co] * X[n+xdn, h+dh, w+dw, ci]  (X out of range = 0).
 * Both operands are the same NHWC fp16 hi/lo tensors og_prep_split writes for the forward / input-gradient kernels
 * (G: [GN][OH][OW][Kp], X: [XN][SH][SW][C]; GN / XN = N or 4N space-to-depth blocks), consumed MN-major by
 * tcgen05.mma; each needs 512 readable bytes after its last element.
 * entries_host: nentries quintuples (gdn, dh, dw, xdn, output tap); dw is [ntaps_out][Kp][C], zero-filled here. */
int og_conv2d_wgrad_tc(const void* gh, const void* gl, const unsigned* amax_g, int N, int GN, int OH, int OW, int Kp,
                       const void* xh, const void* xl, const unsigned* amax_x, int XN, int SH, int SW, int C, float* dw,
                       int ntaps_out, const int* entries_host, int nentries, int nsplit, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * InstanceNorm2d / BatchNorm (train mode) + fused GLU / LeakyReLU / residual
 * -- replaces ref: model.py:19-27 (GLU), 47, 70, 75, 497, 602, 992, 1011 (norm layers).
 * groups = N (instance norm) or 1 (batch norm); P pixels per group; contiguous rows of C channels.
 * amax_out / amax_dy (optional): receive max|out| / max|dy| as float bits, so that a convolution consuming the
 * tensor can skip the amax pass of og_prep_split (amax_ready = 1).
 * ---------------------------------------------------------------------------------------------------- */
int og_norm_stats(const float* x, int groups, long long P, int C, float eps, double* stats, float* mean, float* rstd,
                  float* running_mean, float* running_var, float momentum, int real_c, long long* num_batches_tracked,
                  cudaStream_t stream);
int og_norm_apply(const float* y, int groups, long long P, int Cy, const float* mean, const float* rstd,
                  const float* gamma, const float* beta, const float* res, int act, float slope, float* out,
                  unsigned* amax_out, cudaStream_t stream);
/* og_norm_apply that also emits the fp16 hi / lo tensor-core operand copies of its output (what og_prep_split would
 * make of it, optionally with the reflection halo), scaled by the a-priori bound og_norm_bound leaves in bound_word;
 * out (fp32) may be null when nothing else reads the activation */
int og_norm_bound(const float* gamma, const float* beta, int C, long long count, const unsigned* res_word,
                  unsigned* out_word, cudaStream_t stream);
int og_norm_apply_split(const float* y, int N, int H, int W, int Cy, int instance, const float* mean,
                        const float* rstd, const float* gamma, const float* beta, const float* res, int act,
                        float slope, float* out, const unsigned* bound_word, int pad, void* xh, void* xl,
                        cudaStream_t stream);
int og_norm_backward(const float* y, const float* g, int groups, long long P, int Cy, const float* mean,
                     const float* rstd, const float* gamma, const float* beta, int act, float slope, double* bstats,
                     float* dy, float* dgamma, float* dbeta, int accumulate_param_grads, unsigned* amax_dy,
                     cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Attention -- replaces ref: GlobalAttention.py:32-70 (func_attention), 73-122 (GlobalAttentionGeneral),
 * 125-181 (GlobalBUAttentionGeneral) and miscc/utils.py:401-413 (pprocess_bt_attns).
 * mask: [B][L] bytes, 1 = padding word; the reference's mask-row permutation quirk is reproduced.
 * ---------------------------------------------------------------------------------------------------- */
int og_words_proj(const float* words, const float* W, int B, int idf, int cdf, int L, float* src, cudaStream_t stream);
int og_words_proj_bwd(const float* words, const float* W, const float* gsrc, int B, int idf, int cdf, int L, float* gW,
                      int accumulate, float* gwords, cudaStream_t stream);
int og_att_general_fwd(const float* h, const float* src, const unsigned char* mask, int B, int Q, int idf, int cs,
                       int L, float* wc, float* attn, cudaStream_t stream);
int og_att_general_bwd(const float* h, const float* src, const float* attn, const float* g_wc, const float* g_attn,
                       int B, int Q, int idf, int cs, int L, float* g_h, float* g_src, cudaStream_t stream);
int og_bu_att_fwd(const float* labels, const float* glove, const float* src, const unsigned char* mask, int B, int E,
                  int R, int L, int idf, int norm, float eps, float* wc, float* attn, cudaStream_t stream);
int og_bu_att_bwd(const float* attn, const float* g_wc, int B, int R, int L, int idf, float* g_src,
                  cudaStream_t stream);
int og_paint_max_fwd(const float* f, const float* m, int B, int num, int R, int Rtot, long long P, float* out,
                     int dstride, int doff, cudaStream_t stream);
int og_paint_max_bwd(const float* f, const float* m, const float* g, int gstride, int goff, int B, int num, int R,
                     int Rtot, long long P, float* g_f, cudaStream_t stream);
int og_func_attention_fwd(const float* query, const float* ctx, int B, int ndf, int Lq, int S, float gamma1, float* wc,
                          float* attn, cudaStream_t stream);
int og_func_attention_bwd(const float* query, const float* ctx, const float* attn, const float* g_wc,
                          const float* g_attn, int B, int ndf, int Lq, int S, float gamma1, float* g_query, float* g_ctx,
                          cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Layout, adjoints of pad / upsample, concat, losses, optimiser
 * -- replaces ATen elementwise kernels behind ref: model.py:45 (Upsample), 67 (ReflectionPad2d), torch.cat at
 * model.py:579, 697, 1041; nn.BCELoss in miscc/losses.py:178-208, 378-393; KL_loss miscc/losses.py:533-537;
 * CA_NET.reparametrize model.py:470-478; optim.Adam + EMA trainer.py:197-224, 461-462.
 * ---------------------------------------------------------------------------------------------------- */
int og_nchw_to_nhwc(const float* x, int N, int C, int H, int W, int Cp, float* y, cudaStream_t stream);
int og_nhwc_to_nchw(const float* y, int N, int C, int H, int W, int Cp, float* x, cudaStream_t stream);
int og_act_backward(const float* out, const float* g, long long n, int act, float slope, float* gin,
                    cudaStream_t stream);
int og_channel_sum(const float* x, long long P, int C, double* scratch, float* out, int n_out, int accumulate,
                   cudaStream_t stream);
int og_upsample2x_bwd(const float* gu, int N, int H, int W, int C, float* gx, cudaStream_t stream);
int og_reflect_pad_fwd(const float* x, int N, int H, int W, int C, float* xp, cudaStream_t stream);
int og_reflect_pad_bwd(const float* gpad, int N, int H, int W, int C, float* gx, cudaStream_t stream);
int og_copy_channels(const float* src, int sstride, int soff, float* dst, int dstride, int doff, int nch, long long P,
                     int accumulate, cudaStream_t stream);
int og_broadcast_channels(const float* c, int B, int Cc, float* dst, int dstride, int doff, long long pix_per_img,
                          cudaStream_t stream);
int og_broadcast_channels_bwd(const float* g, int B, int Cc, int gstride, int goff, long long pix_per_img, float* gc,
                              cudaStream_t stream);   /* gc[b][k] = sum_pixels g[pixel][goff + k] */
/* Row gather out[i,:] = x[idx[i],:] (idx: int64 device array) and its adjoint (gx zero-filled, then += g rows):
 * the roi compaction of feat_select (ref: miscc/utils.py:465-499) and the raw_conditions[classes] lookup of
 * objD_loss (ref: miscc/losses.py:280-281) without per-roi host loops. */
int og_gather_rows(const float* x, const long long* idx, long long n_out, long long rowlen, float* out,
                   cudaStream_t stream);
int og_scatter_rows_add(const float* g, const long long* idx, long long n_out, long long n_in, long long rowlen,
                        float* gx, cudaStream_t stream);
int og_add(const float* a, const float* b, float* out, long long n, cudaStream_t stream);
/* F.interpolate(bilinear, align_corners=True) on NHWC (ref: model.py:1217-1218, 1283-1284) and its adjoint */
int og_bilinear_fwd(const float* x, int N, int IH, int IW, int C, int OH, int OW, float* y, cudaStream_t stream);
int og_bilinear_bwd(const float* g, int N, int IH, int IW, int C, int OH, int OW, float* gx, cudaStream_t stream);
int og_glu_fwd(const float* x, long long P, int Ch, float* out, cudaStream_t stream);
int og_glu_bwd(const float* x, const float* g, long long P, int Ch, float* gx, cudaStream_t stream);
int og_reparam_fwd(const float* x, int xs, const float* eps, int B, int D, float* c, int cs, cudaStream_t stream);
int og_reparam_bwd(const float* x, int xs, const float* eps, const float* gc, int gcs, int B, int D, float* gx,
                   cudaStream_t stream);
int og_bce(const float* p, long long n, float target, float weight, float* loss_accum, float* gp, cudaStream_t stream);
int og_kl(const float* x, int xs, int B, int D, float weight, float* loss_accum, float* gx, cudaStream_t stream);
/* step: 1-based Adam step from the host, or (step_dev != NULL) read from a device counter so that a whole training
 * step can be replayed from a CUDA graph; og_inc_i64 bumps such a counter. */
int og_adam_ema(float* p, const float* g, float* m, float* v, float* avg, long long n, double lr, double b1, double b2,
                double eps, int step, const long long* step_dev, float gscale, float decay, cudaStream_t stream);
int og_inc_i64(long long* counter, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * DAMSM matching losses around func_attention (ref: miscc/losses.py:13-159: cosine_similarity, sent_loss,
 * words_loss).  Dense row-major fp32; B <= 64 captions.
 *  og_cosine_cl_*     : out[b][l] = cos(word[:, l], wei[b, :, l]) over the D features (word [D][L] shared by all
 *                       images, wei [B][D][L]); backward w.r.t. wei (losses.py:101-108)
 *  og_expsumlog_*     : out[b] = log(sum_l exp(gamma * s[b][l]))            (Eq. 10, losses.py:112-115)
 *  og_cosine_matrix_* : out[i][j] = <a_i, b_j> / max(|a_i||b_j|, eps); backward w.r.t. a   (losses.py:43-50)
 *  og_ce_pair         : scores = gamma3 * sim with masked entries at -inf; loss0 = CrossEntropy(scores, labels),
 *                       loss1 = CrossEntropy(scores^T, labels) (means), their gradients G0 / G1 w.r.t. sim, and the
 *                       number of top-1 hits of both directions (losses.py:50-68, 131-147)
 *  og_ce_pair_bwd     : gsim = *g0 * G0 + *g1 * G1 (device scalars; NULL = 0)
 * ---------------------------------------------------------------------------------------------------- */
int og_cosine_cl_fwd(const float* word, const float* wei, int B, int D, int L, float eps, float* out,
                     cudaStream_t stream);
int og_cosine_cl_bwd(const float* word, const float* wei, const float* g, int B, int D, int L, float eps, float* gwei,
                     cudaStream_t stream);
int og_expsumlog_fwd(const float* s, int B, int L, float gamma, float* out, cudaStream_t stream);
int og_expsumlog_bwd(const float* s, const float* g, int B, int L, float gamma, float* gs, cudaStream_t stream);
int og_cosine_matrix_fwd(const float* a, const float* b, int Ba, int Bb, int D, float eps, float* out,
                         cudaStream_t stream);
int og_cosine_matrix_bwd(const float* a, const float* b, const float* g, int Ba, int Bb, int D, float eps, float* ga,
                         cudaStream_t stream);
int og_ce_pair(const float* sim, const unsigned char* mask, const long long* labels, int B, float gamma3, float* loss0,
               float* loss1, float* G0, float* G1, float* correct, cudaStream_t stream);
int og_ce_pair_bwd(const float* G0, const float* G1, const float* g0, const float* g1, int n, float* gsim,
                   cudaStream_t stream);

/* All (image, caption) pairs of words_loss in ONE launch (ref: miscc/losses.py:87-127 -- the reference calls
 * func_attention once per caption in a Python loop).  words [NC][ndf][T] (caption i uses its first lens[i] words),
 * ctx [B][ndf][S] image region features (NCHW flattened), lens [NC] int64 on the device.  pair p = b * NC + i.
 *  fwd: wc [B*NC][ndf][T] (weighted contexts), attn [B*NC][T][S], sim [B][NC] = log sum_l exp(gamma2 * cos_l)
 *       (rows / columns beyond a caption's length are not written)
 *  bwd: g_ctx [B][ndf][S] = d (sum_p g_sim[p] * sim[p]) / d ctx  (zero-filled by the call, fp32 atomics) */
int og_words_pairs_fwd(const float* words, const float* ctx, const long long* lens, int B, int NC, int ndf, int T, int S,
                       float gamma1, float gamma2, float eps, float* wc, float* attn, float* sim, cudaStream_t stream);
int og_words_pairs_bwd(const float* words, const float* ctx, const long long* lens, const float* wc, const float* attn,
                       const float* g_sim, int B, int NC, int ndf, int T, int S, float gamma1, float gamma2, float eps,
                       float* g_ctx, cudaStream_t stream);

/* Device-side input preparation (ref: trainDataset.py:79-128, miscc/load.py:160-176, miscc/utils.py:502-522):
 *  og_form_hmaps        : out[b][cls[b][r]][p] += masks[b][r][p] for r < num_rois[b], in roi order (the loader's loop);
 *                         clamp_max > 0 additionally clamps the touched channels (synthetic inputs); out NCHW, zeroed here
 *  og_form_clabels_feat : out[b][e][r] = emb[cls[b][r]][e] for r < num_rois[b], else 0   (B, E, Rmax, 1) */
int og_form_hmaps(const float* masks, const long long* cls, const long long* num_rois, int B, int R, long long P,
                  int ncls, float clamp_max, float* out, cudaStream_t stream);
int og_form_clabels_feat(const float* emb, const long long* cls, const long long* num_rois, int B, int R, int Rmax, int E,
                         int ncls, float* out, cudaStream_t stream);

/* Whole-network weight re-packing after an optimiser step: og_amax + og_pack_weights_f16 for EVERY registered
 * (weight, layout) of a network in one launch each, from a device-resident int64 job table.
 *  amax job (4 fields): x, n, out word, first block index;  a block covers 4096 elements
 *  pack job (14 fields): w (OIHW), hi, lo (or 0), amax word, Co, Ci, taps, Cip, Kp, split, splitp, transposed, total,
 *                        first block index */
int og_amax_multi(const long long* jobs, int njobs, int total_blocks, cudaStream_t stream);
int og_pack_weights_f16_multi(const long long* jobs, int njobs, int total_blocks, cudaStream_t stream);

/* out[b][c] = x[b][perm[b][c]] over NCHW planes of P floats (permute_seg's class-channel shuffle, utils.py:445-462) */
int og_permute_channels(const float* x, const long long* perm, int B, int C, long long P, float* out,
                        cudaStream_t stream);

/* zero-fill (a memset node when captured in a CUDA graph; no kernel launch) */
int og_zero_bytes(float* p, long long bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OBJGAN_B200_H */
