/* emsanet_b200 — C ABI of the B200-native EMSANet forward/backward engine.
 *
 * Plain C: raw device pointers, sizes, a cudaStream_t passed as void*.  Every function returns
 * 0 on success and a non-zero code otherwise; eb200_last_error() returns the message of the last
 * failure on the calling thread.  All tensors are borrowed for the duration of the call and all
 * work is enqueued on the given stream (no host synchronisation), so calls can be captured into a
 * CUDA graph.  Activations are NHWC bf16 unless stated otherwise; parameters arrive in the
 * reference's own layouts (fp32 [Cout,Cin,kh,kw]) and are re-laid-out by eb200_pack_conv_weight.
 *
 * The reference has no FFI (it is pure Python); each entry point below names the
 * reference nn.Module call it replaces.  Paths: MT/ = lib/nicr-multitask-scene-analysis/src/
 * nicr_mt_scene_analysis/.
 */
#ifndef EMSANET_B200_H_
#define EMSANET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EB200_MAX_TAPS 9

/* epilogue flags of eb200_conv2d */
#define EB200_BIAS 1u      /* + bias[c] (fp32)                                            */
#define EB200_RELU 2u      /* max(.,0); applied after the aux add when EB200_AUX_ADD      */
#define EB200_AUX_ADD 4u   /* + aux (bf16, addressed with the aux strides)                */
#define EB200_AUX_MASK 8u  /* keep value only where aux > 0 (ReLU backward)               */
#define EB200_STATS 16u    /* stats[0:C] += sum, stats[C:2C] += sum of squares (fp32)     */
#define EB200_STATS_SUM_ONLY 32u /* with EB200_STATS: only stats[0:C] += sum (bias gradient straight into .grad) */
#define EB200_BN_BWD 64u   /* with EB200_AUX_MASK|EB200_STATS: fused first half of a BatchNorm+ReLU backward on the
                              data gradient this conv produces: aux = the BN's raw input x, the value is kept where
                              x*bn_scale+bn_shift > 0 (the ReLU mask recomputed), stats[0:C] += sum g,
                              stats[C:2C] += sum g*x  (follow with eb200_bn_bwd_apply_raw)            */

/* A strided NHWC view: element (n,h,w,c) lives at ptr + n*sn + h*sh + w*sw + c  (strides in elements).
 * Stride-2 convolutions are expressed as stride-1 convolutions over row/column parity views. */
typedef struct {
  const void* ptr;
  int n, h, w, c;
  long long sn, sh, sw;
} eb200_view;

/* Implicit-GEMM convolution on the tcgen05 tensor cores:
 *   out[n,h,w,co] = epilogue( sum_t sum_ci in[view[t]][n, h+dy[t], w+dx[t], ci] * weight[tap_w[t]][co][ci] )
 * out-of-range (h+dy, w+dx) reads are zero (= zero padding).  Replaces nn.Conv2d.forward for the
 * 1x1 / 3x1 / 1x3 / 3x3 convolutions of MT/model/block.py:174-190, MT/model/utils.py:17-41,59-64,
 * MT/model/decoder/dense_utils.py:21-23, MT/model/decoder/instance.py:51-78 and, with flipped taps and
 * transposed weights, their autograd data-gradient (aten convolution_backward, entered at main.py:598). */
typedef struct {
  eb200_view in[2];
  int n, h, w;                  /* output extent                                         */
  int cin, cout;                /* cin = channels read (rounded up to 64 inside), cout % 8 == 0 */
  int taps;
  int tap_view[EB200_MAX_TAPS], tap_dy[EB200_MAX_TAPS], tap_dx[EB200_MAX_TAPS];
  int tap_w[EB200_MAX_TAPS];    /* weight slice used by tap t (0 <= tap_w[t] < weight_taps)            */
  int weight_taps;              /* slices in `weight`                                                   */
  const void* weight;           /* bf16 [taps][cout_pad][cin_pad] from eb200_pack_conv_weight */
  int cout_pad, cin_pad;
  void* out;                    /* bf16 */
  long long out_sn, out_sh, out_sw;
  const void* aux;              /* bf16 or NULL */
  long long aux_sn, aux_sh, aux_sw;
  const float* bias;            /* fp32 [cout] or NULL */
  float* stats;                 /* fp32 [2*cout] or NULL */
  uint32_t flags;
  const float* bn_scale;        /* EB200_BN_BWD: fp32 [cout] affine of the BatchNorm whose backward is fused */
  const float* bn_shift;
} eb200_conv_desc;

int eb200_conv2d(const eb200_conv_desc* d, void* stream);
/* Two independent convolutions of identical geometry (the RGB / depth encoder branches, the semantic / instance decoders
 * of EMSANet: MT/model/encoder.py:220-261, emsanet/decoder.py) in ONE launch where the halo kernel allows it — even CTAs
 * work on `a`, odd CTAs on `b` — otherwise the same as two eb200_conv2d calls. */
int eb200_conv2d_pair(const eb200_conv_desc* a, const eb200_conv_desc* b, void* stream);

/* Weight gradient on the tensor cores (aten convolution_backward, weight part):
 *   dw[co][ci][t] += sum_{n,h,w} dy[n,h,w,co] * x[view[t]][n, h+dy[t], w+dx[t], ci]
 * dw is fp32 with arbitrary element strides so it can be the parameter's .grad in the reference
 * layout [Cout,Cin,kh,kw] (dw_sco = Cin*taps, dw_sci = taps, dw_st = 1).  Accumulates atomically. */
typedef struct {
  eb200_view dy;                /* c = cout */
  eb200_view x[2];              /* c = cin  */
  int taps;
  int tap_view[EB200_MAX_TAPS], tap_dy[EB200_MAX_TAPS], tap_dx[EB200_MAX_TAPS];
  float* dw;
  long long dw_sco, dw_sci, dw_st;
  float* ws;                    /* optional: ZEROED fp32 staging of >= cout*cin*9 floats for 3x3 filters (left zeroed);   */
  long long ws_floats;          /* lets the three tap groups leave as TMA bulk reductions instead of scalar atomics        */
} eb200_wgrad_desc;

int eb200_conv2d_wgrad(const eb200_wgrad_desc* d, void* stream);
int eb200_conv2d_wgrad_pair(const eb200_wgrad_desc* a, const eb200_wgrad_desc* b, void* stream);

/* fp32 [Cout,Cin,kh,kw] (reference layout) -> bf16 [kh*kw][cout_pad][cin_pad], zero padded.
 * transpose != 0 writes [kh*kw][cin rows][cout cols] (weights for the data gradient; the caller negates the tap
 * offsets instead of flipping the slices).
 * ci_offset/co_offset place the block inside a larger padded matrix (block-diagonal heads). */
int eb200_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, void* packed, int cout_pad,
                           int cin_pad, int transpose, int co_offset, int ci_offset, void* stream);

/* Batched form: one launch re-lays-out every conv weight of the model (a training step changes all of them).
 * `entries` (device) describes each parameter (taps <= 9); block b handles the 32 x 32 (co, ci) tile
 * co0 = 32 * (block_start[b] & 0xffff), ci0 = 32 * (block_start[b] >> 16) of entry block_entry[b], all taps. */
typedef struct {
  const float* w;      /* fp32 [cout][cin][taps] (reference layout)                     */
  void* fwd;           /* bf16 [taps][fwd_rows][fwd_cols], element (t, co+co_off, ci+ci_off) */
  void* bwd;           /* bf16 [taps][bwd_rows][bwd_cols], element (t, ci+ci_off, co+co_off), or NULL */
  int cout, cin, taps;
  int fwd_rows, fwd_cols, bwd_rows, bwd_cols;
  int co_off, ci_off;
  int pad_;
} eb200_pack_entry;
int eb200_pack_conv_weights_batched(const eb200_pack_entry* entries_dev, const int* block_entry_dev,
                                    const int* block_start_dev, int nblocks, void* stream);

/* ---- BatchNorm2d (MT/model/normalization.py:30-31 -> nn.BatchNorm2d, train mode) ------------------------------
 * eb200_conv2d(EB200_STATS) leaves per-channel sum / sum-of-squares in `stats`; finalize turns them into the affine
 * (scale = gamma*rstd, shift = beta - mean*scale), updates the running buffers (momentum, unbiased variance) when
 * they are given (track_running_stats), keeps mean/rstd for backward and zeroes `stats` for the next step. */
int eb200_bn_finalize(float* stats, long long count, const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* rstd,
                      int C, void* stream);
/* y[.., y_coff + c] = relu?((x*scale+shift) * drop[n,c] + res_pre) + res_post ; gap[n,c] += sum_hw y  (all optional)
 * Covers BN+ReLU, BN+Dropout2d+residual+ReLU (MT/model/block.py:201-221), and BN+ReLU+skip-add
 * (MT/model/encoder_decoder_fusion.py:85-87); `gap` is the SE squeeze (MT/model/utils.py:92). */
int eb200_bn_apply(const void* x, void* y, const float* scale, const float* shift, const float* drop,
                   const void* res_pre, const void* res_post, float* gap, int N, int HW, int C, int y_cs, int y_coff,
                   int relu, void* stream);
/* Train-mode form with the finalize folded in: every block derives scale/shift of its channels from the raw sums
 * (`stats`, [2C], left untouched — the caller zeroes its statistics arena once per step); block 0 publishes
 * scale/shift/mean/rstd for the backward pass and updates the running buffers (nn.BatchNorm2d, momentum, unbiased
 * variance).  One launch instead of finalize + apply. */
int eb200_bn_apply_train(const void* x, void* y, const float* stats, long long count, const float* gamma,
                         const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                         float* scale_out, float* shift_out, float* mean_out, float* rstd_out, const float* drop,
                         const void* res_pre, const void* res_post, float* gap, int N, int HW, int C, int y_cs,
                         int y_coff, int relu, void* stream);
/* backward of the above w.r.t. x: g = dy * relu_mask * drop; relu_mode 0 = no ReLU, 1 = mask_src > 0,
 * 2 = recompute (x*scale+shift) > 0.
 * reduce: per-block partial sums go to `partials` (workspace of >= (2*SMs + N) * 2C floats), a second tiny kernel
 *         combines them: sums[0:C] = sum g, sums[C:2C] = sum g*xhat, dbeta += sums[0:C], dgamma += sums[C:2C].
 * apply : dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)); dres (optional) = dy*relu_mask (residual branch). */
int eb200_bn_bwd_reduce(const void* dy, const void* x, const void* mask_src, const float* drop, const float* mean,
                        const float* rstd, const float* scale, const float* shift, float* partials,
                        long long partials_floats, float* sums, float* dgamma, float* dbeta, int N, int HW, int C,
                        int dy_cs, int dy_coff, int relu_mode, void* stream);
int eb200_bn_bwd_apply(const void* dy, const void* x, const void* mask_src, const float* drop, const float* mean,
                       const float* rstd, const float* scale, const float* shift, const float* gamma,
                       const float* sums, void* dx, void* dres, int N, int HW, int C, int dy_cs, int dy_coff,
                       int relu_mode, void* stream);
/* Replica form of the reduce (no partials workspace, no combine launch): `ws` is fp32 [(replicas+1)*2C + 4], ZERO on
 * entry.  Block sums are added to ws[block % replicas][2C]; the last block to finish folds the replicas into
 * ws[replicas][0:C] = sum g, ws[replicas][C:2C] = sum g*xhat and accumulates dbeta / dgamma.  Follow with
 * eb200_bn_bwd_apply(..., sums = ws + replicas*2C, ...). */
int eb200_bn_bwd_reduce_rep(const void* dy, const void* x, const void* mask_src, const float* drop, const float* mean,
                            const float* rstd, const float* scale, const float* shift, float* ws, int replicas,
                            float* dgamma, float* dbeta, int N, int HW, int C, int dy_cs, int dy_coff, int relu_mode,
                            void* stream);
/* reduce + apply of the BatchNorm backward in ONE launch: the replica reduce above, a grid-wide barrier (all blocks are
 * resident; otherwise this entry point falls back to the two launches), then every block re-reads its own pixel range
 * — still in L1/L2 — and writes dx / dres.  `ws` as for eb200_bn_bwd_reduce_rep (ZERO on entry). */
int eb200_bn_bwd_fused(const void* dy, const void* x, const void* mask_src, const float* drop, const float* mean,
                       const float* rstd, const float* scale, const float* shift, const float* gamma, float* ws,
                       int replicas, float* dgamma, float* dbeta, void* dx, void* dres, int N, int HW, int C, int dy_cs,
                       int dy_coff, int relu_mode, void* stream);
/* Second half of the BatchNorm backward fused into a data-gradient epilogue (EB200_BN_BWD): g is already ReLU-masked,
 * raw_sums = (sum g, sum g*x) as left by the conv; every block folds them (sum g*xhat = rstd*(sum gx - mean*sum g)),
 * block 0 accumulates dgamma / dbeta.  dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)). */
int eb200_bn_bwd_apply_raw(const void* g, const void* x, const float* mean, const float* rstd, const float* gamma,
                           const float* raw_sums, float* dgamma, float* dbeta, void* dx, int N, int HW, int C,
                           void* stream);
/* dgamma += sums[C:2C]; dbeta += sums[0:C]; sums = 0 */
int eb200_bn_bwd_param(float* sums, float* dgamma, float* dbeta, int C, void* stream);
/* out[c] += sum_p x[p*cs + coff + c]  (bias gradients) */
int eb200_colsum(const void* x, float* out, long long P, int C, int cs, int coff, void* stream);

/* ---- stem (MT/model/backbone/resnet.py:64-68) ------------------------------------------------------------------
 * im2col of the 7x7/s2/p3 window: fp32 NCHW -> bf16 [N,H/2,W/2,Kpad], k = c*49+ky*7+kx (the weight's own
 * flattening) so the stem conv runs as a 1x1 eb200_conv2d with cin = Cin*49. */
int eb200_im2col_stem(const float* in, void* out, int N, int Cin, int H, int W, int Kpad, void* stream);
/* nn.MaxPool2d(3,2,1), NHWC; idx (uint8, same shape as y) keeps the arg-max tap for backward */
int eb200_maxpool_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, void* stream);
int eb200_maxpool_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, void* stream);

/* ---- SE fusion 'se-add-uni-rgb' (MT/model/utils.py:84-95, MT/model/encoder_fusion.py:63-90) ------------------- */
int eb200_gap(const void* x, float* gap, int N, int HW, int C, void* stream);   /* gap[n,c] += sum_hw x */
/* mean = gap/HW (gap zeroed); hid = relu(W1 mean + b1); wgt = sigmoid(W2 hid + b2); fp32, reference layouts */
int eb200_se_mlp_fwd(float* gap, int HW, const float* w1, const float* b1, const float* w2, const float* b2,
                     float* mean, float* hid, float* wgt, int N, int C, int Cr, void* stream);
int eb200_se_mlp_bwd(float* dwgt, const float* wgt, const float* hid, const float* mean, int HW, const float* w1,
                     const float* w2, float* dw1, float* db1, float* dw2, float* db2, float* dmean, int N, int C,
                     int Cr, void* stream);
int eb200_se_fuse_fwd(const void* a, const void* b, const float* wa, const float* wb, void* out, int N, int HW, int C,
                      void* stream);                                             /* out = a*wa[n,c] + b*wb[n,c] */
int eb200_se_fuse_bwd_reduce(const void* dout, const void* a, const void* b, float* dwa, float* dwb, int N, int HW,
                             int C, void* stream);
int eb200_se_fuse_bwd_apply(const void* dout, const float* wa, const float* wb, const float* dmean_a,
                            const float* dmean_b, const void* db_prev, void* da, void* db, int N, int HW, int C,
                            void* stream);

/* ---- pyramid pooling (MT/model/context_module/ppm.py:57-78) ---------------------------------------------------- */
int eb200_adaptive_pool_fwd(const void* x, void* y, int N, int H, int W, int C, int B, void* stream);
int eb200_adaptive_pool_bwd(const void* dy, void* dx, int N, int H, int W, int C, int B, int accumulate, void* stream);
/* bilinear, align_corners=False; y / dy may be channel slices of a wider tensor (concat fusion) */
int eb200_bilinear_fwd(const void* x, void* y, int N, int Hi, int Wi, int Ho, int Wo, int C, int y_cs, int y_coff,
                       void* stream);
int eb200_bilinear_bwd(const void* dy, void* dx, int N, int Hi, int Wi, int Ho, int Wo, int C, int dy_cs, int dy_coff,
                       void* stream);

/* ---- learned upsampling 'learned-3x3-zeropad' (MT/model/upsampling.py:39-96): nearest x2 + depthwise 3x3 + bias,
 * fused.  w fp32 [Creal,1,3,3], b fp32 [Creal]; tensors have C >= Creal channels (C % 8 == 0, extra channels zero). */
int eb200_upsample_dw_fwd(const void* x, const float* w, const float* b, void* y, int N, int H, int W, int C,
                          int Creal, void* stream);
int eb200_upsample_dw_bwd_input(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int Creal,
                                void* stream);
int eb200_upsample_dw_bwd_weight(const void* dy, const void* x, float* dw, float* db, int N, int H, int W, int C,
                                 int Creal, void* stream);

/* ---- output boundary: NHWC bf16 <-> NCHW fp32 (reference output convention) -------------------
 * act_mode 0: plain copy of Creal channels into y0.  act_mode 1 (C == 8): instance head activations
 * (MT/model/decoder/instance.py:113-119): y0 = sigmoid(ch0), y1 = tanh(ch1..2), y2 = unit-length(ch3..4) or NULL. */
int eb200_nhwc_to_nchw(const void* x, float* y0, float* y1, float* y2, int N, int HW, int C, int Creal, int act_mode,
                       void* stream);
int eb200_nchw_to_nhwc_grad(const float* g0, const float* g1, const float* g2, const void* x, void* dx, int N, int HW,
                            int C, int Creal, int act_mode, void* stream);

/* ---- scene head nn.Linear (MT/model/decoder/scene.py:30,62-63); x bf16 [N,K], w fp32 [M,K] -------------------- */
int eb200_linear_fwd(const void* x, const float* w, const float* b, float* y, int N, int K, int M, void* stream);
int eb200_linear_bwd(const float* dy, const void* x, const float* w, void* dx, float* dw, float* db, int N, int K,
                     int M, void* stream);

/* ---- gradient fan-in / concat helpers ------------------------------------------------------------------------- */
int eb200_add_inplace(void* a, const void* b, long long n, void* stream);       /* a += b, bf16 */
int eb200_copy_channels(const void* src, void* dst, long long P, int C, int scs, int scoff, int dcs, int dcoff,
                        int accumulate, void* stream);

/* ---- inference post-processing ------------------------------------------------------------
 * What the reference runs on the network outputs in eval mode: MT/model/postprocessing/{semantic,instance,panoptic,
 * scene}.py and MT/utils/panoptic_merge.py.  Tensors are the reference's own (fp32 NCHW outputs, int64 index maps,
 * uint8 instance ids, bool/uint8 masks).  Per-instance quantities live in device tables with
 * EB200_PP_MAX_INSTANCES rows per image (row 0 = "no instance"; ids are uint8 like the reference's). */
#define EB200_PP_MAX_INSTANCES 256
#define EB200_PP_ACC_FIELDS 5   /* inst_acc[n][id][]: sum semantic score, pixels, sum cos, sum sin, oriented pixels */

/* softmax over C + (max, first arg-max) of the softmax values: semantic.py:55-56,73-74, scene.py:41-42 (H=W=1).
 * The crop window (y0,x0,Hc,Wc) is the valid region; if (Ho,Wo) != (Hc,Wc) the logits are resampled bilinearly
 * (align_corners=False) first, dense_base.py:15-58.  Optional outputs (NULL to skip): out_logits / scores fp32
 * [N,C,Ho,Wo], score fp32 [N,Ho,Wo], idx int64 [N,Ho,Wo], flag_out uint8 [N,Ho,Wo] = cls_flags[idx] & 1
 * (cls_flags uint8 [C]: bit 0 = thing class, bit 1 = class has orientation; panoptic.py:41-45,123-128). */
int eb200_pp_softmax_argmax(const float* logits, int N, int C, int H, int W, int y0, int x0, int Hc, int Wc, int Ho,
                            int Wo, float* out_logits, float* scores, float* score, long long* idx,
                            const unsigned char* cls_flags, unsigned char* flag_out, void* stream);
/* InstancePostprocessing._get_instance_centers (instance.py:74-160): heat fp32 [N,1,H,W]; threshold, k x k NMS with
 * the first-maximum tie rule, keep everything >= max(top_k-th value, 0); fg (uint8 [N,H,W] or NULL) =
 * heatmap_apply_foreground_mask.  Outputs: centers int32 [N,256,2] (y,x) in row-major order (row r = instance id
 * r+1), center_scores fp32 [N,256], counts int32 [N], status int32 [N] (bit 1: more than 255 centres). */
long long eb200_pp_centers_ws_bytes(int N, int H, int W, int nms_k);
int eb200_pp_instance_centers(const float* heat, int N, int H, int W, float threshold, int nms_k, int top_k,
                              const unsigned char* fg, void* ws, long long ws_bytes, int* centers,
                              float* center_scores, int* counts, int* status, void* stream);
/* InstancePostprocessing._get_instance_segmentation (instance.py:162-273): offset fp32 [N,2,H,W] (scale_y/scale_x
 * undo the normalisation, :357-363), fg uint8 [N,H,W]; seg uint8 [N,H,W] = 1 + first arg-min of the fp32 distance to
 * the centres (0 outside fg / beyond dist_thr >= 0); areas int32 [N,256] (bincount).  votes int32 [N,256,C+1]
 * (optional, needs sem_idx int64 [N,H,W]): semantic-label histogram of every instance for the panoptic merge. */
int eb200_pp_instance_assign(const float* offset, const unsigned char* fg, const int* centers, const int* counts, int N,
                             int H, int W, float scale_y, float scale_x, float dist_thr, const long long* sem_idx,
                             int n_classes, unsigned char* seg, int* areas, int* votes, void* stream);
/* deeplab_merge_batch (panoptic_merge.py:18-41,168-225) + the score / orientation statistics of
 * PanopticPostprocessing._postprocess_inference (panoptic.py:150-236,289-303).  inst_pan int32 [N,256] panoptic id of
 * every instance (class * 2^16 + per-class id), pan / pan_sem int64 [N,H,W]; with scores (softmax fp32 [N,C,H,W]):
 * sem_score / ins_score / pan_score fp32 [N,H,W]; orientation fp32 [N,2,H,W] or NULL; inst_acc fp64
 * [N,256,EB200_PP_ACC_FIELDS]. */
int eb200_pp_panoptic_merge(const unsigned char* seg, const long long* sem_idx, const unsigned char* cls_flags,
                            const int* counts, const int* votes, const float* scores, const float* orientation,
                            const float* center_scores, int N, int C, int H, int W, int* inst_pan, long long* pan,
                            long long* pan_sem, float* sem_score, float* ins_score, float* pan_score, double* inst_acc,
                            void* stream);
/* crop to the valid region + aten upsample_nearest2d (dense_base.py:15-58, mode='nearest'); elements of 1, 4 or 8
 * bytes, [N,H,W] -> [N,Ho,Wo] */
int eb200_pp_nearest_resize(const void* in, void* out, int elem_bytes, int N, int H, int W, int y0, int x0, int Hc,
                            int Wc, int Ho, int Wo, void* stream);

/* InstancePostprocessing._get_instance_orientation on an arbitrary instance map (instance.py:275-323; the dataset
 * evaluation variants on ground-truth maps, :431-451): acc fp64 [N, max_id+1, 3] += (cos, sin, 1) over the pixels with
 * seg == id in 1..max_id inside fg (uint8 / bool [N,H,W] or NULL); seg [N,H,W] of 1 / 2 / 4 / 8-byte integers.
 * NOT YET RUN ON A B200: the host side uses it only with EB200_PP_GT_ORIENTATION=1. */
int eb200_pp_instance_orientation(const float* orientation, const void* seg, int seg_bytes, const unsigned char* fg,
                                  int N, int H, int W, int max_id, double* acc, void* stream);

/* ---- fused semantic cross-entropy ---------------------------------------------------------
 * MT/loss/ce.py:13-68 (CrossEntropyLossSemantic, weighted_reduction=False) = torch.nn.CrossEntropyLoss(weight,
 * reduction='sum', ignore_index=-1, label_smoothing) on `target - 1`, and its autograd backward.  logits fp32
 * [N,C,H,W]; target [N,H,W] of 1 / 4 / 8-byte integers with 0 = void; weights fp32 [C] or NULL.
 * fwd: *loss_acc (fp64) = sum of the per-pixel losses, *count_acc (int64) = non-void pixels (both zeroed first).
 * bwd: dlogits fp32 [N,C,H,W] = *grad_out (a device scalar, the upstream gradient) * dloss/dlogits.
 * NOT YET RUN ON A B200 (round 1 ended without GPU time): parity unverified, see DESIGN.md §10. */
int eb200_ce_loss_fwd(const float* logits, const void* target, int target_bytes, const float* weights,
                      float label_smoothing, int N, int C, int H, int W, double* loss_acc, long long* count_acc,
                      void* stream);
int eb200_ce_loss_bwd(const float* logits, const void* target, int target_bytes, const float* weights,
                      float label_smoothing, const float* grad_out, int N, int C, int H, int W, float* dlogits,
                      void* stream);

/* The LAST learned upsampling of a task head fused with the fp32 NCHW output boundary (semantic head:
 * MT/model/decoder/semantic.py:62-75 -> MT/model/upsampling.py:85-96 -> the tensors EMSANet.forward returns,
 * emsanet/model.py:192-233).  fwd: x bf16 NHWC [N,H,W,C] -> y fp32 NCHW [N,Creal,2H,2W] (values rounded to bf16, as the
 * bf16 activation of the unfused path).  bwd: g fp32 NCHW [N,Creal,2H,2W] read ONCE -> dx bf16 NHWC [N,H,W,C]
 * (padding channels zero), dw [Creal][9] and db [Creal] accumulated. */
int eb200_upsample_dw_fwd_nchw(const void* x, const float* w, const float* b, float* y, int N, int H, int W, int C,
                               int Creal, void* stream);
int eb200_upsample_dw_bwd_nchw(const float* g, const void* x, const float* w, void* dx, float* dw, float* db, int N,
                               int H, int W, int C, int Creal, void* stream);

/* Masked regression losses of the instance / orientation task, one pass each way, no host synchronisation.
 * Replaces (per scale) MSELoss / L1Loss / VonMisesLossBiternion._compute_loss on mask-multiplied predictions and the
 * `mask.sum().cpu().detach().item()` element counts: MT/loss/mse.py:23-41, MT/loss/l1.py:23-41, MT/loss/vonmises.py:29-51,
 * MT/task_helper/instance.py:118-207 (host syncs at :138-139, :161-162, :199-201).
 *   kind 0: loss = sum_p 1/C sum_c (pred_pc * m_p - target_pc)^2      kind 1: ... |pred_pc * m_p - target_pc|
 *   kind 2: loss = sum_{p: m_p} 1 - exp(kappa * (sum_c pred_pc * target_pc - 1))
 *   count = sum_p m_p (mask: one byte per (n, p), NULL = all ones).  Element (n, c, p) at n*sn + c*sc + p*sp (elements).
 * bwd writes dpred (same addressing) = grad_out[0] * dloss/dpred; grad_out is a DEVICE scalar. */
int eb200_masked_loss_fwd(int kind, const float* pred, const float* target, const void* mask, int N, int C,
                          long long P, long long sn, long long sc, long long sp, float kappa, double* loss_acc,
                          long long* count_acc, void* stream);
int eb200_masked_loss_bwd(int kind, const float* pred, const float* target, const void* mask, int N, int C,
                          long long P, long long sn, long long sc, long long sp, float kappa, const float* grad_out,
                          float* dpred, void* stream);

/* ---- GPU-side input normalisation ------------------------------------------------------
 * NormalizeRGB / NormalizeDepth + ToTorchTensors of the reference's data pipeline (MT/data/preprocessing/
 * normalize.py:14-124; `(value - mean) / std` in float32, raw depth keeps `invalid_value`) on raw images of a batch:
 * rgb uint8 NHWC [N,H,W,3] -> fp32 NCHW [N,3,H,W]; depth uint16 (elem_bytes 2) or int32 (4) [N,H,W] -> fp32 [N,1,H,W].
 * mean3 / std3 are HOST arrays.  Bit-identical to the reference's numpy arithmetic. */
int eb200_normalize_rgb(const void* rgb_u8_nhwc, float* out_nchw, int N, int H, int W, const float* mean3,
                        const float* std3, void* stream);
int eb200_normalize_depth(const void* depth, int elem_bytes, float* out, int N, int H, int W, float mean, float std,
                          int raw_depth, float invalid_value, void* stream);

/* ---- fused optimizer step + weight re-layout ---------------------------------------------
 * Replaces torch.optim.{SGD(momentum, nesterov=True), Adam, AdamW}.step() as emsanet/optimizer.py:29-59 builds them
 * (called at main.py:599) AND eb200_pack_conv_weights_batched: one launch updates all parameters from their
 * gradients (fp32 master parameter, momentum / moment buffers in place, torch's arithmetic operation by operation) and
 * writes the bf16 tensor-core layouts of the conv weights from the updated values.
 * `hyper` is a HOST struct, passed to the kernel by value (lr changes per epoch, step per call).
 * Block b works on entry block_entry[b]: plain parameter -> elements [chunk*block_start[b], +chunk), chunk =
 * eb200_optim_chunk(); conv weight (pack >= 0) -> 32 x 32 (co, ci) tile like eb200_pack_conv_weights_batched. */
enum { EB200_OPT_SGD = 0, EB200_OPT_ADAM = 1, EB200_OPT_ADAMW = 2 };
typedef struct {
  float* p;            /* fp32 master parameter, reference layout (the nn.Parameter's storage)      */
  const float* g;      /* its gradient                                                              */
  float* m;            /* momentum buffer / exp_avg (NULL: SGD without momentum)                    */
  float* v;            /* exp_avg_sq (Adam, AdamW) or NULL                                          */
  long long numel;
  int pack;            /* index into the pack entries if this is a tensor-core conv weight, else -1 */
  int pad_;
} eb200_optim_entry;
typedef struct {
  int kind;            /* EB200_OPT_*                                                               */
  float lr, momentum /* = beta1 for Adam */, beta2, eps, weight_decay;
  float bias_correction1, bias_correction2_sqrt;   /* 1 - beta1^step, sqrt(1 - beta2^step)           */
  int nesterov, step /* 1-based: step 1 initialises the state */, flags /* 0; bits 0-2: testing only */;
  /* Adam / AdamW: the scalars torch.optim computes in Python doubles before they reach its kernels as floats */
  float one_minus_beta1, one_minus_beta2, step_size /* lr / bias_correction1 */, decay /* 1 - lr * weight_decay */;
  int pad_;
} eb200_optim_hyper;
int eb200_optim_chunk(void);
int eb200_optim_step(const eb200_optim_entry* entries_dev, const eb200_pack_entry* packs_dev,
                     const int* block_entry_dev, const int* block_start_dev, int nblocks,
                     const eb200_optim_hyper* hyper, void* stream);

/* cudaMemsetAsync(p, 0, bytes) on `stream`: zero-initialised scratch cleared on the stream its consumer runs on */
int eb200_memset_zero(void* p, long long bytes, void* stream);

const char* eb200_last_error(void);
int eb200_version(void);
/* number of kernels launched by this library on the calling process since load (for gpu_launches) */
long long eb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* EMSANET_B200_H_ */
