/*
 * hgr_b200 - C ABI of the B200-native MultiTaskNet forward path.
 *
 * The reference (yingkunwu/hand-gesture-recognition) is pure Python: there is
 * no FFI in it to bind against.  The boundary this library sits behind is the
 * Python class model.multitasknet.MultiTaskNet (reference
 * model/multitasknet.py:8-29) plus libs.utils.get_max_preds (libs/utils.py:4-32)
 * and the crop normalisation of detect.py:106-112.  The host-side mirror of
 * those interfaces lives in hand-gesture-recognition_b200/hgr_b200/ and calls
 * the entry points below through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer, h_* a HOST pointer;
 *   - the caller owns every buffer; nothing here allocates device memory
 *     except hgr_forward_host's internal staging (allocated once per plan);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it
 *     unless stated otherwise;
 *   - return 0 on success, negative on error; hgr_last_error() describes the
 *     last failure of the calling thread;
 *   - dtype codes: HGR_F32 = 0, HGR_BF16 = 1;
 *   - activations between kernels are NHWC bf16; module inputs are NCHW
 *     (fp32 or bf16) and module outputs NCHW / row-major like the reference's.
 */
#ifndef HGR_B200_H_
#define HGR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HGR_API __attribute__((visibility("default")))
#else
#define HGR_API
#endif

#define HGR_F32 0
#define HGR_BF16 1

#define HGR_ACT_NONE 0
#define HGR_ACT_SILU 1
#define HGR_ACT_GELU 2

HGR_API int hgr_version(void);
HGR_API const char* hgr_last_error(void);

/* ------------------------------------------------------------------------
 * Whole-network plan: MultiTaskNet(num_joints, num_classes, [S, S]).forward
 * (reference model/multitasknet.py:24-29).
 * ---------------------------------------------------------------------- */
typedef struct hgr_plan hgr_plan_t;

/* Packed-parameter block: BN-folded bf16 weights in [Cout][tap][Cin] order,
 * fp32 scale/shift/bias vectors, the bf16 sin-cos position table.  The layout
 * is owned by the library; the host packs a state_dict into it entry by entry. */
HGR_API int hgr_param_count(int image_size, int num_joints, int num_classes);
/* Entry i: name (e.g. "encoder.cspelan1.cv2.0.cv1.w"), byte offset, byte size,
 * dtype code and up to 3 dims (unused dims = 1). */
HGR_API int hgr_param_info(int image_size, int num_joints, int num_classes, int index, const char** name, size_t* offset,
                   size_t* nbytes, int* dtype, int64_t dims[3]);
HGR_API size_t hgr_param_bytes(int image_size, int num_joints, int num_classes);

HGR_API size_t hgr_workspace_bytes(int image_size, int batch);

/* Binds a plan to a packed-parameter block and a workspace (both device
 * memory, 1024-byte aligned, owned by the caller, alive as long as the plan). */
HGR_API int hgr_plan_create(hgr_plan_t** out, int image_size, int num_joints, int num_classes, int batch, void* d_params,
                    void* d_workspace, size_t workspace_bytes);
HGR_API void hgr_plan_destroy(hgr_plan_t* plan);

/* forward(x) -> (class logits, pose heatmaps, last-layer attention).
 *   d_x        (B, 3, S, S) NCHW, x_dtype
 *   d_logits   (B, num_classes)            out_dtype
 *   d_heatmaps (B, num_joints, S/4, S/4)   out_dtype
 *   d_attn     (B, 8, T, T) out_dtype or NULL to skip materialising it
 *   batch      must equal the plan's batch (tile grids and tensor maps are built for it) */
HGR_API int hgr_forward(hgr_plan_t* plan, const void* d_x, int x_dtype, int batch, void* d_logits, void* d_heatmaps,
                void* d_attn, int out_dtype, void* stream);

/* The forward with the keypoint decode of libs/utils.py:4-32 fused into the pose head (reference detect.py:143-150
 * runs the classifier and then get_max_preds on its heatmaps): d_preds (B, J, 2) fp32 [x, y] and d_maxvals (B, J, 1)
 * fp32 are bit-identical to hgr_get_max_preds applied to the heatmaps this call would write.  d_heatmaps may be
 * NULL: the heatmaps then never reach HBM (only with the tcgen05 pose head, i.e. image_size <= 320). */
/* 1 when hgr_forward_keypoints can run with d_heatmaps == NULL for this image size (0 otherwise, -1 on a bad size). */
HGR_API int hgr_keypoints_fused(int image_size, int num_joints);

HGR_API int hgr_forward_keypoints(hgr_plan_t* plan, const void* d_x, int x_dtype, int batch, void* d_logits,
                                  void* d_heatmaps, float* d_preds, float* d_maxvals, int out_dtype, void* stream);

/* Same call with HOST buffers: pinned or pageable input is copied to the
 * device, the forward runs, logits and heatmaps are copied back; returns after
 * the results are in host memory.  This is the end-to-end path bench.py times. */
HGR_API int hgr_forward_host(hgr_plan_t* plan, const void* h_x, int x_dtype, int batch, void* h_logits, void* h_heatmaps,
                     int out_dtype, void* stream);

/* Device address and NHWC dims of a named intermediate ("a1", "a2", "g1",
 * "o1", "d1", "g2", "o2", "d2", "g3", "o3", "tokens", "qkv", ...), for the
 * per-stage parity tests.  dims = (N, H, W, C). */
HGR_API int hgr_plan_buffer(hgr_plan_t* plan, const char* name, void** d_ptr, int64_t dims[4]);

/* Number of kernel launches one hgr_forward issues (for bench.py's gpu_launches). */
HGR_API int hgr_plan_launches(hgr_plan_t* plan, int with_attn);

/* Launch i of the sequence: layer name, kind (0 = tcgen05 implicit GEMM, 1 = mma.sync kernel,
 * 2 = memory-bound kernel) and its algorithmic FLOPs / bytes (unpadded, bf16 activations). */
HGR_API int hgr_plan_launch_info(hgr_plan_t* plan, int index, const char** name, int* kind, double* flops, double* bytes);

/* hgr_forward with a CUDA event between consecutive launches: h_ms[i] receives launch i's
 * duration in milliseconds (capacity >= hgr_plan_launches).  Synchronises; returns the count. */
HGR_API int hgr_forward_profile(hgr_plan_t* plan, const void* d_x, int x_dtype, int batch, void* d_logits, void* d_heatmaps,
                        void* d_attn, int out_dtype, void* stream, float* h_ms, int capacity);

/* ------------------------------------------------------------------------
 * Single operators (the building blocks the plan chains; exported so the
 * parity tests can exercise every kernel against the oracle in isolation).
 * ---------------------------------------------------------------------- */

/* Conv(c1, c2, k, s) + folded BatchNorm + activation (+ residual before the
 * activation) - reference model/gelan.py:18-56 and ResBasicBlock :78-87.
 *   d_in   NHWC bf16 (B, H, W, in_ctot); channels [in_coff, in_coff+cin) are read
 *   d_w    bf16 [cout][k*k][cin]
 *   d_out  NHWC bf16 (B, H/s, W/s, out_ctot); channels [out_coff, out_coff+cout) written
 *   d_res  NHWC bf16 (B, H/s, W/s, res_ctot) or NULL; channels [res_coff, ..+cout)
 * k in {1, 3}; s in {1, 2} (s = 2 needs k = 3, even H and W); cin % 64 == 0;
 * cout % 64 == 0. */
HGR_API int hgr_conv_bn_act(const void* d_in, int B, int H, int W, int in_ctot, int in_coff, int cin, const void* d_w,
                    const float* d_scale, const float* d_shift, int k, int s, int act, const void* d_res,
                    int res_ctot, int res_coff, void* d_out, int out_ctot, int out_coff, int cout, void* stream);

/* encoder.conv2 -> encoder.cspelan1.cv1 as one kernel (reference model/gelan.py:156 Conv(64, 128, 3, 2) and :127
 * GELANBlock.cv1 = Conv(128, 128, 1, 1), each conv + folded BatchNorm + SiLU): the 128-channel tensor between the
 * two layers stays in shared memory, rounded to bf16 where the separate launches store it.
 *   d_in   NHWC bf16 (B, H, W, 64), H and W even;  d_w1 bf16 [128][9][64];  d_w2 bf16 [128][128]
 *   d_out  NHWC bf16 (B, H/2, W/2, out_ctot); channels [out_coff, out_coff + 128) written */
HGR_API int hgr_conv_chain(const void* d_in, int B, int H, int W, const void* d_w1, const float* d_scale1,
                           const float* d_shift1, const void* d_w2, const float* d_scale2, const float* d_shift2,
                           void* d_out, int out_ctot, int out_coff, void* stream);

/* encoder.conv1 -> encoder.conv2 -> encoder.cspelan1.cv1 as one kernel (reference model/gelan.py:155 Conv(3, 64, 3, 2),
 * :156 Conv(64, 128, 3, 2), :127 GELANBlock.cv1 = Conv(128, 128, 1, 1), each conv + folded BatchNorm + SiLU): neither
 * the 64-channel nor the 128-channel tensor between the layers reaches HBM; both are rounded to bf16 where the
 * separate launches (hgr_conv1, hgr_conv_chain) store them.
 *   d_x    NCHW bf16 (B, 3, S, S), S a multiple of 64, 16-byte aligned
 *   d_w0   bf16 [64][32] as for hgr_conv1 (BN scale folded in), d_shift0 fp32 [64]
 *   d_w1   bf16 [128][9][64];  d_w2 bf16 [128][128]
 *   d_out  NHWC bf16 (B, S/4, S/4, out_ctot); channels [out_coff, out_coff + 128) written */
HGR_API int hgr_stem_fused(const void* d_x, int B, int S, const void* d_w0, const float* d_shift0, const void* d_w1,
                           const float* d_scale1, const float* d_shift1, const void* d_w2, const float* d_scale2,
                           const float* d_shift2, void* d_out, int out_ctot, int out_coff, void* stream);

/* The tail of the first GELAN block as one kernel (reference model/gelan.py:73-87 ResBasicBlock.forward, second conv +
 * residual + SiLU, and :137-142 GELANBlock.forward's cv4 over the concatenation): y3 = SiLU(BN_h(conv3x3(t)) + y2) is
 * rounded to bf16 where the separate launches store it and stays on the SM as an operand of the 1x1 layer.
 *   d_t    NHWC bf16 (B, H, W, 64), H % 16 == 0, W % 8 == 0;  d_wh bf16 [64][9][64]
 *   d_g    NHWC bf16 (B, H, W, 256): y0 | y1 | y2 in channels 0..191 (y2 is also the residual); 192..255 untouched
 *   d_w4   bf16 [128][256];  d_out NHWC bf16 (B, H, W, 128) */
HGR_API int hgr_gelan_tail(const void* d_t, const void* d_g, int B, int H, int W, const void* d_wh,
                           const float* d_scale_h, const float* d_shift_h, const void* d_w4, const float* d_scale4,
                           const float* d_shift4, void* d_out, void* stream);

/* y = act(scale (.) (x W^T) + bias) (+ residual): nn.Linear of the ViT
 * (reference model/transformer.py:34,37,65,75).  x (rows, cin) bf16,
 * W (cout, cin) bf16, y (rows, cout) bf16; d_scale / d_bias / d_res nullable.
 * LayerNorm folding (cout % 256 == 0, model/transformer.py:33,63 fused into the
 * GEMMs around it):
 *   d_row_stats_out  (rows, 2) fp32: the epilogue also writes (mean, rstd) of every
 *                    output row (cout == 256, residual given, no activation);
 *   d_row_stats_in   (rows, 2) fp32 statistics of the rows of x: with gamma folded
 *                    into W (W' = gamma (.) W), d_scale[n] = sum_k W'[n,k] and
 *                    d_bias[n] = sum_k beta[k] W[n,k] + b[n], the result is
 *                    act(LayerNorm(x) W^T + b). */
HGR_API int hgr_linear(const void* d_x, long long rows, int cin, const void* d_w, const float* d_scale,
                       const float* d_bias, int act, const void* d_res, void* d_y, int cout,
                       const float* d_row_stats_in, float* d_row_stats_out, void* stream);

/* Second half of a ViT layer as one chained kernel (reference model/transformer.py:75 to_out, :93 residual,
 * :29-42 FeedForward, :94 residual):
 *     x1 = attn_out W_out^T + x0;   h = GELU(LayerNorm(x1) W1^T + b1);   x2 = h W2^T + b2 + x1
 * attn_out, x0, x2 (rows, 256) bf16 (x2 may alias x0); W_out, W1', W2 (256, 256) bf16 with the LayerNorm folded
 * into W1' = gamma (.) W1, d_c1[n] = sum_k W1'[n,k], d_d1[n] = sum_k beta[k] W1[n,k] + b1[n] (as in hgr_linear);
 * d_row_stats_out (rows, 2) fp32, nullable: (mean, rstd) of every x2 row for the next layer's folded LayerNorm.
 * x1 and h are rounded to bf16 exactly where the three separate hgr_linear launches store them. */
HGR_API int hgr_vit_block(const void* d_attn_out, const void* d_x0, long long rows, const void* d_w_out,
                          const void* d_w1, const float* d_c1, const float* d_d1, const void* d_w2,
                          const float* d_b2, void* d_x2, float* d_row_stats_out, void* stream);

/* Profiling twin of hgr_vit_block: CTA 0 also writes clock64 marks of its first `trace_tiles` tiles into
 * d_trace[trace_tiles][16] (int64): 0 G0 done, 1 E0 done, 2 G1 done, 3 E1 done, 4 G2 done (all as seen by the
 * first epilogue thread), 5 last output chunk's store issued, 6 epilogue back at the top of the chain, 7 first x0
 * chunk landed, 8..10 the MMA warp starts G0/G1/G2, 11..14 weights of G0's four k-blocks are in shared memory
 * (tools/vit_block_trace.py prints them as cycle deltas). */
HGR_API int hgr_vit_block_trace(const void* d_attn_out, const void* d_x0, long long rows, const void* d_w_out,
                                const void* d_w1, const float* d_c1, const float* d_d1, const void* d_w2,
                                const float* d_b2, void* d_x2, float* d_row_stats_out, long long* d_trace,
                                int trace_tiles, void* stream);

/* encoder.conv1: NCHW (fp32|bf16) -> NHWC bf16 (B, S/2, S/2, 64).
 * d_w bf16 [64][32] (k = (kh*3+kw)*3+c, BN scale folded, zero padded). */
HGR_API int hgr_conv1(const void* d_x, int x_dtype, int B, int S, const void* d_w, const float* d_shift, void* d_out,
              void* stream);

/* nn.LayerNorm(256), eps 1e-5: (rows, 256) bf16 -> bf16. */
HGR_API int hgr_layernorm(const void* d_x, void* d_y, const float* d_gamma, const float* d_beta, long long rows, void* stream);

/* softmax(q k^T / sqrt(32)) v over 8 heads of 32: d_qkv (B, T, 768) bf16 ->
 * d_out (B, T, 256) bf16; d_probs (B, 8, T, T) probs_dtype or NULL. */
HGR_API int hgr_attention(const void* d_qkv, void* d_out, void* d_probs, int probs_dtype, int B, int T, void* stream);

/* The same attention core on tcgen05 tensor cores (csrc/attention_tc.cu): scores, probabilities and outputs in
 * TMEM, P as the TMEM A operand of P V, V as an MN-major operand.  129 <= T <= 160, no probability output.  The
 * forward plan uses it for every attention launch that returns no probabilities (HGR_ATTN_TC=0 falls back to
 * hgr_attention's mma.sync kernels); exported so that the parity suite covers it directly. */
HGR_API int hgr_attention_tc(const void* d_qkv, void* d_out, int B, int T, void* stream);

/* Development aid: the same launch with a cycle timeline of CTA 0.  d_trace receives clock64 marks as
 * [trace_items][*warps][8] long long (see AttnTcParams in csrc/attention_tc.cu for the marks);
 * d_trace == NULL only reports *warps. */
HGR_API int hgr_attention_tc_trace(const void* d_qkv, void* d_out, int B, int T, long long* d_trace, int trace_items,
                                   int* warps, void* stream);

/* mlp_head: Linear(256, C)(LayerNorm(tokens[:, 0])). */
HGR_API int hgr_cls_head(const void* d_tokens, const float* d_gamma, const float* d_beta, const float* d_w,
                 const float* d_bias, void* d_logits, int out_dtype, int B, int T, int num_classes, void* stream);

/* bilinear x4 (align_corners) + ReLU + 1x1 conv + bias on tokens[:, 1:]:
 * d_tokens (B, F*F+1, 256) bf16, d_w bf16 [J][256] -> (B, J, 4F, 4F). */
HGR_API int hgr_pose_head(const void* d_tokens, const void* d_w, const float* d_bias, void* d_heatmaps, int out_dtype, int B,
                  int F, int J, void* stream);

/* The pose head with the keypoint decode fused into its epilogue (see hgr_forward_keypoints); d_heatmaps may be NULL. */
HGR_API int hgr_pose_head_decode(const void* d_tokens, const void* d_w, const float* d_bias, void* d_heatmaps,
                                 int out_dtype, float* d_preds, float* d_maxvals, int B, int F, int J, void* stream);

/* libs.utils.get_max_preds (reference libs/utils.py:4-32):
 * heatmaps (B, J, H, W) -> preds (B, J, 2) fp32 [x, y], maxvals (B, J, 1) fp32. */
HGR_API int hgr_get_max_preds(const void* d_heatmaps, int dtype, int B, int J, int H, int W, float* d_preds,
                      float* d_maxvals, void* stream);

/* detect.py:106-112: (B, H, W, 3) uint8 -> (B, 3, H, W) out_dtype,
 * ((v / 255) - mean[c]) / std[c], ImageNet constants by channel index. */
HGR_API int hgr_crop_normalize(const uint8_t* d_hwc, void* d_chw, int out_dtype, int B, int H, int W, void* stream);

/* libs.metrics.pose_accuracy (reference libs/metrics.py:31-62) after the keypoint decode: d_pred / d_target are
 * the (B, J, 2) fp32 outputs of hgr_get_max_preds on the predicted and the ground-truth heatmaps of size H x W.
 * d_acc: J + 1 doubles (acc[0] = average over the joints with at least one valid sample, acc[j + 1] per joint, -1
 * when no sample is valid); d_avg_cnt: {avg_acc, cnt}; d_counts: 2 * J ints of scratch.  thr is 0.5 in the reference. */
HGR_API int hgr_pose_accuracy(const float* d_pred, const float* d_target, int B, int J, int H, int W, double thr,
                              int* d_counts, double* d_acc, double* d_avg_cnt, void* stream);

/* detect.py:92-117 fused: cv2.warpAffine(frame, trans, (S, S), flags=INTER_LINEAR) (bit-exact fixed-point
 * arithmetic of OpenCV, constant border 0) + the normalisation above, for N crops out of F frames
 * (F, Hf, Wf, 3) uint8.  d_inv_mats: N x 6 doubles, the INVERTED affine maps (crop pixel -> frame pixel) exactly
 * as cv::warpAffine derives them from `trans`; d_frame_index: N ints.  Output (N, 3, S, S). */
HGR_API int hgr_crop_warp_normalize(const uint8_t* d_frames, int F, int Hf, int Wf, const int* d_frame_index,
                                    const double* d_inv_mats, int N, int S, void* d_chw, int out_dtype, void* stream);

/* ------------------------------------------------------------------------
 * Training step (BASELINE.json configs[4]; reference train.py:58-108 with
 * libs/loss.py and torch.optim.AdamW).  Parameters, gradients and BatchNorm
 * running statistics are three flat fp32 device blocks whose layout follows
 * the reference's state_dict order; the host mirror (hgr_b200/training.py)
 * makes the module's nn.Parameters views of the parameter block.
 * ---------------------------------------------------------------------- */
typedef struct hgr_train_plan hgr_train_plan_t;

/* Flat parameter / gradient block: entry i = state_dict key, offset and size in
 * floats (every entry starts on a 256-byte boundary). */
HGR_API int hgr_train_param_count(int num_joints, int num_classes);
HGR_API int hgr_train_param_info(int num_joints, int num_classes, int index, const char** name, size_t* offset_floats,
                                 size_t* numel);
HGR_API size_t hgr_train_param_floats(int num_joints, int num_classes);
/* Flat BatchNorm running-statistics block (running_mean / running_var of the 22 Conv blocks). */
HGR_API int hgr_train_bnstat_count(void);
HGR_API int hgr_train_bnstat_info(int index, const char** name, size_t* offset_floats, size_t* numel);
HGR_API size_t hgr_train_bnstat_floats(void);

HGR_API size_t hgr_train_workspace_bytes(int image_size, int num_joints, int num_classes, int batch);
/* d_pos_embedding_bf16: the (F*F, 256) sin-cos table of transformer.py:9-26 in bf16. */
HGR_API int hgr_train_plan_create(hgr_train_plan_t** out, int image_size, int num_joints, int num_classes, int batch,
                                  float* d_params, float* d_grads, float* d_bnstats, const void* d_pos_embedding_bf16,
                                  void* d_workspace, size_t workspace_bytes);
HGR_API void hgr_train_plan_destroy(hgr_train_plan_t* plan);
/* Named workspace buffer (activations "a1".."o3", "z.<conv>", "x0".."x4", "l<k>.probs", gradients "d_<act>" ...). */
HGR_API int hgr_train_buffer(hgr_train_plan_t* plan, const char* name, void** d_ptr, size_t* nbytes, int64_t dims[4]);

/* MultiTaskNet.forward under .train(): batch-statistics BatchNorm (running statistics updated with `momentum`,
 * unbiased variance; momentum < 0 leaves them untouched), fp32 logits (B, C) and heatmaps (B, J, S/4, S/4). */
HGR_API int hgr_train_forward(hgr_train_plan_t* plan, const void* d_x, int x_dtype, float* d_logits, float* d_heatmaps,
                              float momentum, void* stream);
/* Backward of the last hgr_train_forward of this plan (same d_x): overwrites the whole gradient block.  A few kernels
 * run on a stream the plan owns, forked from and joined back into `stream` by events inside the call, so the call is
 * still ordered on `stream` alone and can be captured into a CUDA graph. */
HGR_API int hgr_train_backward(hgr_train_plan_t* plan, const void* d_x, int x_dtype, const float* d_dlogits,
                               const float* d_dheatmaps, void* stream);
/* The same backward in three parts, to be called in order 0, 1, 2 (train.py:58-108 under DistributedDataParallel
 * overlaps the gradient all-reduce with the backward; the flat gradient block is in state_dict order, so each part
 * completes one contiguous range of it): 0 = heads, transformer, proj (parameters from "proj.weight" to the end),
 * 1 = encoder.cspelan3 and encoder.down2 (from "encoder.down2.conv.weight" to "proj.weight"), 2 = the rest of the
 * backbone (the start of the block). */
HGR_API int hgr_train_backward_part(hgr_train_plan_t* plan, const void* d_x, int x_dtype, const float* d_dlogits,
                                    const float* d_dheatmaps, int part, void* stream);

/* train.py:63-64 / libs/loss.py: total = cls_weight * CrossEntropy(logits, labels) + JointsMSELoss(heatmaps, target,
 * target_weight).  d_loss3 = {total, weighted class loss, joints loss}; d_dlogits / d_dheatmaps (nullable) receive
 * d total / d logits and d total / d heatmaps.  d_scratch: >= 592 floats.  labels are int64. */
HGR_API int hgr_loss(const float* d_logits, const float* d_heatmaps, const long long* d_labels, const float* d_target,
                     const float* d_target_weight, int B, int J, int C, int hw, float cls_weight, float* d_dlogits,
                     float* d_dheatmaps, float* d_scratch, float* d_loss3, void* stream);

/* torch.optim.AdamW (train.py:50-51) over a flat block; grad_scale multiplies the gradient first
 * (1 / world_size after an all-reduce SUM).  step counts from 1. */
HGR_API int hgr_adamw_step(float* d_params, const float* d_grads, float* d_exp_avg, float* d_exp_avg_sq, long long n,
                           float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                           void* stream);

/* Single kernels of the training step (parity tests).
 * hgr_wgrad: dW (cout, cin, k, k) fp32 = sum over output pixels of g[p, :cout]^T x[shift(p), :cin]; g is the
 *   (B*Ho*Wo, g_ctot) bf16 output gradient, x the (B, H, W, x_ctot) bf16 NHWC input, stride s, padding k/2;
 *   d_partial: hgr_wgrad_partial_floats(...) floats of scratch.  cin and cout must be multiples of 64.  Runs on tcgen05
 *   (MN-major operands, split over pixel chunks, fixed-order reduce); HGR_WGRAD_TC=0 selects the mma.sync kernel. */
HGR_API int hgr_wgrad(const void* d_g, int g_ctot, const void* d_x, int x_ctot, int B, int H, int W, int cin, int cout,
                      int k, int s, float* d_partial, float* d_dw, void* stream);
HGR_API size_t hgr_wgrad_partial_floats(int cout, int cin, int k, long long pixels);
/* qkv (B, T, 768), probs (B, 8, T, T), o / do (B, T, 256) -> dqkv (B, T, 768), all bf16 */
HGR_API int hgr_attention_bwd(const void* d_qkv, const void* d_probs, const void* d_o, const void* d_do, void* d_dqkv,
                              int B, int T, void* stream);
/* One parity class (ph, pw) of the input gradient of a 3x3 stride-2 convolution: dz (B, H/2, W/2, cout_fwd) bf16,
 * d_w_parity [cin_fwd][ntaps][cout_fwd] bf16, writes dx[:, ph::2, pw::2, :] of a (B, H, W, cin_fwd) bf16 buffer. */
HGR_API int hgr_dgrad_s2(const void* d_dz, int B, int H, int W, int cout_fwd, const void* d_w_parity, int ph, int pw,
                         void* d_dx, int cin_fwd, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* HGR_B200_H_ */
