// Internal declarations shared by the CUDA translation units of libhgr_b200.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace hgr {

// ----------------------------------------------------------- errors ----
void set_error(const char* fmt, ...);
const char* last_error();

#define HGR_CHECK_CUDA(expr)                                                                  \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::hgr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                              \
    }                                                                                         \
  } while (0)

// ---------------------------------------------------------- launch ----
// Programmatic dependent launch: the kernel may be scheduled while the previous kernel of the
// stream drains, so its prologue (barrier init, TMEM allocation, tensor-map prefetch, affine
// tables) and the launch latency overlap the predecessor's tail.  Every kernel launched this way
// executes griddepcontrol.wait (ptx.cuh: pdl_wait) before it touches memory another kernel wrote.
// Measured on B200 (gpurun r1n, batch 1024): 7.06 ms/step with it against 6.88 ms without - the early-scheduled
// CTAs of the next grid take the SM slots the persistent kernels' stragglers are about to free and gain nothing
// back - so it is off (fixed in common.cu; the launchers keep the attribute code).
bool pdl_enabled();
// launches made while a PdlScope(true) is alive on this thread carry the programmatic-serialization attribute
struct PdlScope {
  explicit PdlScope(bool on);
  ~PdlScope();
  PdlScope(const PdlScope&) = delete;
  PdlScope& operator=(const PdlScope&) = delete;

 private:
  bool prev_;
};
bool train_pdl_enabled();  // HGR_TRAIN_PDL (default off: measured 8 % slower)
// HGR_ZIGZAG=0 disables the alternating tile order of plan.cu.
bool zigzag_enabled();
// HGR_CLUSTER=0 disables the CTA-pair (cta_group::2) mode of the implicit-GEMM kernel.
bool cluster_enabled();
// HGR_ATTN_ONLINE=0 falls back to the strip-in-registers attention kernels when no probabilities are returned.
bool attention_online_enabled();
// HGR_ATTN_TC=0: the 129..160-token attention without probability output falls back from the tcgen05 kernel
// (attention_tc.cu, the default) to the mma.sync kernels of attention.cu.
bool attention_tc_enabled();
// token counts the attention kernels can hold in one CTA's shared memory (checked when a plan is created)
int attention_max_tokens();
bool attention_tokens_supported(int T);
bool attention_tc_supported(int T);
int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, float scale_log2e, int num_sms,
                        cudaStream_t stream, int reverse, long long* trace = nullptr, int trace_items = 0);
int attention_tc_warps();
// fixed to true (common.cu): the online-softmax attention kernel stages Q, K, V by cp.async, not through registers
bool attention_cp_async_enabled();
// fixed to 1 (common.cu): query tiles a warp of the online-softmax attention kernel works on at once
int attention_tiles_per_warp();
// fixed to true (common.cu): one accumulator-release arrive per epilogue warp instead of one per thread
bool warp_arrive_enabled();
// fixed to false (common.cu): the 64-channel halo kernel runs on single CTAs
bool halo_pair_enabled();
// fixed to 0 (common.cu): no L2 prefetch of later activation tiles
int prefetch_distance();

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ------------------------------------------- implicit-GEMM conv / linear ----
//
// One launch computes   OUT[pix, n] = act( scale[n] * sum_k A[pix, k] W[n, k] + shift[n] (+ RES[pix, n]) )
// for a tile grid of 128-pixel x BN-channel tiles.  A is never materialised:
// K runs over (tap, 64-channel chunk) and every k-step is one TMA box load of
// the NHWC activation tensor at tap-shifted coordinates (out-of-bounds rows
// come back as zeros, which is the convolution's zero padding).
enum : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2 };

struct GemmParams {
  int num_taps;        // 1 (1x1 / linear) or 9 (3x3)
  int chunks_per_tap;  // Cin / 64
  int a_c_off;         // first input channel inside the A buffer (concat slices)
  int tap_dc[9];       // per-tap coordinate deltas into the 5-D A map (c, w, p, h, n)
  int tap_dw[9];
  int tap_p[9];
  int tap_dh[9];
  int tiles_w, tiles_h, tiles_n, tiles_nout;
  int bw_log2, bh_log2;  // pixel box = (1<<bw) x (1<<bh) x (128>>(bw+bh)) images
  int W, H, NIMG;        // output extents, for per-row validity of residual reads
  int out_c_off;         // first output channel inside the OUT buffer
  int out_w_off;         // extra w offset of the OUT box (token assembly writes at +1)
  int cout;              // channels produced by this layer (length of scale/shift)
  int act;
  int reverse;           // walk the tile grid back to front (see plan.cu: zig-zag order for L2 reuse)
  int prefetch_dist;     // > 0: L2-prefetch the A tile needed that many tiles ahead (prefetch_distance(), default 2)
  int warp_arrive;       // accumulator stages are released by one arrival per epilogue warp instead of per thread
  int cluster;           // 1, or 2 = CTA-pair mode, cta_group::2 MMAs (the W map's box then holds BN / 2 rows)
  const float* scale;  // nullable: 1
  const float* shift;  // nullable: 0
  const __nv_bfloat16* res;  // nullable; element strides below
  long long res_sn, res_sh, res_sw;
  // LayerNorm folded into the GEMMs around it (BN == cout == 256 only, one thread owns a full row):
  //  stats_out: the epilogue also writes (mean, rstd) of each output row (eps 1e-5, biased variance);
  //  stats_in:  the A rows are the UN-normalised x; with gamma folded into W the epilogue computes
  //             rstd * acc - rstd * mean * scale[n] + shift[n]   (scale = row sums of W', shift = W beta + b).
  float2* stats_out;
  const float2* stats_in;
  long long st_sn;  // stats row = n * st_sn + w + out_w_off
};

// bn is the N tile (64, 128 or 256) and must divide cout.
int launch_gemm(int bn, const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                int num_sms, cudaStream_t stream);
int gemm_smem_bytes(int bn);
// 3x3 s1 64->64 layer with the input halo staged once per 8x16 tile (A map box = (64, 10, 1, 18, 1)).
int launch_conv3x3_halo(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const GemmParams& p,
                        int num_sms, cudaStream_t stream);

// ------------------------------------------- fused ViT layer tail (vit_block.cu) ----
// x2 = FeedForward(LN(x1)) + x1 with x1 = attn_out . Wout^T + x0, one chained tcgen05 kernel per layer
// (transformer.py:75, :93, :29-42, :94).  HGR_VIT_FUSED=0 keeps the three separate GEMM launches.
bool vit_fused_enabled();

struct VitBlockParams {
  long long rows;             // tokens (B * T)
  const __nv_bfloat16* a0;    // attn_out (rows, 256): the A operand of G0 (also reached through the tensor map)
  const __nv_bfloat16* x0;    // residual stream entering the layer tail, (rows, 256)
  const float* c1;            // folded LayerNorm: c[j] = sum_k W1'[j][k]
  const float* d1;            //                   d[j] = sum_k beta[k] W1[j][k] + b1[j]
  const float* b2;            // net.4 bias
  float2* stats_out;          // nullable: (mean, rstd) of every x2 row for the next layer's folded LayerNorm
  int reverse;                // walk the tiles back to front (zig-zag order, see plan.cu)
  long long* trace;           // nullable: [trace_tiles][16] clock64 marks of CTA 0 (hgr_vit_block_trace)
  int trace_tiles;
};

struct VitBlockOp {
  CUtensorMap a, w0, w1, w2, x, o;
  VitBlockParams p;
  double flops, bytes;
};

// x2 may alias x0 (every CTA reads the x0 rows of a tile before it stores the same rows of x2).
int build_vit_block_op(VitBlockOp& op, const void* attn_out, const void* x0, long long rows, const void* w_out,
                       const void* w1, const float* c1, const float* d1, const void* w2, const float* b2, void* x2,
                       float* stats_out);
int launch_vit_block(const VitBlockOp& op, int num_sms, cudaStream_t stream);

// ------------------------------------------------------ tensor maps ----
// rank <= 5; dims innermost first; strides in bytes for dims 1..rank-1.  swizzle_bytes: 128 / 64 = SWIZZLE_128B /
// SWIZZLE_64B (innermost box extent at most that many bytes), 0 = the box lands dense (innermost box extent a multiple
// of 16 bytes).  The innermost start coordinate of a load must sit on a 16-byte boundary of the tensor row.
int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                         const uint32_t* box, int swizzle_bytes = 128);

// --------------------------------------------------- small kernels ----
// w: [64][32] bf16, k = (kh*3+kw)*3 + c, BN scale folded in, k >= 27 zero; shift: [64] fp32.
// HGR_CONV1_TC=0: the stem convolution falls back from the tcgen05 kernel (conv1_tc.cu) to the mma.sync kernel.
bool conv1_tc_enabled();
// HGR_TRAIN_FORK (bit mask, default 1): which parts of the training backward send work to the plan's side stream -
// 1 the class head beside the pose head, 2 the transformer's weight gradients, 4 the backbone's weight gradients
int train_fork_mask();
// HGR_WGRAD_TC=0: weight gradients on mma.sync (train_wgrad.cu) instead of tcgen05 (train_wgrad_tc.cu)
bool wgrad_tc_enabled();
bool conv1_tc_supported(int S);
int launch_conv1_tc(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, const float* shift, int B,
                    int S, bool raw, cudaStream_t stream);
int launch_conv1(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, const float* shift, int B,
                 int S, cudaStream_t stream);

int launch_layernorm(const __nv_bfloat16* x, __nv_bfloat16* y, const float* gamma, const float* beta, long long rows,
                     cudaStream_t stream);

// tokens[b, 0, :] = cls; row_stats[b * T] = cls_stats (mean, rstd) for the folded LayerNorm
int launch_fill_cls(__nv_bfloat16* tokens, const float* cls, const float* cls_stats, float* row_stats, int B, int T,
                    cudaStream_t stream);

// reverse: walk the (image, head) grid back to front (zig-zag order, see plan.cu)
// probs_pitch: elements per row of the probability map (0 = T, dense (B, 8, T, T)); the training step stores
// it with a 16-byte-aligned pitch so that the backward kernel can stage it with 128-bit copies
int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, void* attn_probs, int probs_dtype, int B, int T,
                     cudaStream_t stream, int reverse = 0, int probs_pitch = 0);

int launch_cls_head(const __nv_bfloat16* tokens, const float* gamma, const float* beta, const float* w /*[C][256]*/,
                    const float* bias, void* logits, int out_dtype, int B, int T, int num_classes,
                    cudaStream_t stream);

// Pose head.  heatmaps may be NULL when preds / maxvals are given (keypoints only: the heatmaps never reach HBM);
// preds / maxvals may be NULL (heatmaps only).  The tcgen05 kernel (pose_head_tc.cu) serves F <= 20 unless
// HGR_POSE_TC=0; otherwise pose_head.cu writes the heatmaps and, if asked, tail.cu decodes them.
int launch_pose_head(const __nv_bfloat16* tokens, const __nv_bfloat16* w /*[Jpad][256]*/, const float* bias,
                     void* heatmaps, int out_dtype, int B, int F, int J, cudaStream_t stream, float* preds = nullptr,
                     float* maxvals = nullptr);
bool pose_head_tc_enabled();
bool pose_head_tc_supported(int F, int J);
int launch_pose_head_tc(const __nv_bfloat16* tokens, const __nv_bfloat16* w, const float* bias, void* heatmaps,
                        int out_dtype, float* preds, float* maxvals, int B, int F, int J, int num_sms,
                        cudaStream_t stream);

int launch_get_max_preds(const void* heatmaps, int dtype, long long rows, int hw, int width, float* preds,
                         float* maxvals, cudaStream_t stream);

int launch_crop_normalize(const uint8_t* hwc, void* chw, int out_dtype, int B, int H, int W, cudaStream_t stream);

// libs/metrics.py pose_accuracy on decoded keypoints (B, J, 2) fp32; counts: 2*J ints of scratch; acc: J+1 doubles;
// avg_cnt: {average accuracy, number of joints with a valid accuracy}
int launch_pose_accuracy(const float* pred, const float* target, int B, int J, int H, int W, double thr, int* counts,
                         double* acc, double* avg_cnt, cudaStream_t stream);

// cv2.warpAffine(INTER_LINEAR, constant border 0) + crop normalisation, fused; inv_mats: N x 6 doubles, the
// INVERTED 2x3 maps (dst -> src) as cv::warpAffine computes them; frame_index: N ints into the F frames.
int launch_crop_warp_normalize(const uint8_t* frames, int F, int Hf, int Wf, const int* frame_index,
                               const double* inv_mats, int N, int S, void* out, int out_dtype, cudaStream_t stream);

enum : int { DT_F32 = 0, DT_BF16 = 1 };

}  // namespace hgr
