// Implicit-GEMM launch descriptors shared by the inference plan (plan.cu) and the training plan
// (train_plan.cu): tensor maps + GemmParams + N tile for one conv / linear launch.
#pragma once

#include "hgr_internal.h"

namespace hgr {

struct GemmOp {
  CUtensorMap a, w, o;
  GemmParams p;
  int bn;
  bool halo;     // 64->64 3x3 s1 layer on the halo-staging kernel
  double flops;  // algorithmic: 2 * M * N * K, unpadded
  double bytes;  // algorithmic: A + W + OUT (+ RES) in bf16
};

size_t align_up(size_t v, size_t a);
int device_sm_count();

// Conv / 1x1 / stride-2 layer over NHWC bf16 buffers (weights [Cout][kh][kw][Cin] bf16).
int build_conv_op(GemmOp& op, const void* in, int B, int H, int W, int in_ctot, int in_coff, int cin, const void* wgt,
                  const float* scale, const float* shift, int k, int s, int act, const void* res, int res_ctot,
                  int res_coff, void* out, int out_ctot, int out_coff, int cout);

// y = act(x W^T + b) (+ res) over a (rows, cin) matrix.
int build_linear_op(GemmOp& op, const void* x, long long rows, int cin, const void* wgt, const float* scale,
                    const float* bias, int act, const void* res, void* y, int cout, const float* stats_in = nullptr,
                    float* stats_out = nullptr);

// proj (1x1 conv 512->256) fused with the token assembly: rows land at token index 1 + p, + position table.
int build_proj_op(GemmOp& op, const void* feat, int B, int P, int cin, const void* wgt, const void* pe, void* tokens,
                  int T, float* stats_out);

// One parity class (ph, pw) of the input-gradient of a 3x3 stride-2 convolution (training): reads the
// output-gradient map dz (B, H/2, W/2, cout_fwd) and writes dx[:, ph::2, pw::2, :cin_fwd] of a (B, H, W, cin_fwd)
// buffer.  wgt is the per-parity weight block [cin_fwd][ntaps][cout_fwd] bf16 (see train_plan.cu).
int build_dgrad_s2_op(GemmOp& op, const void* dz, int B, int H, int W, int cout_fwd, const void* wgt, int ph, int pw,
                      void* dx, int cin_fwd);

int run_op(const GemmOp& op, cudaStream_t stream);

// conv2 -> cspelan1.cv1 as one CTA-pair kernel (conv_chain.cu): a 3x3 conv producing 128 channels followed by the
// 1x1 128 -> 128 conv that consumes it, the intermediate tensor never leaves the SM.  HGR_CONV_CHAIN=0 disables.
struct ConvChainOp {
  CUtensorMap a, w, w2, o;
  GemmParams p;  // the first layer's walk; out_c_off / out_w_off of the second layer
  const float* scale2;
  const float* shift2;
  double flops, bytes;
};
bool conv_chain_enabled();
int conv_chain_prefetch();  // fixed to 0 (common.cu): no L2 prefetch of a later item's input patch
int build_conv_chain_op(ConvChainOp& op, const GemmOp& first, const GemmOp& second, const void* w2);
int launch_conv_chain(const ConvChainOp& op, int num_sms, cudaStream_t stream);

// Second version of the chained stem kernel (stem_chain.cu): the stride-2 layer's input patch is staged once per
// 8 x 16 tile as four parity planes, the intermediate tile lives in tensor memory.  HGR_CHAIN_HALO=0 keeps
// conv_chain.cu.  Needs an input map that tiles into 32 x 16 blocks and CTA pairs.
bool stem_chain_enabled();
bool stem_chain_supported(int H, int W);
int run_stem_chain(const void* in, int B, int H, int W, const void* w1, const float* scale1, const float* shift1,
                   const void* w2, const float* scale2, const float* shift2, void* out, int out_ctot, int out_coff,
                   int reverse, int num_sms, cudaStream_t stream);

// conv1 -> conv2 -> cspelan1.cv1 as one CTA-pair kernel, all three contractions on tcgen05 (stem_umma.cu): im2col rows
// built in shared memory, conv1 into tensor memory, SiLU'd pixels stored into the stride-2 layer's parity planes; the
// 64-channel map between the first two layers never reaches HBM.  bf16 NCHW input, image side a multiple of 64, CTA
// pairs.  HGR_STEM_FUSED=0 keeps conv1 + stem_chain.
bool stem_fused_enabled();
bool stem_umma_supported(int S);
int run_stem_umma(const void* x, int B, int S, const void* w0, const float* shift0, const void* w1, const float* scale1,
                  const float* shift1, const void* w2, const float* scale2, const float* shift2, void* out,
                  int out_ctot, int out_coff, int reverse, int num_sms, cudaStream_t stream);

// cspelan1.cv3.0.cv2 (+ residual, SiLU) -> cspelan1.cv4 as one CTA-pair kernel (gelan_tail.cu): y3 stays in tensor
// memory as the A operand of cv4's last K block.  64-channel chunks, maps that tile into 16 x 8 blocks, CTA pairs.
// HGR_GELAN_TAIL=0 keeps the two launches.
bool gelan_tail_enabled();
bool gelan_tail_supported(int H, int W);
int run_gelan_tail(const void* t, const void* g, int B, int H, int W, const void* w_h, const float* scale_h,
                   const float* shift_h, const void* w4, const float* scale_4, const float* shift_4, void* out,
                   int reverse, int num_sms, cudaStream_t stream);

}  // namespace hgr
