// Internal launchers of the training step (train_kernels.cu, train_wgrad.cu, train_vit_bwd.cu), used by
// train_plan.cu.  All pointers are device pointers; `partial` arguments are scratch for the ordered two-stage
// reductions (sized by train_plan.cu).
#pragma once

#include "hgr_internal.h"

namespace hgr {

// number of per-CTA partial rows the column reductions use for `rows` rows (<= 592)
int train_partial_blocks(long long rows);

// ---- train-mode BatchNorm over a dense [rows][C] bf16 matrix (NHWC conv output) ----
int launch_bn_stats(const __nv_bfloat16* z, long long rows, int C, const float* gamma, const float* beta, float* scale,
                    float* shift, float* mean, float* rstd, float* running_mean, float* running_var, float momentum,
                    float* partial, cudaStream_t st);
int launch_bn_act_fwd(const __nv_bfloat16* z, long long rows, int C, const float* scale, const float* shift, int silu,
                      const __nv_bfloat16* res, int res_ctot, __nv_bfloat16* y, int y_ctot, cudaStream_t st);
// dy: gradient of the block's output (channel slice, row pitch dy_ctot); writes dz (dense), dgamma, dbeta and, for
// residual blocks, accumulates du into dres (row pitch dres_ctot).  c1c2: 2*C floats of scratch.
int launch_bn_bwd(const __nv_bfloat16* dy, int dy_ctot, const __nv_bfloat16* z, long long rows, int C, const float* scale,
                  const float* shift, const float* mean, const float* rstd, int silu, const __nv_bfloat16* res,
                  int res_ctot, __nv_bfloat16* dres, int dres_ctot, float* dgamma, float* dbeta, float* c1c2,
                  float* partial, __nv_bfloat16* dz, cudaStream_t st);

int launch_gelu_fwd(const __nv_bfloat16* x, __nv_bfloat16* y, long long n, cudaStream_t st);
int launch_gelu_bwd(const __nv_bfloat16* pre, __nv_bfloat16* dh, long long n, cudaStream_t st);
int launch_colsum(const __nv_bfloat16* g, long long rows, int C, float* dst, float* partial, cudaStream_t st);
int launch_ln_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* x, const float* gamma, const __nv_bfloat16* g_in,
                  __nv_bfloat16* g_out, long long rows, float* dgamma, float* dbeta, float* partial, cudaStream_t st);
int launch_token_bwd(const __nv_bfloat16* g, __nv_bfloat16* dfeat, float* dcls, int B, int T, cudaStream_t st);

// fp32 -> bf16 re-layouts: one descriptor per packed matrix, all packed by ONE launch per step
// (modes documented at pack_jobs_kernel)
struct PackJob {
  long long src_off;  // float offset into the flat parameter block
  __nv_bfloat16* dst;
  long long total;    // elements of dst
  int mode, Co, Ci, k, ph, pw;
};
int launch_pack_jobs(const PackJob* d_jobs, int njobs, const float* params, cudaStream_t st);

int launch_loss(const float* logits, const float* heat, const long long* labels, const float* target,
                const float* weight, int B, int J, int C, int hw, float cls_weight, float* dlogits, float* dheat,
                float* partial, float* out3, cudaStream_t st);
int launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                 float eps, float wd, int step, float grad_scale, cudaStream_t st);
int launch_zero_f32(float* p, long long n, cudaStream_t st);

// ---- weight gradients ----
size_t wgrad_partial_floats(int Cout, int Cin, int k, long long P, int* chunks_out);
// dW (PyTorch (Cout, Cin, k, k) fp32) = sum_p G[p][:]^T X[shift(p)][:];  G: [B*Ho*Wo][g_ctot] (offset applied),
// X: [B][H][W][x_ctot] (offset applied), output map Ho = H / s.
int launch_wgrad(const __nv_bfloat16* g, int g_ctot, const __nv_bfloat16* x, int x_ctot, int B, int H, int W, int Cin,
                 int Cout, int k, int s, float* partial, float* dw, cudaStream_t st);
size_t conv1_wgrad_partial_floats(int B, int S);
int launch_conv1_wgrad(const __nv_bfloat16* dz, const void* x, int x_dtype, int B, int S, float* partial, float* dw,
                       cudaStream_t st);
int launch_partial_sum(const float* partial, int nparts, int n, float* dst, cudaStream_t st);

// ---- ViT backward ----
// probs: (B, 8, T, probs_pitch) bf16 (pitch 0 = T); a pitch that is a multiple of 8 and >= round_up(T, 32) is
// staged with 128-bit copies and requires zero pad columns
int launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* probs, int probs_pitch,
                         const __nv_bfloat16* o, const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B, int T,
                         cudaStream_t st);
int launch_cls_head_bwd(const __nv_bfloat16* tokens, const float* gamma, const float* beta, const float* w,
                        const float* dlogits, int B, int T, int NC, __nv_bfloat16* dtokens, float* dgamma,
                        float* dbeta, float* dw, float* dbias, cudaStream_t st);
// tcgen05 weight gradient (train_wgrad_tc.cu): Cout and Cin multiples of 64; partial layout and reduce of launch_wgrad
bool wgrad_tc_supported(int g_ctot, int x_ctot, int Cin, int Cout, int k, int s, int H, int W);
int wgrad_tc_chunks(int Cout, int Cin, int k, long long P);
int launch_wgrad_tc(const __nv_bfloat16* g, int g_ctot, const __nv_bfloat16* x, int x_ctot, int B, int H, int W, int Cin,
                    int Cout, int k, int s, float* partial, int* chunks_out, cudaStream_t st);
int launch_pose_head_bwd(const __nv_bfloat16* tokens, const float* w, const float* dheat, int B, int F, int J,
                         __nv_bfloat16* dtokens, float* dw_partial, float* dw, cudaStream_t st);
int launch_heat_bias_grad(const float* dheat, int B, int J, int hw, float* dbias, cudaStream_t st);

// conv1 without BatchNorm / activation (train-mode forward)
int launch_conv1_raw(const void* x, int x_dtype, __nv_bfloat16* out, const __nv_bfloat16* w, int B, int S,
                     cudaStream_t stream);

}  // namespace hgr
