// Host side of libhgr_b200: parameter-block layout, workspace layout, the
// per-layer tensor maps, and the launch sequence that realises
// MultiTaskNet.forward (reference model/multitasknet.py:24-29 ->
// model/gelan.py:165-176 -> model/transformer.py:129-152).
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/hgr_b200.h"
#include "gemm_ops.h"
#include "hgr_internal.h"

namespace hgr {

namespace {
constexpr int kDim = 256;
constexpr int kHeads = 8;
constexpr int kDepth = 4;
}  // namespace

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
    // HGR_SM_LIMIT=<k> caps the persistent grids at k CTAs (experiments with two half-batch streams side by side)
    const char* v = getenv("HGR_SM_LIMIT");
    if (v && *v && atoi(v) > 0 && atoi(v) < n) n = atoi(v);
  }
  return n;
}

namespace {
int pick_bn(int cout) { return cout % 256 == 0 ? 256 : (cout % 128 == 0 ? 128 : 64); }

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// CTA-pair mode (cta_group::2, gemm_tcgen05.cu) pays off where the tensor pipe is the bound: measured on B200
// (gpurun r2g, batch 1024) the K >= 576 layers gain 10-20 % (down1 1464 TF/s, cspelan3 3x3 1440 TF/s), the HBM-
// and epilogue-bound K <= 512 layers (1x1 convs of cspelan1/2, ViT linears) lose up to 25 %.
bool use_pair_mode(int bn, int K) { return bn >= 128 && K >= 576 && cluster_enabled(); }

// Pixel box (bw x bh x bi = 128 pixels) for a W x H feature map.
void pick_box(int W, int H, int& bw, int& bh, int& bi) {
  bw = 16;
  while (bw > 1 && W % bw != 0) bw >>= 1;
  bh = 128 / bw;
  if (bh > 8 && W >= 16) bh = 8;  // keep boxes squarish on large maps
  while (bh > 1 && H % bh != 0) bh >>= 1;
  bi = 128 / (bw * bh);
}

}  // namespace

// Conv / 1x1 / stride-2 layer over NHWC bf16 buffers.
int build_conv_op(GemmOp& op, const void* in, int B, int H, int W, int in_ctot, int in_coff, int cin, const void* wgt,
                  const float* scale, const float* shift, int k, int s, int act, const void* res, int res_ctot,
                  int res_coff, void* out, int out_ctot, int out_coff, int cout) {
  if (!((k == 1 && s == 1) || (k == 3 && (s == 1 || s == 2)))) {
    set_error("conv: unsupported kernel %d stride %d", k, s);
    return -1;
  }
  if (cin % 64 != 0 || cout % 64 != 0 || in_coff % 64 != 0 || out_coff % 64 != 0 || in_ctot % 8 != 0 ||
      out_ctot % 8 != 0) {
    set_error("conv: channel counts must be multiples of 64 (cin %d cout %d)", cin, cout);
    return -1;
  }
  if (s == 2 && ((H & 1) || (W & 1))) {
    set_error("conv: stride 2 needs even H and W (%d x %d)", H, W);
    return -1;
  }
  const int Ho = H / s, Wo = W / s;
  int bw, bh, bi;
  pick_box(Wo, Ho, bw, bh, bi);
  memset(&op, 0, sizeof(op));
  op.bn = pick_bn(cout);
  // the halo-staging kernel covers the 64->64 3x3 stride-1 layers whose maps tile exactly into 8 x 16 pixels
  op.halo = k == 3 && s == 1 && cin == 64 && cout == 64 && W % 8 == 0 && H % 16 == 0 && act != ACT_GELU;
  if (op.halo) {
    bw = 8;
    bh = 16;
    bi = 1;
  }
  GemmParams& p = op.p;
  p.warp_arrive = warp_arrive_enabled() ? 1 : 0;
  p.num_taps = k * k;
  p.chunks_per_tap = cin / 64;
  p.a_c_off = in_coff;
  if (s == 1) {
    // A map: (c, w, 1, h, n)
    const uint64_t dims[5] = {(uint64_t)in_ctot, (uint64_t)W, 1, (uint64_t)H, (uint64_t)B};
    const uint64_t row = (uint64_t)in_ctot * 2;
    const uint64_t strides[4] = {row, row * W, row * W, row * W * H};
    const uint32_t box[5] = {64, (uint32_t)(op.halo ? bw + 2 : bw), 1, (uint32_t)(op.halo ? bh + 2 : bh), (uint32_t)bi};
    if (int r = make_tensor_map_bf16(&op.a, in, 5, dims, strides, box)) return r;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int t = kh * k + kw;
        p.tap_dc[t] = 0;
        p.tap_dw[t] = kw - k / 2;
        p.tap_p[t] = 0;
        p.tap_dh[t] = kh - k / 2;
      }
  } else {
    // Stride 2 as a space-to-depth VIEW of the same NHWC buffer:
    // (c2 = pw*C + c, bw = W/2, ph = 2, bh = H/2, n); input column 2*ow + kw - 1
    // is block ow + (kw==0 ? -1 : 0) with parity (kw==1 ? 0 : 1).
    if (in_coff != 0 || in_ctot != cin) {
      set_error("conv: stride-2 layers read a whole buffer (no channel slice)");
      return -1;
    }
    const uint64_t dims[5] = {(uint64_t)(2 * cin), (uint64_t)(W / 2), 2, (uint64_t)(H / 2), (uint64_t)B};
    const uint64_t pix = (uint64_t)cin * 2;
    const uint64_t strides[4] = {2 * pix, pix * W, 2 * pix * W, pix * W * H};
    const uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bi};
    if (int r = make_tensor_map_bf16(&op.a, in, 5, dims, strides, box)) return r;
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int t = kh * 3 + kw;
        p.tap_dc[t] = (kw == 1 ? 0 : 1) * cin;
        p.tap_dw[t] = kw == 0 ? -1 : 0;
        p.tap_p[t] = kh == 1 ? 0 : 1;
        p.tap_dh[t] = kh == 0 ? -1 : 0;
      }
  }
  {
    const uint64_t K = (uint64_t)k * k * cin;
    const uint64_t dims[2] = {K, (uint64_t)cout};
    const uint64_t strides[1] = {K * 2};
    p.cluster = (op.halo ? halo_pair_enabled() : use_pair_mode(op.bn, k * k * cin)) ? 2 : 1;
    p.prefetch_dist = prefetch_distance();
    const uint32_t box[2] = {64, (uint32_t)(op.bn / p.cluster)};
    if (int r = make_tensor_map_bf16(&op.w, wgt, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)out_ctot, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t row = (uint64_t)out_ctot * 2;
    const uint64_t strides[3] = {row, row * Wo, row * Wo * Ho};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bi};
    if (int r = make_tensor_map_bf16(&op.o, out, 4, dims, strides, box)) return r;
  }
  p.tiles_w = (Wo + bw - 1) / bw;
  p.tiles_h = (Ho + bh - 1) / bh;
  p.tiles_n = (B + bi - 1) / bi;
  p.tiles_nout = cout / op.bn;
  p.bw_log2 = ilog2(bw);
  p.bh_log2 = ilog2(bh);
  p.W = Wo;
  p.H = Ho;
  p.NIMG = B;
  p.out_c_off = out_coff;
  p.out_w_off = 0;
  p.cout = cout;
  p.act = act;
  p.scale = scale;
  p.shift = shift;
  if (res != nullptr) {
    p.res = static_cast<const __nv_bfloat16*>(res) + res_coff;
    p.res_sw = res_ctot;
    p.res_sh = (long long)res_ctot * Wo;
    p.res_sn = (long long)res_ctot * Wo * Ho;
  }
  const double M = (double)B * Ho * Wo;
  op.flops = 2.0 * M * cout * k * k * cin;
  op.bytes = 2.0 * ((double)B * H * W * cin + (double)cout * k * k * cin + M * cout * (res ? 2 : 1));
  return 0;
}

// y = act(x W^T + b) (+ res) over a (rows, cin) matrix.
int build_linear_op(GemmOp& op, const void* x, long long rows, int cin, const void* wgt, const float* scale,
                    const float* bias, int act, const void* res, void* y, int cout, const float* stats_in,
                    float* stats_out) {
  if (cin % 64 != 0 || cout % 64 != 0 || rows <= 0 || rows > 0x7fffffffLL) {
    set_error("linear: unsupported shape rows %lld cin %d cout %d", rows, cin, cout);
    return -1;
  }
  memset(&op, 0, sizeof(op));
  op.bn = pick_bn(cout);
  GemmParams& p = op.p;
  p.warp_arrive = warp_arrive_enabled() ? 1 : 0;
  p.num_taps = 1;
  p.chunks_per_tap = cin / 64;
  {
    const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)rows, 1, 1, 1};
    const uint64_t row = (uint64_t)cin * 2;
    const uint64_t strides[4] = {row, row * rows, row * rows, row * rows};
    const uint32_t box[5] = {64, 128, 1, 1, 1};
    if (int r = make_tensor_map_bf16(&op.a, x, 5, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)cin, (uint64_t)cout};
    const uint64_t strides[1] = {(uint64_t)cin * 2};
    p.cluster = use_pair_mode(op.bn, cin) ? 2 : 1;
    p.prefetch_dist = prefetch_distance();
    const uint32_t box[2] = {64, (uint32_t)(op.bn / p.cluster)};
    if (int r = make_tensor_map_bf16(&op.w, wgt, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)rows, 1, 1};
    const uint64_t row = (uint64_t)cout * 2;
    const uint64_t strides[3] = {row, row * rows, row * rows};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int r = make_tensor_map_bf16(&op.o, y, 4, dims, strides, box)) return r;
  }
  p.tiles_w = (int)((rows + 127) / 128);
  p.tiles_h = 1;
  p.tiles_n = 1;
  p.tiles_nout = cout / op.bn;
  p.bw_log2 = 7;
  p.bh_log2 = 0;
  p.W = (int)rows;
  p.H = 1;
  p.NIMG = 1;
  p.cout = cout;
  p.act = act;
  p.scale = scale;
  p.shift = bias;
  p.stats_in = reinterpret_cast<const float2*>(stats_in);
  p.stats_out = reinterpret_cast<float2*>(stats_out);
  p.st_sn = 0;
  if (res != nullptr) {
    p.res = static_cast<const __nv_bfloat16*>(res);
    p.res_sw = cout;
    p.res_sh = 0;
    p.res_sn = 0;
  }
  op.flops = 2.0 * (double)rows * cin * cout;
  op.bytes = 2.0 * ((double)rows * cin + (double)cin * cout + (double)rows * cout * (res ? 2 : 1));
  return 0;
}

// proj (1x1 conv 512->256, no bias) fused with the token assembly of
// ViT.forward (transformer.py:132-139): rows are written at token index 1+p
// and the sin-cos table is added as a batch-broadcast residual.
int build_proj_op(GemmOp& op, const void* feat, int B, int P, int cin, const void* wgt, const void* pe, void* tokens,
                  int T, float* stats_out) {
  if (P % 16 != 0) {
    set_error("proj: %d positions per image is not a multiple of 16", P);
    return -1;
  }
  memset(&op, 0, sizeof(op));
  op.bn = 256;
  GemmParams& p = op.p;
  p.warp_arrive = warp_arrive_enabled() ? 1 : 0;
  p.num_taps = 1;
  p.chunks_per_tap = cin / 64;
  {
    const uint64_t dims[5] = {(uint64_t)cin, (uint64_t)P, 1, 1, (uint64_t)B};
    const uint64_t row = (uint64_t)cin * 2;
    const uint64_t strides[4] = {row, row * P, row * P, row * P};
    const uint32_t box[5] = {64, 16, 1, 1, 8};
    if (int r = make_tensor_map_bf16(&op.a, feat, 5, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[2] = {(uint64_t)cin, (uint64_t)kDim};
    const uint64_t strides[1] = {(uint64_t)cin * 2};
    p.cluster = use_pair_mode(256, cin) ? 2 : 1;
    p.prefetch_dist = prefetch_distance();
    const uint32_t box[2] = {64, (uint32_t)(256 / p.cluster)};
    if (int r = make_tensor_map_bf16(&op.w, wgt, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)kDim, (uint64_t)T, 1, (uint64_t)B};
    const uint64_t row = (uint64_t)kDim * 2;
    const uint64_t strides[3] = {row, row * T, row * T};
    const uint32_t box[4] = {64, 16, 1, 8};
    if (int r = make_tensor_map_bf16(&op.o, tokens, 4, dims, strides, box)) return r;
  }
  p.tiles_w = P / 16;
  p.tiles_h = 1;
  p.tiles_n = (B + 7) / 8;
  p.tiles_nout = 1;
  p.bw_log2 = 4;
  p.bh_log2 = 0;
  p.W = P;
  p.H = 1;
  p.NIMG = B;
  p.out_w_off = 1;
  p.cout = kDim;
  p.act = ACT_NONE;
  p.res = static_cast<const __nv_bfloat16*>(pe);
  p.res_sw = kDim;
  p.res_sh = 0;
  p.res_sn = 0;
  p.stats_out = reinterpret_cast<float2*>(stats_out);
  p.st_sn = T;
  op.flops = 2.0 * (double)B * P * cin * kDim;
  op.bytes = 2.0 * ((double)B * P * cin + (double)cin * kDim + (double)B * P * kDim + (double)P * kDim);
  return 0;
}

// Input gradient of a 3x3 stride-2 convolution, one parity class per launch (training, see train_plan.cu).
// Forward: z[oh, ow] = sum_{kh,kw} W[kh][kw] x[2 oh + kh - 1, 2 ow + kw - 1].  For dx[2i + ph, 2j + pw] only the
// taps with (ph - kh + 1) even contribute: ph = 0 -> kh = 1 (oh = i); ph = 1 -> kh = 0 (oh = i + 1) and kh = 2
// (oh = i); the same along w.  That is a 1-, 2-, 2- or 4-tap stride-1 "convolution" of dz whose output lands on
// the (ph, pw) sub-grid of dx, which a strided TMA store map addresses directly.  Tap order (and therefore
// the K order of the packed weights): kh-candidates outer, kw-candidates inner, each in the order listed above.
int build_dgrad_s2_op(GemmOp& op, const void* dz, int B, int H, int W, int cout_fwd, const void* wgt, int ph, int pw,
                      void* dx, int cin_fwd) {
  if (cout_fwd % 64 != 0 || cin_fwd % 64 != 0 || (H & 1) || (W & 1) || ph < 0 || ph > 1 || pw < 0 || pw > 1) {
    set_error("dgrad_s2: unsupported shape (cout %d cin %d %dx%d parity %d,%d)", cout_fwd, cin_fwd, H, W, ph, pw);
    return -1;
  }
  const int Ho = H / 2, Wo = W / 2;
  int bw, bh, bi;
  pick_box(Wo, Ho, bw, bh, bi);
  memset(&op, 0, sizeof(op));
  op.bn = pick_bn(cin_fwd);
  GemmParams& p = op.p;
  p.warp_arrive = warp_arrive_enabled() ? 1 : 0;
  const int nh = ph ? 2 : 1, nw = pw ? 2 : 1;
  p.num_taps = nh * nw;
  p.chunks_per_tap = cout_fwd / 64;
  {
    const uint64_t dims[5] = {(uint64_t)cout_fwd, (uint64_t)Wo, 1, (uint64_t)Ho, (uint64_t)B};
    const uint64_t row = (uint64_t)cout_fwd * 2;
    const uint64_t strides[4] = {row, row * Wo, row * Wo, row * Wo * Ho};
    const uint32_t box[5] = {64, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bi};
    if (int r = make_tensor_map_bf16(&op.a, dz, 5, dims, strides, box)) return r;
  }
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) {
      const int t = a * nw + b;
      p.tap_dc[t] = 0;
      p.tap_p[t] = 0;
      p.tap_dh[t] = (ph && a == 0) ? 1 : 0;
      p.tap_dw[t] = (pw && b == 0) ? 1 : 0;
    }
  {
    const uint64_t K = (uint64_t)p.num_taps * cout_fwd;
    const uint64_t dims[2] = {K, (uint64_t)cin_fwd};
    const uint64_t strides[1] = {K * 2};
    p.cluster = use_pair_mode(op.bn, (int)K) ? 2 : 1;
    p.prefetch_dist = prefetch_distance();
    const uint32_t box[2] = {64, (uint32_t)(op.bn / p.cluster)};
    if (int r = make_tensor_map_bf16(&op.w, wgt, 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {(uint64_t)cin_fwd, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    const uint64_t pix = (uint64_t)cin_fwd * 2;
    const uint64_t strides[3] = {2 * pix, 2 * pix * W, pix * W * H};
    const uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bi};
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(dx) + ((size_t)ph * W + pw) * cin_fwd;
    if (int r = make_tensor_map_bf16(&op.o, base, 4, dims, strides, box)) return r;
  }
  p.tiles_w = (Wo + bw - 1) / bw;
  p.tiles_h = (Ho + bh - 1) / bh;
  p.tiles_n = (B + bi - 1) / bi;
  p.tiles_nout = cin_fwd / op.bn;
  p.bw_log2 = ilog2(bw);
  p.bh_log2 = ilog2(bh);
  p.W = Wo;
  p.H = Ho;
  p.NIMG = B;
  p.cout = cin_fwd;
  p.act = ACT_NONE;
  const double M = (double)B * Ho * Wo;
  op.flops = 2.0 * M * cin_fwd * p.num_taps * cout_fwd;
  op.bytes = 2.0 * (M * cout_fwd + (double)cin_fwd * p.num_taps * cout_fwd + M * cin_fwd);
  return 0;
}

int run_op(const GemmOp& op, cudaStream_t stream) {
  if (op.halo) return launch_conv3x3_halo(op.a, op.w, op.o, op.p, device_sm_count(), stream);
  return launch_gemm(op.bn, op.a, op.w, op.o, op.p, device_sm_count(), stream);
}

namespace {

// ------------------------------------------------------------------------
// Parameter block layout
// ------------------------------------------------------------------------
struct ParamEntry {
  std::string name;
  size_t offset, nbytes;
  int dtype;
  int64_t dims[3];
};

struct ConvSpec {
  const char* name;
  int cin, cout, k, s;
};

// The 21 tensor-core convolutions of GELANNet('small') (gelan.py:155-160 and
// GELANBlock :127-135) in execution order; conv1 is handled separately.
const ConvSpec kConvs[] = {
    {"encoder.conv2", 64, 128, 3, 2},
    {"encoder.cspelan1.cv1", 128, 128, 1, 1},
    {"encoder.cspelan1.cv2.0.cv1", 64, 64, 3, 1},
    {"encoder.cspelan1.cv2.0.cv2", 64, 64, 3, 1},
    {"encoder.cspelan1.cv3.0.cv1", 64, 64, 3, 1},
    {"encoder.cspelan1.cv3.0.cv2", 64, 64, 3, 1},
    {"encoder.cspelan1.cv4", 256, 128, 1, 1},
    {"encoder.down1", 128, 256, 3, 2},
    {"encoder.cspelan2.cv1", 256, 256, 1, 1},
    {"encoder.cspelan2.cv2.0.cv1", 128, 128, 3, 1},
    {"encoder.cspelan2.cv2.0.cv2", 128, 128, 3, 1},
    {"encoder.cspelan2.cv3.0.cv1", 128, 128, 3, 1},
    {"encoder.cspelan2.cv3.0.cv2", 128, 128, 3, 1},
    {"encoder.cspelan2.cv4", 512, 256, 1, 1},
    {"encoder.down2", 256, 512, 3, 2},
    {"encoder.cspelan3.cv1", 512, 512, 1, 1},
    {"encoder.cspelan3.cv2.0.cv1", 256, 256, 3, 1},
    {"encoder.cspelan3.cv2.0.cv2", 256, 256, 3, 1},
    {"encoder.cspelan3.cv3.0.cv1", 256, 256, 3, 1},
    {"encoder.cspelan3.cv3.0.cv2", 256, 256, 3, 1},
    {"encoder.cspelan3.cv4", 1024, 512, 1, 1},
};
constexpr int kNumConvs = sizeof(kConvs) / sizeof(kConvs[0]);

std::vector<ParamEntry> param_layout(int S, int J, int C) {
  std::vector<ParamEntry> v;
  size_t off = 0;
  auto add = [&](const std::string& name, int dtype, int64_t d0, int64_t d1 = 1, int64_t d2 = 1) {
    ParamEntry e;
    e.name = name;
    e.dtype = dtype;
    e.dims[0] = d0;
    e.dims[1] = d1;
    e.dims[2] = d2;
    e.nbytes = (size_t)(d0 * d1 * d2) * (dtype == DT_F32 ? 4 : 2);
    e.offset = off;
    off = align_up(off + e.nbytes, 1024);
    v.push_back(e);
  };
  add("encoder.conv1.w", DT_BF16, 64, 32);
  add("encoder.conv1.shift", DT_F32, 64);
  for (int i = 0; i < kNumConvs; ++i) {
    const ConvSpec& c = kConvs[i];
    add(std::string(c.name) + ".w", DT_BF16, c.cout, c.k * c.k, c.cin);
    add(std::string(c.name) + ".scale", DT_F32, c.cout);
    add(std::string(c.name) + ".shift", DT_F32, c.cout);
  }
  const int F = S / 16;
  add("proj.w", DT_BF16, kDim, 512);
  add("decoder.pos_embedding", DT_BF16, (int64_t)F * F, kDim);
  add("decoder.cls_token", DT_F32, kDim);
  add("decoder.cls_token.stats", DT_F32, 2);  // (mean, rstd) of the bf16 class token, for the folded LayerNorm
  for (int l = 0; l < kDepth; ++l) {
    const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
    const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
    // LayerNorm is folded into the GEMM that consumes it: W' = gamma (.) W in bf16,
    // c[n] = sum_k W'[n, k], d[n] = sum_k beta[k] W[n, k] (+ bias)
    add(a + "to_qkv.w", DT_BF16, 3 * kDim, kDim);
    add(a + "to_qkv.c", DT_F32, 3 * kDim);
    add(a + "to_qkv.d", DT_F32, 3 * kDim);
    add(a + "to_out.w", DT_BF16, kDim, kDim);
    add(f + "1.w", DT_BF16, kDim, kDim);
    add(f + "1.c", DT_F32, kDim);
    add(f + "1.d", DT_F32, kDim);
    add(f + "4.w", DT_BF16, kDim, kDim);
    add(f + "4.bias", DT_F32, kDim);
  }
  add("decoder.mlp_head.0.weight", DT_F32, kDim);
  add("decoder.mlp_head.0.bias", DT_F32, kDim);
  add("decoder.mlp_head.1.weight", DT_F32, C, kDim);
  add("decoder.mlp_head.1.bias", DT_F32, C);
  add("decoder.simple_decoder.1.w", DT_BF16, J, kDim);
  add("decoder.simple_decoder.1.bias", DT_F32, J);
  return v;
}

size_t param_total(const std::vector<ParamEntry>& v) {
  return v.empty() ? 0 : align_up(v.back().offset + v.back().nbytes, 1024);
}

// ------------------------------------------------------------------------
// Workspace layout (all NHWC bf16)
// ------------------------------------------------------------------------
struct Buf {
  std::string name;
  size_t offset;
  int64_t dims[4];  // N, H, W, C
};

std::vector<Buf> workspace_layout(int S, int B, size_t* total) {
  std::vector<Buf> v;
  size_t off = 0;
  auto add = [&](const char* name, int64_t h, int64_t w, int64_t c) {
    Buf b;
    b.name = name;
    b.offset = off;
    b.dims[0] = B;
    b.dims[1] = h;
    b.dims[2] = w;
    b.dims[3] = c;
    off = align_up(off + (size_t)B * h * w * c * 2, 1024);
    v.push_back(b);
  };
  const int H1 = S / 2, H2 = S / 4, H3 = S / 8, H4 = S / 16, T = H4 * H4 + 1;
  add("a1", H1, H1, 64);
  add("a2", H2, H2, 128);
  add("g1", H2, H2, 256);
  add("t1", H2, H2, 64);
  add("o1", H2, H2, 128);
  add("d1", H3, H3, 256);
  add("g2", H3, H3, 512);
  add("t2", H3, H3, 128);
  add("o2", H3, H3, 256);
  add("d2", H4, H4, 512);
  add("g3", H4, H4, 1024);
  add("t3", H4, H4, 256);
  add("o3", H4, H4, 512);
  add("tokens", 1, T, kDim);
  add("tokens_b", 1, T, kDim);
  add("row_stats", 1, T, 4);  // (mean, rstd) fp32 per token row = 8 bytes
  add("qkv", 1, T, 3 * kDim);
  add("attn_out", 1, T, kDim);
  add("hidden", 1, T, kDim);
  *total = off;
  return v;
}

}  // namespace

}  // namespace hgr

// ==========================================================================
// C ABI
// ==========================================================================
using namespace hgr;

struct hgr_plan {
  int S, F, T, B, J, C;
  uint8_t* params;
  uint8_t* ws;
  std::vector<ParamEntry> playout;
  std::vector<Buf> bufs;
  std::vector<GemmOp> convs;        // kNumConvs backbone layers
  GemmOp proj;
  GemmOp qkv[kDepth], out[kDepth], ff1[kDepth], ff2[kDepth];
  ConvChainOp stem_chain;  // conv2 -> cspelan1.cv1 (conv_chain.cu)
  bool conv_chain = false;
  VitBlockOp blk[kDepth];  // to_out + residual + FeedForward + residual as one chained kernel (vit_block.cu)
  bool vit_fused = false;
  // launch sequence of one forward pass
  struct Io {
    const void* x;
    int x_dtype;
    void* logits;
    void* heat;
    void* attn;
    int out_dtype;
    float* preds = nullptr;    // hgr_forward_keypoints: decoded keypoints instead of / beside the heatmaps
    float* maxvals = nullptr;
  };
  struct Step {
    std::string name;
    int kind;      // 0 = tcgen05 implicit GEMM, 1 = mma.sync kernel, 2 = memory-bound kernel, 3 = fused tcgen05 kernel
                   // (attention, pose head: tensor-core contractions inside a kernel bound by something else)
    double flops;  // algorithmic
    double bytes;  // algorithmic
    std::function<int(cudaStream_t, const Io&)> run;
  };
  std::vector<Step> steps;
  cudaEvent_t* events = nullptr;
  int num_events = 0;
  // host-path staging
  void* d_x = nullptr;
  void* d_logits = nullptr;
  void* d_heat = nullptr;
  size_t d_x_bytes = 0, d_logits_bytes = 0, d_heat_bytes = 0;

  const ParamEntry* param(const std::string& n) const {
    for (auto& e : playout)
      if (e.name == n) return &e;
    return nullptr;
  }
  const Buf* buf(const std::string& n) const {
    for (auto& b : bufs)
      if (b.name == n) return &b;
    return nullptr;
  }
  template <typename T>
  T* pp(const std::string& n) const {
    const ParamEntry* e = param(n);
    return e ? reinterpret_cast<T*>(params + e->offset) : nullptr;
  }
  __nv_bfloat16* bp(const std::string& n) const {
    const Buf* b = buf(n);
    return b ? reinterpret_cast<__nv_bfloat16*>(ws + b->offset) : nullptr;
  }
};

namespace {

int check_config(int S, int J, int C) {
  if (S < 64 || S > 1024 || S % 64 != 0) {
    set_error("image_size %d unsupported: must be a multiple of 64 in [64, 1024]", S);
    return -1;
  }
  if (J < 1 || J > 24 || C < 1 || C > 4096) {
    set_error("num_joints %d / num_classes %d unsupported", J, C);
    return -1;
  }
  // the attention kernels keep Q, K and V of one (image, head) in shared memory: refuse a token count they cannot
  // hold HERE, not in the middle of a forward whose backbone has already been launched
  const int T = (S / 16) * (S / 16) + 1;
  if (!attention_tokens_supported(T)) {
    set_error("image_size %d unsupported: %d tokens exceed the attention kernel's shared-memory limit (max %d)", S, T,
              attention_max_tokens());
    return -1;
  }
  return 0;
}

}  // namespace

extern "C" {

int hgr_version(void) { return 100; }

const char* hgr_last_error(void) { return last_error(); }

int hgr_param_count(int S, int J, int C) {
  if (check_config(S, J, C)) return -1;
  return (int)param_layout(S, J, C).size();
}

int hgr_param_info(int S, int J, int C, int index, const char** name, size_t* offset, size_t* nbytes, int* dtype,
                   int64_t dims[3]) {
  if (check_config(S, J, C)) return -1;
  static thread_local std::vector<ParamEntry> cache;
  static thread_local int cs = 0, cj = 0, cc = 0;
  if (cs != S || cj != J || cc != C) {
    cache = param_layout(S, J, C);
    cs = S;
    cj = J;
    cc = C;
  }
  if (index < 0 || index >= (int)cache.size()) {
    set_error("param index %d out of range", index);
    return -1;
  }
  const ParamEntry& e = cache[index];
  *name = e.name.c_str();
  *offset = e.offset;
  *nbytes = e.nbytes;
  *dtype = e.dtype;
  dims[0] = e.dims[0];
  dims[1] = e.dims[1];
  dims[2] = e.dims[2];
  return 0;
}

size_t hgr_param_bytes(int S, int J, int C) {
  if (check_config(S, J, C)) return 0;
  return param_total(param_layout(S, J, C));
}

size_t hgr_workspace_bytes(int S, int batch) {
  if (check_config(S, 21, 19) || batch < 1) return 0;
  size_t total = 0;
  workspace_layout(S, batch, &total);
  return total;
}

int hgr_plan_create(hgr_plan_t** out, int S, int J, int C, int batch, void* d_params, void* d_workspace,
                    size_t workspace_bytes) {
  if (!out) return -1;
  *out = nullptr;
  if (check_config(S, J, C)) return -1;
  if (batch < 1) {
    set_error("batch %d", batch);
    return -1;
  }
  if (!d_params || !d_workspace || (reinterpret_cast<uintptr_t>(d_params) & 1023) ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 1023)) {
    set_error("params/workspace must be non-null, 1024-byte aligned device pointers");
    return -1;
  }
  hgr_plan* pl = new hgr_plan();
  pl->S = S;
  pl->F = S / 16;
  pl->T = pl->F * pl->F + 1;
  pl->B = batch;
  pl->J = J;
  pl->C = C;
  pl->params = static_cast<uint8_t*>(d_params);
  pl->ws = static_cast<uint8_t*>(d_workspace);
  pl->playout = param_layout(S, J, C);
  size_t need = 0;
  pl->bufs = workspace_layout(S, batch, &need);
  if (workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    delete pl;
    return -1;
  }
  const int B = batch;
  const int H1 = S / 2, H2 = S / 4, H3 = S / 8, H4 = S / 16;
  pl->convs.resize(kNumConvs);
  int ci = 0;
  int rc = 0;
  auto conv = [&](const char* in, int H, int in_ctot, int in_coff, const char* res, int res_ctot, int res_coff,
                  const char* outb, int out_ctot, int out_coff, int act) {
    if (rc) return;
    const ConvSpec& c = kConvs[ci];
    const std::string n = c.name;
    rc = build_conv_op(pl->convs[ci], pl->bp(in), B, H, H, in_ctot, in_coff, c.cin, pl->pp<void>(n + ".w"),
                       pl->pp<float>(n + ".scale"), pl->pp<float>(n + ".shift"), c.k, c.s, act,
                       res ? pl->bp(res) : nullptr, res_ctot, res_coff, pl->bp(outb), out_ctot, out_coff, c.cout);
    ++ci;
  };
  // GELANNet.forward (gelan.py:165-176); GELANBlock.forward (:137-142) with the
  // chunk/cat realised as channel slices of one buffer; ResBasicBlock (:78-87).
  conv("a1", H1, 64, 0, nullptr, 0, 0, "a2", 128, 0, ACT_SILU);          // conv2
  auto gelan = [&](const char* in, int H, int cin, const char* g, const char* t, const char* o, int hid1, int hid2) {
    const int gtot = hid1 + 2 * hid2;
    conv(in, H, cin, 0, nullptr, 0, 0, g, gtot, 0, ACT_SILU);                        // cv1
    conv(g, H, gtot, hid1 / 2, nullptr, 0, 0, t, hid2, 0, ACT_SILU);                 // cv2.0.cv1
    conv(t, H, hid2, 0, g, gtot, hid1 / 2, g, gtot, hid1, ACT_SILU);                 // cv2.0.cv2 + x, SiLU
    conv(g, H, gtot, hid1, nullptr, 0, 0, t, hid2, 0, ACT_SILU);                     // cv3.0.cv1
    conv(t, H, hid2, 0, g, gtot, hid1, g, gtot, hid1 + hid2, ACT_SILU);              // cv3.0.cv2 + x, SiLU
    conv(g, H, gtot, 0, nullptr, 0, 0, o, cin, 0, ACT_SILU);                         // cv4
  };
  gelan("a2", H2, 128, "g1", "t1", "o1", 128, 64);
  conv("o1", H2, 128, 0, nullptr, 0, 0, "d1", 256, 0, ACT_SILU);         // down1
  gelan("d1", H3, 256, "g2", "t2", "o2", 256, 128);
  conv("o2", H3, 256, 0, nullptr, 0, 0, "d2", 512, 0, ACT_SILU);         // down2
  gelan("d2", H4, 512, "g3", "t3", "o3", 512, 256);
  if (!rc && conv_chain_enabled() && pl->convs[0].p.cluster == 2) {
    rc = build_conv_chain_op(pl->stem_chain, pl->convs[0], pl->convs[1], pl->pp<void>(std::string(kConvs[1].name) + ".w"));
    pl->conv_chain = rc == 0;
  }
  float* stats = reinterpret_cast<float*>(pl->bp("row_stats"));
  if (!rc)
    rc = build_proj_op(pl->proj, pl->bp("o3"), B, H4 * H4, 512, pl->pp<void>("proj.w"),
                       pl->pp<void>("decoder.pos_embedding"), pl->bp("tokens"), pl->T, stats);
  const long long rows = (long long)B * pl->T;
  for (int l = 0; l < kDepth && !rc; ++l) {
    const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
    const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
    // Attention.forward (transformer.py:62-77): LN folded into to_qkv through the row statistics of x
    rc = build_linear_op(pl->qkv[l], pl->bp("tokens"), rows, kDim, pl->pp<void>(a + "to_qkv.w"),
                         pl->pp<float>(a + "to_qkv.c"), pl->pp<float>(a + "to_qkv.d"), ACT_NONE, nullptr,
                         pl->bp("qkv"), 3 * kDim, stats, nullptr);
    // x1 = to_out(attn) + x  (:93); its epilogue leaves the statistics of x1 for the FeedForward's LN
    if (!rc)
      rc = build_linear_op(pl->out[l], pl->bp("attn_out"), rows, kDim, pl->pp<void>(a + "to_out.w"), nullptr, nullptr,
                           ACT_NONE, pl->bp("tokens"), pl->bp("tokens_b"), kDim, nullptr, stats);
    // FeedForward (transformer.py:32-42): LN folded into net.1, GELU in its epilogue
    if (!rc)
      rc = build_linear_op(pl->ff1[l], pl->bp("tokens_b"), rows, kDim, pl->pp<void>(f + "1.w"),
                           pl->pp<float>(f + "1.c"), pl->pp<float>(f + "1.d"), ACT_GELU, nullptr, pl->bp("hidden"),
                           kDim, stats, nullptr);
    // x2 = net.4(h) + b + x1 (:94), statistics of x2 for the next layer's attention LN
    if (!rc)
      rc = build_linear_op(pl->ff2[l], pl->bp("hidden"), rows, kDim, pl->pp<void>(f + "4.w"), nullptr,
                           pl->pp<float>(f + "4.bias"), ACT_NONE, pl->bp("tokens_b"), pl->bp("tokens"), kDim, nullptr,
                           stats);
    // the same three layers as ONE chained kernel; x2 overwrites x0 in place, x1 and h stay in shared memory
    if (!rc)
      rc = build_vit_block_op(pl->blk[l], pl->bp("attn_out"), pl->bp("tokens"), rows, pl->pp<void>(a + "to_out.w"),
                              pl->pp<void>(f + "1.w"), pl->pp<float>(f + "1.c"), pl->pp<float>(f + "1.d"),
                              pl->pp<void>(f + "4.w"), pl->pp<float>(f + "4.bias"), pl->bp("tokens"), stats);
  }
  pl->vit_fused = vit_fused_enabled();
  if (rc) {
    delete pl;
    return rc;
  }
  // ---- the launch sequence -------------------------------------------------
  using Io = hgr_plan::Io;
  auto add = [&](const std::string& name, int kind, double flops, double bytes,
                 std::function<int(cudaStream_t, const Io&)> fn) {
    pl->steps.push_back({name, kind, flops, bytes, std::move(fn)});
  };
  // Zig-zag tile order: every activation buffer of a batch-1024 step is larger than the L2, so a consumer
  // that walks its tile grid in the producer's order starts on the lines that were evicted first.  Alternating
  // the direction from launch to launch makes each kernel begin with what the previous one wrote last, which
  // is still L2-resident.  (HGR_ZIGZAG=0 keeps every launch front to back.)
  const bool zig = zigzag_enabled();
  int dir = 0;  // direction of the previous launch; conv1 runs front to back
  auto add_gemm = [&](const std::string& name, GemmOp* op) {
    dir = zig ? !dir : 0;
    op->p.reverse = dir;
    add(name, 0, op->flops, op->bytes, [op](cudaStream_t st, const Io&) { return run_op(*op, st); });
  };
  const int T = pl->T;
  const double dB = B, dS = S;
  const double conv1_flops = 2.0 * dB * (dS / 2) * (dS / 2) * 64 * 27;
  const double conv1_bytes = dB * 3 * dS * dS * 2 + dB * (dS / 2) * (dS / 2) * 64 * 2;
  auto run_conv1 = [pl, B](cudaStream_t st, const Io& io) {
    return launch_conv1(io.x, io.x_dtype, pl->bp("a1"), pl->pp<__nv_bfloat16>("encoder.conv1.w"),
                        pl->pp<float>("encoder.conv1.shift"), B, pl->S, st);
  };
  int first_conv = 0;
  if (pl->conv_chain) {
    ConvChainOp* ch = &pl->stem_chain;
    const std::string n0 = kConvs[0].name, n1 = kConvs[1].name;
    const int H1s = pl->S / 2;
    auto run_chain_halo = [pl, B, H1s, n0, n1](cudaStream_t st, int rev) {
      return run_stem_chain(pl->bp("a1"), B, H1s, H1s, pl->pp<void>(n0 + ".w"), pl->pp<float>(n0 + ".scale"),
                            pl->pp<float>(n0 + ".shift"), pl->pp<void>(n1 + ".w"), pl->pp<float>(n1 + ".scale"),
                            pl->pp<float>(n1 + ".shift"), pl->bp("g1"), 256, 0, rev, device_sm_count(), st);
    };
    const bool halo = stem_chain_enabled() && stem_chain_supported(H1s, H1s);
    if (halo && stem_fused_enabled() && stem_umma_supported(pl->S)) {
      // conv1 -> conv2 -> cspelan1.cv1 in one launch (stem_umma.cu): a1 never reaches HBM.  The kernel reads a bf16
      // batch through a TMA box; an fp32 batch (decided per call) takes the two launches it replaces.
      dir = zig ? !dir : 0;
      const int rev = dir;
      add("encoder.conv1+" + std::string(kConvs[0].name + 8) + "+" + (kConvs[1].name + 8), 0, conv1_flops + ch->flops,
          ch->bytes - dB * (dS / 2) * (dS / 2) * 64 * 2 + dB * 3 * dS * dS * 2,
          [pl, B, rev, n0, n1, run_conv1, run_chain_halo](cudaStream_t st, const Io& io) {
            if (io.x_dtype != DT_BF16 || (reinterpret_cast<uintptr_t>(io.x) & 15u) != 0) {
              if (int rc = run_conv1(st, io)) return rc;
              return run_chain_halo(st, rev);
            }
            return run_stem_umma(io.x, B, pl->S, pl->pp<void>("encoder.conv1.w"), pl->pp<float>("encoder.conv1.shift"),
                                  pl->pp<void>(n0 + ".w"), pl->pp<float>(n0 + ".scale"), pl->pp<float>(n0 + ".shift"),
                                  pl->pp<void>(n1 + ".w"), pl->pp<float>(n1 + ".scale"), pl->pp<float>(n1 + ".shift"),
                                  pl->bp("g1"), 256, 0, rev, device_sm_count(), st);
          });
    } else {
      add("encoder.conv1", 1, conv1_flops, conv1_bytes, run_conv1);
      dir = zig ? !dir : 0;
      ch->p.reverse = dir;
      if (halo) {
        // halo-staged version (stem_chain.cu): same operands, the patch of a tile is loaded once
        const int rev = dir;
        add(n0 + "+" + (kConvs[1].name + 8), 0, ch->flops, ch->bytes,
            [run_chain_halo, rev](cudaStream_t st, const Io&) { return run_chain_halo(st, rev); });
      } else {
        add(n0 + "+" + (kConvs[1].name + 8), 0, ch->flops, ch->bytes,
            [ch](cudaStream_t st, const Io&) { return launch_conv_chain(*ch, device_sm_count(), st); });
      }
    }
    first_conv = 2;
  } else {
    add("encoder.conv1", 1, conv1_flops, conv1_bytes, run_conv1);
  }
  // cspelan1.cv3.0.cv2 + cspelan1.cv4 (kConvs[5], [6]) as one launch (gelan_tail.cu): y3 never reaches HBM
  const bool tail = gelan_tail_enabled() && cluster_enabled() && gelan_tail_supported(S / 4, S / 4);
  for (int i = first_conv; i < kNumConvs; ++i) {
    if (tail && i == 5) {
      dir = zig ? !dir : 0;
      const int rev = dir, H2s = S / 4;
      const std::string nh = kConvs[5].name, n4 = kConvs[6].name;
      GemmOp *oh = &pl->convs[5], *o4 = &pl->convs[6];
      add(nh + "+" + (kConvs[6].name + 8), 0, oh->flops + o4->flops, oh->bytes + o4->bytes - 2.0 * dB * H2s * H2s * 64 * 2,
          [pl, B, H2s, rev, nh, n4](cudaStream_t st, const Io&) {
            return run_gelan_tail(pl->bp("t1"), pl->bp("g1"), B, H2s, H2s, pl->pp<void>(nh + ".w"),
                                  pl->pp<float>(nh + ".scale"), pl->pp<float>(nh + ".shift"), pl->pp<void>(n4 + ".w"),
                                  pl->pp<float>(n4 + ".scale"), pl->pp<float>(n4 + ".shift"), pl->bp("o1"), rev,
                                  device_sm_count(), st);
          });
      ++i;  // cv4 is part of the launch
      continue;
    }
    add_gemm(kConvs[i].name, &pl->convs[i]);
  }
  add("decoder.cls_token", 2, 0, dB * kDim * 2, [pl, B, T](cudaStream_t st, const Io&) {
    return launch_fill_cls(pl->bp("tokens"), pl->pp<float>("decoder.cls_token"),
                           pl->pp<float>("decoder.cls_token.stats"), reinterpret_cast<float*>(pl->bp("row_stats")), B,
                           T, st);
  });
  add_gemm("proj+pos_embedding", &pl->proj);
  for (int l = 0; l < kDepth; ++l) {
    const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
    const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
    add_gemm(a + "norm+to_qkv", &pl->qkv[l]);
    const bool last = l == kDepth - 1;
    dir = zig ? !dir : 0;
    const int attn_dir = dir;
    // the last layer returns probabilities only when the caller asks for them (decided per call): label by the kernel
    // the benchmark configuration runs
    add(a + "attention", attention_tc_enabled() && attention_tc_supported(T) ? 3 : 1, 4.0 * dB * kHeads * T * T * 32,
        (double)rows * kDim * 2 * 4,
        [pl, B, T, last, attn_dir](cudaStream_t st, const Io& io) {
          return launch_attention(pl->bp("qkv"), pl->bp("attn_out"), last ? io.attn : nullptr, io.out_dtype, B, T, st,
                                  attn_dir);
        });
    if (pl->vit_fused) {
      dir = zig ? !dir : 0;
      VitBlockOp* blk = &pl->blk[l];
      blk->p.reverse = dir;
      add(a + "to_out+residual+feedforward+residual", 0, blk->flops, blk->bytes,
          [blk](cudaStream_t st, const Io&) { return launch_vit_block(*blk, device_sm_count(), st); });
    } else {
      add_gemm(a + "to_out+residual", &pl->out[l]);
      add_gemm(f + "0+1+gelu", &pl->ff1[l]);
      add_gemm(f + "4+residual", &pl->ff2[l]);
    }
  }
  add("decoder.mlp_head", 2, 2.0 * dB * kDim * C, dB * kDim * 2, [pl, B, T](cudaStream_t st, const Io& io) {
    return launch_cls_head(pl->bp("tokens"), pl->pp<float>("decoder.mlp_head.0.weight"),
                           pl->pp<float>("decoder.mlp_head.0.bias"), pl->pp<float>("decoder.mlp_head.1.weight"),
                           pl->pp<float>("decoder.mlp_head.1.bias"), io.logits, io.out_dtype, B, T, pl->C, st);
  });
  add("decoder.simple_decoder", pose_head_tc_enabled() && pose_head_tc_supported(pl->F, J) ? 3 : 1,
      2.0 * dB * (dS / 4) * (dS / 4) * kDim * J,
      (double)rows * kDim * 2 + dB * J * (dS / 4) * (dS / 4) * 2, [pl, B](cudaStream_t st, const Io& io) {
        return launch_pose_head(pl->bp("tokens"), pl->pp<__nv_bfloat16>("decoder.simple_decoder.1.w"),
                                pl->pp<float>("decoder.simple_decoder.1.bias"), io.heat, io.out_dtype, B, pl->F, pl->J,
                                st, io.preds, io.maxvals);
      });
  *out = pl;
  return 0;
}

void hgr_plan_destroy(hgr_plan_t* plan) {
  if (!plan) return;
  if (plan->d_x) cudaFree(plan->d_x);
  if (plan->d_logits) cudaFree(plan->d_logits);
  if (plan->d_heat) cudaFree(plan->d_heat);
  for (int i = 0; i < plan->num_events; ++i) cudaEventDestroy(plan->events[i]);
  free(plan->events);
  delete plan;
}

int hgr_plan_launches(hgr_plan_t* plan, int with_attn) {
  (void)with_attn;
  if (!plan) return -1;
  return (int)plan->steps.size();
}

int hgr_plan_launch_info(hgr_plan_t* plan, int index, const char** name, int* kind, double* flops, double* bytes) {
  if (!plan || index < 0 || index >= (int)plan->steps.size()) {
    set_error("launch index %d out of range", index);
    return -1;
  }
  const hgr_plan::Step& s = plan->steps[index];
  *name = s.name.c_str();
  *kind = s.kind;
  *flops = s.flops;
  *bytes = s.bytes;
  return 0;
}

static int check_forward_args(hgr_plan_t* pl, const void* d_x, int x_dtype, int batch, void* d_logits,
                              void* d_heatmaps, int out_dtype) {
  if (!pl || !d_x || !d_logits || !d_heatmaps) {
    set_error("hgr_forward: null argument");
    return -1;
  }
  if (batch != pl->B) {
    // tensor maps and tile grids are built for the plan's batch; the host mirror keeps one plan per batch size.
    set_error("hgr_forward: batch %d != plan batch %d", batch, pl->B);
    return -1;
  }
  if ((x_dtype != DT_F32 && x_dtype != DT_BF16) || (out_dtype != DT_F32 && out_dtype != DT_BF16)) {
    set_error("hgr_forward: bad dtype code");
    return -1;
  }
  return 0;
}

int hgr_forward(hgr_plan_t* pl, const void* d_x, int x_dtype, int batch, void* d_logits, void* d_heatmaps,
                void* d_attn, int out_dtype, void* stream_v) {
  if (int rc = check_forward_args(pl, d_x, x_dtype, batch, d_logits, d_heatmaps, out_dtype)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const hgr_plan::Io io{d_x, x_dtype, d_logits, d_heatmaps, d_attn, out_dtype};
  for (const auto& step : pl->steps)
    if (int rc = step.run(st, io)) return rc;
  return 0;
}

int hgr_keypoints_fused(int S, int J) {
  if (check_config(S, J, 1)) return -1;
  return pose_head_tc_enabled() && pose_head_tc_supported(S / 16, J) ? 1 : 0;
}

int hgr_forward_keypoints(hgr_plan_t* pl, const void* d_x, int x_dtype, int batch, void* d_logits, void* d_heatmaps,
                          float* d_preds, float* d_maxvals, int out_dtype, void* stream_v) {
  if (!d_preds || !d_maxvals) {
    set_error("hgr_forward_keypoints: null keypoint outputs");
    return -1;
  }
  // d_heatmaps may be NULL here; the argument check only needs a non-null placeholder
  if (int rc = check_forward_args(pl, d_x, x_dtype, batch, d_logits, d_heatmaps ? d_heatmaps : d_preds, out_dtype))
    return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  hgr_plan::Io io{d_x, x_dtype, d_logits, d_heatmaps, nullptr, out_dtype};
  io.preds = d_preds;
  io.maxvals = d_maxvals;
  for (const auto& step : pl->steps)
    if (int rc = step.run(st, io)) return rc;
  return 0;
}

int hgr_forward_profile(hgr_plan_t* pl, const void* d_x, int x_dtype, int batch, void* d_logits, void* d_heatmaps,
                        void* d_attn, int out_dtype, void* stream_v, float* h_ms, int capacity) {
  if (int rc = check_forward_args(pl, d_x, x_dtype, batch, d_logits, d_heatmaps, out_dtype)) return rc;
  const int n = (int)pl->steps.size();
  if (!h_ms || capacity < n) {
    set_error("hgr_forward_profile: need room for %d launch times", n);
    return -1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  if (pl->num_events < n + 1) {
    pl->events = static_cast<cudaEvent_t*>(realloc(pl->events, sizeof(cudaEvent_t) * (n + 1)));
    for (int i = pl->num_events; i < n + 1; ++i) HGR_CHECK_CUDA(cudaEventCreate(&pl->events[i]));
    pl->num_events = n + 1;
  }
  const hgr_plan::Io io{d_x, x_dtype, d_logits, d_heatmaps, d_attn, out_dtype};
  HGR_CHECK_CUDA(cudaEventRecord(pl->events[0], st));
  for (int i = 0; i < n; ++i) {
    if (int rc = pl->steps[i].run(st, io)) return rc;
    HGR_CHECK_CUDA(cudaEventRecord(pl->events[i + 1], st));
  }
  HGR_CHECK_CUDA(cudaEventSynchronize(pl->events[n]));
  for (int i = 0; i < n; ++i) HGR_CHECK_CUDA(cudaEventElapsedTime(&h_ms[i], pl->events[i], pl->events[i + 1]));
  return n;
}

int hgr_forward_host(hgr_plan_t* pl, const void* h_x, int x_dtype, int batch, void* h_logits, void* h_heatmaps,
                     int out_dtype, void* stream_v) {
  if (!pl || !h_x || !h_logits || !h_heatmaps) {
    set_error("hgr_forward_host: null argument");
    return -1;
  }
  // validate BEFORE sizing any copy from the arguments: a smaller batch than the plan's would over-read h_x
  if (batch != pl->B) {
    set_error("hgr_forward_host: batch %d does not match the plan's batch %d", batch, pl->B);
    return -1;
  }
  if ((x_dtype != DT_F32 && x_dtype != DT_BF16) || (out_dtype != DT_F32 && out_dtype != DT_BF16)) {
    set_error("hgr_forward_host: dtype must be HGR_F32 or HGR_BF16");
    return -1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const size_t xin = (size_t)pl->B * 3 * pl->S * pl->S * (x_dtype == DT_F32 ? 4 : 2);
  const size_t osz = out_dtype == DT_F32 ? 4 : 2;
  const size_t lb = (size_t)pl->B * pl->C * osz;
  const size_t hb = (size_t)pl->B * pl->J * (pl->S / 4) * (pl->S / 4) * osz;
  if (pl->d_x_bytes < xin) {
    if (pl->d_x) cudaFree(pl->d_x);
    HGR_CHECK_CUDA(cudaMalloc(&pl->d_x, xin));
    pl->d_x_bytes = xin;
  }
  if (pl->d_logits_bytes < lb) {
    if (pl->d_logits) cudaFree(pl->d_logits);
    HGR_CHECK_CUDA(cudaMalloc(&pl->d_logits, lb));
    pl->d_logits_bytes = lb;
  }
  if (pl->d_heat_bytes < hb) {
    if (pl->d_heat) cudaFree(pl->d_heat);
    HGR_CHECK_CUDA(cudaMalloc(&pl->d_heat, hb));
    pl->d_heat_bytes = hb;
  }
  HGR_CHECK_CUDA(cudaMemcpyAsync(pl->d_x, h_x, xin, cudaMemcpyHostToDevice, st));
  if (int rc = hgr_forward(pl, pl->d_x, x_dtype, batch, pl->d_logits, pl->d_heat, nullptr, out_dtype, stream_v))
    return rc;
  HGR_CHECK_CUDA(cudaMemcpyAsync(h_logits, pl->d_logits, lb, cudaMemcpyDeviceToHost, st));
  HGR_CHECK_CUDA(cudaMemcpyAsync(h_heatmaps, pl->d_heat, hb, cudaMemcpyDeviceToHost, st));
  HGR_CHECK_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int hgr_plan_buffer(hgr_plan_t* pl, const char* name, void** d_ptr, int64_t dims[4]) {
  if (!pl || !name) return -1;
  const Buf* b = pl->buf(name);
  if (!b) {
    set_error("no workspace buffer named '%s'", name);
    return -1;
  }
  *d_ptr = pl->ws + b->offset;
  for (int i = 0; i < 4; ++i) dims[i] = b->dims[i];
  return 0;
}

// ---------------------------------------------------------- single ops ----

int hgr_conv_bn_act(const void* d_in, int B, int H, int W, int in_ctot, int in_coff, int cin, const void* d_w,
                    const float* d_scale, const float* d_shift, int k, int s, int act, const void* d_res,
                    int res_ctot, int res_coff, void* d_out, int out_ctot, int out_coff, int cout, void* stream) {
  GemmOp op;
  if (int rc = build_conv_op(op, d_in, B, H, W, in_ctot, in_coff, cin, d_w, d_scale, d_shift, k, s, act, d_res,
                             res_ctot, res_coff, d_out, out_ctot, out_coff, cout))
    return rc;
  return run_op(op, static_cast<cudaStream_t>(stream));
}

int hgr_conv_chain(const void* d_in, int B, int H, int W, const void* d_w1, const float* d_scale1,
                   const float* d_shift1, const void* d_w2, const float* d_scale2, const float* d_shift2, void* d_out,
                   int out_ctot, int out_coff, void* stream) {
  if (!cluster_enabled()) {
    set_error("hgr_conv_chain: the chained kernel runs on CTA pairs (HGR_CLUSTER=0 disables them)");
    return -1;
  }
  if (stem_chain_enabled() && stem_chain_supported(H, W))
    return run_stem_chain(d_in, B, H, W, d_w1, d_scale1, d_shift1, d_w2, d_scale2, d_shift2, d_out, out_ctot, out_coff, 0,
                          device_sm_count(), static_cast<cudaStream_t>(stream));
  // the intermediate buffer pointer of the two layer descriptors is never dereferenced: only their tile walk,
  // tensor maps of the outer tensors and epilogue parameters are taken over
  GemmOp first, second;
  if (int rc = build_conv_op(first, d_in, B, H, W, 64, 0, 64, d_w1, d_scale1, d_shift1, 3, 2, ACT_SILU, nullptr, 0, 0,
                             d_out, out_ctot, out_coff, 128))
    return rc;
  if (int rc = build_conv_op(second, d_out, B, H / 2, W / 2, out_ctot, out_coff, 128, d_w2, d_scale2, d_shift2, 1, 1,
                             ACT_SILU, nullptr, 0, 0, d_out, out_ctot, out_coff, 128))
    return rc;
  ConvChainOp op;
  if (int rc = build_conv_chain_op(op, first, second, d_w2)) return rc;
  return launch_conv_chain(op, device_sm_count(), static_cast<cudaStream_t>(stream));
}

int hgr_stem_fused(const void* d_x, int B, int S, const void* d_w0, const float* d_shift0, const void* d_w1,
                   const float* d_scale1, const float* d_shift1, const void* d_w2, const float* d_scale2,
                   const float* d_shift2, void* d_out, int out_ctot, int out_coff, void* stream) {
  if (!cluster_enabled()) {
    set_error("hgr_stem_fused: the fused stem kernel runs on CTA pairs (HGR_CLUSTER=0 disables them)");
    return -1;
  }
  return run_stem_umma(d_x, B, S, d_w0, d_shift0, d_w1, d_scale1, d_shift1, d_w2, d_scale2, d_shift2, d_out, out_ctot, out_coff, 0,
             device_sm_count(), static_cast<cudaStream_t>(stream));
}

int hgr_gelan_tail(const void* d_t, const void* d_g, int B, int H, int W, const void* d_wh, const float* d_scale_h,
                   const float* d_shift_h, const void* d_w4, const float* d_scale4, const float* d_shift4, void* d_out,
                   void* stream) {
  if (!cluster_enabled()) {
    set_error("hgr_gelan_tail: the fused kernel runs on CTA pairs (HGR_CLUSTER=0 disables them)");
    return -1;
  }
  return run_gelan_tail(d_t, d_g, B, H, W, d_wh, d_scale_h, d_shift_h, d_w4, d_scale4, d_shift4, d_out, 0,
                        device_sm_count(), static_cast<cudaStream_t>(stream));
}

int hgr_linear(const void* d_x, long long rows, int cin, const void* d_w, const float* d_scale, const float* d_bias,
               int act, const void* d_res, void* d_y, int cout, const float* d_row_stats_in, float* d_row_stats_out,
               void* stream) {
  GemmOp op;
  if (int rc = build_linear_op(op, d_x, rows, cin, d_w, d_scale, d_bias, act, d_res, d_y, cout, d_row_stats_in,
                               d_row_stats_out))
    return rc;
  return run_op(op, static_cast<cudaStream_t>(stream));
}

int hgr_vit_block(const void* d_attn_out, const void* d_x0, long long rows, const void* d_w_out, const void* d_w1,
                  const float* d_c1, const float* d_d1, const void* d_w2, const float* d_b2, void* d_x2,
                  float* d_row_stats_out, void* stream) {
  VitBlockOp op;
  if (int rc = build_vit_block_op(op, d_attn_out, d_x0, rows, d_w_out, d_w1, d_c1, d_d1, d_w2, d_b2, d_x2,
                                  d_row_stats_out))
    return rc;
  return launch_vit_block(op, device_sm_count(), static_cast<cudaStream_t>(stream));
}

int hgr_vit_block_trace(const void* d_attn_out, const void* d_x0, long long rows, const void* d_w_out, const void* d_w1,
                        const float* d_c1, const float* d_d1, const void* d_w2, const float* d_b2, void* d_x2,
                        float* d_row_stats_out, long long* d_trace, int trace_tiles, void* stream) {
  VitBlockOp op;
  if (int rc = build_vit_block_op(op, d_attn_out, d_x0, rows, d_w_out, d_w1, d_c1, d_d1, d_w2, d_b2, d_x2,
                                  d_row_stats_out))
    return rc;
  op.p.trace = d_trace;
  op.p.trace_tiles = trace_tiles;
  return launch_vit_block(op, device_sm_count(), static_cast<cudaStream_t>(stream));
}

int hgr_conv1(const void* d_x, int x_dtype, int B, int S, const void* d_w, const float* d_shift, void* d_out,
              void* stream) {
  return launch_conv1(d_x, x_dtype, static_cast<__nv_bfloat16*>(d_out), static_cast<const __nv_bfloat16*>(d_w),
                      d_shift, B, S, static_cast<cudaStream_t>(stream));
}

int hgr_layernorm(const void* d_x, void* d_y, const float* d_gamma, const float* d_beta, long long rows,
                  void* stream) {
  return launch_layernorm(static_cast<const __nv_bfloat16*>(d_x), static_cast<__nv_bfloat16*>(d_y), d_gamma, d_beta,
                          rows, static_cast<cudaStream_t>(stream));
}

int hgr_attention(const void* d_qkv, void* d_out, void* d_probs, int probs_dtype, int B, int T, void* stream) {
  return launch_attention(static_cast<const __nv_bfloat16*>(d_qkv), static_cast<__nv_bfloat16*>(d_out), d_probs,
                          probs_dtype, B, T, static_cast<cudaStream_t>(stream));
}

int hgr_attention_tc(const void* d_qkv, void* d_out, int B, int T, void* stream) {
  return launch_attention_tc(static_cast<const __nv_bfloat16*>(d_qkv), static_cast<__nv_bfloat16*>(d_out), B, T,
                             0.17677669529663687f * 1.4426950408889634f, device_sm_count(),
                             static_cast<cudaStream_t>(stream), 0);
}

int hgr_attention_tc_trace(const void* d_qkv, void* d_out, int B, int T, long long* d_trace, int trace_items,
                           int* warps, void* stream) {
  if (warps) *warps = attention_tc_warps();
  if (!d_trace) return 0;  // size query
  return launch_attention_tc(static_cast<const __nv_bfloat16*>(d_qkv), static_cast<__nv_bfloat16*>(d_out), B, T,
                             0.17677669529663687f * 1.4426950408889634f, device_sm_count(),
                             static_cast<cudaStream_t>(stream), 0, d_trace, trace_items);
}

int hgr_cls_head(const void* d_tokens, const float* d_gamma, const float* d_beta, const float* d_w,
                 const float* d_bias, void* d_logits, int out_dtype, int B, int T, int num_classes, void* stream) {
  return launch_cls_head(static_cast<const __nv_bfloat16*>(d_tokens), d_gamma, d_beta, d_w, d_bias, d_logits,
                         out_dtype, B, T, num_classes, static_cast<cudaStream_t>(stream));
}

int hgr_pose_head_decode(const void* d_tokens, const void* d_w, const float* d_bias, void* d_heatmaps, int out_dtype,
                         float* d_preds, float* d_maxvals, int B, int F, int J, void* stream) {
  if (!d_preds || !d_maxvals) {
    set_error("hgr_pose_head_decode: null keypoint outputs");
    return -1;
  }
  return launch_pose_head(static_cast<const __nv_bfloat16*>(d_tokens), static_cast<const __nv_bfloat16*>(d_w), d_bias,
                          d_heatmaps, out_dtype, B, F, J, static_cast<cudaStream_t>(stream), d_preds, d_maxvals);
}

int hgr_pose_head(const void* d_tokens, const void* d_w, const float* d_bias, void* d_heatmaps, int out_dtype, int B,
                  int F, int J, void* stream) {
  return launch_pose_head(static_cast<const __nv_bfloat16*>(d_tokens), static_cast<const __nv_bfloat16*>(d_w), d_bias,
                          d_heatmaps, out_dtype, B, F, J, static_cast<cudaStream_t>(stream));
}

int hgr_get_max_preds(const void* d_heatmaps, int dtype, int B, int J, int H, int W, float* d_preds,
                      float* d_maxvals, void* stream) {
  if (B < 0 || J < 0 || H <= 0 || W <= 0) {
    set_error("get_max_preds: bad shape (%d, %d, %d, %d)", B, J, H, W);
    return -1;
  }
  return launch_get_max_preds(d_heatmaps, dtype, (long long)B * J, H * W, W, d_preds, d_maxvals,
                              static_cast<cudaStream_t>(stream));
}

int hgr_crop_normalize(const uint8_t* d_hwc, void* d_chw, int out_dtype, int B, int H, int W, void* stream) {
  return launch_crop_normalize(d_hwc, d_chw, out_dtype, B, H, W, static_cast<cudaStream_t>(stream));
}

int hgr_pose_accuracy(const float* d_pred, const float* d_target, int B, int J, int H, int W, double thr, int* d_counts,
                      double* d_acc, double* d_avg_cnt, void* stream) {
  if (!d_pred || !d_target || !d_counts || !d_acc || !d_avg_cnt) {
    set_error("hgr_pose_accuracy: null argument");
    return -1;
  }
  return launch_pose_accuracy(d_pred, d_target, B, J, H, W, thr, d_counts, d_acc, d_avg_cnt,
                              static_cast<cudaStream_t>(stream));
}

int hgr_crop_warp_normalize(const uint8_t* d_frames, int F, int Hf, int Wf, const int* d_frame_index,
                            const double* d_inv_mats, int N, int S, void* d_chw, int out_dtype, void* stream) {
  if (!d_frames || !d_frame_index || !d_inv_mats || !d_chw) {
    set_error("hgr_crop_warp_normalize: null argument");
    return -1;
  }
  return launch_crop_warp_normalize(d_frames, F, Hf, Wf, d_frame_index, d_inv_mats, N, S, d_chw, out_dtype,
                                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"
