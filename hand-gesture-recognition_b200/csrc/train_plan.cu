// Training step of MultiTaskNet on one GPU (BASELINE.json configs[4]):
// train-mode forward (batch-statistics BatchNorm, reference model/gelan.py:46,56 under .train()), backward
// through the whole network, and the flat fp32 gradient block a data-parallel trainer all-reduces.
//
// Storage.  Parameters, gradients and BatchNorm running statistics are three flat fp32 blocks whose layout
// (hgr_train_param_info / hgr_train_bnstat_info) follows the reference's state_dict order; the host mirror
// makes the nn.Parameters views of the parameter block, so torch.optim.AdamW (train.py:50-51) or the fused
// hgr_adamw_step update the same memory and ONE NCCL all-reduce covers every gradient.
//
// Arithmetic.  Same kernels as inference wherever the math is the same: every convolution and Linear (forward
// and input-gradient) is the tcgen05 implicit-GEMM kernel of gemm_tcgen05.cu on bf16 operands re-packed from
// the fp32 master weights once per step; stride-2 input-gradients are four parity-class launches writing
// through strided TMA store maps; weight-gradients are the split-K mma.sync kernel of train_wgrad.cu; BatchNorm,
// LayerNorm, GELU, softmax-attention and the two heads have dedicated backward kernels.  Gradients travel
// between kernels in bf16, parameter gradients are fp32.
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hgr_b200.h"
#include "gemm_ops.h"
#include "hgr_internal.h"
#include "train.h"

namespace hgr {

namespace {

constexpr int kDim = 256;
constexpr int kHeads = 8;
constexpr int kDepth = 4;

typedef __nv_bfloat16 bf16;

struct ConvDef {
  const char* name;
  int cin, cout, k, s;
  const char* in;    // activation buffer read (nullptr: the network input)
  int in_ctot, in_coff;
  const char* out;   // activation buffer written
  int out_ctot, out_coff;
  const char* res;   // residual source (activation buffer) or nullptr
  int res_ctot, res_coff;
  int level;         // input map side = S >> level
  bool dx_acc;       // the input-gradient accumulates into d(in) instead of overwriting it
};

// GELANNet('small') in execution order (gelan.py:155-160, 127-142); chunk/cat are channel slices of g*.
const ConvDef kDefs[] = {
    {"encoder.conv1", 3, 64, 3, 2, nullptr, 0, 0, "a1", 64, 0, nullptr, 0, 0, 0, false},
    {"encoder.conv2", 64, 128, 3, 2, "a1", 64, 0, "a2", 128, 0, nullptr, 0, 0, 1, false},
    {"encoder.cspelan1.cv1", 128, 128, 1, 1, "a2", 128, 0, "g1", 256, 0, nullptr, 0, 0, 2, false},
    {"encoder.cspelan1.cv2.0.cv1", 64, 64, 3, 1, "g1", 256, 64, "t1a", 64, 0, nullptr, 0, 0, 2, true},
    {"encoder.cspelan1.cv2.0.cv2", 64, 64, 3, 1, "t1a", 64, 0, "g1", 256, 128, "g1", 256, 64, 2, false},
    {"encoder.cspelan1.cv3.0.cv1", 64, 64, 3, 1, "g1", 256, 128, "t1b", 64, 0, nullptr, 0, 0, 2, true},
    {"encoder.cspelan1.cv3.0.cv2", 64, 64, 3, 1, "t1b", 64, 0, "g1", 256, 192, "g1", 256, 128, 2, false},
    {"encoder.cspelan1.cv4", 256, 128, 1, 1, "g1", 256, 0, "o1", 128, 0, nullptr, 0, 0, 2, false},
    {"encoder.down1", 128, 256, 3, 2, "o1", 128, 0, "d1", 256, 0, nullptr, 0, 0, 2, false},
    {"encoder.cspelan2.cv1", 256, 256, 1, 1, "d1", 256, 0, "g2", 512, 0, nullptr, 0, 0, 3, false},
    {"encoder.cspelan2.cv2.0.cv1", 128, 128, 3, 1, "g2", 512, 128, "t2a", 128, 0, nullptr, 0, 0, 3, true},
    {"encoder.cspelan2.cv2.0.cv2", 128, 128, 3, 1, "t2a", 128, 0, "g2", 512, 256, "g2", 512, 128, 3, false},
    {"encoder.cspelan2.cv3.0.cv1", 128, 128, 3, 1, "g2", 512, 256, "t2b", 128, 0, nullptr, 0, 0, 3, true},
    {"encoder.cspelan2.cv3.0.cv2", 128, 128, 3, 1, "t2b", 128, 0, "g2", 512, 384, "g2", 512, 256, 3, false},
    {"encoder.cspelan2.cv4", 512, 256, 1, 1, "g2", 512, 0, "o2", 256, 0, nullptr, 0, 0, 3, false},
    {"encoder.down2", 256, 512, 3, 2, "o2", 256, 0, "d2", 512, 0, nullptr, 0, 0, 3, false},
    {"encoder.cspelan3.cv1", 512, 512, 1, 1, "d2", 512, 0, "g3", 1024, 0, nullptr, 0, 0, 4, false},
    {"encoder.cspelan3.cv2.0.cv1", 256, 256, 3, 1, "g3", 1024, 256, "t3a", 256, 0, nullptr, 0, 0, 4, true},
    {"encoder.cspelan3.cv2.0.cv2", 256, 256, 3, 1, "t3a", 256, 0, "g3", 1024, 512, "g3", 1024, 256, 4, false},
    {"encoder.cspelan3.cv3.0.cv1", 256, 256, 3, 1, "g3", 1024, 512, "t3b", 256, 0, nullptr, 0, 0, 4, true},
    {"encoder.cspelan3.cv3.0.cv2", 256, 256, 3, 1, "t3b", 256, 0, "g3", 1024, 768, "g3", 1024, 512, 4, false},
    {"encoder.cspelan3.cv4", 1024, 512, 1, 1, "g3", 1024, 0, "o3", 512, 0, nullptr, 0, 0, 4, false},
};
constexpr int kNumDefs = sizeof(kDefs) / sizeof(kDefs[0]);

struct TParam {
  std::string name;
  size_t off, numel;  // floats
};

size_t pad64(size_t v) { return (v + 63) / 64 * 64; }

// The 114 parameters in the reference's state_dict order, offsets in floats.
std::vector<TParam> train_param_layout(int J, int C) {
  std::vector<TParam> v;
  size_t off = 0;
  auto add = [&](const std::string& n, size_t numel) {
    v.push_back({n, off, numel});
    off += pad64(numel);
  };
  for (int i = 0; i < kNumDefs; ++i) {
    const ConvDef& d = kDefs[i];
    add(std::string(d.name) + ".conv.weight", (size_t)d.cout * d.cin * d.k * d.k);
    add(std::string(d.name) + ".bn.weight", d.cout);
    add(std::string(d.name) + ".bn.bias", d.cout);
  }
  add("proj.weight", (size_t)kDim * 512);
  add("decoder.cls_token", kDim);
  for (int l = 0; l < kDepth; ++l) {
    const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
    const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
    add(a + "norm.weight", kDim);
    add(a + "norm.bias", kDim);
    add(a + "to_qkv.weight", (size_t)3 * kDim * kDim);
    add(a + "to_out.weight", (size_t)kDim * kDim);
    add(f + "0.weight", kDim);
    add(f + "0.bias", kDim);
    add(f + "1.weight", (size_t)kDim * kDim);
    add(f + "1.bias", kDim);
    add(f + "4.weight", (size_t)kDim * kDim);
    add(f + "4.bias", kDim);
  }
  add("decoder.mlp_head.0.weight", kDim);
  add("decoder.mlp_head.0.bias", kDim);
  add("decoder.mlp_head.1.weight", (size_t)C * kDim);
  add("decoder.mlp_head.1.bias", C);
  add("decoder.simple_decoder.1.weight", (size_t)J * kDim);
  add("decoder.simple_decoder.1.bias", J);
  return v;
}

std::vector<TParam> train_bnstat_layout() {
  std::vector<TParam> v;
  size_t off = 0;
  for (int i = 0; i < kNumDefs; ++i) {
    v.push_back({std::string(kDefs[i].name) + ".bn.running_mean", off, (size_t)kDefs[i].cout});
    off += pad64(kDefs[i].cout);
    v.push_back({std::string(kDefs[i].name) + ".bn.running_var", off, (size_t)kDefs[i].cout});
    off += pad64(kDefs[i].cout);
  }
  return v;
}

size_t layout_total(const std::vector<TParam>& v) { return v.empty() ? 0 : v.back().off + pad64(v.back().numel); }

struct WsBuf {
  std::string name;
  size_t off, bytes;
  int64_t dims[4];
};

struct ConvRt {
  bf16 *z, *w_fwd, *w_dg[4], *dz;  // dz: one of the plan's two pre-BatchNorm gradient buffers (layer parity)
  float *scale, *shift, *mean, *rstd;
  GemmOp fwd, dg[4];
  int ndg;
};

struct LayerRt {
  bf16 *ln1, *qkv, *probs, *attn_out, *xmid, *ln2, *hpre, *hact;
  bf16 *wqkv, *wqkv_t, *wo, *wo_t, *w1, *w1_t, *w4, *w4_t;
  GemmOp f_qkv, f_out, f_ff1, f_ff2;    // forward
  GemmOp b_dh, b_dln2, b_dattn, b_dln1;  // input gradients
};

}  // namespace

}  // namespace hgr

using namespace hgr;

struct hgr_train_plan {
  int S, F, T, B, J, C;
  float *params, *grads, *bnstats;
  const bf16* pos_emb;
  uint8_t* ws;
  std::vector<TParam> playout, slayout;
  std::vector<WsBuf> bufs;
  ConvRt conv[kNumDefs];
  LayerRt layer[kDepth];
  bf16 *x[kDepth + 1];
  bf16 *gA, *gB, *dln, *dh, *dattn, *dqkv, *dfeat, *dz;
  bf16 *wproj, *wproj_t, *wpose, *wconv1;
  GemmOp f_proj, b_do3;
  float *partial, *wpartial, *c1c2;
  PackJob* d_jobs;
  int njobs;
  // Weight-gradient branch: at batch 32 a backward kernel runs 10-25 us and rarely fills the GPU, so the weight
  // gradients (wgrad + split reduce, 1.3 ms of kernel time per step) run on a plan-owned side stream beside the
  // chain that carries the input gradient.  Fork and join are events, so the branch is captured into the
  // trainer's CUDA graph like any launch.  side == nullptr (HGR_TRAIN_FORK=0): everything on the caller's stream.
  static constexpr int kEvents = 160;
  cudaStream_t side = nullptr;
  int fork_mask = 0;
  bool pdl = false;  // programmatic dependent launch for the step's kernels (HGR_TRAIN_PDL, read at creation)
  cudaEvent_t events[kEvents] = {};
  int next_event = 0;
  bf16* dz2 = nullptr;

  ~hgr_train_plan() {
    for (cudaEvent_t e : events)
      if (e) cudaEventDestroy(e);
    if (side) cudaStreamDestroy(side);
  }

  size_t poff(const std::string& n) const {
    for (auto& e : playout)
      if (e.name == n) return e.off;
    return (size_t)-1;
  }
  float* P(const std::string& n) const { return params + poff(n); }
  float* G(const std::string& n) const { return grads + poff(n); }
  const WsBuf* buf(const std::string& n) const {
    for (auto& b : bufs)
      if (b.name == n) return &b;
    return nullptr;
  }
  bf16* bp(const std::string& n) const {
    const WsBuf* b = buf(n);
    return b ? reinterpret_cast<bf16*>(ws + b->off) : nullptr;
  }
};

namespace {

int check_train_config(int S, int J, int C, int B) {
  if (S < 64 || S > 1024 || S % 64 != 0) {
    set_error("image_size %d unsupported: must be a multiple of 64 in [64, 1024]", S);
    return -1;
  }
  if (J < 1 || J > 24 || C < 1 || C > 128 || B < 2) {
    set_error("training: num_joints %d (1..24) / num_classes %d (1..128) / batch %d (>= 2) unsupported", J, C, B);
    return -1;
  }
  const int T = (S / 16) * (S / 16) + 1;
  if (!attention_tokens_supported(T)) {
    set_error("image_size %d unsupported: %d tokens exceed the attention kernel's shared-memory limit (max %d)", S, T,
              attention_max_tokens());
    return -1;
  }
  return 0;
}

// row pitch of the stored attention probabilities: 16-byte aligned rows, covers the 32-padded token count
int probs_pitch(int T) { return (T + 31) / 32 * 32 + 8; }

// workspace plan: every named buffer of the training step
std::vector<WsBuf> train_workspace(int S, int J, int C, int B, size_t* total) {
  std::vector<WsBuf> v;
  size_t off = 0;
  auto add = [&](const std::string& name, size_t bytes, int64_t d0 = 0, int64_t d1 = 0, int64_t d2 = 0, int64_t d3 = 0) {
    v.push_back({name, off, bytes, {d0, d1, d2, d3}});
    off = align_up(off + bytes, 1024);
  };
  auto act = [&](const std::string& name, int64_t h, int64_t c) { add(name, (size_t)B * h * h * c * 2, B, h, h, c); };
  const int H1 = S / 2, H2 = S / 4, H3 = S / 8, H4 = S / 16, F = H4, T = F * F + 1;
  const struct { const char* n; int h, c; } acts[] = {
      {"a1", H1, 64},  {"a2", H2, 128}, {"g1", H2, 256}, {"t1a", H2, 64},  {"t1b", H2, 64},  {"o1", H2, 128},
      {"d1", H3, 256}, {"g2", H3, 512}, {"t2a", H3, 128}, {"t2b", H3, 128}, {"o2", H3, 256},
      {"d2", H4, 512}, {"g3", H4, 1024}, {"t3a", H4, 256}, {"t3b", H4, 256}, {"o3", H4, 512}};
  for (auto& a : acts) act(a.n, a.h, a.c);
  for (auto& a : acts) act(std::string("d_") + a.n, a.h, a.c);
  size_t zmax = 0;
  for (int i = 0; i < kNumDefs; ++i) {
    const ConvDef& d = kDefs[i];
    const int ho = (S >> d.level) / d.s;
    const size_t zb = (size_t)B * ho * ho * d.cout * 2;
    add(std::string("z.") + d.name, zb, B, ho, ho, d.cout);
    zmax = zb > zmax ? zb : zmax;
    add(std::string("bn.") + d.name, (size_t)4 * d.cout * 4);  // scale, shift, mean, rstd
    const size_t wn = (size_t)d.cout * d.cin * d.k * d.k;
    if (i == 0) {
      add(std::string("w.") + d.name, 64 * 32 * 2);
    } else {
      add(std::string("w.") + d.name, wn * 2);
      if (d.s == 1) {
        add(std::string("wdg.") + d.name, wn * 2);
      } else {
        const int taps[4] = {1, 2, 2, 4};
        for (int q = 0; q < 4; ++q) add(std::string("wdg") + std::to_string(q) + "." + d.name, (size_t)d.cin * taps[q] * d.cout * 2);
      }
    }
  }
  add("dz", zmax);
  add("dz2", zmax);  // layers alternate, so the side stream's wgrad of layer i reads its dz while layer i - 1 writes the other
  const size_t R = (size_t)B * T;
  for (int l = 0; l <= kDepth; ++l) add("x" + std::to_string(l), R * kDim * 2, B, 1, T, kDim);
  for (int l = 0; l < kDepth; ++l) {
    const std::string p = "l" + std::to_string(l) + ".";
    add(p + "ln1", R * kDim * 2, B, 1, T, kDim);
    add(p + "qkv", R * 3 * kDim * 2, B, 1, T, 3 * kDim);
    add(p + "probs", (size_t)B * kHeads * T * probs_pitch(T) * 2, B, kHeads, T, probs_pitch(T));
    add(p + "attn_out", R * kDim * 2, B, 1, T, kDim);
    add(p + "xmid", R * kDim * 2, B, 1, T, kDim);
    add(p + "ln2", R * kDim * 2, B, 1, T, kDim);
    add(p + "hpre", R * kDim * 2, B, 1, T, kDim);
    add(p + "hact", R * kDim * 2, B, 1, T, kDim);
    for (const char* m : {"wqkv", "wqkv_t"}) add(p + m, (size_t)3 * kDim * kDim * 2);
    for (const char* m : {"wo", "wo_t", "w1", "w1_t", "w4", "w4_t"}) add(p + m, (size_t)kDim * kDim * 2);
  }
  add("gA", R * kDim * 2, B, 1, T, kDim);
  add("gB", R * kDim * 2, B, 1, T, kDim);
  add("dln", R * kDim * 2);
  add("dh", R * kDim * 2);
  add("dattn", R * kDim * 2);
  add("dqkv", R * 3 * kDim * 2, B, 1, T, 3 * kDim);
  add("dfeat", (size_t)B * (T - 1) * kDim * 2);
  add("wproj", (size_t)kDim * 512 * 2);
  add("wproj_t", (size_t)kDim * 512 * 2);
  add("wpose", (size_t)24 * kDim * 2);
  // reduction scratch
  add("partial", (size_t)592 * 2 * 1024 * 4);
  add("c1c2", (size_t)2 * 1024 * 4);
  size_t wp = conv1_wgrad_partial_floats(B, S);
  for (int i = 1; i < kNumDefs; ++i) {
    const ConvDef& d = kDefs[i];
    const int ho = (S >> d.level) / d.s;
    const size_t n = wgrad_partial_floats(d.cout, d.cin, d.k, (long long)B * ho * ho, nullptr);
    wp = n > wp ? n : wp;
  }
  {
    const size_t n1 = wgrad_partial_floats(3 * kDim, kDim, 1, (long long)R, nullptr);
    const size_t n2 = wgrad_partial_floats(kDim, 512, 1, (long long)B * (T - 1), nullptr);
    const size_t n3 = (size_t)B * F * J * kDim;  // pose head per-CTA partials
    wp = n1 > wp ? n1 : wp;
    wp = n2 > wp ? n2 : wp;
    wp = n3 > wp ? n3 : wp;
  }
  add("wpartial", wp * 4);
  add("packjobs", 256 * sizeof(PackJob));
  (void)C;
  *total = off;
  return v;
}

int run(const GemmOp& op, cudaStream_t st) { return run_op(op, st); }

}  // namespace

extern "C" {

int hgr_train_param_count(int J, int C) { return (int)train_param_layout(J, C).size(); }

int hgr_train_param_info(int J, int C, int index, const char** name, size_t* offset_floats, size_t* numel) {
  static thread_local std::vector<TParam> cache;
  static thread_local int cj = -1, cc = -1;
  if (cj != J || cc != C) {
    cache = train_param_layout(J, C);
    cj = J;
    cc = C;
  }
  if (index < 0 || index >= (int)cache.size()) {
    set_error("train param index %d out of range", index);
    return -1;
  }
  *name = cache[index].name.c_str();
  *offset_floats = cache[index].off;
  *numel = cache[index].numel;
  return 0;
}

size_t hgr_train_param_floats(int J, int C) { return layout_total(train_param_layout(J, C)); }

int hgr_train_bnstat_count(void) { return 2 * kNumDefs; }

int hgr_train_bnstat_info(int index, const char** name, size_t* offset_floats, size_t* numel) {
  static thread_local std::vector<TParam> cache;
  if (cache.empty()) cache = train_bnstat_layout();
  if (index < 0 || index >= (int)cache.size()) {
    set_error("bn stat index %d out of range", index);
    return -1;
  }
  *name = cache[index].name.c_str();
  *offset_floats = cache[index].off;
  *numel = cache[index].numel;
  return 0;
}

size_t hgr_train_bnstat_floats(void) { return layout_total(train_bnstat_layout()); }

size_t hgr_train_workspace_bytes(int S, int J, int C, int batch) {
  if (check_train_config(S, J, C, batch)) return 0;
  size_t total = 0;
  train_workspace(S, J, C, batch, &total);
  return total;
}

int hgr_train_plan_create(hgr_train_plan_t** out, int S, int J, int C, int batch, float* d_params, float* d_grads,
                          float* d_bnstats, const void* d_pos_embedding_bf16, void* d_workspace,
                          size_t workspace_bytes) {
  if (!out) return -1;
  *out = nullptr;
  if (check_train_config(S, J, C, batch)) return -1;
  if (!d_params || !d_grads || !d_bnstats || !d_pos_embedding_bf16 || !d_workspace ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 1023) || (reinterpret_cast<uintptr_t>(d_params) & 255) ||
      (reinterpret_cast<uintptr_t>(d_grads) & 255) || (reinterpret_cast<uintptr_t>(d_pos_embedding_bf16) & 15)) {
    set_error("train plan: null or misaligned pointer (workspace 1024 B, params/grads 256 B)");
    return -1;
  }
  hgr_train_plan* pl = new hgr_train_plan();
  pl->S = S;
  pl->F = S / 16;
  pl->T = pl->F * pl->F + 1;
  pl->B = batch;
  pl->J = J;
  pl->C = C;
  pl->params = d_params;
  pl->grads = d_grads;
  pl->bnstats = d_bnstats;
  pl->pos_emb = static_cast<const bf16*>(d_pos_embedding_bf16);
  pl->ws = static_cast<uint8_t*>(d_workspace);
  pl->playout = train_param_layout(J, C);
  pl->slayout = train_bnstat_layout();
  size_t need = 0;
  pl->bufs = train_workspace(S, J, C, batch, &need);
  if (workspace_bytes < need) {
    set_error("train workspace too small: %zu < %zu", workspace_bytes, need);
    delete pl;
    return -1;
  }
  const int B = batch, T = pl->T, F = pl->F;
  const long long R = (long long)B * T;
  pl->dz = pl->bp("dz");
  pl->dz2 = pl->bp("dz2");
  pl->pdl = train_pdl_enabled();
  pl->fork_mask = train_fork_mask();
  if (pl->fork_mask) {
    bool ok = cudaStreamCreateWithFlags(&pl->side, cudaStreamNonBlocking) == cudaSuccess;
    for (int e = 0; ok && e < hgr_train_plan::kEvents; ++e)
      ok = cudaEventCreateWithFlags(&pl->events[e], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      set_error("train plan: creating the weight-gradient stream and its events failed");
      delete pl;
      return -2;
    }
  }
  pl->partial = reinterpret_cast<float*>(pl->bp("partial"));
  pl->c1c2 = reinterpret_cast<float*>(pl->bp("c1c2"));
  pl->wpartial = reinterpret_cast<float*>(pl->bp("wpartial"));
  std::vector<PackJob> jobs;
  auto job = [&](const std::string& pname, bf16* dst, int mode, int Co, int Ci, int k, int ph, int pw, long long total) {
    PackJob j;
    j.src_off = (long long)pl->poff(pname);
    j.dst = dst;
    j.total = total;
    j.mode = mode;
    j.Co = Co;
    j.Ci = Ci;
    j.k = k;
    j.ph = ph;
    j.pw = pw;
    jobs.push_back(j);
  };
  int rc = 0;
  // ---- backbone: forward convolution (raw output z) and input-gradient launches ----
  for (int i = 0; i < kNumDefs && !rc; ++i) {
    const ConvDef& d = kDefs[i];
    ConvRt& r = pl->conv[i];
    const std::string n = d.name;
    const int H = S >> d.level, Ho = H / d.s;
    r.z = pl->bp("z." + n);
    float* bnp = reinterpret_cast<float*>(pl->bp("bn." + n));
    r.scale = bnp;
    r.shift = bnp + d.cout;
    r.mean = bnp + 2 * d.cout;
    r.rstd = bnp + 3 * d.cout;
    r.w_fwd = pl->bp("w." + n);
    r.dz = (i & 1) ? pl->dz2 : pl->dz;
    r.ndg = 0;
    if (i == 0) {
      job(n + ".conv.weight", r.w_fwd, 3, 64, 3, 3, 0, 0, 64 * 32);
      continue;  // conv1 has its own forward kernel and needs no input gradient
    }
    const long long wn = (long long)d.cout * d.cin * d.k * d.k;
    job(n + ".conv.weight", r.w_fwd, 0, d.cout, d.cin, d.k, 0, 0, wn);
    rc = build_conv_op(r.fwd, pl->bp(d.in), B, H, H, d.in_ctot, d.in_coff, d.cin, r.w_fwd, nullptr, nullptr, d.k, d.s,
                       ACT_NONE, nullptr, 0, 0, r.z, d.cout, 0, d.cout);
    if (rc) break;
    bf16* dx = pl->bp(std::string("d_") + d.in);
    if (d.s == 1) {
      // dx = conv(dz, flipped / transposed weights), same geometry; "+=" through the residual epilogue
      r.w_dg[0] = pl->bp("wdg." + n);
      job(n + ".conv.weight", r.w_dg[0], 1, d.cout, d.cin, d.k, 0, 0, wn);
      r.ndg = 1;
      rc = build_conv_op(r.dg[0], r.dz, B, Ho, Ho, d.cout, 0, d.cout, r.w_dg[0], nullptr, nullptr, d.k, 1, ACT_NONE,
                         d.dx_acc ? dx : nullptr, d.in_ctot, d.in_coff, dx, d.in_ctot, d.in_coff, d.cin);
    } else {
      r.ndg = 4;
      for (int q = 0; q < 4 && !rc; ++q) {
        const int ph = q >> 1, pw = q & 1;
        r.w_dg[q] = pl->bp("wdg" + std::to_string(q) + "." + n);
        job(n + ".conv.weight", r.w_dg[q], 2, d.cout, d.cin, 3, ph, pw,
            (long long)d.cin * (ph ? 2 : 1) * (pw ? 2 : 1) * d.cout);
        rc = build_dgrad_s2_op(r.dg[q], r.dz, B, H, H, d.cout, r.w_dg[q], ph, pw, dx, d.cin);
      }
    }
  }
  // ---- ViT ----
  for (int l = 0; l <= kDepth; ++l) pl->x[l] = pl->bp("x" + std::to_string(l));
  pl->gA = pl->bp("gA");
  pl->gB = pl->bp("gB");
  pl->dln = pl->bp("dln");
  pl->dh = pl->bp("dh");
  pl->dattn = pl->bp("dattn");
  pl->dqkv = pl->bp("dqkv");
  pl->dfeat = pl->bp("dfeat");
  pl->wproj = pl->bp("wproj");
  pl->wproj_t = pl->bp("wproj_t");
  pl->wpose = pl->bp("wpose");
  pl->wconv1 = pl->conv[0].w_fwd;
  job("proj.weight", pl->wproj, 4, kDim, 512, 1, 0, 0, (long long)kDim * 512);
  job("proj.weight", pl->wproj_t, 5, kDim, 512, 1, 0, 0, (long long)kDim * 512);
  job("decoder.simple_decoder.1.weight", pl->wpose, 4, J, kDim, 1, 0, 0, (long long)J * kDim);
  if (!rc)
    rc = build_proj_op(pl->f_proj, pl->bp("o3"), B, F * F, 512, pl->wproj, pl->pos_emb, pl->x[0], T, nullptr);
  if (!rc)
    rc = build_linear_op(pl->b_do3, pl->dfeat, (long long)B * F * F, kDim, pl->wproj_t, nullptr, nullptr, ACT_NONE,
                         nullptr, pl->bp("d_o3"), 512);
  for (int l = 0; l < kDepth && !rc; ++l) {
    LayerRt& L = pl->layer[l];
    const std::string p = "l" + std::to_string(l) + ".";
    const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
    const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
    L.ln1 = pl->bp(p + "ln1");
    L.qkv = pl->bp(p + "qkv");
    L.probs = pl->bp(p + "probs");
    // pad columns of the probability maps are never written by the forward kernel and must read as zero
    if (cudaMemset(L.probs, 0, (size_t)B * kHeads * T * probs_pitch(T) * 2) != cudaSuccess) rc = -2;
    L.attn_out = pl->bp(p + "attn_out");
    L.xmid = pl->bp(p + "xmid");
    L.ln2 = pl->bp(p + "ln2");
    L.hpre = pl->bp(p + "hpre");
    L.hact = pl->bp(p + "hact");
    L.wqkv = pl->bp(p + "wqkv");
    L.wqkv_t = pl->bp(p + "wqkv_t");
    L.wo = pl->bp(p + "wo");
    L.wo_t = pl->bp(p + "wo_t");
    L.w1 = pl->bp(p + "w1");
    L.w1_t = pl->bp(p + "w1_t");
    L.w4 = pl->bp(p + "w4");
    L.w4_t = pl->bp(p + "w4_t");
    job(a + "to_qkv.weight", L.wqkv, 4, 3 * kDim, kDim, 1, 0, 0, (long long)3 * kDim * kDim);
    job(a + "to_qkv.weight", L.wqkv_t, 5, 3 * kDim, kDim, 1, 0, 0, (long long)3 * kDim * kDim);
    job(a + "to_out.weight", L.wo, 4, kDim, kDim, 1, 0, 0, (long long)kDim * kDim);
    job(a + "to_out.weight", L.wo_t, 5, kDim, kDim, 1, 0, 0, (long long)kDim * kDim);
    job(f + "1.weight", L.w1, 4, kDim, kDim, 1, 0, 0, (long long)kDim * kDim);
    job(f + "1.weight", L.w1_t, 5, kDim, kDim, 1, 0, 0, (long long)kDim * kDim);
    job(f + "4.weight", L.w4, 4, kDim, kDim, 1, 0, 0, (long long)kDim * kDim);
    job(f + "4.weight", L.w4_t, 5, kDim, kDim, 1, 0, 0, (long long)kDim * kDim);
    // forward (Attention.forward transformer.py:62-77, FeedForward :32-42, residuals :93-94)
    rc = build_linear_op(L.f_qkv, L.ln1, R, kDim, L.wqkv, nullptr, nullptr, ACT_NONE, nullptr, L.qkv, 3 * kDim);
    if (!rc) rc = build_linear_op(L.f_out, L.attn_out, R, kDim, L.wo, nullptr, nullptr, ACT_NONE, pl->x[l], L.xmid, kDim);
    if (!rc)
      rc = build_linear_op(L.f_ff1, L.ln2, R, kDim, L.w1, nullptr, pl->P(f + "1.bias"), ACT_NONE, nullptr, L.hpre, kDim);
    if (!rc)
      rc = build_linear_op(L.f_ff2, L.hact, R, kDim, L.w4, nullptr, pl->P(f + "4.bias"), ACT_NONE, L.xmid, pl->x[l + 1],
                           kDim);
    // input gradients: y = x W^T  =>  dx = dy W, i.e. a Linear with the transposed weight
    if (!rc) rc = build_linear_op(L.b_dh, pl->gA, R, kDim, L.w4_t, nullptr, nullptr, ACT_NONE, nullptr, pl->dh, kDim);
    if (!rc) rc = build_linear_op(L.b_dln2, pl->dh, R, kDim, L.w1_t, nullptr, nullptr, ACT_NONE, nullptr, pl->dln, kDim);
    if (!rc)
      rc = build_linear_op(L.b_dattn, pl->gB, R, kDim, L.wo_t, nullptr, nullptr, ACT_NONE, nullptr, pl->dattn, kDim);
    if (!rc)
      rc = build_linear_op(L.b_dln1, pl->dqkv, R, 3 * kDim, L.wqkv_t, nullptr, nullptr, ACT_NONE, nullptr, pl->dln, kDim);
  }
  if (rc) {
    delete pl;
    return rc;
  }
  if (jobs.size() > 256) {
    set_error("pack job table overflow (%zu)", jobs.size());
    delete pl;
    return -1;
  }
  pl->njobs = (int)jobs.size();
  pl->d_jobs = reinterpret_cast<PackJob*>(pl->bp("packjobs"));
  if (cudaMemcpy(pl->d_jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("train plan: copying the pack job table failed");
    delete pl;
    return -2;
  }
  // the memsets and the copy above ran on the legacy default stream, the first forward runs on the caller's
  // (possibly non-blocking) stream: make them complete before anyone can launch against this plan
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("train plan: device synchronisation after initialisation failed");
    delete pl;
    return -2;
  }
  *out = pl;
  return 0;
}

void hgr_train_plan_destroy(hgr_train_plan_t* plan) { delete plan; }

int hgr_train_buffer(hgr_train_plan_t* pl, const char* name, void** d_ptr, size_t* nbytes, int64_t dims[4]) {
  if (!pl || !name) return -1;
  const WsBuf* b = pl->buf(name);
  if (!b) {
    set_error("no training workspace buffer named '%s'", name);
    return -1;
  }
  *d_ptr = pl->ws + b->off;
  *nbytes = b->bytes;
  for (int i = 0; i < 4; ++i) dims[i] = b->dims[i];
  return 0;
}

// Train-mode forward: logits (B, C) fp32, heatmaps (B, J, S/4, S/4) fp32.  Activations stay in the workspace
// for hgr_train_backward.  momentum < 0 leaves the running statistics untouched.
int hgr_train_forward(hgr_train_plan_t* pl, const void* d_x, int x_dtype, float* d_logits, float* d_heatmaps,
                      float momentum, void* stream) {
  if (!pl || !d_x || !d_logits || !d_heatmaps || (x_dtype != DT_F32 && x_dtype != DT_BF16)) {
    set_error("hgr_train_forward: bad argument");
    return -1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PdlScope pdl(pl->pdl);
  const int B = pl->B, S = pl->S, T = pl->T, F = pl->F;
  const long long R = (long long)B * T;
  if (int rc = launch_pack_jobs(pl->d_jobs, pl->njobs, pl->params, st)) return rc;
  for (int i = 0; i < kNumDefs; ++i) {
    const ConvDef& d = kDefs[i];
    ConvRt& r = pl->conv[i];
    const std::string n = d.name;
    const int Ho = (S >> d.level) / d.s;
    const long long rows = (long long)B * Ho * Ho;
    if (i == 0) {
      if (int rc = launch_conv1_raw(d_x, x_dtype, r.z, r.w_fwd, B, S, st)) return rc;
    } else {
      if (int rc = run(r.fwd, st)) return rc;
    }
    float* rm = momentum >= 0.f ? pl->bnstats + pl->slayout[2 * i].off : nullptr;
    float* rv = momentum >= 0.f ? pl->bnstats + pl->slayout[2 * i + 1].off : nullptr;
    if (int rc = launch_bn_stats(r.z, rows, d.cout, pl->P(n + ".bn.weight"), pl->P(n + ".bn.bias"), r.scale, r.shift,
                                 r.mean, r.rstd, rm, rv, momentum, pl->partial, st))
      return rc;
    const bf16* res = d.res ? pl->bp(d.res) + d.res_coff : nullptr;
    if (int rc = launch_bn_act_fwd(r.z, rows, d.cout, r.scale, r.shift, 1, res, d.res_ctot,
                                   pl->bp(d.out) + d.out_coff, d.out_ctot, st))
      return rc;
  }
  // token assembly (ViT.forward transformer.py:132-139): class token row + proj (+ position table) rows
  if (int rc = launch_fill_cls(pl->x[0], pl->P("decoder.cls_token"), nullptr, nullptr, B, T, st)) return rc;
  if (int rc = run(pl->f_proj, st)) return rc;
  for (int l = 0; l < kDepth; ++l) {
    LayerRt& L = pl->layer[l];
    const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
    const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
    if (int rc = launch_layernorm(pl->x[l], L.ln1, pl->P(a + "norm.weight"), pl->P(a + "norm.bias"), R, st)) return rc;
    if (int rc = run(L.f_qkv, st)) return rc;
    if (int rc = launch_attention(L.qkv, L.attn_out, L.probs, DT_BF16, B, T, st, 0, probs_pitch(T))) return rc;
    if (int rc = run(L.f_out, st)) return rc;
    if (int rc = launch_layernorm(L.xmid, L.ln2, pl->P(f + "0.weight"), pl->P(f + "0.bias"), R, st)) return rc;
    if (int rc = run(L.f_ff1, st)) return rc;
    if (int rc = launch_gelu_fwd(L.hpre, L.hact, R * kDim, st)) return rc;
    if (int rc = run(L.f_ff2, st)) return rc;
  }
  if (int rc = launch_cls_head(pl->x[kDepth], pl->P("decoder.mlp_head.0.weight"), pl->P("decoder.mlp_head.0.bias"),
                               pl->P("decoder.mlp_head.1.weight"), pl->P("decoder.mlp_head.1.bias"), d_logits, DT_F32, B,
                               T, pl->C, st))
    return rc;
  return launch_pose_head(pl->x[kDepth], pl->wpose, pl->P("decoder.simple_decoder.1.bias"), d_heatmaps, DT_F32, B, F,
                          pl->J, st);
}

// Backward of the forward pass last run on this plan: d_dlogits (B, C) fp32 and d_dheatmaps (B, J, S/4, S/4)
// fp32 are the loss gradients; every entry of the flat gradient block is overwritten.
// The backward in three parts, in the order the gradients become final (the flat gradient block is in state_dict
// order, so each part completes one contiguous range of it and the trainer can start that range's all-reduce while
// the next part runs): 0 = heads, transformer, token assembly, proj (parameters from "proj.weight" on);
// 1 = cspelan3 and down2 (kDefs[kSplitConv ..]); 2 = the rest of the backbone down to conv1.
constexpr int kSplitConv = 15;  // "encoder.down2"
static_assert(kNumDefs == 22, "the split index assumes GELANNet('small')'s 22 convolutions");

int hgr_train_backward_part(hgr_train_plan_t* pl, const void* d_x, int x_dtype, const float* d_dlogits,
                            const float* d_dheatmaps, int part, void* stream) {
  if (!pl || !d_x || !d_dlogits || !d_dheatmaps || (x_dtype != DT_F32 && x_dtype != DT_BF16) || part < 0 || part > 2) {
    set_error("hgr_train_backward_part: bad argument");
    return -1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PdlScope pdl(pl->pdl);
  const int B = pl->B, S = pl->S, T = pl->T, F = pl->F;
  const long long R = (long long)B * T;
  // fork / join of the weight-gradient branch (see hgr_train_plan::side).  ws = the stream weight gradients go to.
  const int section = part == 0 ? 2 : 4;  // HGR_TRAIN_FORK bits: 1 class head, 2 transformer, 4 backbone
  cudaStream_t ws = pl->side && (pl->fork_mask & section) ? pl->side : st;
  pl->next_event = 0;
  bool ev_ok = true, forked = false;
  auto record = [&](cudaStream_t on) -> cudaEvent_t {  // an event marking "everything launched on `on` so far"
    if (ws == st) return nullptr;
    if (pl->next_event >= hgr_train_plan::kEvents) {
      ev_ok = false;
      return nullptr;
    }
    cudaEvent_t e = pl->events[pl->next_event++];
    ev_ok = ev_ok && cudaEventRecord(e, on) == cudaSuccess;
    return e;
  };
  auto wait = [&](cudaStream_t on, cudaEvent_t e) {
    if (e) ev_ok = ev_ok && cudaStreamWaitEvent(on, e, 0) == cudaSuccess;
  };
  auto fork = [&]() {  // the side stream sees what the chain has produced so far
    wait(ws, record(st));
    forked = ws != st;
  };
  auto finish = [&]() -> int {  // the call returns with the branch joined
    if (forked) wait(st, record(ws));
    if (!ev_ok) {
      set_error("hgr_train_backward_part: event record / wait failed (%s)", cudaGetErrorString(cudaGetLastError()));
      return -2;
    }
    return 0;
  };
  if (part == 0) {
    // heads: the class head (one CTA, token 0 of gA) and the pose head's bias gradient beside the pose head
    // (tokens 1.. of gA)
    cudaStream_t ws_layers = ws;
    ws = pl->side && (pl->fork_mask & 1) ? pl->side : st;
    fork();
    if (int rc = launch_cls_head_bwd(pl->x[kDepth], pl->P("decoder.mlp_head.0.weight"), pl->P("decoder.mlp_head.0.bias"),
                                     pl->P("decoder.mlp_head.1.weight"), d_dlogits, B, T, pl->C, pl->gA,
                                     pl->G("decoder.mlp_head.0.weight"), pl->G("decoder.mlp_head.0.bias"),
                                     pl->G("decoder.mlp_head.1.weight"), pl->G("decoder.mlp_head.1.bias"), ws))
      return rc;
    if (int rc = launch_heat_bias_grad(d_dheatmaps, B, pl->J, 16 * F * F, pl->G("decoder.simple_decoder.1.bias"), ws))
      return rc;
    cudaEvent_t cls_done = record(ws);
    if (int rc = launch_pose_head_bwd(pl->x[kDepth], pl->P("decoder.simple_decoder.1.weight"), d_dheatmaps, B, F, pl->J,
                                      pl->gA, pl->wpartial, pl->G("decoder.simple_decoder.1.weight"), st))
      return rc;
    wait(st, cls_done);
    ws = ws_layers;
    // Within a layer the four weight gradients leave for the side stream as soon as their operand exists; the chain
    // waits for one of them only where it is about to overwrite that operand: gA at the layer's last kernel, dh /
    // gB / dqkv in the NEXT layer (w_ff1 / w_out / w_qkv below carry those events across the iteration).  wpartial,
    // the split-reduction scratch of every wgrad, is used by the pose head above on the chain: the first fork
    // orders the branch after it.
    cudaEvent_t w_ff2 = nullptr, w_ff1 = nullptr, w_out = nullptr, w_qkv = nullptr;
    // transformer layers, last to first; gA holds d(x[l+1]) on entry and d(x[l]) on exit
    for (int l = kDepth - 1; l >= 0; --l) {
      LayerRt& L = pl->layer[l];
      const std::string a = "decoder.transformer.layers." + std::to_string(l) + ".0.";
      const std::string f = "decoder.transformer.layers." + std::to_string(l) + ".1.net.";
      // FeedForward: x[l+1] = xmid + W4 gelu(W1 LN2(xmid) + b1) + b4
      if (int rc = launch_colsum(pl->gA, R, kDim, pl->G(f + "4.bias"), pl->partial, st)) return rc;
      fork();
      if (int rc = launch_wgrad(pl->gA, kDim, L.hact, kDim, (int)R, 1, 1, kDim, kDim, 1, 1, pl->wpartial,
                                pl->G(f + "4.weight"), ws))
        return rc;
      w_ff2 = record(ws);
      wait(st, w_ff1);  // the previous layer's wgrad still reads dh
      if (int rc = run(L.b_dh, st)) return rc;
      if (int rc = launch_gelu_bwd(L.hpre, pl->dh, R * kDim, st)) return rc;
      if (int rc = launch_colsum(pl->dh, R, kDim, pl->G(f + "1.bias"), pl->partial, st)) return rc;
      fork();
      if (int rc = launch_wgrad(pl->dh, kDim, L.ln2, kDim, (int)R, 1, 1, kDim, kDim, 1, 1, pl->wpartial,
                                pl->G(f + "1.weight"), ws))
        return rc;
      w_ff1 = record(ws);
      if (int rc = run(L.b_dln2, st)) return rc;
      wait(st, w_out);  // ... gB
      if (int rc = launch_ln_bwd(pl->dln, L.xmid, pl->P(f + "0.weight"), pl->gA, pl->gB, R, pl->G(f + "0.weight"),
                                 pl->G(f + "0.bias"), pl->partial, st))
        return rc;
      // Attention: xmid = x[l] + Wo attn(LN1(x[l]))
      fork();
      if (int rc = launch_wgrad(pl->gB, kDim, L.attn_out, kDim, (int)R, 1, 1, kDim, kDim, 1, 1, pl->wpartial,
                                pl->G(a + "to_out.weight"), ws))
        return rc;
      w_out = record(ws);
      if (int rc = run(L.b_dattn, st)) return rc;
      wait(st, w_qkv);  // ... dqkv
      if (int rc = launch_attention_bwd(L.qkv, L.probs, probs_pitch(T), L.attn_out, pl->dattn, pl->dqkv, B, T, st)) return rc;
      fork();
      if (int rc = launch_wgrad(pl->dqkv, 3 * kDim, L.ln1, kDim, (int)R, 1, 1, kDim, 3 * kDim, 1, 1, pl->wpartial,
                                pl->G(a + "to_qkv.weight"), ws))
        return rc;
      w_qkv = record(ws);
      if (int rc = run(L.b_dln1, st)) return rc;
      wait(st, w_ff2);  // this layer's first wgrad reads gA, which the next kernel rewrites
      if (int rc = launch_ln_bwd(pl->dln, pl->x[l], pl->P(a + "norm.weight"), pl->gB, pl->gA, R,
                                 pl->G(a + "norm.weight"), pl->G(a + "norm.bias"), pl->partial, st))
        return rc;
    }
    // token assembly and proj
    if (int rc = launch_token_bwd(pl->gA, pl->dfeat, pl->G("decoder.cls_token"), B, T, st)) return rc;
    fork();
    if (int rc = launch_wgrad(pl->dfeat, kDim, pl->bp("o3"), 512, B, F, F, 512, kDim, 1, 1, pl->wpartial,
                              pl->G("proj.weight"), ws))
      return rc;
    if (int rc = run(pl->b_do3, st)) return rc;
    return finish();
  }
  // backbone, last conv to first.  Layer i's BatchNorm backward writes dz buffer (i & 1); its wgrad reads that buffer
  // on the side stream while the chain goes on to the input gradient and to layer i - 1 (other buffer), and layer
  // i - 2 waits for it before it writes the buffer again.
  const int i_hi = part == 1 ? kNumDefs - 1 : kSplitConv - 1, i_lo = part == 1 ? kSplitConv : 0;
  cudaEvent_t w_done[2] = {nullptr, nullptr};
  for (int i = i_hi; i >= i_lo; --i) {
    const ConvDef& d = kDefs[i];
    ConvRt& r = pl->conv[i];
    const std::string n = d.name;
    const int H = S >> d.level, Ho = H / d.s;
    const long long rows = (long long)B * Ho * Ho;
    const bf16* dy = pl->bp(std::string("d_") + d.out) + d.out_coff;
    const bf16* res = d.res ? pl->bp(d.res) + d.res_coff : nullptr;
    bf16* dres = d.res ? pl->bp(std::string("d_") + d.res) + d.res_coff : nullptr;
    wait(st, w_done[i & 1]);
    if (int rc = launch_bn_bwd(dy, d.out_ctot, r.z, rows, d.cout, r.scale, r.shift, r.mean, r.rstd, 1, res, d.res_ctot,
                               dres, d.res_ctot, pl->G(n + ".bn.weight"), pl->G(n + ".bn.bias"), pl->c1c2, pl->partial,
                               r.dz, st))
      return rc;
    fork();
    if (i == 0) {
      if (int rc = launch_conv1_wgrad(r.dz, d_x, x_dtype, B, S, pl->wpartial, pl->G(n + ".conv.weight"), ws)) return rc;
      break;
    }
    if (int rc = launch_wgrad(r.dz, d.cout, pl->bp(d.in) + d.in_coff, d.in_ctot, B, H, H, d.cin, d.cout, d.k, d.s,
                              pl->wpartial, pl->G(n + ".conv.weight"), ws))
      return rc;
    w_done[i & 1] = record(ws);
    for (int q = 0; q < r.ndg; ++q)
      if (int rc = run(r.dg[q], st)) return rc;
  }
  return finish();
}

int hgr_train_backward(hgr_train_plan_t* pl, const void* d_x, int x_dtype, const float* d_dlogits,
                       const float* d_dheatmaps, void* stream) {
  for (int part = 0; part < 3; ++part)
    if (int rc = hgr_train_backward_part(pl, d_x, x_dtype, d_dlogits, d_dheatmaps, part, stream)) return rc;
  return 0;
}

/* train.py:63-64: total = cls_weight * CE + JointsMSE; writes d_loss3 = {total, class, joints} and the
 * gradients of `total` with respect to logits and heatmaps (either may be NULL). */
int hgr_loss(const float* d_logits, const float* d_heatmaps, const long long* d_labels, const float* d_target,
             const float* d_target_weight, int B, int J, int C, int hw, float cls_weight, float* d_dlogits,
             float* d_dheatmaps, float* d_scratch, float* d_loss3, void* stream) {
  if (!d_logits || !d_heatmaps || !d_labels || !d_target || !d_target_weight || !d_scratch || !d_loss3) {
    set_error("hgr_loss: null argument");
    return -1;
  }
  PdlScope pdl(train_pdl_enabled());
  return launch_loss(d_logits, d_heatmaps, d_labels, d_target, d_target_weight, B, J, C, hw, cls_weight, d_dlogits,
                     d_dheatmaps, d_scratch, d_loss3, static_cast<cudaStream_t>(stream));
}

int hgr_adamw_step(float* d_params, const float* d_grads, float* d_exp_avg, float* d_exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  if (!d_params || !d_grads || !d_exp_avg || !d_exp_avg_sq || n < 0 || step < 1) {
    set_error("hgr_adamw_step: bad argument");
    return -1;
  }
  PdlScope pdl(train_pdl_enabled());
  return launch_adamw(d_params, d_grads, d_exp_avg, d_exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step,
                      grad_scale, static_cast<cudaStream_t>(stream));
}

/* single-kernel entry points of the training step, for parity tests */
int hgr_wgrad(const void* d_g, int g_ctot, const void* d_x, int x_ctot, int B, int H, int W, int cin, int cout, int k,
              int s, float* d_partial, float* d_dw, void* stream) {
  return launch_wgrad(static_cast<const bf16*>(d_g), g_ctot, static_cast<const bf16*>(d_x), x_ctot, B, H, W, cin, cout,
                      k, s, d_partial, d_dw, static_cast<cudaStream_t>(stream));
}

size_t hgr_wgrad_partial_floats(int cout, int cin, int k, long long pixels) {
  return wgrad_partial_floats(cout, cin, k, pixels, nullptr);
}

int hgr_attention_bwd(const void* d_qkv, const void* d_probs, const void* d_o, const void* d_do, void* d_dqkv, int B,
                      int T, void* stream) {
  return launch_attention_bwd(static_cast<const bf16*>(d_qkv), static_cast<const bf16*>(d_probs), 0,
                              static_cast<const bf16*>(d_o), static_cast<const bf16*>(d_do), static_cast<bf16*>(d_dqkv),
                              B, T, static_cast<cudaStream_t>(stream));
}

int hgr_dgrad_s2(const void* d_dz, int B, int H, int W, int cout_fwd, const void* d_w_parity, int ph, int pw, void* d_dx,
                 int cin_fwd, void* stream) {
  GemmOp op;
  if (int rc = build_dgrad_s2_op(op, d_dz, B, H, W, cout_fwd, d_w_parity, ph, pw, d_dx, cin_fwd)) return rc;
  return run_op(op, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
