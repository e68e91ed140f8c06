// Error reporting and TMA tensor-map construction for libhgr_b200.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include <cudaTypedefs.h>

#include "gemm_ops.h"
#include "hgr_internal.h"

namespace hgr {

namespace {
thread_local char g_err[1024] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

static bool env_flag(const char* name, bool dflt) {
  const char* v = getenv(name);
  if (v == nullptr || *v == 0) return dflt;
  return *v != '0';
}

bool pdl_enabled() {
  static const bool on = env_flag("HGR_PDL", false);
  return on;
}

int prefetch_distance() {
  static const int d = [] {
    const char* v = getenv("HGR_PREFETCH");
    return (v && *v) ? atoi(v) : 0;  // measured on B200: distances 1, 2, 4 are 1-5 % SLOWER than none
  }();
  return d;
}

bool attention_online_enabled() {
  static const bool on = env_flag("HGR_ATTN_ONLINE", true);
  return on;
}

int attention_tiles_per_warp() {
  static const int v = [] {
    const char* e = getenv("HGR_ATTN_MT");
    return (e && e[0] == '2') ? 2 : 1;  // measured at batch 1024, T = 145: 0.156 (1) vs 0.166 ms (2) per launch
  }();
  return v;
}

bool conv_chain_enabled() {
  static const bool on = env_flag("HGR_CONV_CHAIN", true);
  return on && cluster_enabled();
}

int conv_chain_prefetch() {
  static const int d = [] {
    const char* v = getenv("HGR_CHAIN_PREFETCH");
    return (v && *v) ? atoi(v) : 0;
  }();
  return d;
}

bool attention_tc_enabled() {
  static const bool on = env_flag("HGR_ATTN_TC", true);
  return on;
}

bool stem_fused_enabled() {
  static const bool on = env_flag("HGR_STEM_FUSED", true);
  return on;
}

bool gelan_tail_enabled() {
  static const bool on = env_flag("HGR_GELAN_TAIL", true);
  return on;
}

bool stem_chain_enabled() {
  static const bool on = env_flag("HGR_CHAIN_HALO", true);
  return on && cluster_enabled();
}

bool conv1_tc_enabled() {
  static const bool on = env_flag("HGR_CONV1_TC", false);  // measured: 0.320 ms against 0.30 ms (mma.sync), see conv1_tc.cu
  return on;
}

bool pose_head_tc_enabled() {
  static const bool on = env_flag("HGR_POSE_TC", true);
  return on;
}

bool warp_arrive_enabled() {
  static const bool on = env_flag("HGR_WARP_ARRIVE", true);
  return on;
}

bool attention_cp_async_enabled() {
  static const bool on = env_flag("HGR_ATTN_CPASYNC", true);
  return on;
}

bool vit_fused_enabled() {
  static const bool on = env_flag("HGR_VIT_FUSED", true);
  return on;
}

bool halo_pair_enabled() {
  static const bool on = env_flag("HGR_HALO_PAIR", false);  // measured: 0.168 -> 0.192 ms per layer, slower
  return on && cluster_enabled();
}

bool cluster_enabled() {
  static const bool on = env_flag("HGR_CLUSTER", true);
  return on;
}

bool zigzag_enabled() {
  static const bool on = env_flag("HGR_ZIGZAG", true);
  return on;
}

namespace {

// The driver entry point is resolved at run time so that the library has no
// link-time dependency on libcuda (it must load, and export its symbols, on a
// machine without a GPU driver).
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

}  // namespace

int make_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                         const uint32_t* box, int swizzle_bytes) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return -3;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstride[4];
  cuuint32_t bdim[5];
  cuuint32_t estride[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estride[i] = 1;
    if (i > 0) gstride[i - 1] = strides[i - 1];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstride, bdim, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                       : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu %llu %llu %llu %llu box %u %u %u %u %u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return -4;
  }
  return 0;
}

}  // namespace hgr
