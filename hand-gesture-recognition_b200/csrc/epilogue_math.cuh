// Epilogue arithmetic shared by the tcgen05 kernels (gemm_tcgen05.cu, vit_block.cu).
#pragma once

#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

// Fast activations for the epilogue.  The epilogue runs with ONE warp per SM
// sub-partition and accumulator stage, so its cost is counted in issue slots:
// everything that can be decided at compile time (activation, residual) is a
// template parameter, which also keeps the loop body inside the instruction cache.
template <int ACT>
__device__ __forceinline__ float apply_act(float v) {
  if constexpr (ACT == ACT_SILU) {
    // v arrives pre-halved (scale and shift are stored * 0.5): x*sigmoid(x) = h + h*tanh(h), h = x/2
    return fmaf(v, tanh_approx(v), v);
  } else if constexpr (ACT == ACT_GELU) {
    // exact-erf GELU, 0.5 v (1 + erf(v / sqrt 2)), written as relu(v) - |v| / 2 * erfc(|v| / sqrt 2) so that the
    // negative tail has no cancellation, with erfc(x) = (1 + a1 x + ... + a6 x^6)^-16 (Abramowitz-Stegun 7.1.28,
    // |err| <= 3e-7; measured |gelu err| <= 7.1e-7 over [-8, 8] in fp32).  ONE MUFU op (the reciprocal) per
    // element: the epilogues that apply it are bound by the MUFU pipe, and 7.1.26 (rcp + ex2) needs two.
    const float ax = fabsf(v) * 0.70710678118654752f;
    float dsum = fmaf(ax, 0.0000430638f, 0.0002765672f);
    dsum = fmaf(ax, dsum, 0.0001520143f);
    dsum = fmaf(ax, dsum, 0.0092705272f);
    dsum = fmaf(ax, dsum, 0.0422820123f);
    dsum = fmaf(ax, dsum, 0.0705230784f);
    dsum = fmaf(ax, dsum, 1.0f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dsum));
    r *= r;
    r *= r;
    r *= r;
    r *= r;
    return fmaf(fabsf(v) * -0.5f, r, fmaxf(v, 0.0f));
  } else {
    return v;
  }
}

}  // namespace hgr
