// Second half of a ViT layer as ONE chained tcgen05 kernel
// (reference model/transformer.py:75 `to_out`, :93 `x = attn(x) + x`, :29-42 FeedForward, :94 `x = ff(x) + x`):
//
//     x1 = attn_out . Wout^T + x0                      (G0, E0)
//     h  = GELU(LN(x1) . W1^T + b1)                    (G1, E1; LN folded: rstd * (x1 . W1'^T) - rstd * mean * c + d)
//     x2 = h . W2^T + b2 + x1                          (G2, E2)
//
// As three separate launches (gemm_tcgen05.cu) x1 and h make an HBM round trip each and x0 / x1 are read a second
// time as residuals: 9 row-passes of 512 B per token.  Here a CTA owns a 128-token tile through the whole chain:
// x1 and h exist only as bf16 tiles in shared memory, already in the K-major SWIZZLE_128B layout the next MMA reads
// its A operand from, so per token 512 B come in twice (attn_out, x0) and go out once (x2).  The rounding points are
// the ones of the separate launches (x1 and h are rounded to bf16 exactly where they used to be stored), so the
// results are the same numbers.
//
// Shared memory (one CTA per SM):  P 64 KiB (attn_out tile, later h) | Q 64 KiB (x0, replaced in place by x1, replaced
// in place by x2 = the store staging) | 3 x 32 KiB ring of [256 x 64] weight k-blocks streamed from L2 | 2 KiB
// row-statistic exchange | barriers.  TMEM: 512 columns = two 256-column accumulators used alternately by the chain
// (GEMM n -> accumulator n & 1).
//
// Warps (640 threads): 0 weight producer (TMA), 3 activation traffic (attn_out / x0 loads, x2 stores, L2 prefetch of
// the next tile), 1 MMA issuer, 2 TMEM allocator, 4-19 epilogue (warp w touches TMEM lanes 32 (w % 4)..; the four
// warps of a lane quarter split every 64-column chunk).
//
// What the timeline (tools/vit_block_trace.py) showed and the structure answers:
//  * the chain is EPILOGUE-bound (the three GEMMs need 3 x 2 350 cycles per tile, the epilogues ~13 000), so G1 / G2
//    start on k-block kb as soon as the epilogue has produced the 64-column chunk kb of x1 / h and run underneath
//    E0 / E1; G0 of the next tile runs underneath E2;
//  * sixteen epilogue warps (four per scheduler) instead of eight: 0.121 -> 0.094 ms per layer at batch 1024;
//  * all CTAs reach their load phase together, so the next tile's inputs are prefetched into L2 one tile ahead;
//  * x0 read by the row owners through the LSU cost one 32-byte sector per request (~2 300 stalled cycles per
//    tile): it comes in by TMA, straight into Q, chunk by chunk as the previous tile's store releases Q.
#include <cstdio>
#include <cstring>

#include "epilogue_math.cuh"
#include "hgr_internal.h"
#include "ptx.cuh"

namespace hgr {

namespace {

constexpr int kThreads = 640;  // 4 service warps + 16 epilogue warps
constexpr int kChunkBytes = 128 * 64 * 2;      // one 64-column k-block of a 128-row bf16 tile
constexpr int kActBytes = 4 * kChunkBytes;     // 128 x 256 bf16
constexpr int kWBytes = 256 * 64 * 2;          // one k-block of a [256 x 256] weight matrix
constexpr int kWStages = 3;
constexpr int kOffP = 0;
constexpr int kOffQ = kActBytes;
constexpr int kOffW = 2 * kActBytes;
constexpr int kOffStats = kOffW + kWStages * kWBytes;  // float2 [2 halves][128 rows]
constexpr int kOffBars = kOffStats + 2 * 128 * 8;
constexpr int kNumBars = 2 * kWStages + 16;
constexpr int kOffTmemPtr = kOffBars + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kSmemBytes <= 227 * 1024, "vit_block shared-memory plan exceeds one CTA");

// Optional per-tile timeline of CTA 0 (hgr_vit_block_trace): clock64 at the hand-over points of the chain.
__device__ __forceinline__ void trace_mark(const VitBlockParams& p, int tile_iter, int event) {
  if (p.trace != nullptr && blockIdx.x == 0 && tile_iter < p.trace_tiles) p.trace[tile_iter * 16 + event] = clock64();
}

__global__ void __launch_bounds__(kThreads, 1)
vit_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW0,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmO,
                 const VitBlockParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  uint64_t* w_full = bars;                       // [kWStages]
  uint64_t* w_empty = bars + kWStages;           // [kWStages]
  uint64_t* a_full = bars + 2 * kWStages;        // attn_out tile landed in P
  uint64_t* p_free = bars + 2 * kWStages + 1;    // G2 has finished reading P
  uint64_t* acc_full = bars + 2 * kWStages + 2;  // [2]
  uint64_t* chunk_done = bars + 2 * kWStages + 4;  // [4] E0 / E1 wrote 64-column chunk kb of x1 / h (256 arrivals)
  uint64_t* x_full = bars + 2 * kWStages + 8;      // [4] chunk j of x0 landed in Q (issued once the store of chunk j has read Q)
  uint64_t* out_ready = bars + 2 * kWStages + 12;  // [4] E2 staged chunk j of x2 in Q (one arrival per epilogue warp)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);

  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
    printf("hgr: dynamic smem base not 1024-byte aligned\n");
    __trap();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW0);
    prefetch_tensormap(&tmW1);
    prefetch_tensormap(&tmW2);
    prefetch_tensormap(&tmX);
    prefetch_tensormap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(a_full, 1);
    mbar_init(p_free, 1);
    mbar_init(&acc_full[0], 1);
    mbar_init(&acc_full[1], 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&out_ready[i], 16);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&chunk_done[i], 16);  // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  const int total_tiles = (int)((p.rows + 127) / 128);
  const int first = p.reverse ? total_tiles - 1 - (int)blockIdx.x : (int)blockIdx.x;
  const int step = p.reverse ? -(int)gridDim.x : (int)gridDim.x;
  const int my_tiles = (int)blockIdx.x < total_tiles ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    // ================= weight producer: 12 k-blocks per tile through the ring =================
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const CUtensorMap* tm = g == 0 ? &tmW0 : (g == 1 ? &tmW1 : &tmW2);
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&w_empty[stage], phase ^ 1);
            mbar_expect_tx(&w_full[stage], kWBytes);
            tma_load_2d(smem + kOffW + stage * kWBytes, tm, &w_full[stage], kb * 64, 0);
            if (++stage == kWStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ================= activation traffic: attn_out tile -> P, x0 tile -> Q, x2 tile Q -> global =================
    // One thread owns every bulk copy of the activations.  Q is recycled chunk by chunk: E2 hands over chunk j of
    // x2 as soon as it is staged, its store is issued at once, and as soon as that store has READ the chunk the
    // same 16 KiB take chunk j of the next tile's x0 - so the store of a tile and the x0 load of the next one run
    // underneath E2 instead of after it (as one 64 KiB store + load they cost ~4 700 cycles between two tiles).
    if (elect_one_sync()) {
      int tile = first;
      if (my_tiles > 0) {
        mbar_expect_tx(a_full, kActBytes);
        tma_load_5d(smem + kOffP, &tmA, a_full, 0, tile * 128, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mbar_expect_tx(&x_full[j], kChunkBytes);
          tma_load_5d(smem + kOffQ + j * kChunkBytes, &tmX, &x_full[j], 0, tile * 128, j, 0, 0);
        }
      }
      for (int i = 0; i < my_tiles; ++i, tile += step) {
        const bool more = i + 1 < my_tiles;
        // Every CTA of the grid reaches its load phase at about the same time; without help the whole wave then
        // waits on HBM (128 KiB per CTA, measured ~5 000 cycles per tile).  The NEXT tile's attn_out and x0 rows
        // (two contiguous 64 KiB blocks) are pulled into L2 a whole tile period ahead.
        if (more) {
          const long long r0 = (long long)(tile + step) * 128;
          const long long nrows = p.rows - r0 < 128 ? p.rows - r0 : 128;
          bulk_prefetch_l2(p.a0 + r0 * 256, (uint32_t)(nrows * 512));
          bulk_prefetch_l2(p.x0 + r0 * 256, (uint32_t)(nrows * 512));
        }
        mbar_wait(p_free, i & 1);  // G2 of this tile has read h: P may take the next attn_out tile
        if (more) {
          mbar_expect_tx(a_full, kActBytes);
          tma_load_5d(smem + kOffP, &tmA, a_full, 0, (tile + step) * 128, 0, 0, 0);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&out_ready[j], i & 1);  // chunk j of x2 is staged in Q
          tma_store_4d(&tmO, smem + kOffQ + j * kChunkBytes, 0, tile * 128, j, 0);
          tma_store_commit();
          if (more && j > 0) {
            tma_store_wait_read<1>();  // the store of chunk j - 1 has read its 16 KiB
            mbar_expect_tx(&x_full[j - 1], kChunkBytes);
            tma_load_5d(smem + kOffQ + (j - 1) * kChunkBytes, &tmX, &x_full[j - 1], 0, (tile + step) * 128, j - 1, 0, 0);
          }
        }
        trace_mark(p, i, 5);  // last store issued
        if (more) {
          tma_store_wait_read<0>();
          mbar_expect_tx(&x_full[3], kChunkBytes);
          tma_load_5d(smem + kOffQ + 3 * kChunkBytes, &tmX, &x_full[3], 0, (tile + step) * 128, 3, 0, 0);
        }
      }
      tma_store_wait_all();
    }
  } else if (warp == 1) {
    // ================= MMA issuer: G0, G1, G2 of every tile =================
    // G1 / G2 start k-block kb as soon as the epilogue has written the 64-column chunk kb of x1 / h.
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
    int stage = 0;
    uint32_t phase = 0, chunk_phase = 0, n = 0;
    for (int i = 0; i < my_tiles; ++i) {
#pragma unroll
      for (int g = 0; g < 3; ++g, ++n) {
        if (g == 0) {
          mbar_wait(a_full, i & 1);
          tc_fence_after();
        }
        if (lane == 0) trace_mark(p, i, 8 + g);
        const uint32_t a_off = g == 1 ? kOffQ : kOffP;
        const uint32_t tmem_d = tmem_base + (n & 1) * 256;
        for (int kb = 0; kb < 4; ++kb) {
          if (g != 0) {
            mbar_wait(&chunk_done[kb], chunk_phase);
            tc_fence_after();
          }
          mbar_wait(&w_full[stage], phase);
          tc_fence_after();
          if (g == 0 && lane == 0) trace_mark(p, i, 11 + kb);  // weights of G0 k-block kb are in shared memory
          const uint64_t a_base = umma_desc_sw128(smem_u32(smem + a_off + kb * kChunkBytes), 1024);
          const uint64_t b_base = umma_desc_sw128(smem_u32(smem + kOffW + stage * kWBytes), 1024);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_d, a_base + 2 * k, b_base + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&w_empty[stage]);
            if (kb == 3) {
              umma_commit(&acc_full[n & 1]);
              if (g == 2) umma_commit(p_free);
            }
          }
          __syncwarp();
          if (++stage == kWStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (g != 0) chunk_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: 512 threads (four warps per scheduler, so one warp's TMEM / MUFU / shared-memory
    // latencies hide behind the others) work through the four 64-column chunks of a tile together; thread =
    // (row, 16-column quarter of the chunk), so a chunk is complete - and its k-block can be issued - after a
    // quarter of the epilogue =================
    const int ew = warp - 4;
    const int q = ew & 3;     // TMEM lane quarter this warp may touch
    const int cq = ew >> 2;   // 16-column quarter of every chunk
    const int row = q * 32 + lane;
    const int etid = threadIdx.x - 128;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cq * 16;
    // this thread's two 16-byte pieces inside a chunk buffer: row * 128 + ((cq * 2 + v) ^ sw) * 16
    const uint32_t off0 = row * 128 + ((static_cast<uint32_t>(cq * 2) ^ sw) * 16);
    const uint32_t off1 = row * 128 + ((static_cast<uint32_t>(cq * 2 + 1) ^ sw) * 16);
    float2* s_buf = reinterpret_cast<float2*>(smem + kOffStats);  // [2][128]
    const float4* c1v = reinterpret_cast<const float4*>(p.c1 + cq * 16);
    const float4* d1v = reinterpret_cast<const float4*>(p.d1 + cq * 16);
    const float4* b2v = reinterpret_cast<const float4*>(p.b2 + cq * 16);
    uint32_t n = 0;
    int tile = first;
    for (int i = 0; i < my_tiles; ++i, tile += step) {
      const long long grow = (long long)tile * 128 + row;
      const bool valid = grow < p.rows;

      // ---------------- E0: x1 = acc + x0 -> Q, row statistics of x1 ----------------
      // x0 is already in Q in the chunk layout (TMA, warp 3); x1 replaces it in place
      if (etid == 0) trace_mark(p, i, 6);  // epilogue is back at the top of the chain
      mbar_wait(&acc_full[n & 1], (n >> 1) & 1);
      tc_fence_after();
      if (etid == 0) trace_mark(p, i, 0);  // G0 done
      bar_sync(1, 512);  // every thread is done with the previous tile's statistics exchange
      float s1 = 0.f, s2 = 0.f;
      // the accumulator columns of chunk j + 1 are already on their way out of TMEM while chunk j is computed
      uint32_t accb[2][16];
      tmem_ld_32x32b_x16(t_lane + (n & 1) * 256, accb[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint8_t* buf = smem + kOffQ + j * kChunkBytes;
        const uint32_t(&acc)[16] = accb[j & 1];
        mbar_wait(&x_full[j], i & 1);  // chunk j of x0 is in Q (so the previous tile's store of it has read Q)
        if (etid == 0 && j == 0) trace_mark(p, i, 7);
        uint4 x0v[2];
        x0v[0] = *reinterpret_cast<const uint4*>(buf + off0);
        x0v[1] = *reinterpret_cast<const uint4*>(buf + off1);
        tmem_ld_wait();
        if (j < 3) tmem_ld_32x32b_x16(t_lane + (n & 1) * 256 + (j + 1) * 64, accb[(j + 1) & 1]);
        uint32_t packed[8];
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(x0v);
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          const float v0 = __uint_as_float(acc[e]) + bf16_lo(rw[e >> 1]);
          const float v1 = __uint_as_float(acc[e + 1]) + bf16_hi(rw[e >> 1]);
          s1 += v0 + v1;
          s2 = fmaf(v0, v0, fmaf(v1, v1, s2));
          packed[e >> 1] = pack_bf16x2(v0, v1);
        }
        *reinterpret_cast<uint4*>(buf + off0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4*>(buf + off1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&chunk_done[j]);
      }
      ++n;
      if (etid == 0) trace_mark(p, i, 1);  // E0 done (this thread)
      // row totals over the four column quarters, always summed as (q0 + q2) + (q1 + q3); the 2 KiB exchange
      // buffer is all the shared memory that is left, hence the three rounds
      float rstd, rmean;
      {
        if (cq >= 2) s_buf[(cq - 2) * 128 + row] = make_float2(s1, s2);
        bar_sync(1, 512);
        if (cq < 2) {
          const float2 o = s_buf[cq * 128 + row];
          s1 += o.x;
          s2 += o.y;
          if (cq == 1) s_buf[128 + row] = make_float2(s1, s2);
        }
        bar_sync(1, 512);
        if (cq == 0) {
          const float2 o = s_buf[128 + row];
          const float mean = (s1 + o.x) * (1.0f / 256);
          const float var = fmaxf(fmaf(s2 + o.y, 1.0f / 256, -mean * mean), 0.0f);
          const float rs = rsqrtf(var + 1e-5f);
          s_buf[row] = make_float2(rs, mean * rs);
        }
        bar_sync(1, 512);
        const float2 st = s_buf[row];
        rstd = st.x;
        rmean = st.y;
      }

      // ---------------- E1: h = GELU(rstd * acc - rstd * mean * c + d) -> P ----------------
      mbar_wait(&acc_full[n & 1], (n >> 1) & 1);
      tc_fence_after();
      if (etid == 0) trace_mark(p, i, 2);  // G1 done
      tmem_ld_32x32b_x16(t_lane + (n & 1) * 256, accb[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint8_t* buf = smem + kOffP + j * kChunkBytes;
        const uint32_t(&acc)[16] = accb[j & 1];
        float4 sc[4], sh[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          sc[e] = __ldg(c1v + j * 16 + e);
          sh[e] = __ldg(d1v + j * 16 + e);
        }
        tmem_ld_wait();
        if (j < 3) tmem_ld_32x32b_x16(t_lane + (n & 1) * 256 + (j + 1) * 64, accb[(j + 1) & 1]);
        uint32_t packed[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = apply_act<ACT_GELU>(fmaf(__uint_as_float(acc[4 * e]), rstd, fmaf(-rmean, sc[e].x, sh[e].x)));
          const float v1 = apply_act<ACT_GELU>(fmaf(__uint_as_float(acc[4 * e + 1]), rstd, fmaf(-rmean, sc[e].y, sh[e].y)));
          const float v2 = apply_act<ACT_GELU>(fmaf(__uint_as_float(acc[4 * e + 2]), rstd, fmaf(-rmean, sc[e].z, sh[e].z)));
          const float v3 = apply_act<ACT_GELU>(fmaf(__uint_as_float(acc[4 * e + 3]), rstd, fmaf(-rmean, sc[e].w, sh[e].w)));
          packed[2 * e] = pack_bf16x2(v0, v1);
          packed[2 * e + 1] = pack_bf16x2(v2, v3);
        }
        *reinterpret_cast<uint4*>(buf + off0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4*>(buf + off1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&chunk_done[j]);
      }
      ++n;
      if (etid == 0) trace_mark(p, i, 3);  // E1 done (this thread)

      // ---------------- E2: x2 = acc + b2 + x1 -> Q in place -> TMA store, row statistics of x2 ----------------
      mbar_wait(&acc_full[n & 1], (n >> 1) & 1);
      tc_fence_after();
      if (etid == 0) trace_mark(p, i, 4);  // G2 done
      s1 = 0.f;
      s2 = 0.f;
      tmem_ld_32x32b_x16(t_lane + (n & 1) * 256, accb[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint8_t* buf = smem + kOffQ + j * kChunkBytes;
        const uint32_t(&acc)[16] = accb[j & 1];
        uint4 x1[2];
        x1[0] = *reinterpret_cast<const uint4*>(buf + off0);
        x1[1] = *reinterpret_cast<const uint4*>(buf + off1);
        float4 bb[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) bb[e] = __ldg(b2v + j * 16 + e);
        tmem_ld_wait();
        if (j < 3) tmem_ld_32x32b_x16(t_lane + (n & 1) * 256 + (j + 1) * 64, accb[(j + 1) & 1]);
        uint32_t packed[8];
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(x1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v0 = (__uint_as_float(acc[4 * e]) + bb[e].x) + bf16_lo(rw[2 * e]);
          const float v1 = (__uint_as_float(acc[4 * e + 1]) + bb[e].y) + bf16_hi(rw[2 * e]);
          const float v2 = (__uint_as_float(acc[4 * e + 2]) + bb[e].z) + bf16_lo(rw[2 * e + 1]);
          const float v3 = (__uint_as_float(acc[4 * e + 3]) + bb[e].w) + bf16_hi(rw[2 * e + 1]);
          s1 += (v0 + v1) + (v2 + v3);
          s2 = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, s2))));
          packed[2 * e] = pack_bf16x2(v0, v1);
          packed[2 * e + 1] = pack_bf16x2(v2, v3);
        }
        *reinterpret_cast<uint4*>(buf + off0) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4*>(buf + off1) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_ready[j]);  // warp 3 stores the chunk once all sixteen warps are here
      }
      ++n;
      if (cq >= 2) s_buf[(cq - 2) * 128 + row] = make_float2(s1, s2);
      if (p.stats_out != nullptr) {
        bar_sync(1, 512);
        if (cq < 2) {
          const float2 o = s_buf[cq * 128 + row];
          s1 += o.x;
          s2 += o.y;
          if (cq == 1) s_buf[128 + row] = make_float2(s1, s2);
        }
        bar_sync(1, 512);
        if (cq == 0 && valid) {
          const float2 o = s_buf[128 + row];
          const float mean = (s1 + o.x) * (1.0f / 256);
          const float var = fmaxf(fmaf(s2 + o.y, 1.0f / 256, -mean * mean), 0.0f);
          p.stats_out[grow] = make_float2(mean, rsqrtf(var + 1e-5f));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

int launch_vit_block(const VitBlockOp& op, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    HGR_CHECK_CUDA(cudaFuncSetAttribute(vit_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int total_tiles = (int)((op.p.rows + 127) / 128);
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (grid <= 0) return 0;
  HGR_CHECK_CUDA(launch_pdl(vit_block_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, op.a, op.w0, op.w1,
                            op.w2, op.x, op.o, op.p));
  return 0;
}

int build_vit_block_op(VitBlockOp& op, const void* attn_out, const void* x0, long long rows, const void* w_out,
                       const void* w1, const float* c1, const float* d1, const void* w2, const float* b2, void* x2,
                       float* stats_out) {
  memset(&op, 0, sizeof(op));
  // (rows, 256) seen as (64 columns, rows, 4 k-blocks): one box = the four [128 x 64] SWIZZLE_128B chunk buffers
  {
    const uint64_t dims[5] = {64, (uint64_t)rows, 4, 1, 1};
    const uint64_t strides[4] = {512, 128, 512 * (uint64_t)rows, 512 * (uint64_t)rows};
    const uint32_t box[5] = {64, 128, 4, 1, 1};
    if (int r = make_tensor_map_bf16(&op.a, attn_out, 5, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[5] = {64, (uint64_t)rows, 4, 1, 1};
    const uint64_t strides[4] = {512, 128, 512 * (uint64_t)rows, 512 * (uint64_t)rows};
    const uint32_t box[5] = {64, 128, 1, 1, 1};  // one chunk per request (Q is recycled chunk by chunk)
    if (int r = make_tensor_map_bf16(&op.x, x0, 5, dims, strides, box)) return r;
  }
  const void* ws[3] = {w_out, w1, w2};
  CUtensorMap* wm[3] = {&op.w0, &op.w1, &op.w2};
  for (int g = 0; g < 3; ++g) {
    const uint64_t dims[2] = {256, 256};
    const uint64_t strides[1] = {512};
    const uint32_t box[2] = {64, 256};
    if (int r = make_tensor_map_bf16(wm[g], ws[g], 2, dims, strides, box)) return r;
  }
  {
    const uint64_t dims[4] = {64, (uint64_t)rows, 4, 1};
    const uint64_t strides[3] = {512, 128, 512 * (uint64_t)rows};
    const uint32_t box[4] = {64, 128, 1, 1};
    if (int r = make_tensor_map_bf16(&op.o, x2, 4, dims, strides, box)) return r;
  }
  op.p.rows = rows;
  op.p.a0 = static_cast<const __nv_bfloat16*>(attn_out);
  op.p.x0 = static_cast<const __nv_bfloat16*>(x0);
  op.p.c1 = c1;
  op.p.d1 = d1;
  op.p.b2 = b2;
  op.p.stats_out = reinterpret_cast<float2*>(stats_out);
  op.p.reverse = 0;
  op.p.trace = nullptr;
  op.p.trace_tiles = 0;
  op.flops = 3 * 2.0 * (double)rows * 256 * 256;
  op.bytes = 2.0 * ((double)rows * 256 * 3 + 3.0 * 256 * 256);
  return 0;
}

}  // namespace hgr
