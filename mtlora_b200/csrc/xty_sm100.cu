// Adapter / weight gradient reductions over the token dimension on the 5th-gen tensor cores (sm_100a):
//
//     out[w, r] (+)= sum_m  Wide[m, w] * (rowscale(m) * Rank[m, r])
//
// for a list of job groups sharing one launch: dB_a = dY_a^T U_a, dA_a = G_a^T X_in(a) of every adapter of an
// MTLoRALinear (autograd of reference models/lora.py:260-265), or dW = dY^T X of a trainable dense weight
// (PatchMerging.reduction, lora.py:599-600). M (tokens) is huge, the outputs are tiny, every input element is
// needed once: the kernel is a pure HBM stream.
//
//   * Both operands are consumed straight from their row-major [m][.] layout: a [64 rows x 64 cols] TMA box with
//     the 128-byte swizzle IS the canonical MN-major UMMA operand atom (8 rows x 128 B), so tcgen05.mma runs with
//     a_major = b_major = MN and no thread ever touches the data (the mma.sync version this replaces was bound by
//     ldmatrix / shared-memory bandwidth at 40 % of HBM speed).
//   * job = (group, 128-column tile of the wide operand, 64-column chunk of the rank operand); unit = (job, row
//     range). Persistent CTAs walk the units; the 8-stage TMA ring keeps ~190 KB of loads in flight per SM.
//   * D[128 x 64] fp32 accumulates in TMEM (double-buffered so the reduction epilogue of one unit overlaps the
//     stream of the next) and is added to the fp32 result with vector / coalesced red.global.
//   * DropPath row scales (per-sample) are applied by four helper warps to the (narrow) rank tile in shared
//     memory between TMA arrival and MMA issue.
#include "kernels.cuh"
#include "linear_sm100.cuh"

#include <string.h>
#include <mutex>

namespace mtl {

namespace {

constexpr int XG_ROWS = 64;                 // contraction rows per pipeline stage
constexpr int XG_BOX_BYTES = XG_ROWS * 128; // one [64 x 64] bf16 box
constexpr int XG_STAGE_BYTES = 3 * XG_BOX_BYTES;   // wide box 0, wide box 1, rank box (rank_boxes == 1)
constexpr int XG_STAGES = 8;                // ring depth with one rank box per stage; fewer when the stages are wider
constexpr int XG_RING_BYTES = XG_STAGES * XG_STAGE_BYTES;
constexpr int XG_MAX_RANK_BOXES = 4;        // UMMA N = 64 .. 256
constexpr int XG_MAX_GROUPS = 24;
constexpr int XG_THREADS = 384;             // 4 control warps, 4 epilogue warps, 4 row-scale warps
constexpr int XG_TMEM_COLS = 512;           // 2 accumulators x (64 * rank_boxes <= 256) columns

struct XGroup {
  int wide_map, wide_stream, width, n_wt;   // wide operand: tensor map 0/1, stream (3rd coordinate), columns, 128-col tiles
  int rank_map, r0, rlen, n_rc;             // rank operand: tensor map 0/1, column range, 64-col chunks
  int kind;                                 // 0: out[w * ld + r], 1: out[r * ld + w]
  int out_ld;
  int job0;                                 // first job of the group
  int pad_;
  float* out;
  const float* rowscale;                    // [n_samples] or null
};

struct XParams {
  long M;
  long m_chunk;            // rows per unit (multiple of XG_ROWS)
  int n_groups, n_jobs, n_splits, n_units;
  int rows_per_sample, n_samples, any_rowscale;
  // 64-column boxes of the rank operand per pipeline stage = per accumulator (1..4). A wide "rank" operand (mtl_xty on
  // a trainable dense weight: dW = dY^T X, 192..768 columns on both sides) then re-streams the other operand once per
  // 256 columns instead of once per 64 (the kernel is bound by operand traffic from L2)
  int rank_boxes;
  XGroup g[XG_MAX_GROUPS];
};

struct Unit {
  int grp;
  int wide_c0, lane_lo, rank_c0, rank_valid;
  long m_begin;
  int n_steps;
};

__device__ __forceinline__ Unit decode_unit(const XParams& p, int u) {
  Unit t;
  const int s = u / p.n_jobs;
  const int job = u - s * p.n_jobs;
  int gi = 0;
  while (gi + 1 < p.n_groups && job >= p.g[gi + 1].job0) ++gi;
  const XGroup& g = p.g[gi];
  const int local = job - g.job0;
  const int wt = local / g.n_rc;
  const int rc = local - wt * g.n_rc;
  t.grp = gi;
  // the last tile of a wide operand is shifted left to stay inside the tensor; its duplicate lanes are skipped
  int c0 = wt * 128;
  int lim = g.width - 128;
  if (lim < 0) lim = 0;
  t.wide_c0 = c0 < lim ? c0 : lim;
  t.lane_lo = c0 - t.wide_c0;
  const int rcols = 64 * p.rank_boxes;
  t.rank_c0 = g.r0 + rc * rcols;
  int rv = g.r0 + g.rlen - t.rank_c0;
  t.rank_valid = rv < rcols ? rv : rcols;
  t.m_begin = static_cast<long>(s) * p.m_chunk;
  long m_end = t.m_begin + p.m_chunk;
  if (m_end > p.M) m_end = p.M;
  t.n_steps = static_cast<int>((m_end - t.m_begin + XG_ROWS - 1) / XG_ROWS);
  return t;
}

// UMMA shared-memory descriptor of an MN-major operand built from [64 rows x 128 B] SW128 boxes:
// canonical layout ((8 chunks, n_atoms), (8 rows, k)) : ((16 B, LBO), (128 B, SBO)), SBO = 1024 B between 8-row
// groups, LBO = distance between 64-column atoms (cute::UMMA make_umma_desc<Major::MN>, mma_traits_sm100.hpp).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* ptr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* ptr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(ptr), "f"(a) : "memory");
}

__global__ void __launch_bounds__(XG_THREADS, 1)
xty_umma_kernel(const __grid_constant__ CUtensorMap tm_w0, const __grid_constant__ CUtensorMap tm_w1,
                const __grid_constant__ CUtensorMap tm_r0, const __grid_constant__ CUtensorMap tm_r1,
                const __grid_constant__ XParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + XG_RING_BYTES;
  const uint32_t stage_bytes = (2 + p.rank_boxes) * XG_BOX_BYTES;
  const int n_stages = XG_RING_BYTES / stage_bytes;   // 8, 6, 4, 4
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (XG_STAGES + s); };
  auto scaled_bar = [&](int s) { return bar_base + 8u * (2 * XG_STAGES + s); };
  auto acc_full = [&](int b) { return bar_base + 8u * (3 * XG_STAGES + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (3 * XG_STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * XG_STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + XG_RING_BYTES + 8u * (3 * XG_STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w0);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_r0);
    tma_prefetch_desc(&tm_r1);
    for (int s = 0; s < XG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(scaled_bar(s), 128);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, XG_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp == 0) {
    // ============================================ TMA producer ============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const Unit t = decode_unit(p, u);
        const XGroup& g = p.g[t.grp];
        const CUtensorMap* wm = g.wide_map ? &tm_w1 : &tm_w0;
        const CUtensorMap* rm = g.rank_map ? &tm_r1 : &tm_r0;
        for (int st = 0; st < t.n_steps; ++st) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t dst = smem_base + stage * stage_bytes;
          const int row = static_cast<int>(t.m_begin) + st * XG_ROWS;
          mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
          tma_load_3d(dst, wm, full_bar(stage), t.wide_c0, row, g.wide_stream);
          tma_load_3d(dst + XG_BOX_BYTES, wm, full_bar(stage), t.wide_c0 + 64, row, g.wide_stream);
          // (boxes past the last column of the operand are zero-filled by TMA; columns past the group's range are
          // computed and dropped by the epilogue)
          for (int b = 0; b < p.rank_boxes; ++b)
            tma_load_2d(dst + (2 + b) * XG_BOX_BYTES, rm, full_bar(stage), t.rank_c0 + 64 * b, row);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================================= MMA issuer =============================================
    if (lane == 0) {
      // M = 128 (wide columns), N = 64 * rank_boxes (rank columns), both operands MN-major
      const uint32_t idesc = umma_idesc_bf16_m128(64 * p.rank_boxes) | (1u << 15) | (1u << 16);
      const uint32_t acc_cols = 64 * p.rank_boxes;
      int stage = 0;
      uint32_t phase = 0, ub = 0;
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ub) {
        const Unit t = decode_unit(p, u);
        const uint32_t buf = ub & 1u;
        mbar_wait(acc_empty(buf), ((ub >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * acc_cols;
        for (int st = 0; st < t.n_steps; ++st) {
          mbar_wait(p.any_rowscale ? scaled_bar(stage) : full_bar(stage), phase);
          tc_fence_after();
          const uint32_t base = smem_base + stage * stage_bytes;
#pragma unroll
          for (int ks = 0; ks < XG_ROWS / 16; ++ks) {
            const uint64_t adesc = umma_desc_mn_sw128(base + ks * 2048, XG_BOX_BYTES);
            const uint64_t bdesc = umma_desc_mn_sw128(base + 2 * XG_BOX_BYTES + ks * 2048, XG_BOX_BYTES);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (st | ks) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(acc_full(buf));
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ========================================= reduction epilogue =========================================
    const int q4 = warp & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    uint32_t ub = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++ub) {
      const Unit t = decode_unit(p, u);
      const XGroup& g = p.g[t.grp];
      const uint32_t buf = ub & 1u;
      mbar_wait(acc_full(buf), (ub >> 1) & 1u);
      tc_fence_after();
      const int li = q4 * 32 + lane;
      const int w = t.wide_c0 + li;
      const bool ok = li >= t.lane_lo && w < g.width;
#pragma unroll 1
      for (int c = 0; c < 4 * p.rank_boxes; ++c) {
        if (c * 16 >= t.rank_valid) break;
        uint32_t r[16];
        tmem_ld16(t_lane + buf * 64 * p.rank_boxes + c * 16, r);
        tmem_ld_wait();
        if (ok) {
          if (g.kind == 0) {
            float* o = g.out + static_cast<size_t>(w) * g.out_ld + t.rank_c0 + c * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int left = t.rank_valid - (c * 16 + 4 * i);
              if (left >= 4 && ((reinterpret_cast<uintptr_t>(o + 4 * i) & 15) == 0)) {
                red_add_v4(o + 4 * i, __uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                           __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (e < left) red_add(o + 4 * i + e, __uint_as_float(r[4 * i + e]));
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (c * 16 + i < t.rank_valid)
                red_add(g.out + static_cast<size_t>(t.rank_c0 + c * 16 + i) * g.out_ld + w, __uint_as_float(r[i]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty(buf));
    }
  } else if (warp >= 8 && p.any_rowscale) {
    // ====================================== DropPath row scaling of the rank tile ======================================
    const int tid = threadIdx.x - 256;   // 0..127: row = tid / 2, half of the 128-byte row = tid % 2
    const int row = tid >> 1, half = tid & 1;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const Unit t = decode_unit(p, u);
      const float* rs = p.g[t.grp].rowscale;
      for (int st = 0; st < t.n_steps; ++st) {
        mbar_wait(full_bar(stage), phase);
        if (rs != nullptr) {
          const long grow = t.m_begin + static_cast<long>(st) * XG_ROWS + row;
          long smp = grow / p.rows_per_sample;
          if (smp >= p.n_samples) smp = p.n_samples - 1;
          const float s = rs[smp];
          if (s != 1.f) {
            uint8_t* rowp = smem_gen + stage * stage_bytes + 2 * XG_BOX_BYTES + row * 128 + half * 64;   // rank_boxes == 1
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 v = *reinterpret_cast<uint4*>(rowp + 16 * i);
              uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) wv[e] = pack_bf16x2(bf16lo_to_f32(wv[e]) * s, bf16hi_to_f32(wv[e]) * s);
              *reinterpret_cast<uint4*>(rowp + 16 * i) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
          }
          fence_proxy_async_smem();
        }
        mbar_arrive(scaled_bar(stage));
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, XG_TMEM_COLS);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int launch_xty_groups(const XtyOperand* wide, const XtyOperand* rank, const XtyJobGroup* groups, int n_groups,
                      long M, int rows_per_sample, cudaStream_t stream) {
  MTL_REQUIRE(M > 0 && M < (1l << 31), "xty: M=%ld out of range", M);
  MTL_REQUIRE(n_groups > 0 && n_groups <= XG_MAX_GROUPS, "xty: %d job groups (max %d)", n_groups, XG_MAX_GROUPS);
  XParams p;
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.n_groups = n_groups;
  p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  p.n_samples = static_cast<int>((M + p.rows_per_sample - 1) / p.rows_per_sample);
  // rank boxes per stage: as many as the widest rank range needs (DropPath row scaling rewrites ONE box per stage)
  p.rank_boxes = 1;
  {
    int widest = 0;
    bool scaled = false;
    for (int i = 0; i < n_groups; ++i) {
      if (groups[i].rlen > widest) widest = groups[i].rlen;
      if (groups[i].rowscale != nullptr) scaled = true;
    }
    if (!scaled) {
      p.rank_boxes = (widest + 63) / 64;
      if (p.rank_boxes > XG_MAX_RANK_BOXES) p.rank_boxes = XG_MAX_RANK_BOXES;
      if (p.rank_boxes < 1) p.rank_boxes = 1;
    }
  }
  int jobs = 0;
  for (int i = 0; i < n_groups; ++i) {
    const XtyJobGroup& s = groups[i];
    XGroup& g = p.g[i];
    MTL_REQUIRE(s.wide_op >= 0 && s.wide_op < 2 && s.rank_op >= 0 && s.rank_op < 2, "xty: bad operand index");
    const XtyOperand& w = wide[s.wide_op];
    const XtyOperand& r = rank[s.rank_op];
    MTL_REQUIRE(w.base != nullptr && r.base != nullptr && s.out != nullptr, "xty: NULL operand");
    MTL_REQUIRE(w.cols >= 72, "xty: wide operand needs >= 72 columns (got %d)", w.cols);
    MTL_REQUIRE(s.wide_stream >= 0 && s.wide_stream < w.streams, "xty: stream %d out of range", s.wide_stream);
    MTL_REQUIRE(s.r0 >= 0 && s.rlen > 0 && s.r0 + s.rlen <= r.cols && s.r0 % 8 == 0 && s.rlen % 4 == 0,
                "xty: bad rank range [%d, +%d) of %d", s.r0, s.rlen, r.cols);
    g.wide_map = s.wide_op;
    g.wide_stream = s.wide_stream;
    g.width = w.cols;
    g.n_wt = (w.cols + 127) / 128;
    g.rank_map = s.rank_op;
    g.r0 = s.r0;
    g.rlen = s.rlen;
    g.n_rc = (s.rlen + 64 * p.rank_boxes - 1) / (64 * p.rank_boxes);
    g.kind = s.out_rank_major ? 1 : 0;
    g.out_ld = s.out_ld;
    g.out = s.out;
    g.rowscale = s.rowscale;
    if (s.rowscale != nullptr) p.any_rowscale = 1;
    g.job0 = jobs;
    jobs += g.n_wt * g.n_rc;
  }
  p.n_jobs = jobs;

  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  // row splits: enough units to fill and balance the persistent grid; a unit costs its rows plus ~6 steps' worth of
  // pipeline fill, accumulator drain and red.global epilogue (measured ~3 us), so small problems take ONE unit per CTA
  const long max_splits = (M + 4 * XG_ROWS - 1) / (4 * XG_ROWS);
  long best_splits = 1;
  double best_cost = 1e300;
  const long lo = (1L * n_sm + jobs - 1) / jobs, hi = (8L * n_sm + jobs - 1) / jobs;
  for (long sp = (lo < 1 ? 1 : lo); sp <= (hi < 1 ? 1 : hi); ++sp) {
    const long s = sp > max_splits ? max_splits : sp;
    long mc = (M + s - 1) / s;
    mc = (mc + XG_ROWS - 1) / XG_ROWS * XG_ROWS;
    const long ns = (M + mc - 1) / mc;
    const long units = ns * jobs;
    const long waves = (units + n_sm - 1) / n_sm;
    const double cost = static_cast<double>(waves) * (mc + 6.0 * XG_ROWS);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best_splits = ns;
      p.m_chunk = mc;
    }
    if (s == max_splits) break;
  }
  p.n_splits = static_cast<int>(best_splits);
  p.n_units = p.n_splits * p.n_jobs;

  CUtensorMap tm[4];
  for (int i = 0; i < 2; ++i) {
    const XtyOperand& w = wide[i].base != nullptr ? wide[i] : wide[0];
    MTL_REQUIRE(w.base != nullptr, "xty: wide operand 0 missing");
    if (int e = make_tmap(&tm[i], w.base, w.cols, M, w.streams, 64, XG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B, w.pitch))
      return e;
    const XtyOperand& r = rank[i].base != nullptr ? rank[i] : rank[0];
    MTL_REQUIRE(r.base != nullptr, "xty: rank operand 0 missing");
    if (int e = make_tmap(&tm[2 + i], r.base, r.cols, M, 0, 64, XG_ROWS, CU_TENSOR_MAP_SWIZZLE_128B, r.pitch))
      return e;
  }
  const uint32_t smem_bytes = XG_STAGES * XG_STAGE_BYTES + 1024 + 1024;
  static bool attr_done[64] = {};
  MTL_CHECK_CUDA(ensure_max_dyn_smem(attr_done, xty_umma_kernel, XG_STAGES * XG_STAGE_BYTES + 2048));
  const int grid = p.n_units < n_sm ? p.n_units : n_sm;
  xty_umma_kernel<<<grid, XG_THREADS, smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], p);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
