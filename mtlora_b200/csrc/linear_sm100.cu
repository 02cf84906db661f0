// Fused MTLoRALinear GEMM for sm_100a: frozen dense product + (1+T) low-rank adapters in one persistent kernel.
//
// Reference semantics: models/lora.py:253-284 (MTLoRALinear.forward) — the reference evaluates F.linear(x, W, b),
// 2(1+T) skinny matmuls and (1+T) mul/add passes as separate kernels. Here (see linear_sm100.cuh for the algebra):
//   * one persistent CTA per SM walks the (128-row tile, column split) work list;
//   * TMA stages X / W / A_cat / B_cat tiles in a shared-memory ring, tcgen05.mma accumulates in TMEM;
//   * the rank-space activations U = X.A_cat^T are formed once per tile, converted to a bf16 K-major operand in
//     shared memory and replayed against B_cat for every output stream ("stream-sequential" delta accumulators);
//   * the dense accumulator P of a column chunk is shared by all (1+T) stream epilogues;
//   * two epilogue groups of 8 warps (4 TMEM lane quadrants x 2 column halves) ping-pong over the items
//     (chunk, stream): TMEM -> registers -> bias / DropPath scale / GELU, GELU' / residual -> bf16 -> 64-byte-swizzled
//     2 KB smem slab -> TMA store, so every output element is written exactly once by coalesced bulk stores;
//   * one leader warp per group polls the accumulator mbarriers, the others block on a named barrier; the producer
//     prefetches the epilogue inputs (residual, GELU' factor) of its tile into L2.
//
// Warp roles (640 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-19 = epilogue
// (warp % 4 = TMEM lane quadrant); setmaxnreg gives the epilogue warpgroups 104 registers, the control one 64.
// Planner notes (plan_linear): items of 128 columns wherever TMEM allows (64-column items cost 14-26 % more at any K),
// except the GELU pair, whose long items need the double-buffered accumulators; a third epilogue group was slower.
#include "linear_sm100.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>
#include <mutex>

namespace mtl {

namespace {

constexpr int kGroups = 2;                        // epilogue groups: each owns one accumulator item at a time
constexpr int kColSplit = 2;                      // warps per TMEM lane quadrant inside a group: each takes 32 of the
                                                  // 64 columns of a half (the epilogue is latency-bound per warp, so
                                                  // 4 warps per scheduler instead of 2 nearly halve the item time)
constexpr int kGroupWarps = 4 * kColSplit;
constexpr int kEpiWarps = kGroupWarps * kGroups;
constexpr int kThreads = 128 + 32 * kEpiWarps;   // 4 control warps + epilogue warps
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kTileABytes = LIN_BM * LIN_BK * 2;  // 16 KiB: one [128 x 64] bf16 K-major SW128 tile
constexpr int kPieceCols = 64 / kColSplit;        // columns of a half one warp converts and stores
constexpr int kPieceGran = kPieceCols / 16;
constexpr int kSlabBytes = 32 * kPieceCols * 2;   // 2 KiB: one warp's [32 rows x 32 cols] bf16 store slab (SW64)
// launch: 640 threads x 96 registers; the control warpgroup drops to kCtrlRegs, the epilogue warpgroups grow to kEpiRegs
constexpr int kCtrlRegs = 64, kEpiRegs = 104;
static_assert(128 * kCtrlRegs + kEpiThreads * kEpiRegs <= kThreads * 96, "register pool of the CTA exceeded");
static_assert(kColSplit == 2, "slab swizzle / tensor-map boxes are written for 32-column pieces");
constexpr int kMaxStages = 8;
constexpr int kMaxUAtoms = LIN_MAX_GRAN / 4;

struct SmemLayout {
  uint32_t usm;    // U operand: n_uatoms * 16 KiB, after the ring stages
  uint32_t slabs;  // kEpiWarps * n_slabs * kSlabBytes
  uint32_t bars;
  uint32_t total;
};

__host__ __device__ inline SmemLayout smem_layout(int n_stages, int stage_bytes, int r_pad, int n_slabs) {
  SmemLayout l;
  const uint32_t n_uatoms = (r_pad + 63) / 64;
  l.usm = n_stages * stage_bytes;
  l.slabs = l.usm + n_uatoms * kTileABytes;
  l.bars = l.slabs + kEpiWarps * n_slabs * kSlabBytes;
  l.total = l.bars + 1024;
  return l;
}

// Shared-memory address of the i-th Up tile ([BN x 64] bf16, SW128) held by the ring stage starting at `stage_addr`:
// tile 0 sits in the stage's B half, tiles 1.. in its A half (16 KiB = two 64-row tiles or one 128-row tile).
__device__ __forceinline__ uint32_t up_tile_addr(uint32_t stage_addr, int i, int bn) {
  return i == 0 ? stage_addr + kTileABytes : stage_addr + (i - 1) * bn * 128;
}

__device__ __forceinline__ bool col_in_out(const LinPlan& p, int j, int col) {
  return (col >= p.out_r0[j][0] && col < p.out_r0[j][0] + p.out_len[j][0]) ||
         (col >= p.out_r0[j][1] && col < p.out_r0[j][1] + p.out_len[j][1]);
}

// ---- TMA stores (shared -> global), bulk-group completion ------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const void* tmap, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tmap)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void group_bar_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(32 * kGroupWarps) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEpiThreads) : "memory");
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 on sm_100) -----------------------------------------------------------
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
#define MTL_C2(c) pack2(c, c)

// Branch-free normal CDF for the GELU epilogues: Phi(x) = 0.5 + x * Q(x^2), Q an odd-minimax fit of
// 0.5 * erf(x / sqrt 2) on |x| <= 4.25 (max abs error 1.1e-5, far below the bf16 resolution of the outputs); the
// argument is clamped, so Phi saturates at 1 - 1e-5 / 1e-5. nn.GELU() is the exact-erf GELU (reference :45):
// GELU(x) = x * Phi(x), GELU'(x) = Phi(x) + x * phi(x). Two elements per instruction.
// Evaluated breadth-first over the 8 element pairs of a 16-column granule so that every Horner step issues 8
// independent FFMA2 (the epilogue runs with only two warps per scheduler: dependent chains would stall the issue port).
// xc / t (optional outputs): the clamped arguments and their squares, reused by the derivative epilogue
__device__ __forceinline__ void phi16(const float (&v)[16], uint64_t (&phi)[8], uint64_t (&xc)[8], uint64_t (&t)[8]) {
  uint64_t q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    xc[i] = pack2(fminf(fmaxf(v[2 * i], -4.25f), 4.25f), fminf(fmaxf(v[2 * i + 1], -4.25f), 4.25f));
#pragma unroll
  for (int i = 0; i < 8; ++i) t[i] = mul2(xc[i], xc[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = fma2(MTL_C2(5.565109802e-11f), t[i], MTL_C2(-5.327931323e-09f));
#define MTL_HORNER(C)              \
  _Pragma("unroll") for (int i = 0; i < 8; ++i) q[i] = fma2(q[i], t[i], MTL_C2(C));
  MTL_HORNER(2.255476184e-07f)
  MTL_HORNER(-5.626496851e-06f)
  MTL_HORNER(9.341922331e-05f)
  MTL_HORNER(-1.108562514e-03f)
  MTL_HORNER(9.815966060e-03f)
  MTL_HORNER(-6.634444852e-02f)
  MTL_HORNER(3.989024332e-01f)
#undef MTL_HORNER
#pragma unroll
  for (int i = 0; i < 8; ++i) phi[i] = fma2(xc[i], q[i], MTL_C2(0.5f));
}
__device__ __forceinline__ void phi16(const float (&v)[16], uint64_t (&phi)[8]) {
  uint64_t xc[8], t[8];
  phi16(v, phi, xc, t);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GELU(v) for a granule, packed to bf16x2
__device__ __forceinline__ void gelu16_pack(const float (&v)[16], uint32_t (&pk)[8]) {
  uint64_t phi[8];
  phi16(v, phi);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float g0, g1;
    unpack2(mul2(pack2(v[2 * i], v[2 * i + 1]), phi[i]), g0, g1);
    pk[i] = pack_bf16x2(g0, g1);
  }
}
// GELU(v) and GELU'(v) = Phi(v) + v * pdf(v) for a granule, both packed to bf16x2 (fc1 forward when the consuming
// fc2 backward wants the derivative factor instead of the pre-activation)
__device__ __forceinline__ void gelu_and_grad16_pack(const float (&v)[16], uint32_t (&pk_gelu)[8], uint32_t (&pk_grad)[8]) {
  uint64_t phi[8], xc[8], t[8];
  phi16(v, phi, xc, t);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float g0, g1, d0, d1, e0, e1;
    const uint64_t v2 = pack2(v[2 * i], v[2 * i + 1]);
    unpack2(mul2(v2, phi[i]), g0, g1);
    pk_gelu[i] = pack_bf16x2(g0, g1);
    // x pdf(x) = x exp(-x^2 / 2) / sqrt(2 pi) on the clamped argument (its square comes from the Phi polynomial): beyond
    // |x| = 4.25 the true value and the clamped one (2e-4) both vanish against GELU' ~ 0 / 1 at bf16 resolution
    unpack2(mul2(t[i], MTL_C2(-0.72134752044448170368f)), e0, e1);
    const uint64_t pdf = pack2(ex2_approx(e0), ex2_approx(e1));
    unpack2(fma2(mul2(xc[i], MTL_C2(0.39894228040143267794f)), pdf, phi[i]), d0, d1);
    pk_grad[i] = pack_bf16x2(d0, d1);
  }
}
// v *= GELU'(a) for a granule; a given as 8 packed bf16x2 words. GELU'(x) = Phi(x) + x * pdf(x)
__device__ __forceinline__ void gelu_grad16_mul(float (&v)[16], const uint32_t (&aw)[8]) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[2 * i] = bf16lo_to_f32(aw[i]);
    a[2 * i + 1] = bf16hi_to_f32(aw[i]);
  }
  uint64_t phi[8];
  phi16(a, phi);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // pdf = exp(-x^2 / 2) / sqrt(2 pi)
    const float p0 = 0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * a[2 * i] * a[2 * i]);
    const float p1 = 0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * a[2 * i + 1] * a[2 * i + 1]);
    float d0, d1;
    unpack2(fma2(pack2(a[2 * i], a[2 * i + 1]), pack2(p0, p1), phi[i]), d0, d1);
    v[2 * i] *= d0;
    v[2 * i + 1] *= d1;
  }
}

struct WorkItem {
  int m0, split, n_my_chunks;
};
// Debug timeline (MTL_LINEAR_TRACE=<file>): CTA 0 records (globaltimer, code) pairs per role.
template <bool ON>
struct Tracer {
  unsigned long long* buf;
  int n;
  __device__ __forceinline__ void init(unsigned long long* base, int role) {
    if constexpr (ON) {
      buf = (base != nullptr && blockIdx.x == 0) ? base + role * 2048 : nullptr;
      n = 0;
    }
  }
  __device__ __forceinline__ void ev(unsigned long long code) {
    if constexpr (ON) {
      if (buf != nullptr && n < 1023) {
        buf[2 * n] = globaltimer_ns();
        buf[2 * n + 1] = code;
        ++n;
      }
    }
  }
};
__device__ __forceinline__ WorkItem get_work(const LinPlan& p, int w) {
  WorkItem it;
  if (p.n_splits == 1) {   // the common case (many row tiles): no integer divisions on the per-item path
    it.split = 0;
    it.m0 = w * LIN_BM;
    it.n_my_chunks = p.n_chunks;
    return it;
  }
  const int m_tile = w / p.n_splits;
  it.split = w - m_tile * p.n_splits;
  it.m0 = m_tile * LIN_BM;
  it.n_my_chunks = (p.n_chunks - it.split + p.n_splits - 1) / p.n_splits;
  return it;
}

// EP: LinEpilogue; HAS_RES: a residual tensor is added in the epilogue (compile-time specialisation keeps the
// per-element instruction count of the epilogue — the bottleneck of the wide stage-0/1 layers — minimal)
template <int EP, bool HAS_RES, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
mtl_linear_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_down, const __grid_constant__ CUtensorMap tm_up,
                  const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_y2,
                  const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_in,
                  const __grid_constant__ CUtensorMap tm_pf, const __grid_constant__ LinPlan p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t stage_bytes = kTileABytes + p.stage_b_bytes;
  const SmemLayout L = smem_layout(p.n_stages, stage_bytes, p.u_in ? 0 : p.R_pad, p.n_slabs);
  const uint32_t usm_base = smem_base + L.usm;
  const uint32_t bar_base = smem_base + L.bars;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t u_full = bar_base + 8u * 16;
  const uint32_t u_ready = bar_base + 8u * 17;
  const uint32_t usm_free = bar_base + 8u * 18;
  auto p_full = [&](int b) { return bar_base + 8u * (20 + b); };
  auto p_empty = [&](int b) { return bar_base + 8u * (22 + b); };
  auto d_full = [&](int b) { return bar_base + 8u * (24 + b); };    // b = group * 2 + buffer
  auto d_empty = [&](int b) { return bar_base + 8u * (32 + b); };
  auto in_bar = [&](int ew, int slab) { return bar_base + 8u * (48 + ew * 4 + slab); };  // epilogue-input slabs
  const uint32_t tmem_slot = bar_base + 8u * 40;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L.bars + 8u * 40);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kb = (p.Kc + LIN_BK - 1) / LIN_BK;
  const int n_uatoms = (p.R_pad + 63) / 64;
  const bool multi = p.n_regions > 1;
  const int n_work = p.n_work;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_y);
    if (p.R_pad > 0) {
      tma_prefetch_desc(&tm_down);
      tma_prefetch_desc(&tm_up);
    }
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(u_full, 1);
    mbar_init(u_ready, kEpiThreads);
    mbar_init(usm_free, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(p_full(b), 1);
      mbar_init(p_empty(b), kEpiThreads);
    }
    for (int b = 0; b < 2 * kGroups; ++b) {
      mbar_init(d_full(b), 1);
      mbar_init(d_empty(b), 32 * kGroupWarps);
    }
    for (int e = 0; e < kEpiWarps; ++e)
      for (int b = 0; b < 4; ++b) mbar_init(in_bar(e, b), 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // TMEM columns: [0, u_cols) rank-space accumulators U; then P[n_pbuf] (multi); then D[group][n_dbuf]
  // item G -> group G % kGroups, that group's k-th item (k = G / kGroups) -> buffer k % n_dbuf, use k / n_dbuf
  const uint32_t p_col0 = p.acc_col0;
  const uint32_t d_col0 = p.acc_col0 + (multi ? p.n_pbuf * p.BN : 0);

  // Register rebalancing: the 4 control warps (one warpgroup) need few registers, the epilogue warps many.
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
  if (warp == 0) {
    // ============================================ TMA producer ============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      Tracer<TRACE> tr;
      tr.init(p.trace, 0);
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const WorkItem it = get_work(p, w);
        tr.ev(1000000ull + w);
        // epilogue inputs of this work item (residual / GELU' argument): pull them into L2 now, their consumers run
        // several microseconds from now
        for (int sj = 0; sj < p.in_streams; ++sj)
          for (int ci = 0; ci < it.n_my_chunks; ++ci) {
            const int c0 = (it.split + ci * p.n_splits) * p.BN;
            for (int hcol = c0; hcol < c0 + p.BN && hcol < p.Nn; hcol += 64) tma_prefetch_l2_3d(&tm_pf, hcol, it.m0, sj);
          }
        // phase 1: rank-space ("down") products (u_in: U was produced by an earlier launch)
        for (int g = 0; g < (p.u_in ? 0 : p.n_groups); ++g) {
          const int len = p.grp_len[g];
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.wait_hint_ns);
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            const uint32_t b_dst = a_dst + kTileABytes;
            mbar_arrive_expect_tx(full_bar(stage), kTileABytes + len * 128);
            tma_load_3d(a_dst, &tm_x, full_bar(stage), kb * LIN_BK, it.m0, p.grp_in[g]);
            for (int i = 0; i < len / 16; ++i)
              tma_load_2d(b_dst + i * 2048, &tm_down, full_bar(stage), kb * LIN_BK, p.grp_r0[g] + 16 * i);
            advance();
          }
        }
        // phase 2: dense product + adapter replay per output chunk
        for (int ci = 0; ci < it.n_my_chunks; ++ci) {
          const int c = it.split + ci * p.n_splits;
          for (int i = 0; i < p.n_main; ++i) {
            for (int kb = 0; kb < n_kb; ++kb) {
              mbar_wait(empty_bar(stage), phase ^ 1u, p.wait_hint_ns);
              const uint32_t a_dst = smem_base + stage * stage_bytes;
              const uint32_t b_dst = a_dst + kTileABytes;
              mbar_arrive_expect_tx(full_bar(stage), kTileABytes + p.BN * 128);
              tma_load_3d(a_dst, &tm_x, full_bar(stage), kb * LIN_BK, it.m0, p.main_in[i]);
              tma_load_2d(b_dst, &tm_w, full_bar(stage), kb * LIN_BK, c * p.BN);
              tr.ev(2000000ull + ci * 1000 + kb);
              advance();
            }
          }
          // Up tiles of the chunk, up_pack rank atoms per ring stage: the first in the B half, further ones (wide rank
          // spaces only) in the A half, which these stages do not otherwise use
          for (int a0 = 0; a0 < n_uatoms; a0 += p.up_pack) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.wait_hint_ns);
            const int na = n_uatoms - a0 < p.up_pack ? n_uatoms - a0 : p.up_pack;
            mbar_arrive_expect_tx(full_bar(stage), na * p.BN * 128 + (p.u_in ? kTileABytes : 0));
            for (int i = 0; i < na; ++i)
              tma_load_2d(up_tile_addr(smem_base + stage * stage_bytes, i, p.BN), &tm_up, full_bar(stage), (a0 + i) * 64,
                          c * p.BN);
            // u_in (up_pack == 1): the [128 x 64] U atom of this rank block rides in the A half of the same stage
            if (p.u_in) tma_load_2d(smem_base + stage * stage_bytes, &tm_u, full_bar(stage), a0 * 64, it.m0);
            advance();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================================= MMA issuer =============================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      uint32_t lw = 0, Cn = 0, G = 0;  // local work / chunk / item counters (same sequence in the epilogue warps)
      Tracer<TRACE> tr;
      tr.init(p.trace, 1);
      for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++lw) {
        const WorkItem it = get_work(p, w);
        tr.ev(1000000ull + w);
        // ---- phase 1: U[:, group] = X[in] . Down[group]^T (the U columns were drained before u_ready of the
        //      previous work item, which this thread has already waited for)
        for (int g = 0; g < (p.u_in ? 0 : p.n_groups); ++g) {
          const uint32_t idesc = umma_idesc_bf16_m128(p.grp_len[g]);
          const uint32_t d_tmem = tmem_base + p.grp_r0[g];
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(full_bar(stage), phase, p.wait_hint_ns);
            tc_fence_after();
            const uint32_t a_src = smem_base + stage * stage_bytes;
            const uint64_t adesc = umma_desc_sw128(a_src);
            const uint64_t bdesc = umma_desc_sw128(a_src + kTileABytes);
            for (int q = 0; q < 4; ++q) {
              if (kb * LIN_BK + q * 16 >= p.Kc) break;
              umma_bf16(d_tmem, adesc + 2 * q, bdesc + 2 * q, idesc, ((kb | q) != 0 || p.grp_acc[g]) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
            advance();
          }
        }
        if (p.R_pad > 0 && !p.u_in) umma_commit(u_full);
        bool u_waited = (p.R_pad == 0) || p.u_in;

        for (int ci = 0; ci < it.n_my_chunks; ++ci) {
          const int c = it.split + ci * p.n_splits;
          int n_eff = p.Nn - c * p.BN;
          if (n_eff > p.BN) n_eff = p.BN;
          const uint32_t idesc_bn = umma_idesc_bf16_m128(n_eff);
          uint32_t acc_dense;  // where the dense product accumulates
          uint32_t g0 = 0;
          if (multi) {
            const uint32_t pb = Cn % p.n_pbuf;
            mbar_wait(p_empty(pb), ((Cn / p.n_pbuf) & 1u) ^ 1u, p.wait_hint_ns);
            acc_dense = tmem_base + p_col0 + pb * p.BN;
          } else {
            const uint32_t k = G / kGroups;
            g0 = (G % kGroups) * 2 + k % p.n_dbuf;
            mbar_wait(d_empty(g0), ((k / p.n_dbuf) & 1u) ^ 1u, p.wait_hint_ns);
            acc_dense = tmem_base + d_col0 + ((G % kGroups) * p.n_dbuf + k % p.n_dbuf) * p.BN;
          }
          tc_fence_after();
          tr.ev(3000000ull + ci);   // accumulator acquired
          bool first = true;
          for (int i = 0; i < p.n_main; ++i) {
            for (int kb = 0; kb < n_kb; ++kb) {
              mbar_wait(full_bar(stage), phase, p.wait_hint_ns);
              tr.ev(2000000ull + ci * 1000 + kb);
              tc_fence_after();
              const uint32_t a_src = smem_base + stage * stage_bytes;
              const uint64_t adesc = umma_desc_sw128(a_src);
              const uint64_t bdesc = umma_desc_sw128(a_src + kTileABytes);
              for (int q = 0; q < 4; ++q) {
                if (kb * LIN_BK + q * 16 >= p.Kc) break;
                umma_bf16(acc_dense, adesc + 2 * q, bdesc + 2 * q, idesc_bn, first ? 0u : 1u);
                first = false;
              }
              umma_commit(empty_bar(stage));
              advance();
            }
          }
          if (multi) umma_commit(p_full(Cn % p.n_pbuf));
          if (n_uatoms > 0) {
            if (!u_waited) {
              mbar_wait(u_ready, lw & 1u, p.wait_hint_ns);
              u_waited = true;
            }
            // the Up tiles of all rank atoms of this chunk occupy n_up_stages consecutive ring stages
            uint32_t up_addr[kMaxUAtoms], ua_addr[kMaxUAtoms];
            {
              int s = stage;
              uint32_t ph = phase;
              for (int a0 = 0; a0 < n_uatoms; a0 += p.up_pack) {
                mbar_wait(full_bar(s), ph, p.wait_hint_ns);
                for (int i = 0; i < p.up_pack && a0 + i < n_uatoms; ++i) {
                  up_addr[a0 + i] = up_tile_addr(smem_base + s * stage_bytes, i, p.BN);
                  ua_addr[a0 + i] = p.u_in ? smem_base + s * stage_bytes : usm_base + (a0 + i) * kTileABytes;
                }
                if (++s == p.n_stages) {
                  s = 0;
                  ph ^= 1u;
                }
              }
            }
            tc_fence_after();
            const int n_items = multi ? p.S_out : 1;
            for (int j = 0; j < n_items; ++j) {
              uint32_t acc;
              bool started;
              uint32_t db = 0;
              if (multi) {
                const uint32_t k = G / kGroups;
                if (p.d_shared) {
                  // One delta accumulator serves both groups (item G is its G-th use). This thread sees every phase
                  // of d_empty(0) in order; the groups alternate, so each gets its OWN full barrier (a group waiting on a
                  // shared one would skip every other phase and mistake an old completion for its own).
                  mbar_wait(d_empty(0), (G & 1u) ^ 1u, p.wait_hint_ns);
                  tc_fence_after();
                  db = (G % kGroups) * 2;
                  acc = tmem_base + d_col0;
                } else {
                  db = (G % kGroups) * 2 + k % p.n_dbuf;
                  mbar_wait(d_empty(db), ((k / p.n_dbuf) & 1u) ^ 1u, p.wait_hint_ns);
                  tc_fence_after();
                  acc = tmem_base + d_col0 + ((G % kGroups) * p.n_dbuf + k % p.n_dbuf) * p.BN;
                }
                started = false;
              } else {
                acc = acc_dense;
                started = !first;
              }
              for (int a = 0; a < n_uatoms; ++a) {
                const uint64_t adesc = umma_desc_sw128(ua_addr[a]);
                const uint64_t bdesc = umma_desc_sw128(up_addr[a]);
                for (int q = 0; q < 4; ++q) {
                  const int col = a * 64 + q * 16;
                  if (col >= p.R_pad) break;
                  if (!col_in_out(p, j, col)) continue;
                  umma_bf16(acc, adesc + 2 * q, bdesc + 2 * q, idesc_bn, started ? 1u : 0u);
                  started = true;
                }
              }
              if (multi) {
                umma_commit(d_full(db));
                ++G;
              }
            }
            for (int a0 = 0; a0 < n_uatoms; a0 += p.up_pack) {
              umma_commit(empty_bar(stage));
              advance();
            }
          }
          tr.ev(4000000ull + ci);   // chunk issued
          if (multi) {
            ++Cn;
          } else {
            umma_commit(d_full(g0));
            ++G;
          }
        }
        if (p.R_pad > 0 && !p.u_in) umma_commit(usm_free);
      }
    }
    __syncwarp();
  }
  } else {
    // ======================================= U converter + epilogue =======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int q4 = warp & 3;                 // TMEM lane quadrant
    const int ew = warp - 4;                 // 0 .. kEpiWarps-1
    const uint32_t grp = (ew >> 2) % kGroups;   // epilogue group
    const int sel = (ew >> 2) / kGroups;        // which 32-column piece of every 64-column half this warp owns
    const int row = q4 * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    uint8_t* slab_gen = smem_gen + L.slabs + ew * p.n_slabs * kSlabBytes;
    const uint32_t slab_base = smem_base + L.slabs + ew * p.n_slabs * kSlabBytes;
    int slab_k = 0;
    bool first_half = true;
    uint32_t in_ph = 0;   // phase bit of every epilogue-input slab barrier of this warp
    auto nxt_slab = [&](int k) { return k + 1 == p.n_slabs ? 0 : k + 1; };
    const bool grp_leader = (sel == 0 && q4 == 0);   // the one warp of the group that polls the accumulator barriers
    const uint32_t dsh = p.n_dbuf - 1, psh = p.n_pbuf > 0 ? p.n_pbuf - 1 : 0;   // n_dbuf, n_pbuf in {1, 2}
    const bool is_t0 = (warp == 4 && lane == 0);
    const uint32_t thr = dropout_threshold(p.drop_p);
    const float keep_scale = 1.f / (1.f - p.drop_p);
    const float out_scale = p.out_scale != 0.f ? p.out_scale : 1.f;
    const bool scale_rows = p.rowscale_out != nullptr || out_scale != 1.f;
    constexpr bool dual = EP == LIN_EP_GELU_DUAL || EP == LIN_EP_GELU_DUAL_GRAD;

    uint32_t lw = 0, Cn = 0, G = 0;
    Tracer<TRACE> tr;
    tr.init((lane == 0 && q4 == 0 && sel == 0) ? p.trace : nullptr, 2 + static_cast<int>(grp));
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++lw) {
      const WorkItem it = get_work(p, w);
      tr.ev(1000000ull + w);
      const int grow = it.m0 + row;
      const bool row_ok = grow < p.M;
      const int sample = (p.rows_per_sample > 0) ? min(grow / p.rows_per_sample, p.n_samples - 1) : 0;
      const int row0 = it.m0 + q4 * 32;

      if (p.R_pad > 0 && !p.u_in) {
        if (warp == 4) {   // one warp polls, the other epilogue warps block on the named barrier (no issue slots)
          mbar_wait(u_full, lw & 1u, p.wait_hint_ns);
          if (lw > 0) mbar_wait(usm_free, (lw - 1) & 1u, p.wait_hint_ns);  // previous item's delta MMAs finished reading usm
          if (p.u_save != nullptr && lane == 0) bulk_wait_read<0>();       // ... and so did its u_save bulk stores
        }
        epi_bar_sync(1);
        tc_fence_after();
        for (int gq = ew >> 2; gq < p.R_pad / 16; gq += kEpiWarps / 4) {
          uint32_t r[16];
          tmem_ld16(t_lane + gq * 16, r);
          tmem_ld_wait();
          float s = p.gran_scale[gq];
          if (p.rowscale_in != nullptr) s *= p.rowscale_in[p.gran_in[gq] * p.n_samples + sample];
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]) * s, __uint_as_float(r[2 * i + 1]) * s);
          uint8_t* atom = smem_gen + L.usm + (gq >> 2) * kTileABytes;
          const uint32_t c0 = (gq & 3) * 16;
          *reinterpret_cast<uint4*>(atom + sw128_offset(row, c0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(atom + sw128_offset(row, c0 + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        tr.ev(8000000ull);   // U converted
        fence_proxy_async_smem();
        tc_fence_before();
        if (p.u_save != nullptr && it.split == 0) {
          epi_bar_sync(2);
          if (is_t0) {
            for (int a = 0; a < n_uatoms; ++a) tma_store_2d(&tm_u, usm_base + a * kTileABytes, a * 64, it.m0);
            bulk_commit();
          }
        }
        mbar_arrive(u_ready);
      }

      for (int ci = 0; ci < it.n_my_chunks; ++ci) {
        const int c = it.split + ci * p.n_splits;
        int n_eff = p.Nn - c * p.BN;
        if (n_eff > p.BN) n_eff = p.BN;
        const int n_items = multi ? p.S_out : 1;
        const uint32_t pb = multi ? (Cn & psh) : 0;
        bool p_waited = false;
        for (int j = 0; j < n_items; ++j, ++G) {
          if ((G % kGroups) != grp) continue;
          // d_shared: full barrier of this group (grp * 2), its kk-th use; the empty barrier is the shared d_empty(0)
          const uint32_t kk = G / kGroups, dbuf = kk & dsh, db = p.d_shared ? grp * 2 : grp * 2 + dbuf;
          const uint32_t d_parity = p.d_shared ? (kk & 1u) : ((kk >> dsh) & 1u);
          // Epilogue inputs that do not depend on the accumulators (the GELU' argument of the fc2 backward, the residual
          // of proj / fc2 forward) are staged by TMA into the very slab the half's output will be written to, one
          // 64-column half ahead (across item and tile boundaries), so their HBM latency never stalls the math.
          constexpr bool need_aux = EP == LIN_EP_GELU_BWD || EP == LIN_EP_MUL_AUX, need_res = HAS_RES;
          constexpr bool has_in = need_aux || need_res;
          int nx_w = w, nx_ci = ci, nx_j = j;      // lookahead: this group's next item
          bool nx_valid = true;
          if (has_in) {
            uint32_t g2 = G;
            int nchunks = it.n_my_chunks;
            do {
              ++nx_j; ++g2;
              if (nx_j == n_items) {
                nx_j = 0;
                if (++nx_ci == nchunks) {
                  nx_ci = 0;
                  nx_w += gridDim.x;
                  if (nx_w >= n_work) { nx_valid = false; break; }
                  nchunks = get_work(p, nx_w).n_my_chunks;
                }
              }
            } while ((g2 % kGroups) != grp);
          }
          auto issue_in = [&](int slab, int col, int r0, int strm) {   // lane 0 only; col = first column of the half
            int pc = col + sel * kPieceCols;
            if (pc >= p.Nn) pc = 0;   // this warp's piece lies beyond N (ragged last chunk): load anything, it is ignored
            mbar_arrive_expect_tx(in_bar(ew, slab), kSlabBytes);
            tma_load_3d(slab_base + slab * kSlabBytes, &tm_in, in_bar(ew, slab), pc, r0, need_res && p.res_streams == 1 ? 0 : strm);
          };
          if (has_in && first_half && lane == 0) issue_in(0, c * p.BN, row0, j);   // very first half of this warp
          tr.ev(5000000ull + ci * 10 + j);   // waiting for accumulator
          const bool use_p = multi && p.out_useP[j];
          if (grp_leader) {
            mbar_wait(d_full(db), d_parity, p.wait_hint_ns);
            if (use_p && !p_waited) mbar_wait(p_full(pb), (Cn >> psh) & 1u, p.wait_hint_ns);
          }
          if (use_p) p_waited = true;
          group_bar_sync(3 + static_cast<int>(grp));
          tr.ev(6000000ull + ci * 10 + j);   // got it
          tc_fence_after();
          const uint32_t acc_d = t_lane + d_col0 + (p.d_shared ? 0u : (grp * p.n_dbuf + dbuf) * p.BN);
          const uint32_t acc_p = t_lane + p_col0 + pb * p.BN;
          const bool mask_delta = multi && (p.drop_mode == 2) && j == 0;
          const float rs = ((p.rowscale_out != nullptr) ? p.rowscale_out[j * p.n_samples + sample] : 1.f) * out_scale;
          const int n_half = (n_eff + 63) >> 6;
          for (int h = 0; h < n_half; ++h) {
            const int col_h = c * p.BN + h * 64;
            const int col_p = col_h + sel * kPieceCols;   // first column of this warp's piece
            // slabs of this half: y -> ks_y, GELU(y) -> ks_y2. A slab is reused every n_slabs stores; with at most
            // n_slabs - n_out bulk groups still pending the ones that used these slabs have been read.
            constexpr int n_out = dual ? 2 : 1;
            const int ks_y = slab_k;
            const int ks_y2 = nxt_slab(slab_k);
            if (has_in) {
              // next half of this warp: same item, or the first half of the group's next item
              const int ks_n = ks_y2;
              if (lane == 0) {
                if (p.n_slabs >= 3) bulk_wait_read<1>(); else bulk_wait_read<0>();   // slab ks_n's last store was read
                if (h + 1 < n_half) {
                  issue_in(ks_n, col_h + 64, row0, j);
                } else if (nx_valid) {
                  const WorkItem ni = get_work(p, nx_w);
                  issue_in(ks_n, (ni.split + nx_ci * p.n_splits) * p.BN, ni.m0 + q4 * 32, nx_j);
                }
              }
              mbar_wait(in_bar(ew, ks_y), (in_ph >> ks_y) & 1u, p.wait_hint_ns);
              in_ph ^= 1u << ks_y;
            } else {
              if (lane == 0) {
                if (p.n_slabs - n_out >= 1) bulk_wait_read<1>(); else bulk_wait_read<0>();
              }
            }
            __syncwarp();
            tr.ev(9000000ull + h);   // slabs free / epilogue inputs landed
            uint8_t* sy = slab_gen + ks_y * kSlabBytes;
            uint8_t* sy2 = slab_gen + ks_y2 * kSlabBytes;
#pragma unroll 1
            for (int g2 = 0; g2 < kPieceGran; ++g2) {
              const int gq = sel * kPieceGran + g2;   // granule inside the half
              const int n0 = col_h + gq * 16;
              if (n0 >= p.Nn) break;
              const int tcol = h * 64 + gq * 16;
              // every load of the granule is issued before the TMEM wait so that the latencies overlap
              uint32_t pr[16], dr[16];
              if (use_p) tmem_ld16(acc_p + tcol, pr);
              tmem_ld16(acc_d + tcol, dr);
              float4 bv[4];
              if (p.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
              }
              uint32_t aw[8];
              if (has_in) {
                const uint4 a0 = *reinterpret_cast<const uint4*>(sy + sw64_offset(lane, g2 * 16));
                const uint4 a1 = *reinterpret_cast<const uint4*>(sy + sw64_offset(lane, g2 * 16 + 8));
                aw[0] = a0.x; aw[1] = a0.y; aw[2] = a0.z; aw[3] = a0.w;
                aw[4] = a1.x; aw[5] = a1.y; aw[6] = a1.z; aw[7] = a1.w;
              }
              tmem_ld_wait();
              if (mask_delta) {
                const uint64_t e0 = (static_cast<uint64_t>(grow) * p.Nn + n0) >> 1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const uint32_t hsh = dropout_pair_bits(p.drop_seed, e0 + i);
                  dr[2 * i] = (hsh << 16) >= thr ? __float_as_uint(__uint_as_float(dr[2 * i]) * keep_scale) : 0u;
                  dr[2 * i + 1] =
                      (hsh & 0xffff0000u) >= thr ? __float_as_uint(__uint_as_float(dr[2 * i + 1]) * keep_scale) : 0u;
                }
              }
              // packed fp32x2 arithmetic: two columns per instruction
              uint64_t v2[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v2[i] = pack2(__uint_as_float(dr[2 * i]), __uint_as_float(dr[2 * i + 1]));
                if (use_p) v2[i] = add2(v2[i], pack2(__uint_as_float(pr[2 * i]), __uint_as_float(pr[2 * i + 1])));
              }
              if (p.bias != nullptr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v2[2 * i] = add2(v2[2 * i], pack2(bv[i].x, bv[i].y));
                  v2[2 * i + 1] = add2(v2[2 * i + 1], pack2(bv[i].z, bv[i].w));
                }
              }
              if (scale_rows) {
                const uint64_t rs2 = pack2(rs, rs);
#pragma unroll
                for (int i = 0; i < 8; ++i) v2[i] = mul2(v2[i], rs2);
              }
              if (has_in && !need_aux) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v2[i] = add2(v2[i], pack2(bf16lo_to_f32(aw[i]), bf16hi_to_f32(aw[i])));
              }
              float v[16];
#pragma unroll
              for (int i = 0; i < 8; ++i) unpack2(v2[i], v[2 * i], v[2 * i + 1]);
              if (EP == LIN_EP_GELU_BWD) gelu_grad16_mul(v, aw);
              if (EP == LIN_EP_MUL_AUX) {   // aux already holds GELU'(pre-activation)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  v[2 * i] *= bf16lo_to_f32(aw[i]);
                  v[2 * i + 1] *= bf16hi_to_f32(aw[i]);
                }
              }
              uint32_t pk[8];
              [[maybe_unused]] uint32_t pk_act[8];
              if (EP == LIN_EP_GELU_DUAL_GRAD) {
                gelu_and_grad16_pack(v, pk_act, pk);   // y <- GELU'(v), y2 <- GELU(v)
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
              }
              *reinterpret_cast<uint4*>(sy + sw64_offset(lane, g2 * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(sy + sw64_offset(lane, g2 * 16 + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              if (dual) {
                if (EP == LIN_EP_GELU_DUAL_GRAD) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) pk[i] = pk_act[i];
                } else {
                  gelu16_pack(v, pk);
                }
                *reinterpret_cast<uint4*>(sy2 + sw64_offset(lane, g2 * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(sy2 + sw64_offset(lane, g2 * 16 + 8)) =
                    make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
            fence_proxy_async_smem();
            __syncwarp();
            tr.ev(9500000ull + h);   // half computed
            const bool piece_ok = row0 < p.M && col_p < p.Nn;
            if (lane == 0) {   // (an empty bulk group keeps the slab-reuse accounting uniform for skipped pieces)
              if (piece_ok) tma_store_3d(&tm_y, slab_base + ks_y * kSlabBytes, col_p, row0, j);
              bulk_commit();
              if (dual) {
                if (piece_ok) tma_store_3d(&tm_y2, slab_base + ks_y2 * kSlabBytes, col_p, row0, j);
                bulk_commit();
              }
            }
            slab_k = dual ? nxt_slab(ks_y2) : ks_y2;
            first_half = false;
            if (dual && p.drop_mode == 1 && j == 0) {
              // D(m) of the shared stream (LoRA dropout of the consuming fc2, drawn with seed + 1): derived from the
              // bf16-rounded activation in the y2 slab so that it equals dropout(y2, seed + 1) exactly
              const int ks_d = slab_k;
              uint8_t* sd = slab_gen + ks_d * kSlabBytes;
              if (lane == 0) {   // all but the (n_slabs - 1) most recent stores have been read
                if (p.n_slabs >= 3) bulk_wait_read<2>(); else bulk_wait_read<1>();
              }
              __syncwarp();
#pragma unroll 1
              for (int g2 = 0; g2 < kPieceGran; ++g2) {
                const int n0 = col_p + g2 * 16;
                if (n0 >= p.Nn) break;
                const uint64_t e0 = static_cast<uint64_t>(grow) * p.Nn + n0;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const uint4 mv = *reinterpret_cast<const uint4*>(sy2 + sw64_offset(lane, g2 * 16 + hh * 8));
                  const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
                  uint32_t ow[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    ow[i] = dropout_apply_pair(mw[i], p.drop_seed + 1, e0 + hh * 8 + 2 * i, thr, keep_scale);
                  *reinterpret_cast<uint4*>(sd + sw64_offset(lane, g2 * 16 + hh * 8)) =
                      make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (piece_ok) tma_store_3d(&tm_y2, slab_base + ks_d * kSlabBytes, col_p, row0, p.S_out);
                bulk_commit();
              }
              slab_k = nxt_slab(slab_k);
            }
          }
          tc_fence_before();
          mbar_arrive(d_empty(p.d_shared ? 0u : db));
          tr.ev(7000000ull + ci * 10 + j);   // item done
        }
        if (multi) {
          tc_fence_before();
          mbar_arrive(p_empty(pb));
          ++Cn;
        }
      }
    }
    if (lane == 0) bulk_wait_all();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
namespace {
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

}  // namespace

namespace {
// Descriptor cache: a training step re-launches the same ~100 layer shapes on buffers the caching
// allocator hands back at the same addresses, so the 128-byte tensor maps are looked up by (base, dims, box, swizzle,
// pitch) instead of being re-encoded by the driver nine times per launch. Small direct-mapped table behind a mutex; a
// descriptor only describes addresses and strides, so a stale entry is impossible — equal keys give equal maps.
struct TmapKey {
  const void* base;
  uint64_t d0, d1, d2, pitch;
  uint32_t b0, b1, swz;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && pitch == o.pitch && b0 == o.b0 && b1 == o.b1 &&
           swz == o.swz;
  }
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int kTmapSlots = 4096;
TmapSlot* tmap_table() {
  static TmapSlot* t = new TmapSlot[kTmapSlots]();
  return t;
}
std::mutex g_tmap_mu;
inline size_t tmap_hash(const TmapKey& k) {
  uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
  h ^= (k.d0 * 0xBF58476D1CE4E5B9ull) ^ (k.d1 * 0x94D049BB133111EBull) ^ (k.d2 << 7) ^ (k.pitch << 29) ^
       (static_cast<uint64_t>(k.b0) << 40) ^ (static_cast<uint64_t>(k.b1) << 52) ^ (static_cast<uint64_t>(k.swz) << 60);
  h ^= h >> 31;
  return static_cast<size_t>(h) % kTmapSlots;
}
}  // namespace

// bf16 row-major [d2][d1][d0] tensor (d0 contiguous), box = (b0, b1, 1), 128-byte swizzle unless stated.
// pitch = elements between consecutive rows (0: d0, i.e. densely packed).
int make_tmap(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
              uint32_t b1, CUtensorMapSwizzle swz, uint64_t pitch) {
  if (pitch == 0) pitch = d0;
  const TmapKey key{base, d0, d1, d2, pitch, b0, b1, static_cast<uint32_t>(swz)};
  const size_t slot = tmap_hash(key);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    const TmapSlot& s = tmap_table()[slot];
    if (s.valid && s.key == key) {
      *tm = s.map;
      return 0;
    }
  }
  auto fn = get_encode_fn();
  MTL_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  MTL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer not 16-byte aligned");
  MTL_REQUIRE((pitch * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (pitch %llu elements)",
              (unsigned long long)pitch);
  const int rank = d2 > 0 ? 3 : 2;
  cuuint64_t dims[3] = {d0, d1, d2 > 0 ? d2 : 1};
  cuuint64_t strides[2] = {pitch * 2, pitch * d1 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu box %u,%u)",
              (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    TmapSlot& s = tmap_table()[slot];
    s.key = key;
    s.map = *tm;
    s.valid = true;
  }
  return 0;
}

namespace {
int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Tuning / debugging switches of the planner, read from the environment once per process.
struct PlanEnv {
  bool no_dshared;
  int force_bn, force_splits;
  PlanEnv() {
    no_dshared = getenv("MTL_LINEAR_NO_DSHARED") != nullptr;
    const char* e = getenv("MTL_LINEAR_BN");
    force_bn = e ? atoi(e) : 0;
    e = getenv("MTL_LINEAR_SPLITS");
    force_splits = e ? atoi(e) : 0;
  }
};
const PlanEnv& plan_env() {
  static const PlanEnv env;
  return env;
}
}  // namespace

int plan_linear(LinPlan& p, int n_sm, uint32_t* smem_bytes_out) {
  MTL_REQUIRE(p.M > 0 && p.Kc > 0 && p.Nn > 0, "linear: empty problem (M=%d K=%d N=%d)", p.M, p.Kc, p.Nn);
  MTL_REQUIRE(p.Kc % 16 == 0 && p.Nn % 16 == 0, "linear: K (%d) and N (%d) must be multiples of 16", p.Kc, p.Nn);
  MTL_REQUIRE(p.R_pad % 16 == 0 && p.R_pad <= 16 * LIN_MAX_GRAN, "linear: rank space %d unsupported (<= %d)",
              p.R_pad, 16 * LIN_MAX_GRAN);
  MTL_REQUIRE(p.S_in >= 1 && p.S_in <= LIN_MAX_STREAMS && p.S_out >= 1 && p.S_out <= LIN_MAX_STREAMS,
              "linear: stream counts out of range");
  MTL_REQUIRE(p.n_main > 0, "linear: nothing to compute");

  // ---- TMEM plan -------------------------------------------------------------------------------
  // merged : one accumulator per item (dense + adapters), double-buffered between the two epilogue groups
  // multi  : dense accumulator P per chunk (n_pbuf buffers) + per-stream delta accumulators D[2]
  if (p.u_in) {
    MTL_REQUIRE(p.R_pad > 0 && p.R_pad <= 128 && p.u_save != nullptr, "linear: u_in needs a rank space of 16..128 columns and U");
    MTL_REQUIRE(p.S_out == 1, "linear: u_in serves single-output-stream layers");
  }
  const int u_cols = p.u_in ? 0 : round_up(p.R_pad, 32);   // u_in: no rank-space accumulators in TMEM
  const int r_smem = p.u_in ? 0 : p.R_pad;                 // ... and no U operand region in shared memory
  const bool multi = p.R_pad > 0 && (p.S_out > 1 || p.force_split || p.drop_mode == 2);
  p.n_regions = multi ? 1 + p.S_out : 1;
  // Column-chunk width BN and accumulator buffering. Short contractions (K_eff < 256: stages 0-1) are epilogue / HBM
  // bound: narrow chunks, two accumulators per epilogue group so the MMA warp runs ahead. Long contractions are
  // bound by operand traffic (every chunk re-streams the X tile from L2): wide chunks, one accumulator per group.
  const bool heavy = p.Kc >= 256;
  // TMEM columns = u_cols + n_pbuf * BN (multi) + kGroups * n_dbuf * BN <= 512.
  auto cols = [&](int bn_, int pb, int db) { return u_cols + (pb + kGroups * db) * bn_; };
  const int n_uatoms = (p.R_pad + 63) / 64;
  MTL_REQUIRE(n_uatoms <= kMaxUAtoms, "linear: too many rank atoms");
  int min_stages = n_uatoms > 0 ? n_uatoms + 2 : 2;   // the Up tiles of a chunk are resident together, + 2 to stream
  // epilogues with two outputs per half (GELU pair) or with TMA-staged inputs want a third store slab per warp:
  // with two, every half waits for the bulk store that just left (measured: 12.7 us instead of ~5 us per item)
  const bool ep_dual = p.ep_mode == LIN_EP_GELU_DUAL || p.ep_mode == LIN_EP_GELU_DUAL_GRAD;
  const bool ep_aux = p.ep_mode == LIN_EP_GELU_BWD || p.ep_mode == LIN_EP_MUL_AUX;
  const int want_slabs = (ep_dual || ep_aux || p.res != nullptr) ? 3 : 2;
  auto fits_smem = [&](int bn_, int slabs) {   // with the minimum ring depth
    return smem_layout(min_stages, kTileABytes + bn_ * 128, r_smem, slabs).total + 1024 <= 227u * 1024;
  };
  p.n_pbuf = 0;
  p.n_dbuf = 1;
  int bn = 64;
  if (!multi) {
    // one accumulator per item: wide chunks amortise the per-chunk cost of the single MMA-issuing thread
    // (~100 ns per tcgen05 op, measured) and the X-tile re-reads; the groups take the items round-robin
    if (p.Nn > 64 && cols(128, 0, 1) <= 512 && fits_smem(128, want_slabs)) bn = 128;
    if (p.Nn > 128 && cols(192, 0, 1) <= 512 && fits_smem(192, want_slabs)) bn = 192;
    if (cols(bn, 0, 2) <= 512) p.n_dbuf = 2;
  } else {
    // dense accumulator P per chunk + delta accumulators per group
    p.n_pbuf = 2;
    // (measured: for long contractions 128-column chunks with two slabs beat 64-column chunks with three)
    // ... and for short contractions unless the epilogue writes two outputs per half (GELU pair: it needs the third
    // slab more than the wider items): proj forward 0.38 -> 0.28 ms, fc2 backward 1.20 -> 1.06 ms at stage 0
    if ((heavy || !ep_dual) && p.S_out > 1 && p.Nn > 64 && cols(128, 1, 1) <= 512 && fits_smem(128, 2)) {
      bn = 128;
      p.n_pbuf = 1;
    }
    // (measured: the GELU pair is slower with 128-column items even with three slabs — its long items need the
    // double-buffered accumulators more than the wider chunks)
    if (cols(bn, p.n_pbuf, 1) > 512) p.n_pbuf = 1;
    if (cols(bn, p.n_pbuf, 2) <= 512) p.n_dbuf = 2;
    // single output stream with separate dense / adapter accumulators (input gradient with LoRA dropout) and a long
    // contraction: 128-column chunks halve the re-streaming of the X tile from L2. The dense accumulator stays
    // double-buffered; ONE delta accumulator serves both epilogue groups (its MMAs are tiny, so waiting for the
    // previous item's epilogue before issuing them costs next to nothing).
    // (measured at K = 96 too: 128-column items beat 64-column ones by 14-23 % — the per-item cost of the epilogue
    // protocol outweighs the lost overlap of the two groups on the tiny delta products)
    if (p.S_out == 1 && p.Nn > 64 && u_cols + 3 * 128 <= 512 && fits_smem(128, want_slabs) &&
        !plan_env().no_dshared) {
      bn = 128;
      p.n_pbuf = 2;
      p.n_dbuf = 1;
      p.d_shared = 1;
    }
  }
  if (plan_env().force_bn > 0) {   // tuning aid (MTL_LINEAR_BN): force the chunk width (merged mode only)
    const int fb = plan_env().force_bn;
    if (!multi && (fb == 64 || fb == 128 || fb == 192) && cols(fb, 0, 1) <= 512) {
      bn = fb;
      p.n_dbuf = cols(fb, 0, 2) <= 512 ? 2 : 1;
    }
  }
  p.BN = bn;
  p.n_chunks = (p.Nn + bn - 1) / bn;
  p.acc_col0 = u_cols;
  const int need_cols = p.d_shared ? u_cols + (p.n_pbuf + 1) * bn : cols(bn, p.n_pbuf, p.n_dbuf);
  MTL_REQUIRE(need_cols <= 512, "linear: TMEM budget exceeded (R_pad=%d, S_out=%d)", p.R_pad, p.S_out);
  p.tmem_cols = 32;
  while (p.tmem_cols < need_cols) p.tmem_cols *= 2;

  // ---- phase-1 groups must fit the B half of a ring stage (BN rows) ----------------------------------------
  {
    const LinPlan q = p;
    int n = 0;
    for (int g = 0; g < q.n_groups; ++g) {
      int r0 = q.grp_r0[g], len = q.grp_len[g];
      MTL_REQUIRE(len % 16 == 0 && len >= 16, "linear: down group length %d invalid", len);
      while (len > 0) {
        MTL_REQUIRE(n < LIN_MAX_GROUPS, "linear: too many adapter groups");
        const int l = len > bn ? bn : len;
        p.grp_in[n] = q.grp_in[g];
        p.grp_r0[n] = r0;
        p.grp_len[n] = l;
        p.grp_acc[n] = q.grp_acc[g];
        ++n;
        r0 += l;
        len -= l;
      }
    }
    p.n_groups = n;
  }
  p.stage_b_bytes = bn * 128;
  const int stage_bytes = kTileABytes + p.stage_b_bytes;
  // Wide rank spaces (R_pad > 128: the U operand alone takes 48-80 KiB): one Up tile per stage no longer fits next
  // to the store slabs. Pack the tiles — the A half of those stages is free — and, if still short, stream with one
  // spare stage instead of two.
  p.up_pack = 1;
  if (n_uatoms > 0 && smem_layout(min_stages, stage_bytes, r_smem, 2).total + 1024 > 227u * 1024) {
    p.up_pack = 1 + kTileABytes / (bn * 128);
    const int up_stages = (n_uatoms + p.up_pack - 1) / p.up_pack;
    min_stages = up_stages + 2;
    if (smem_layout(min_stages, stage_bytes, r_smem, 2).total + 1024 > 227u * 1024) min_stages = up_stages + 1;
  }

  // ---- work decomposition: (128-row tile, column split), persistent CTAs ---------------------------------------
  const int m_tiles = (p.M + LIN_BM - 1) / LIN_BM;
  // Column splits: with few row tiles, split the chunks of a tile over several work items so that the persistent
  // grid is balanced (every split repeats the rank-space phase, modelled as one extra 64-column chunk).
  p.n_splits = 1;
  if (m_tiles < 4 * n_sm && p.n_chunks > 1) {
    double best = -1.0;
    for (int sp = 1; sp <= p.n_chunks; ++sp) {
      const long work = static_cast<long>(m_tiles) * sp;
      const long waves = (work + n_sm - 1) / n_sm;
      const int chunks_max = (p.n_chunks + sp - 1) / sp;                       // chunks of the largest split
      const double item_cost = chunks_max * static_cast<double>(bn) + (p.R_pad > 0 ? 64.0 : 0.0) + 16.0;
      const double t = waves * item_cost;                                      // ~ time of the slowest CTA
      const double score = 1.0 / t;
      if (score > best * 1.0001) {
        best = score;
        p.n_splits = sp;
      }
    }
  }
  if (plan_env().force_splits > 0) {   // tuning aid (MTL_LINEAR_SPLITS)
    const int fs = plan_env().force_splits;
    if (fs >= 1 && fs <= p.n_chunks) p.n_splits = fs;
  }
  p.n_work = m_tiles * p.n_splits;

  // ---- shared memory: store slabs + U operand + as many ring stages as fit ---------------------------------------
  const bool has_in = ep_aux || p.res != nullptr;   // epilogue inputs staged through the slabs
  p.n_slabs = (ep_dual || has_in) ? 3 : 2;   // reduced to 2 below when smem is short
  if (smem_layout(min_stages, stage_bytes, r_smem, p.n_slabs).total + 1024 > 227u * 1024) p.n_slabs = 2;
  p.n_stages = kMaxStages;
  while (p.n_stages > min_stages &&
         smem_layout(p.n_stages, stage_bytes, r_smem, p.n_slabs).total + 1024 > 227u * 1024)
    --p.n_stages;
  const uint32_t smem_bytes = smem_layout(p.n_stages, stage_bytes, r_smem, p.n_slabs).total + 1024;
  MTL_REQUIRE(!p.u_in || p.up_pack == 1, "linear: u_in with packed Up tiles");
  MTL_REQUIRE(smem_bytes <= 227 * 1024, "linear: shared memory %u exceeds 227 KiB", smem_bytes);
  MTL_REQUIRE(p.ep_mode >= 0 && p.ep_mode <= 4, "linear: unknown epilogue mode %d", p.ep_mode);
  MTL_REQUIRE(!((ep_aux || ep_dual) && p.res != nullptr), "linear: GELU epilogues cannot take a residual");
  *smem_bytes_out = smem_bytes;
  return 0;
}

int launch_linear(LinPlan p, const void* x, const void* wm, const void* down, const void* up,
                  cudaStream_t stream) {
  int dev = 0, n_sm = 148;
  MTL_CHECK_CUDA(cudaGetDevice(&dev));
  MTL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  uint32_t smem_bytes = 0;
  if (int e = plan_linear(p, n_sm, &smem_bytes)) return e;
  const bool ep_dual = p.ep_mode == LIN_EP_GELU_DUAL || p.ep_mode == LIN_EP_GELU_DUAL_GRAD;
  const bool ep_aux = p.ep_mode == LIN_EP_GELU_BWD || p.ep_mode == LIN_EP_MUL_AUX;

  // ---- tensor maps ---------------------------------------------------------------------------
  CUtensorMap tm_x, tm_w, tm_down, tm_up, tm_y, tm_y2, tm_u, tm_in, tm_pf;
  if (int e = make_tmap(&tm_x, x, p.Kc, p.M, p.S_in, LIN_BK, LIN_BM)) return e;
  if (int e = make_tmap(&tm_w, wm, p.Kc, p.Nn, 0, LIN_BK, p.BN)) return e;
  if (p.R_pad > 0) {
    if (int e = make_tmap(&tm_down, down, p.Kc, p.R_pad, 0, LIN_BK, 16)) return e;
    if (int e = make_tmap(&tm_up, up, p.R_pad, p.Nn, 0, LIN_BK, p.BN)) return e;
  } else {
    tm_down = tm_w;
    tm_up = tm_w;
  }
  // epilogue slabs: [32 rows x 32 columns] pieces, 64-byte swizzle
  constexpr CUtensorMapSwizzle kSw64 = CU_TENSOR_MAP_SWIZZLE_64B;
  if (int e = make_tmap(&tm_y, p.y, p.Nn, p.M, p.S_out, kPieceCols, 32, kSw64)) return e;
  if (ep_dual) {
    MTL_REQUIRE(p.y2 != nullptr, "linear: GELU epilogue needs y2");
    if (int e = make_tmap(&tm_y2, p.y2, p.Nn, p.M, p.S_out + (p.drop_mode == 1 ? 1 : 0), kPieceCols, 32, kSw64)) return e;
  } else {
    tm_y2 = tm_y;
  }
  if (p.u_save != nullptr && p.R_pad > 0) {
    if (int e = make_tmap(&tm_u, p.u_save, p.R_pad, p.M, 0, 64, LIN_BM)) return e;
  } else {
    tm_u = tm_y;
  }
  // tm_pf: the same tensor with [128 rows x 64 columns] boxes, used by the producer to prefetch the epilogue inputs
  // of a work item into L2 while its accumulators are still being formed
  if (ep_aux) {
    MTL_REQUIRE(p.aux != nullptr, "linear: GELU' epilogue needs aux");
    if (int e = make_tmap(&tm_in, p.aux, p.Nn, p.M, p.S_out, kPieceCols, 32, kSw64)) return e;
    if (int e = make_tmap(&tm_pf, p.aux, p.Nn, p.M, p.S_out, 64, LIN_BM, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    p.in_streams = p.S_out;
  } else if (p.res != nullptr) {
    if (int e = make_tmap(&tm_in, p.res, p.Nn, p.M, p.res_streams, kPieceCols, 32, kSw64)) return e;
    if (int e = make_tmap(&tm_pf, p.res, p.Nn, p.M, p.res_streams, 64, LIN_BM, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    p.in_streams = p.res_streams;
  } else {
    tm_in = tm_y;
    tm_pf = tm_y;
    p.in_streams = 0;
  }

  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap,
                            CUtensorMap, CUtensorMap, LinPlan);
  // [epilogue mode][trace]; the residual variant exists for LIN_EP_NONE only (index 5)
  static KernelFn kernels[6][2] = {
      {mtl_linear_kernel<LIN_EP_NONE, false, false>, mtl_linear_kernel<LIN_EP_NONE, false, true>},
      {mtl_linear_kernel<LIN_EP_GELU_DUAL, false, false>, mtl_linear_kernel<LIN_EP_GELU_DUAL, false, true>},
      {mtl_linear_kernel<LIN_EP_GELU_BWD, false, false>, mtl_linear_kernel<LIN_EP_GELU_BWD, false, true>},
      {mtl_linear_kernel<LIN_EP_GELU_DUAL_GRAD, false, false>, mtl_linear_kernel<LIN_EP_GELU_DUAL_GRAD, false, true>},
      {mtl_linear_kernel<LIN_EP_MUL_AUX, false, false>, mtl_linear_kernel<LIN_EP_MUL_AUX, false, true>},
      {mtl_linear_kernel<LIN_EP_NONE, true, false>, mtl_linear_kernel<LIN_EP_NONE, true, true>}};
  static std::once_flag attr_once;
  static int max_ctas = -1;
  static uint32_t wait_hint = 1000u;
  std::call_once(attr_once, []() {
    const char* e = getenv("MTL_LINEAR_MAX_CTAS");  // debugging aid: 0 = one CTA per work item (non-persistent)
    max_ctas = e ? atoi(e) : -1;
    const char* h = getenv("MTL_WAIT_HINT_NS");      // mbarrier.try_wait suspend-time hint (tuning aid)
    if (h) wait_hint = static_cast<uint32_t>(atoi(h));
  });
  p.wait_hint_ns = wait_hint;
  static const char* trace_path = getenv("MTL_LINEAR_TRACE");
  static bool attr_done[12][64] = {};   // per kernel variant and device
  {
    const int ki = (p.res != nullptr ? 5 : p.ep_mode) * 2 + (trace_path != nullptr ? 1 : 0);
    MTL_CHECK_CUDA(ensure_max_dyn_smem(attr_done[ki], kernels[ki / 2][ki % 2], 227 * 1024));
  }
  p.trace = nullptr;
  if (trace_path != nullptr) {
    MTL_CHECK_CUDA(cudaMalloc(&p.trace, 4 * 2048 * sizeof(unsigned long long)));
    MTL_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, 4 * 2048 * sizeof(unsigned long long), stream));
  }

  int grid = p.n_work < n_sm ? p.n_work : n_sm;
  if (max_ctas == 0) grid = p.n_work;
  else if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  kernels[p.res != nullptr ? 5 : p.ep_mode][p.trace != nullptr ? 1 : 0]<<<grid, kThreads, smem_bytes, stream>>>(
      tm_x, tm_w, tm_down, tm_up, tm_y, tm_y2, tm_u, tm_in, tm_pf, p);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  if (p.trace != nullptr) {   // debugging only: synchronous dump of CTA 0's timeline (last launch wins)
    static unsigned long long host[4 * 2048];
    MTL_CHECK_CUDA(cudaStreamSynchronize(stream));
    MTL_CHECK_CUDA(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      fprintf(f, "# M=%d K=%d N=%d BN=%d chunks=%d splits=%d stages=%d slabs=%d pbuf=%d dbuf=%d multi=%d\n", p.M, p.Kc, p.Nn,
              p.BN, p.n_chunks, p.n_splits, p.n_stages, p.n_slabs, p.n_pbuf, p.n_dbuf, p.n_regions > 1);
      for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 1024 && host[r * 2048 + 2 * i] != 0; ++i)
          fprintf(f, "%d %llu %llu\n", r, host[r * 2048 + 2 * i], host[r * 2048 + 2 * i + 1]);
      fclose(f);
    }
  }
  return 0;
}

}  // namespace mtl
