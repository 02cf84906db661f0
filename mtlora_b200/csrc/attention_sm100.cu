// Shifted-window attention forward on the 5th-generation tensor cores (tcgen05 + TMEM) for sm_100a.
//
// Reference semantics (models/swin_transformer_mtlora.py): torch.roll + window_partition (:338-342, or
// kernels/window_process), q*scale, q@k^T + relative-position bias + SW-MSA mask, softmax, attn@v (:194-220),
// window_reverse + roll back (:365-377). Same contract as win_attn_fwd_kernel (attention.cu), which stays as the path for
// explicit masks: activations stay in (B, H, W, C) token order, the shift / partition / reverse are index math of the
// gather and the scatter, the saved log-sum-exp feeds the unchanged backward kernel.
//
// Work unit = TWO windows (2 x 49 tokens in two 64-row slots of one 128-row UMMA tile) x a group of heads:
//   * C = 96 (stage 0 of Swin-T/S): all three heads — the unit gathers WHOLE 576-byte qkv rows (the per-head kernel
//     fetched 64-byte slices and read 1.8x the algorithmic bytes);
//   * wider stages: a pair of heads (128-byte, 128-byte-aligned q / k / v segments).
// Shared-memory operand layout: the gathered columns are laid out as [128 rows x 64 columns] bf16 atoms with the
// 128-byte swizzle — exactly the K-major UMMA operand atom for Q and K (a head is a 32-column half of an atom, selected by
// the descriptor's start address) and the MN-major operand atom for V (tokens along K, both heads of the atom along N).
//   S_h = Q_h K_h^T        tcgen05.mma M=128 N=128 K=32      (both windows at once; the two off-diagonal 64x64 blocks are
//                                                             never read)
//   softmax                one thread per query row: tcgen05.ld of its window's 64 columns, bias + analytic shift mask,
//                          base-2 softmax in registers, P (bf16) -> K-major swizzled smem tile (zero off-diagonal blocks)
//   O_h = P_h V            tcgen05.mma M=128 N=64 K=128      (N covers the head's atom; its 32-column half is read back)
// Roles (256 threads): warps 0-3 softmax / epilogue (TMEM lane quadrants 0-3), warp 4 lane 0 MMA issuer, warps 5-7
// gather producers (cp.async into the swizzled atoms, double-buffered across units). TMEM: S double-buffered (2 x 128
// columns) + one 64-column O accumulator per head of the unit.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {
namespace {

constexpr int kAtomBytes = 128 * 128;     // [128 rows x 64 bf16], SW128
constexpr int kThreads = 256;
constexpr int kProducers = 96;            // warps 5-7
constexpr int kMaxUnitHeads = 4;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major SW128 operand: 8-row groups 1024 B apart (tokens along K), see xty_sm100.cu
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((kAtomBytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

struct UAttnParams {
  const __nv_bfloat16* qkv;   // [B*H*W, 3C]
  const float* rpb;           // [(2ws-1)^2, nH]
  __nv_bfloat16* out;         // [B*H*W, C]
  __nv_bfloat16* out_drop;    // optional dropped copy
  float* lse;                 // [B*nW, nH, 64]
  uint64_t drop_seed;
  float drop_p;
  float scale;
  int B, C, nH, H, W, ws, shift, nwh, nww, N;
  int whole_row;    // 1: unit = all heads, local columns = the qkv row; 0: unit = head pair, local columns = 3 x 64
  int heads_per_unit, n_hg, n_atoms, n_units, n_win;
  int koff, voff;   // local column of k / v of the unit's first head
  int cpr;          // 16-byte chunks per gathered row
};

__global__ void __launch_bounds__(kThreads, 1) win_attn_fwd_umma_kernel(const UAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int buf_bytes = p.n_atoms * kAtomBytes;
  const uint32_t qkv_s[2] = {base, base + static_cast<uint32_t>(buf_bytes)};
  const uint32_t p_s = base + 2 * buf_bytes;                       // P tile: 2 atoms
  uint8_t* p_gen = gen + 2 * buf_bytes;
  const int tbl = (2 * p.ws - 1) * (2 * p.ws - 1);
  int* koff_s = reinterpret_cast<int*>(gen + 2 * buf_bytes + 2 * kAtomBytes);   // [64]: ky*(2ws-1)+kx of key j, or -1
  int* reg_s = koff_s + 64;                                        // [2][128] region ids of the unit's tokens (seam)
  float* bias_s = reinterpret_cast<float*>(reg_s + 256);           // [nH][tbl], pre-multiplied by log2 e
  const uint32_t bar0 = (base + 2 * buf_bytes + 2 * kAtomBytes + (p.nH * tbl + 64 + 256) * 4 + 15u) & ~15u;
  auto qkv_full = [&](int b) { return bar0 + 8u * b; };
  auto qkv_empty = [&](int b) { return bar0 + 8u * (2 + b); };
  auto s_full = [&](int b) { return bar0 + 8u * (4 + b); };
  auto s_empty = [&](int b) { return bar0 + 8u * (6 + b); };
  const uint32_t p_full = bar0 + 8u * 8, p_empty = bar0 + 8u * 9, o_empty = bar0 + 8u * 10;
  auto o_full = [&](int j) { return bar0 + 8u * (11 + j); };
  const uint32_t tmem_slot = bar0 + 8u * 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nW = p.nwh * p.nww;

  // ---- one-time setup: zero the operand tiles (padding rows / off-diagonal P blocks stay zero), tables, barriers ----
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(2 * buf_bytes + 2 * kAtomBytes) / 16; i += kThreads)
    reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < p.nH * tbl; i += kThreads) {
    const int h = i / tbl, e = i - h * tbl;
    bias_s[i] = p.rpb[e * p.nH + h] * kLog2e;
  }
  if (threadIdx.x < 64) {
    const int j = threadIdx.x, jy = j / p.ws, jx = j - jy * p.ws;
    koff_s[j] = j < p.N ? jy * (2 * p.ws - 1) + jx : -1;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(qkv_full(b), kProducers);
      mbar_init(qkv_empty(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), 128);
    }
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    mbar_init(o_empty, 128);
    for (int j = 0; j < kMaxUnitHeads; ++j) mbar_init(o_full(j), 1);
    mbar_fence_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp >= 5) {
    // ================================================ gather producers ================================================
    const int tid = threadIdx.x - 160;
    const int C3 = 3 * p.C;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(qkv_empty(b), ((it >> 1) & 1u) ^ 1u);
      const int wp = u / p.n_hg, hg = u - wp * p.n_hg;
      const int n_chunks = 2 * p.N * p.cpr;
      for (int c = tid; c < n_chunks; c += kProducers) {
        const int t = c / p.cpr, cc = c - t * p.cpr;          // t: token slot 0 .. 2N-1
        const int wsel = t >= p.N ? 1 : 0, i = t - wsel * p.N;
        const int win = 2 * wp + wsel;
        if (win >= p.n_win) continue;
        const int bimg = win / nW, wi = win - bimg * nW;
        const int wy = wi / p.nww, wx = wi - wy * p.nww;
        const int iy = i / p.ws, ix = i - iy * p.ws;
        int r = wy * p.ws + iy + p.shift, col = wx * p.ws + ix + p.shift;
        if (r >= p.H) r -= p.H;
        if (col >= p.W) col -= p.W;
        const size_t grow = static_cast<size_t>(bimg * p.H + r) * p.W + col;
        int lc, gc;                                          // local / global column of this chunk
        if (p.whole_row) {
          lc = gc = cc * 8;
        } else {
          const int per_seg = 4 * p.heads_per_unit;          // chunks per q / k / v segment (pairs: 8; a lone head: 4)
          const int seg = cc / per_seg, w8 = (cc - seg * per_seg) * 8;
          lc = seg * 64 + w8;
          gc = seg * p.C + hg * 32 * p.heads_per_unit + w8;
          if (gc >= (seg + 1) * p.C) continue;               // last head group of an odd head count
        }
        const uint32_t dst = qkv_s[b] + (lc >> 6) * kAtomBytes + sw128_offset(wsel * 64 + i, lc & 63);
        cp_async_16(dst, p.qkv + grow * C3 + gc);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(qkv_full(b));
    }
  } else if (warp == 4) {
    // ================================================== MMA issuer ===================================================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16_m128(128);
      const uint32_t idesc_o = umma_idesc_bf16_m128(64) | (1u << 16);     // B operand (V) MN-major
      uint32_t it = 0, sc = 0, pc = 0;                                     // unit / S-buffer / P-buffer use counters
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
        const int b = it & 1;
        const int hg = u % p.n_hg;
        int g = p.heads_per_unit;
        if (!p.whole_row && (hg + 1) * p.heads_per_unit > p.nH) g = p.nH - hg * p.heads_per_unit;
        mbar_wait(qkv_full(b), (it >> 1) & 1u);
        tc_fence_after();
        auto issue_s = [&](int j) {
          const uint32_t sb = sc & 1u;
          mbar_wait(s_empty(sb), ((sc >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const int lq = 32 * j, lk = p.koff + 32 * j;
          const uint32_t qa = qkv_s[b] + (lq >> 6) * kAtomBytes + ((lq >> 5) & 1) * 64;
          const uint32_t ka = qkv_s[b] + (lk >> 6) * kAtomBytes + ((lk >> 5) & 1) * 64;
          for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tmem_base + sb * 128, umma_desc_sw128(qa + ks * 32), umma_desc_sw128(ka + ks * 32), idesc_s,
                      ks ? 1u : 0u);
          umma_commit(s_full(sb));
          ++sc;
        };
        issue_s(0);
        for (int j = 0; j < g; ++j) {
          if (j + 1 < g) issue_s(j + 1);
          mbar_wait(p_full, pc & 1u);
          if (j == 0) mbar_wait(o_empty, (it & 1u) ^ 1u);
          tc_fence_after();
          const int lv = p.voff + 32 * j;
          const uint32_t va = qkv_s[b] + (lv >> 6) * kAtomBytes;
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16(tmem_base + 256 + 64 * j, umma_desc_sw128(p_s + (ks >> 2) * kAtomBytes + (ks & 3) * 32),
                      desc_mn_sw128(va + ks * 2048), idesc_o, ks ? 1u : 0u);
          umma_commit(p_empty);
          umma_commit(o_full(j));
          ++pc;
        }
        umma_commit(qkv_empty(b));
      }
    }
    __syncwarp();
  } else {
    // ============================================ softmax + output (warps 0-3) ============================================
    const int row = threadIdx.x;                  // 0..127 = TMEM lane = row of the unit's 128-row tile
    const int wsel = row >> 6, i = row & 63;
    const bool tok_ok = i < p.N;
    const int iy = i / p.ws, ix = i - iy * p.ws;
    const int bq = tok_ok ? (iy + p.ws - 1) * (2 * p.ws - 1) + ix + p.ws - 1 : 0;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float scale2 = p.scale * kLog2e;
    const uint32_t thr = dropout_threshold(p.drop_p);
    const float keep_scale = 1.f / (1.f - p.drop_p);
    uint32_t it = 0, sc = 0, pc = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int wp = u / p.n_hg, hg = u - wp * p.n_hg;
      int g = p.heads_per_unit;
      if (!p.whole_row && (hg + 1) * p.heads_per_unit > p.nH) g = p.nH - hg * p.heads_per_unit;
      const int head0 = p.whole_row ? 0 : hg * p.heads_per_unit;
      const int win = 2 * wp + wsel;
      const bool win_ok = win < p.n_win;
      const int bimg = win_ok ? win / nW : 0, wi = win_ok ? win - bimg * nW : 0;
      const int wy = wi / p.nww, wx = wi - wy * p.nww;
      int r = wy * p.ws + iy + p.shift, col = wx * p.ws + ix + p.shift;
      if (r >= p.H) r -= p.H;
      if (col >= p.W) col -= p.W;
      const size_t grow = static_cast<size_t>(bimg * p.H + r) * p.W + col;
      const bool store_ok = win_ok && tok_ok;
      // analytic SW-MSA mask (:297-319): 3x3 region ids on the rolled grid; only the last window row / column has seams
      const bool seam_unit = p.shift > 0;          // cheap enough to evaluate for every unit of a shifted block
      int my_reg = 0;
      int* regs = reg_s + (it & 1) * 128;
      if (seam_unit) {
        const int rr0 = wy * p.ws + iy, cc0 = wx * p.ws + ix;
        const int rr = rr0 < p.H - p.ws ? 0 : (rr0 < p.H - p.shift ? 1 : 2);
        const int rc = cc0 < p.W - p.ws ? 0 : (cc0 < p.W - p.shift ? 1 : 2);
        my_reg = rr * 3 + rc;
        regs[row] = my_reg;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      auto finish_head = [&](int j) {    // O_j -> bf16 -> global (64 bytes of this token's row) (+ dropped copy)
        mbar_wait(o_full(j), it & 1u);
        tc_fence_after();
        const int lv = p.voff + 32 * j;
        uint32_t o0[16], o1[16];
        tmem_ld16(t_lane + 256 + 64 * j + ((lv >> 5) & 1) * 32, o0);
        tmem_ld16(t_lane + 256 + 64 * j + ((lv >> 5) & 1) * 32 + 16, o1);
        tmem_ld_wait();
        if (store_ok) {
          const size_t off = grow * p.C + (head0 + j) * 32;
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            w[e] = pack_bf16x2(__uint_as_float(o0[2 * e]), __uint_as_float(o0[2 * e + 1]));
            w[8 + e] = pack_bf16x2(__uint_as_float(o1[2 * e]), __uint_as_float(o1[2 * e + 1]));
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4*>(p.out + off + q4 * 8) = make_uint4(w[4 * q4], w[4 * q4 + 1], w[4 * q4 + 2], w[4 * q4 + 3]);
          if (p.out_drop != nullptr) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint32_t d[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                d[e] = dropout_apply_pair(w[4 * q4 + e], p.drop_seed, off + q4 * 8 + 2 * e, thr, keep_scale);
              *reinterpret_cast<uint4*>(p.out_drop + off + q4 * 8) = make_uint4(d[0], d[1], d[2], d[3]);
            }
          }
        }
      };
      for (int j = 0; j < g; ++j) {
        const uint32_t sb = sc & 1u;
        mbar_wait(s_full(sb), (sc >> 1) & 1u);
        tc_fence_after();
        // this row's 64 scores against the keys of its own window
        uint32_t sr[4][16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) tmem_ld16(t_lane + sb * 128 + wsel * 64 + q4 * 16, sr[q4]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty(sb));
        ++sc;
        const float* bias_h = bias_s + (head0 + j) * tbl + bq;
        float mx = -INFINITY;
        float s[64];
#pragma unroll
        for (int q4 = 0; q4 < 16; ++q4) {
          const int4 ko = *reinterpret_cast<const int4*>(koff_s + q4 * 4);
          const int kk[4] = {ko.x, ko.y, ko.z, ko.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int jj = q4 * 4 + e;
            float v = -INFINITY;
            if (kk[e] >= 0) {
              v = fmaf(__uint_as_float(sr[jj >> 4][jj & 15]), scale2, bias_h[-kk[e]]);
              if (seam_unit && regs[wsel * 64 + jj] != my_reg) v += -100.0f * kLog2e;
            }
            s[jj] = v;
            mx = fmaxf(mx, v);
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int jj = 0; jj < 64; ++jj) {
          s[jj] = ex2f(s[jj] - mx);
          sum += s[jj];
        }
        const float inv = rcpf(sum);
        if (p.lse != nullptr && store_ok)
          p.lse[(static_cast<size_t>(win) * p.nH + head0 + j) * 64 + i] = (mx + lg2f(sum)) * 0.6931471805599453f;
        // P row -> K-major swizzled tile: keys of window `wsel` live in atom `wsel`
        mbar_wait(p_empty, (pc & 1u) ^ 1u);
        uint8_t* pa = p_gen + wsel * kAtomBytes;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2(s[c8 * 8 + 2 * e] * inv, s[c8 * 8 + 2 * e + 1] * inv);
          *reinterpret_cast<uint4*>(pa + sw128_offset(row, c8 * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(p_full);
        ++pc;
        if (j > 0) finish_head(j - 1);
      }
      finish_head(g - 1);
      tc_fence_before();
      mbar_arrive(o_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace

// Returns false when the shape is not served by this kernel (the caller falls back to win_attn_fwd_kernel).
bool win_attn_fwd_umma_supported(int C, int nH, int ws) {
  const bool whole_row = 3 * C <= 320 && nH <= kMaxUnitHeads;
  return C == nH * 32 && ws >= 2 && ws * ws <= 64 && nH >= 1 && nH <= 64 && (whole_row || nH % 2 == 0);
}

int launch_win_attn_fwd_umma(const void* qkv, const float* rpb, void* out, void* out_drop, float* lse, int B, int H,
                             int W, int C, int nH, int ws, int shift, float scale, float drop_p, uint64_t drop_seed,
                             cudaStream_t stream) {
  UAttnParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.rpb = rpb;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_drop = static_cast<__nv_bfloat16*>(out_drop);
  p.lse = lse;
  p.drop_seed = drop_seed;
  p.drop_p = drop_p;
  p.scale = scale;
  p.B = B; p.C = C; p.nH = nH; p.H = H; p.W = W; p.ws = ws; p.shift = shift;
  p.nwh = H / ws; p.nww = W / ws; p.N = ws * ws;
  p.n_win = B * p.nwh * p.nww;
  p.whole_row = (3 * C <= 320 && nH <= kMaxUnitHeads) ? 1 : 0;
  if (p.whole_row) {
    p.heads_per_unit = nH;
    p.n_hg = 1;
    p.n_atoms = (3 * C + 63) / 64;
    p.koff = C;
    p.voff = 2 * C;
    p.cpr = 3 * C / 8;
  } else {
    p.heads_per_unit = 2;
    p.n_hg = (nH + 1) / 2;
    p.n_atoms = 3;
    p.koff = 64;
    p.voff = 128;
    p.cpr = 3 * 8;
  }
  p.n_units = ((p.n_win + 1) / 2) * p.n_hg;
  const int tbl = (2 * ws - 1) * (2 * ws - 1);
  const size_t smem = 1024 + 2 * static_cast<size_t>(p.n_atoms) * kAtomBytes + 2 * kAtomBytes +
                      (static_cast<size_t>(nH) * tbl + 64 + 256) * 4 + 16 + 32 * 8;
  MTL_REQUIRE(smem <= 227 * 1024, "attention (tcgen05): shared memory %zu exceeds 227 KiB", smem);
  MTL_REQUIRE(p.whole_row || nH % 2 == 0, "attention (tcgen05): head pairs need an even head count (got %d)", nH);
  int dev = 0, n_sm = 148;
  MTL_CHECK_CUDA(cudaGetDevice(&dev));
  MTL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_done[64] = {};
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    MTL_CHECK_CUDA(cudaFuncSetAttribute(win_attn_fwd_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done[dev] = true;
  }
  const int grid = p.n_units < n_sm ? p.n_units : n_sm;
  win_attn_fwd_umma_kernel<<<grid, kThreads, smem, stream>>>(p);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
