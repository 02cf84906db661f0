// Shifted-window attention forward on the 5th-generation tensor cores (tcgen05 + TMEM) for sm_100a.
//
// Reference semantics (models/swin_transformer_mtlora.py): torch.roll + window_partition (:338-342, or
// kernels/window_process), q*scale, q@k^T + relative-position bias + SW-MSA mask, softmax, attn@v (:194-220),
// window_reverse + roll back (:365-377). Same contract as win_attn_fwd_kernel (attention.cu), which stays as the path for
// explicit masks: activations stay in (B, H, W, C) token order, the shift / partition / reverse are index math of the
// gather and the scatter, the saved log-sum-exp feeds the unchanged backward kernel.
//
// Work unit = TWO windows (2 x 49 tokens in two 64-row slots of one 128-row UMMA tile) x a PAIR of heads: the gather
// fetches 128-byte q / k / v segments per token (the per-head kernel fetched 64-byte slices and read 1.8x the
// algorithmic bytes at stage 0); an odd head count leaves a lone head in the last unit of a window pair.
// Shared-memory operand layout: the gathered columns are laid out as [128 rows x 64 columns] bf16 atoms with the
// 128-byte swizzle — exactly the K-major UMMA operand atom for Q and K (a head is a 32-column half of an atom, selected by
// the descriptor's start address) and the MN-major operand atom for V (tokens along K, both heads of the atom along N).
//   S_h = Q_h K_h^T        tcgen05.mma M=128 N=128 K=32      (both windows at once; the two off-diagonal 64x64 blocks are
//                                                             never read)
//   softmax                one thread per query row: tcgen05.ld of its window's 64 columns, bias + analytic shift mask,
//                          base-2 softmax in registers, P (bf16) -> K-major swizzled smem tile (zero off-diagonal blocks)
//   O_h = P_h V            tcgen05.mma M=128 N=64 K=128      (N covers the head's atom; its 32-column half is read back)
// Roles (416 threads): warps 0-7 softmax / epilogue (TMEM lane quadrant = warp % 4, 32-column half = warp / 4; the two
// halves of a row exchange max / sum through shared memory), warp 8 lane 0 MMA issuer, warps 9-12 gather producers (one
// token per thread, cp.async into the swizzled atoms, double-buffered across units). S (TMEM) and P (smem) are
// double-buffered so the softmax of head h+1 overlaps the PV product of head h; TMEM: 2 x 128 (S) + 2 x 64 (O) columns.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {
namespace {

constexpr int kAtomBytes = 128 * 128;     // [128 rows x 64 bf16], SW128
constexpr int kSoftmaxWarps = 8;          // warps 0-7: TMEM lane quadrant = warp % 4, column half = warp / 4
constexpr int kProducerWarps = 4;         // warps 9-12
constexpr int kThreads = 32 * (kSoftmaxWarps + 1 + kProducerWarps);
constexpr int kProducers = 32 * kProducerWarps;
constexpr int kSoftmax = 32 * kSoftmaxWarps;
constexpr float kLog2e = 1.4426950408889634f;
constexpr uint32_t kWaitHintNs = 2000u;  // mbarrier.try_wait suspend-time hint (the hardware wakes the thread on completion;
                                         // a short hint only multiplies the polling instructions: 200 ns cost ~20 M of 118 M)

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
  int n_hg, n_units, n_win;
};

// Unit = (pair of windows, pair of heads). Shared memory: 2 x (Q, K, V atoms) | 2 x P tile (2 atoms) | tables | barriers
__global__ void __launch_bounds__(kThreads, 1) win_attn_fwd_umma_kernel(const UAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  constexpr int kBufBytes = 3 * kAtomBytes;
  constexpr int kTileBytes = 2 * kBufBytes + 4 * kAtomBytes;
  const uint32_t qkv_s[2] = {base, base + kBufBytes};
  const uint32_t p_s[2] = {base + 2 * kBufBytes, base + 2 * kBufBytes + 2 * kAtomBytes};
  uint8_t* p_gen = gen + 2 * kBufBytes;
  const int tbl = (2 * p.ws - 1) * (2 * p.ws - 1);
  int* reg_s = reinterpret_cast<int*>(gen + kTileBytes);            // [2][128] region ids of the unit's tokens (seam)
  float* xm_s = reinterpret_cast<float*>(reg_s + 256);              // [2][128] row maxima of the two column halves
  float* xl_s = xm_s + 256;                                         // [2][128] row sums
  float* bias_s = xl_s + 256;                // [nH][tbl + 1], pre-multiplied by log2 e; entry tbl = -inf (padded keys)
  const uint32_t bar0 = (base + kTileBytes + (768 + p.nH * (tbl + 1)) * 4 + 15u) & ~15u;
  auto qkv_full = [&](int b) { return bar0 + 8u * b; };
  auto qkv_empty = [&](int b) { return bar0 + 8u * (2 + b); };
  auto s_full = [&](int b) { return bar0 + 8u * (4 + b); };
  auto s_empty = [&](int b) { return bar0 + 8u * (6 + b); };
  auto p_full = [&](int b) { return bar0 + 8u * (8 + b); };
  auto p_empty = [&](int b) { return bar0 + 8u * (10 + b); };
  auto o_full = [&](int j) { return bar0 + 8u * (12 + j); };
  const uint32_t o_empty = bar0 + 8u * 14;
  const uint32_t tmem_slot = bar0 + 8u * 15;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nW = p.nwh * p.nww;

  // ---- one-time setup: zero the operand tiles (padding rows / off-diagonal P blocks stay zero), tables, barriers ----
  for (uint32_t i = threadIdx.x; i < static_cast<uint32_t>(kTileBytes) / 16; i += kThreads)
    reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < p.nH * (tbl + 1); i += kThreads) {
    const int h = i / (tbl + 1), e = i - h * (tbl + 1);
    bias_s[i] = e < tbl ? p.rpb[e * p.nH + h] * kLog2e : -INFINITY;
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(qkv_full(b), kProducers);
      mbar_init(qkv_empty(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), kSoftmax);
      mbar_init(p_full(b), kSoftmax);
      mbar_init(p_empty(b), 1);
      mbar_init(o_full(b), 1);
    }
    mbar_init(o_empty, kSoftmax);
    mbar_fence_init();
  }
  if (warp == kSoftmaxWarps) {
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp > kSoftmaxWarps) {
    // ================================================ gather producers ================================================
    // thread t owns token slot t (two windows x N tokens <= 128): the row index is computed once per unit, the 24 (12
    // for a lone head) 16-byte chunks of its q / k / v segments stream in with cp.async
    const int t = threadIdx.x - 32 * (kSoftmaxWarps + 1);
    const int wsel = t >> 6, i = t & 63;
    const bool tok_ok = i < p.N;
    const int iy = i / p.ws, ix = i - iy * p.ws;
    const int C3 = 3 * p.C;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int b = it & 1;
      mbar_wait(qkv_empty(b), ((it >> 1) & 1u) ^ 1u, kWaitHintNs);
      const int wp = u / p.n_hg, hg = u - wp * p.n_hg;
      const int win = 2 * wp + wsel;
      if (tok_ok && win < p.n_win) {
        const int bimg = win / nW, wi = win - bimg * nW;
        const int wy = wi / p.nww, wx = wi - wy * p.nww;
        int r = wy * p.ws + iy + p.shift, col = wx * p.ws + ix + p.shift;
        if (r >= p.H) r -= p.H;
        if (col >= p.W) col -= p.W;
        const __nv_bfloat16* src = p.qkv + (static_cast<size_t>(bimg * p.H + r) * p.W + col) * C3 + hg * 64;
        const int g = (2 * hg + 2 <= p.nH) ? 2 : 1;
        const uint32_t dst = qkv_s[b];
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c < 4 * g) cp_async_16(dst + seg * kAtomBytes + sw128_offset(t, c * 8), src + seg * p.C + c * 8);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(qkv_full(b));
    }
  } else if (warp == kSoftmaxWarps) {
    // ================================================== MMA issuer ===================================================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16_m128(128);
      const uint32_t idesc_o = umma_idesc_bf16_m128(64) | (1u << 16);     // B operand (V) MN-major
      uint32_t it = 0, sc = 0, pc = 0, oc[2] = {0, 0};
      for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
        const int b = it & 1;
        const int hg = u % p.n_hg;
        const int g = (2 * hg + 2 <= p.nH) ? 2 : 1;
        mbar_wait(qkv_full(b), (it >> 1) & 1u, kWaitHintNs);
        tc_fence_after();
        auto issue_s = [&](int j) {
          const uint32_t sb = sc & 1u;
          mbar_wait(s_empty(sb), ((sc >> 1) & 1u) ^ 1u, kWaitHintNs);
          tc_fence_after();
          const uint32_t qa = qkv_s[b] + j * 64, ka = qkv_s[b] + kAtomBytes + j * 64;
          for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tmem_base + sb * 128, umma_desc_sw128(qa + ks * 32), umma_desc_sw128(ka + ks * 32), idesc_s,
                      ks ? 1u : 0u);
          umma_commit(s_full(sb));
          ++sc;
        };
        issue_s(0);
        for (int j = 0; j < g; ++j) {
          if (j + 1 < g) issue_s(j + 1);
          const uint32_t pb = pc & 1u;
          mbar_wait(p_full(pb), (pc >> 1) & 1u, kWaitHintNs);
          if (j == 0) mbar_wait(o_empty, (it & 1u) ^ 1u, kWaitHintNs);
          tc_fence_after();
          const uint32_t va = qkv_s[b] + 2 * kAtomBytes;
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16(tmem_base + 256 + 64 * j, umma_desc_sw128(p_s[pb] + (ks >> 2) * kAtomBytes + (ks & 3) * 32),
                      desc_mn_sw128(va + ks * 2048), idesc_o, ks ? 1u : 0u);
          umma_commit(p_empty(pb));
          umma_commit(o_full(j));
          ++pc;
          ++oc[j];
        }
        umma_commit(qkv_empty(b));
      }
    }
    __syncwarp();
  } else {
    // ========================================== softmax + output (warps 0-7) ==========================================
    const int quad = warp & 3, half = warp >> 2;   // TMEM lane quadrant, 32-column half of the row's 64 scores
    const int row = quad * 32 + lane;              // row of the unit's 128-row tile = TMEM lane
    const int wsel = row >> 6, i = row & 63;
    const bool tok_ok = i < p.N;
    const int iy = i / p.ws, ix = i - iy * p.ws;
    const int bq = tok_ok ? (iy + p.ws - 1) * (2 * p.ws - 1) + ix + p.ws - 1 : 0;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float scale2 = p.scale * kLog2e;
    const uint32_t thr = dropout_threshold(p.drop_p);
    const float keep_scale = 1.f / (1.f - p.drop_p);
    // bias-table entry of (this query, key column c) for this thread's 32 key columns — the same for every unit and head
    // (a head only shifts the table base); padded key columns point at the table's -inf sentinel
    int bidx[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int j = half * 32 + c, jy = j / p.ws, jx = j - jy * p.ws;
      bidx[c] = j < p.N ? (tok_ok ? bq - (jy * (2 * p.ws - 1) + jx) : 0) : tbl;
    }
    // this row's four 16-byte chunk slots of the swizzled P tile (fixed per thread)
    uint32_t poff[4];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) poff[c8] = wsel * kAtomBytes + sw128_offset(row, half * 32 + c8 * 8);
    const int pair_bar = 2 + quad;                 // named barrier of the two warps sharing this quadrant's rows
    uint32_t it = 0, sc = 0, pc = 0, oc[2] = {0, 0};
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++it) {
      const int wp = u / p.n_hg, hg = u - wp * p.n_hg;
      const int g = (2 * hg + 2 <= p.nH) ? 2 : 1;
      const int head0 = 2 * hg;
      const int win = 2 * wp + wsel;
      const bool win_ok = win < p.n_win;
      const int bimg = win_ok ? win / nW : 0, wi = win_ok ? win - bimg * nW : 0;
      const int wy = wi / p.nww, wx = wi - wy * p.nww;
      int r = wy * p.ws + iy + p.shift, col = wx * p.ws + ix + p.shift;
      if (r >= p.H) r -= p.H;
      if (col >= p.W) col -= p.W;
      const size_t grow = static_cast<size_t>(bimg * p.H + r) * p.W + col;
      const bool store_ok = win_ok && tok_ok;
      // analytic SW-MSA mask (:297-319): 3x3 region ids on the rolled grid; bit c of `diff` = key column c lies in another
      // region than this query (only windows in the last window row / column have more than one region)
      uint32_t diff = 0;
      if (p.shift > 0) {
        const int rr0 = wy * p.ws + iy, cc0 = wx * p.ws + ix;
        const int rr = rr0 < p.H - p.ws ? 0 : (rr0 < p.H - p.shift ? 1 : 2);
        const int rc = cc0 < p.W - p.ws ? 0 : (cc0 < p.W - p.shift ? 1 : 2);
        const int my_reg = rr * 3 + rc;
        int* regs = reg_s + (it & 1) * 128;
        if (half == 0) regs[row] = my_reg;
        asm volatile("bar.sync 1, %0;" ::"n"(kSoftmax) : "memory");
        const bool seam = wy == p.nwh - 1 || wx == p.nww - 1;
        if (seam) {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const int4 rj = *reinterpret_cast<const int4*>(regs + wsel * 64 + half * 32 + c4 * 4);
            diff |= (rj.x != my_reg ? 1u : 0u) << (c4 * 4) | (rj.y != my_reg ? 1u : 0u) << (c4 * 4 + 1) |
                    (rj.z != my_reg ? 1u : 0u) << (c4 * 4 + 2) | (rj.w != my_reg ? 1u : 0u) << (c4 * 4 + 3);
          }
        }
      }
      auto finish_head = [&](int j) {    // this thread's 16 of the 32 output columns of head j -> bf16 -> global
        mbar_wait(o_full(j), oc[j] & 1u, kWaitHintNs);
        ++oc[j];
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(t_lane + 256 + 64 * j + j * 32 + half * 16, o);
        tmem_ld_wait();
        if (store_ok) {
          const size_t off = grow * p.C + (head0 + j) * 32 + half * 16;
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) w[e] = pack_bf16x2(__uint_as_float(o[2 * e]), __uint_as_float(o[2 * e + 1]));
          *reinterpret_cast<uint4*>(p.out + off) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(p.out + off + 8) = make_uint4(w[4], w[5], w[6], w[7]);
          if (p.out_drop != nullptr) {
            uint32_t d[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] = dropout_apply_pair(w[e], p.drop_seed, off + 2 * e, thr, keep_scale);
            *reinterpret_cast<uint4*>(p.out_drop + off) = make_uint4(d[0], d[1], d[2], d[3]);
            *reinterpret_cast<uint4*>(p.out_drop + off + 8) = make_uint4(d[4], d[5], d[6], d[7]);
          }
        }
      };
      for (int j = 0; j < g; ++j) {
        const uint32_t sb = sc & 1u;
        mbar_wait(s_full(sb), (sc >> 1) & 1u, kWaitHintNs);
        tc_fence_after();
        uint32_t sr[2][16];
        tmem_ld16(t_lane + sb * 128 + wsel * 64 + half * 32, sr[0]);
        tmem_ld16(t_lane + sb * 128 + wsel * 64 + half * 32 + 16, sr[1]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty(sb));
        ++sc;
        const float* bias_h = bias_s + (head0 + j) * (tbl + 1);
        float s[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = fmaf(__uint_as_float(sr[c >> 4][c & 15]), scale2, bias_h[bidx[c]]);
        if (diff != 0) {   // seam windows only: keys of another region get the reference's -100 (:318-319)
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if ((diff >> c) & 1u) s[c] += -100.0f * kLog2e;
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, s[c]);
        xm_s[half * 128 + row] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        mx = fmaxf(mx, xm_s[(half ^ 1) * 128 + row]);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          s[c] = ex2f(s[c] - mx);
          sum += s[c];
        }
        xl_s[half * 128 + row] = sum;
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        sum += xl_s[(half ^ 1) * 128 + row];
        const float inv = rcpf(sum);
        if (p.lse != nullptr && store_ok && half == 0)
          p.lse[(static_cast<size_t>(win) * p.nH + head0 + j) * 64 + i] = (mx + lg2f(sum)) * 0.6931471805599453f;
        // P row -> K-major swizzled tile: keys of window `wsel` live in atom `wsel`
        const uint32_t pb = pc & 1u;
        mbar_wait(p_empty(pb), ((pc >> 1) & 1u) ^ 1u, kWaitHintNs);
        uint8_t* pa = p_gen + pb * 2 * kAtomBytes;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16x2(s[c8 * 8 + 2 * e] * inv, s[c8 * 8 + 2 * e + 1] * inv);
          *reinterpret_cast<uint4*>(pa + poff[c8]) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(p_full(pb));
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
  if (warp == kSoftmaxWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

}  // namespace

// Returns false when the shape is not served by this kernel (the caller falls back to win_attn_fwd_kernel).
bool win_attn_fwd_umma_supported(int C, int nH, int ws) {
  return C == nH * 32 && ws >= 2 && ws * ws <= 64 && nH >= 1 && nH <= 64;
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
  p.n_hg = (nH + 1) / 2;
  p.n_units = ((p.n_win + 1) / 2) * p.n_hg;
  const int tbl = (2 * ws - 1) * (2 * ws - 1);
  const size_t smem = 1024 + 10 * static_cast<size_t>(kAtomBytes) + (768 + static_cast<size_t>(nH) * (tbl + 1)) * 4 + 16 + 32 * 8;
  MTL_REQUIRE(smem <= 227 * 1024, "attention (tcgen05): shared memory %zu exceeds 227 KiB", smem);
  int dev = 0, n_sm = 148;
  MTL_CHECK_CUDA(cudaGetDevice(&dev));
  MTL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  static bool attr_done[64] = {};
  MTL_CHECK_CUDA(ensure_max_dyn_smem(attr_done, win_attn_fwd_umma_kernel, 227 * 1024));
  const int grid = p.n_units < n_sm ? p.n_units : n_sm;
  win_attn_fwd_umma_kernel<<<grid, kThreads, smem, stream>>>(p);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
