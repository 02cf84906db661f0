// Fused shifted-window attention for sm_100a.
//
// Replaces, in one kernel per direction (reference models/swin_transformer_mtlora.py):
//   torch.roll + window_partition      :338-342   (or kernels/window_process WindowProcess)
//   q*scale, q@k^T, + relative-position bias gather, + attn_mask, softmax, attn@v   :194-220
//   window_reverse + torch.roll back   :365-377   (or WindowProcessReverse)
// The cyclic shift / partition / reverse are pure index math on the gather of q,k,v rows and on the
// scatter of the output rows, so activations stay in (B, H, W, C) token order end to end.
//
// A window is N = ws*ws <= 64 tokens with head_dim 32: one CTA (4 warps x 16 query rows) handles a
// (window, head) pair per iteration on mma.sync m16n8k16 tiles. Arithmetic intensity is ~24 FLOP/B,
// so the kernel is bound by the HBM gather/scatter, not the tensor pipe.
#include "common.cuh"
#include "kernels.cuh"

#include <stdlib.h>

namespace mtl {

namespace {

constexpr int HD = 32;        // head dim (fixed in Swin: C / num_heads == 32)
constexpr int NP = 64;        // padded tokens per window
constexpr int QS = 40;        // smem row stride (bf16) for [NP][HD] tiles: 80 B, conflict-free ldmatrix
constexpr int PS = 72;        // smem row stride (bf16) for [NP][NP] tiles

// single-instruction MUFU forms (arguments are max-subtracted / positive sums: no range handling needed;
// ex2.approx(-inf) = 0, relative error 2^-22)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct WinGeom {
  int H, W, ws, shift, nwh, nww, N;
};

// token index (within window) -> row in the (B*H*W) token matrix, with the cyclic shift folded in
__device__ __forceinline__ int token_row(const WinGeom& g, int b, int wy, int wx, int i) {
  const int iy = i / g.ws, ix = i - iy * g.ws;
  int r = wy * g.ws + iy + g.shift;
  int c = wx * g.ws + ix + g.shift;
  if (r >= g.H) r -= g.H;
  if (c >= g.W) c -= g.W;
  return (b * g.H + r) * g.W + c;
}
// region id used by the analytic SW-MSA mask (:297-319), on rolled coordinates
__device__ __forceinline__ int region_id(const WinGeom& g, int wy, int wx, int i) {
  const int iy = i / g.ws, ix = i - iy * g.ws;
  const int r = wy * g.ws + iy, c = wx * g.ws + ix;
  const int rr = r < g.H - g.ws ? 0 : (r < g.H - g.shift ? 1 : 2);
  const int rc = c < g.W - g.ws ? 0 : (c < g.W - g.shift ? 1 : 2);
  return rr * 3 + rc;
}

// S (16 x 64 per warp, c-layout) = scale * Q K^T + bias + mask ; padded key columns -> -inf
__device__ __forceinline__ void scores_16x64(float (&s)[8][4], const __nv_bfloat16* Qs, const __nv_bfloat16* Ks,
                                             int warp, int lane) {
  uint32_t aq[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
    ldmatrix_x4(aq[ks], smem_u32(Qs + (warp * 16 + (lane & 15)) * QS + ks * 16 + (lane >> 4) * 8));
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    uint32_t bk[4];
    ldmatrix_x4(bk, smem_u32(Ks + (nt * 8 + (lane & 7)) * QS + (lane >> 3) * 8));
    s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    const uint32_t b0[2] = {bk[0], bk[1]}, b1[2] = {bk[2], bk[3]};
    mma_bf16_16816(s[nt], aq[0], b0);
    mma_bf16_16816(s[nt], aq[1], b1);
  }
}

struct AttnParams {
  const __nv_bfloat16* qkv;   // [B*H*W, 3C]
  const float* rpb;           // [(2ws-1)^2, nH]
  const float* mask;          // optional explicit mask [nW_mask, N, N] (else analytic when shift > 0)
  int n_mask;
  __nv_bfloat16* out;         // [B*H*W, C]
  __nv_bfloat16* out_drop;    // optional dropped copy (LoRA dropout for proj), same shape
  float* lse;                 // [B*nW, nH, NP]
  uint64_t drop_seed;
  float drop_p;
  int B, C, nH;
  float scale;
  WinGeom g;
};

__global__ void __launch_bounds__(128, 4) win_attn_fwd_kernel(const AttnParams p) {
  // q / k / v tiles are double-buffered: the gather of the CTA's next window runs under the math of the current one
  __shared__ __align__(16) __nv_bfloat16 QKV[2][3][NP * QS];
  __shared__ float bias_s[225];
  __shared__ int2 tok_yx[NP];   // (row, column) of every window token: one division per token per CTA, not per pair
  __shared__ __align__(8) int reg_s[NP];

  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g4 = lane >> 2, t4 = lane & 3;
  const WinGeom g = p.g;
  const int nW = g.nwh * g.nww;
  const int n_win = p.B * nW;
  const int tbl = (2 * g.ws - 1) * (2 * g.ws - 1);
  for (int i = threadIdx.x; i < tbl; i += blockDim.x) bias_s[i] = p.rpb[i * p.nH + head];
  if (threadIdx.x < NP) {
    const int ty = threadIdx.x / g.ws;
    tok_yx[threadIdx.x] = make_int2(ty, threadIdx.x - ty * g.ws);
  }
  if (threadIdx.x < NP) reg_s[threadIdx.x] = 0;
  const int C3 = 3 * p.C;
  __syncthreads();
  // The (query, key) pairs a thread owns are the same for every window, so the relative-position bias of this head
  // (swin_transformer_mtlora.py:202-207) is gathered once into registers, pre-multiplied by log2(e); padded key
  // columns get -inf. Per window only the optional SW-MSA mask is added.
  constexpr float kLog2e = 1.4426950408889634f;
  const int i0 = warp * 16 + g4, i1 = i0 + 8;
  float bias_r[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = (e < 2) ? i0 : i1;
      const int j = nt * 8 + t4 * 2 + (e & 1);
      float b = -INFINITY;
      if (j < g.N) {
        const int ic = i < g.N ? i : 0;
        const int iy = tok_yx[ic].x, ix = tok_yx[ic].y, jy = tok_yx[j].x, jx = tok_yx[j].y;
        b = bias_s[(iy - jy + g.ws - 1) * (2 * g.ws - 1) + (ix - jx + g.ws - 1)] * kLog2e;
      }
      bias_r[nt][e] = b;
    }
  }
  const float scale2 = p.scale * kLog2e;
  // gather / scatter ownership: thread t moves the 16-byte chunk (t & 3) of window tokens (t >> 2) and (t >> 2) + 32
  // of every tile, so the token -> row index math is done twice per window and thread, in registers
  const int ch8 = (threadIdx.x & 3) * 8;
  const int tok_a = threadIdx.x >> 2, tok_b = tok_a + 32;
  const int ay = tok_a / g.ws, ax = tok_a - ay * g.ws, by = tok_b / g.ws, bx = tok_b - by * g.ws;
  const bool a_ok = tok_a < g.N, b_ok = tok_b < g.N;
  // token (within window) -> row of the token matrix for window `win`, cyclic shift folded in
  auto rows_of = [&](int win, int& ra, int& rb) {
    const int b = win / nW, wi = win - b * nW;
    const int wy = wi / g.nww, wx = wi - wy * g.nww;
    int r = wy * g.ws + ay + g.shift, c = wx * g.ws + ax + g.shift;
    if (r >= g.H) r -= g.H;
    if (c >= g.W) c -= g.W;
    ra = (b * g.H + r) * g.W + c;
    r = wy * g.ws + by + g.shift; c = wx * g.ws + bx + g.shift;
    if (r >= g.H) r -= g.H;
    if (c >= g.W) c -= g.W;
    rb = (b * g.H + r) * g.W + c;
  };
  // gather q,k,v rows of this head into buffer `bi`: 3 tiles x 2 rows x one 16-byte chunk per thread
  auto gather = [&](int bi, int ra, int rb) {
    const __nv_bfloat16* sa = p.qkv + static_cast<size_t>(a_ok ? ra : 0) * C3 + head * HD + ch8;
    const __nv_bfloat16* sb = p.qkv + static_cast<size_t>(b_ok ? rb : 0) * C3 + head * HD + ch8;
#pragma unroll
    for (int w3 = 0; w3 < 3; ++w3) {
      cp_async_16_zfill(smem_u32(QKV[bi][w3] + tok_a * QS + ch8), sa + w3 * p.C, a_ok);
      cp_async_16_zfill(smem_u32(QKV[bi][w3] + tok_b * QS + ch8), sb + w3 * p.C, b_ok);
    }
    cp_async_commit();
  };

  int buf = 0;
  int ra = 0, rb = 0, ra_n = 0, rb_n = 0;
  if (static_cast<int>(blockIdx.x) < n_win) {
    rows_of(blockIdx.x, ra, rb);
    gather(0, ra, rb);
  }
  for (int win = blockIdx.x; win < n_win; win += gridDim.x, buf ^= 1, ra = ra_n, rb = rb_n) {
    const int b = win / nW, wi = win - b * nW;
    const int wy = wi / g.nww, wx = wi - wy * g.nww;
    // only windows in the last window row / column straddle the shift seam (:297-319)
    const bool seam = p.mask == nullptr && g.shift > 0 && (wy == g.nwh - 1 || wx == g.nww - 1);
    __nv_bfloat16* Qs = QKV[buf][0];
    __nv_bfloat16* Ks = QKV[buf][1];
    __nv_bfloat16* Vs = QKV[buf][2];
    __syncthreads();  // previous iteration done with the other buffer (its output staging) and with reg_s
    if (seam && threadIdx.x < NP) reg_s[threadIdx.x] = threadIdx.x < g.N ? region_id(g, wy, wx, threadIdx.x) : 0;
    const int next = win + gridDim.x;
    if (next < n_win) {
      rows_of(next, ra_n, rb_n);
      gather(buf ^ 1, ra_n, rb_n);
      cp_async_wait<1>();   // this window's tiles have landed; the next window's are in flight
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float s[8][4];
    scores_16x64(s, Qs, Ks, warp, lane);

    // bias + mask + softmax (base 2) on rows i0 and i1
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = fmaf(s[nt][e], scale2, bias_r[nt][e]);
    }
    if (p.mask != nullptr) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (e < 2) ? i0 : i1;
          const int j = nt * 8 + t4 * 2 + (e & 1);
          if (i < g.N && j < g.N)
            s[nt][e] += p.mask[(static_cast<size_t>(win % p.n_mask) * g.N + i) * g.N + j] * kLog2e;
        }
      }
    } else if (seam) {
      const int r0 = reg_s[i0], r1 = reg_s[i1];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int2 rj = *reinterpret_cast<const int2*>(&reg_s[nt * 8 + t4 * 2]);
        if (r0 != rj.x) s[nt][0] += -100.0f * kLog2e;
        if (r0 != rj.y) s[nt][1] += -100.0f * kLog2e;
        if (r1 != rj.x) s[nt][2] += -100.0f * kLog2e;
        if (r1 != rj.y) s[nt][3] += -100.0f * kLog2e;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = (e < 2) ? mx0 : mx1;
        const float pv = fast_ex2(s[nt][e] - m);
        s[nt][e] = pv;
        if (e < 2) sum0 += pv; else sum1 += pv;
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = fast_rcp(sum0), inv1 = fast_rcp(sum1);
    if (p.lse != nullptr && t4 == 0) {
      float* l = p.lse + (static_cast<size_t>(win) * p.nH + head) * NP;
      l[i0] = (mx0 + fast_lg2(sum0)) * 0.6931471805599453f;   // natural-log units
      l[i1] = (mx1 + fast_lg2(sum1)) * 0.6931471805599453f;
    }

    // O = P V  (P from registers, V via transposed ldmatrix)
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ap[4];
      ap[0] = pack_bf16x2(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
      ap[1] = pack_bf16x2(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
      ap[2] = pack_bf16x2(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
      ap[3] = pack_bf16x2(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bv[4];
        ldmatrix_x4_trans(bv, smem_u32(Vs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * QS + np * 16 + (lane >> 4) * 8));
        const uint32_t b0[2] = {bv[0], bv[1]}, b1[2] = {bv[2], bv[3]};
        mma_bf16_16816(o[np * 2], ap, b0);
        mma_bf16_16816(o[np * 2 + 1], ap, b1);
      }
    }
    __syncthreads();  // everyone done reading Qs -> reuse as output staging
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<uint32_t*>(Qs + i0 * QS + nt * 8 + t4 * 2) = pack_bf16x2(o[nt][0], o[nt][1]);
      *reinterpret_cast<uint32_t*>(Qs + i1 * QS + nt * 8 + t4 * 2) = pack_bf16x2(o[nt][2], o[nt][3]);
    }
    __syncthreads();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (!(hh ? b_ok : a_ok)) continue;
      const int i = hh ? tok_b : tok_a;
      const uint4 v = *reinterpret_cast<const uint4*>(Qs + i * QS + ch8);
      const size_t off = static_cast<size_t>(hh ? rb : ra) * p.C + head * HD + ch8;
      *reinterpret_cast<uint4*>(p.out + off) = v;
      if (p.out_drop != nullptr) {
        const float keep_scale = 1.f / (1.f - p.drop_p);
        const uint32_t thr = dropout_threshold(p.drop_p);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o4[e] = dropout_apply_pair(w4[e], p.drop_seed, off + 2 * e, thr, keep_scale);
        *reinterpret_cast<uint4*>(p.out_drop + off) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

struct AttnBwdParams {
  const __nv_bfloat16* qkv;   // [B*H*W, 3C]
  const __nv_bfloat16* dout;  // [B*H*W, C]
  const float* rpb;
  const float* mask;
  int n_mask;
  const float* lse;           // [B*nW, nH, NP]
  __nv_bfloat16* dqkv;        // [B*H*W, 3C]
  float* drpb;                // [(2ws-1)^2, nH] accumulated with atomics (caller zero-fills), or null
  int B, C, nH;
  float scale;
  WinGeom g;
};

constexpr int kBwdTile = NP * QS;                 // one [64 x 32] tile (padded rows), in bf16 elements
constexpr int kBwdSmemBytes = (2 * 4 * kBwdTile + 2 * NP * PS) * 2;   // q,k,v,dO double-buffered + P, dS

__global__ void __launch_bounds__(128, 3) win_attn_bwd_kernel(const AttnBwdParams p) {
  // q / k / v / dO tiles are double-buffered (dynamic shared memory): the gather of the CTA's next window runs under the
  // math of the current one
  extern __shared__ __align__(16) __nv_bfloat16 bwd_smem[];
  __nv_bfloat16* const Ps = bwd_smem + 2 * 4 * kBwdTile;
  __nv_bfloat16* const dSs = Ps + NP * PS;
  __shared__ float bias_s[225];
  __shared__ int2 tok_yx[NP];   // (row, column) of every window token: one division per token per CTA, not per pair
  __shared__ float dbias_s[225];
  __shared__ __align__(8) int reg_s[NP];

  const int head = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g4 = lane >> 2, t4 = lane & 3;
  const WinGeom g = p.g;
  const int nW = g.nwh * g.nww;
  const int n_win = p.B * nW;
  const int tbl = (2 * g.ws - 1) * (2 * g.ws - 1);
  for (int i = threadIdx.x; i < tbl; i += blockDim.x) {
    bias_s[i] = p.rpb[i * p.nH + head];
    dbias_s[i] = 0.f;
  }
  if (threadIdx.x < NP) {
    const int ty = threadIdx.x / g.ws;
    tok_yx[threadIdx.x] = make_int2(ty, threadIdx.x - ty * g.ws);
  }
  const int C3 = 3 * p.C;
  float dsacc[8][4];  // sum over this CTA's windows of dS (for d relative_position_bias_table)
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) dsacc[nt][0] = dsacc[nt][1] = dsacc[nt][2] = dsacc[nt][3] = 0.f;
  __syncthreads();
  // per-thread relative-position bias (log2 units), gathered once: the owned (query, key) pairs never change
  constexpr float kLog2e = 1.4426950408889634f;
  const int i0 = warp * 16 + g4, i1 = i0 + 8;
  float bias_r[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = (e < 2) ? i0 : i1;
      const int j = nt * 8 + t4 * 2 + (e & 1);
      float b = -INFINITY;
      if (j < g.N) {
        const int ic = i < g.N ? i : 0;
        const int iy = tok_yx[ic].x, ix = tok_yx[ic].y, jy = tok_yx[j].x, jx = tok_yx[j].y;
        b = bias_s[(iy - jy + g.ws - 1) * (2 * g.ws - 1) + (ix - jx + g.ws - 1)] * kLog2e;
      }
      bias_r[nt][e] = b;
    }
  }
  const float scale2 = p.scale * kLog2e;
  // gather / scatter ownership as in the forward kernel: chunk (t & 3) of tokens (t >> 2) and (t >> 2) + 32
  const int ch8 = (threadIdx.x & 3) * 8;
  const int tok_a = threadIdx.x >> 2, tok_b = tok_a + 32;
  const int ay = tok_a / g.ws, ax = tok_a - ay * g.ws, by = tok_b / g.ws, bx = tok_b - by * g.ws;
  const bool a_ok = tok_a < g.N, b_ok = tok_b < g.N;

  auto rows_of = [&](int win, int& ra, int& rb) {
    const int b = win / nW, wi = win - b * nW;
    const int wy = wi / g.nww, wx = wi - wy * g.nww;
    int r = wy * g.ws + ay + g.shift, c = wx * g.ws + ax + g.shift;
    if (r >= g.H) r -= g.H;
    if (c >= g.W) c -= g.W;
    ra = (b * g.H + r) * g.W + c;
    r = wy * g.ws + by + g.shift; c = wx * g.ws + bx + g.shift;
    if (r >= g.H) r -= g.H;
    if (c >= g.W) c -= g.W;
    rb = (b * g.H + r) * g.W + c;
  };
  auto gather = [&](int bi, int ra, int rb) {
    __nv_bfloat16* t0 = bwd_smem + bi * 4 * kBwdTile;
    const size_t oa = static_cast<size_t>(a_ok ? ra : 0), ob = static_cast<size_t>(b_ok ? rb : 0);
    const __nv_bfloat16* sa = p.qkv + oa * C3 + head * HD + ch8;
    const __nv_bfloat16* sb = p.qkv + ob * C3 + head * HD + ch8;
#pragma unroll
    for (int w3 = 0; w3 < 3; ++w3) {
      cp_async_16_zfill(smem_u32(t0 + w3 * kBwdTile + tok_a * QS + ch8), sa + w3 * p.C, a_ok);
      cp_async_16_zfill(smem_u32(t0 + w3 * kBwdTile + tok_b * QS + ch8), sb + w3 * p.C, b_ok);
    }
    cp_async_16_zfill(smem_u32(t0 + 3 * kBwdTile + tok_a * QS + ch8), p.dout + oa * p.C + head * HD + ch8, a_ok);
    cp_async_16_zfill(smem_u32(t0 + 3 * kBwdTile + tok_b * QS + ch8), p.dout + ob * p.C + head * HD + ch8, b_ok);
    cp_async_commit();
  };

  int buf = 0;
  int ra = 0, rb = 0, ra_n = 0, rb_n = 0;
  if (static_cast<int>(blockIdx.x) < n_win) {
    rows_of(blockIdx.x, ra, rb);
    gather(0, ra, rb);
  }
  for (int win = blockIdx.x; win < n_win; win += gridDim.x, buf ^= 1, ra = ra_n, rb = rb_n) {
    const int b = win / nW, wi = win - b * nW;
    const int wy = wi / g.nww, wx = wi - wy * g.nww;
    const bool seam = p.mask == nullptr && g.shift > 0 && (wy == g.nwh - 1 || wx == g.nww - 1);
    __nv_bfloat16* const Qs = bwd_smem + buf * 4 * kBwdTile;
    __nv_bfloat16* const Ks = Qs + kBwdTile;
    __nv_bfloat16* const Vs = Ks + kBwdTile;
    __nv_bfloat16* const dOs = Vs + kBwdTile;
    __syncthreads();   // previous iteration done with the other buffer (dq / dk / dv staging) and with reg_s
    if (seam && threadIdx.x < NP) reg_s[threadIdx.x] = threadIdx.x < g.N ? region_id(g, wy, wx, threadIdx.x) : 0;
    const int next = win + gridDim.x;
    if (next < n_win) {
      rows_of(next, ra_n, rb_n);
      gather(buf ^ 1, ra_n, rb_n);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // ---- recompute P = exp(S - lse) --------------------------------------------------------
    float s[8][4];
    scores_16x64(s, Qs, Ks, warp, lane);
    const float* l = p.lse + (static_cast<size_t>(win) * p.nH + head) * NP;
    // padded query rows: lse = +inf -> P = 0
    const float lse0 = i0 < g.N ? l[i0] * kLog2e : INFINITY, lse1 = i1 < g.N ? l[i1] * kLog2e : INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = fmaf(s[nt][e], scale2, bias_r[nt][e]);
    }
    if (p.mask != nullptr) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (e < 2) ? i0 : i1;
          const int j = nt * 8 + t4 * 2 + (e & 1);
          if (i < g.N && j < g.N)
            s[nt][e] += p.mask[(static_cast<size_t>(win % p.n_mask) * g.N + i) * g.N + j] * kLog2e;
        }
      }
    } else if (seam) {
      const int r0 = reg_s[i0], r1 = reg_s[i1];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int2 rj = *reinterpret_cast<const int2*>(&reg_s[nt * 8 + t4 * 2]);
        if (r0 != rj.x) s[nt][0] += -100.0f * kLog2e;
        if (r0 != rj.y) s[nt][1] += -100.0f * kLog2e;
        if (r1 != rj.x) s[nt][2] += -100.0f * kLog2e;
        if (r1 != rj.y) s[nt][3] += -100.0f * kLog2e;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = fast_ex2(s[nt][0] - lse0);
      s[nt][1] = fast_ex2(s[nt][1] - lse0);
      s[nt][2] = fast_ex2(s[nt][2] - lse1);
      s[nt][3] = fast_ex2(s[nt][3] - lse1);
    }
    // ---- dP = dO V^T -------------------------------------------------------------------------
    float dp[8][4];
    scores_16x64(dp, dOs, Vs, warp, lane);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      d0 += s[nt][0] * dp[nt][0] + s[nt][1] * dp[nt][1];
      d1 += s[nt][2] * dp[nt][2] + s[nt][3] * dp[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    // dS = P * (dP - delta); stash P and dS (bf16) for the transposed products
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float ds[4];
      ds[0] = s[nt][0] * (dp[nt][0] - d0);
      ds[1] = s[nt][1] * (dp[nt][1] - d0);
      ds[2] = s[nt][2] * (dp[nt][2] - d1);
      ds[3] = s[nt][3] * (dp[nt][3] - d1);
#pragma unroll
      for (int e = 0; e < 4; ++e) dsacc[nt][e] += ds[e];
      const int col = nt * 8 + t4 * 2;
      *reinterpret_cast<uint32_t*>(Ps + i0 * PS + col) = pack_bf16x2(s[nt][0], s[nt][1]);
      *reinterpret_cast<uint32_t*>(Ps + i1 * PS + col) = pack_bf16x2(s[nt][2], s[nt][3]);
      *reinterpret_cast<uint32_t*>(dSs + i0 * PS + col) = pack_bf16x2(ds[0], ds[1]);
      *reinterpret_cast<uint32_t*>(dSs + i1 * PS + col) = pack_bf16x2(ds[2], ds[3]);
      dp[nt][0] = ds[0]; dp[nt][1] = ds[1]; dp[nt][2] = ds[2]; dp[nt][3] = ds[3];
    }
    // ---- dQ = scale * dS K  (A from registers, K via transposed ldmatrix) --------------------
    float dq[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(dp[2 * kk][0], dp[2 * kk][1]);
      a[1] = pack_bf16x2(dp[2 * kk][2], dp[2 * kk][3]);
      a[2] = pack_bf16x2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      a[3] = pack_bf16x2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bk[4];
        ldmatrix_x4_trans(bk, smem_u32(Ks + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * QS + np * 16 + (lane >> 4) * 8));
        const uint32_t b0[2] = {bk[0], bk[1]}, b1[2] = {bk[2], bk[3]};
        mma_bf16_16816(dq[np * 2], a, b0);
        mma_bf16_16816(dq[np * 2 + 1], a, b1);
      }
    }
    __syncthreads();  // Ps / dSs complete; all warps finished reading Ks,Vs for S, dP, dQ
    // ---- dK = scale * dS^T Q ; dV = P^T dO  (this warp owns key rows warp*16 .. +15) ----------
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
      dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // reduction over query rows i in steps of 16
      uint32_t ads[4], apt[4];
      const int mi = lane >> 3;
      const int krow = kk * 16 + (lane & 7) + (mi >> 1) * 8;
      const int mcol = warp * 16 + (mi & 1) * 8;
      ldmatrix_x4_trans(ads, smem_u32(dSs + krow * PS + mcol));
      ldmatrix_x4_trans(apt, smem_u32(Ps + krow * PS + mcol));
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bq[4], bo[4];
        const int brow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int bcol = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(bq, smem_u32(Qs + brow * QS + bcol));
        ldmatrix_x4_trans(bo, smem_u32(dOs + brow * QS + bcol));
        const uint32_t q0[2] = {bq[0], bq[1]}, q1[2] = {bq[2], bq[3]};
        const uint32_t o0[2] = {bo[0], bo[1]}, o1[2] = {bo[2], bo[3]};
        mma_bf16_16816(dk[np * 2], ads, q0);
        mma_bf16_16816(dk[np * 2 + 1], ads, q1);
        mma_bf16_16816(dv[np * 2], apt, o0);
        mma_bf16_16816(dv[np * 2 + 1], apt, o1);
      }
    }
    __syncthreads();  // all reads of Qs/Ks/Vs/dOs done -> reuse Qs,Ks,Vs as dq,dk,dv staging
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = nt * 8 + t4 * 2;
      *reinterpret_cast<uint32_t*>(Qs + i0 * QS + col) = pack_bf16x2(dq[nt][0] * p.scale, dq[nt][1] * p.scale);
      *reinterpret_cast<uint32_t*>(Qs + i1 * QS + col) = pack_bf16x2(dq[nt][2] * p.scale, dq[nt][3] * p.scale);
      *reinterpret_cast<uint32_t*>(Ks + i0 * QS + col) = pack_bf16x2(dk[nt][0] * p.scale, dk[nt][1] * p.scale);
      *reinterpret_cast<uint32_t*>(Ks + i1 * QS + col) = pack_bf16x2(dk[nt][2] * p.scale, dk[nt][3] * p.scale);
      *reinterpret_cast<uint32_t*>(Vs + i0 * QS + col) = pack_bf16x2(dv[nt][0], dv[nt][1]);
      *reinterpret_cast<uint32_t*>(Vs + i1 * QS + col) = pack_bf16x2(dv[nt][2], dv[nt][3]);
    }
    __syncthreads();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      if (!(hh ? b_ok : a_ok)) continue;
      const int i = hh ? tok_b : tok_a;
      __nv_bfloat16* dst = p.dqkv + static_cast<size_t>(hh ? rb : ra) * C3 + head * HD + ch8;
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(Qs + i * QS + ch8);
      *reinterpret_cast<uint4*>(dst + p.C) = *reinterpret_cast<const uint4*>(Ks + i * QS + ch8);
      *reinterpret_cast<uint4*>(dst + 2 * p.C) = *reinterpret_cast<const uint4*>(Vs + i * QS + ch8);
    }
  }

  if (p.drpb != nullptr) {
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (e < 2) ? i0 : i1;
        const int j = nt * 8 + t4 * 2 + (e & 1);
        if (i < g.N && j < g.N) {
          const int iy = tok_yx[i].x, ix = tok_yx[i].y, jy = tok_yx[j].x, jx = tok_yx[j].y;
          atomicAdd(&dbias_s[(iy - jy + g.ws - 1) * (2 * g.ws - 1) + (ix - jx + g.ws - 1)], dsacc[nt][e]);
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < tbl; i += blockDim.x) atomicAdd(p.drpb + i * p.nH + head, dbias_s[i]);
  }
}

int check_geom(int B, int H, int W, int C, int nH, int ws, int shift) {
  MTL_REQUIRE(B > 0 && H > 0 && W > 0, "attention: empty input");
  MTL_REQUIRE(C == nH * HD, "attention: head_dim must be 32 (C=%d, heads=%d)", C, nH);
  MTL_REQUIRE(ws >= 1 && ws <= 8, "attention: window size %d unsupported (1..8)", ws);
  MTL_REQUIRE(H % ws == 0 && W % ws == 0, "attention: H,W (%d,%d) not divisible by window %d", H, W, ws);
  MTL_REQUIRE(shift >= 0 && shift < ws, "attention: shift_size must be in [0, window)");
  return 0;
}

}  // namespace

int launch_win_attn_fwd(const void* qkv, const float* rpb, const float* mask, int n_mask, void* out, void* out_drop,
                        float* lse, int B, int H, int W, int C, int nH, int ws, int shift, float scale,
                        float drop_p, uint64_t drop_seed, cudaStream_t stream) {
  if (int e = check_geom(B, H, W, C, nH, ws, shift)) return e;
  // MTL_ATTN_UMMA=1 selects the tcgen05 / TMEM forward (attention_sm100.cu). It is parity-green but measured SLOWER than
  // the mma.sync kernel below on every Swin stage (B200, batch 32: 0.34 vs 0.15 ms at stage 0, profiles/r02_ncu_attn_umma.json):
  // a 49-token window with head_dim 32 is 0.3 MFLOP of tensor work per (window, head), so the kernel is bound by the
  // CUDA-core softmax / bias work and by the TMEM -> registers -> smem round trip of P, which the register-resident
  // mma.sync formulation does not have. The product path therefore stays on the kernel below.
  const char* um = getenv("MTL_ATTN_UMMA");
  if (mask == nullptr && um != nullptr && um[0] == '1' && win_attn_fwd_umma_supported(C, nH, ws))
    return launch_win_attn_fwd_umma(qkv, rpb, out, out_drop, lse, B, H, W, C, nH, ws, shift, scale, drop_p, drop_seed,
                                    stream);
  AttnParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.rpb = rpb;
  p.mask = mask;
  p.n_mask = n_mask > 0 ? n_mask : 1;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_drop = static_cast<__nv_bfloat16*>(out_drop);
  p.lse = lse;
  p.drop_seed = drop_seed;
  p.drop_p = drop_p;
  p.B = B; p.C = C; p.nH = nH; p.scale = scale;
  p.g = WinGeom{H, W, ws, shift, H / ws, W / ws, ws * ws};
  const int n_win = B * p.g.nwh * p.g.nww;
  int gx = n_win;
  const int cap = (sm_count() * 8 + nH - 1) / nH;  // ~8 CTAs per SM in flight across heads
  if (gx > cap) gx = cap;
  win_attn_fwd_kernel<<<dim3(gx, nH), 128, 0, stream>>>(p); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_win_attn_bwd(const void* qkv, const void* dout, const float* rpb, const float* mask, int n_mask,
                        const float* lse, void* dqkv, float* drpb, int B, int H, int W, int C, int nH, int ws,
                        int shift, float scale, cudaStream_t stream) {
  if (int e = check_geom(B, H, W, C, nH, ws, shift)) return e;
  AttnBwdParams p;
  p.qkv = static_cast<const __nv_bfloat16*>(qkv);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.rpb = rpb;
  p.mask = mask;
  p.n_mask = n_mask > 0 ? n_mask : 1;
  p.lse = lse;
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.drpb = drpb;
  p.B = B; p.C = C; p.nH = nH; p.scale = scale;
  p.g = WinGeom{H, W, ws, shift, H / ws, W / ws, ws * ws};
  const int n_win = B * p.g.nwh * p.g.nww;
  int gx = n_win;
  const int cap = (sm_count() * 3 + nH - 1) / nH;  // 3 resident CTAs per SM (register-limited), strided over windows
  if (gx > cap) gx = cap;
  static bool attr_done[64] = {};
  MTL_CHECK_CUDA(ensure_max_dyn_smem(attr_done, win_attn_bwd_kernel, kBwdSmemBytes));
  win_attn_bwd_kernel<<<dim3(gx, nH), 128, kBwdSmemBytes, stream>>>(p); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
