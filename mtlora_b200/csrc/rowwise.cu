// Row-wise / gather kernels of the Swin block: LayerNorm (optionally fused with the PatchMerging 2x2
// gather), the roll+window-partition gathers of kernels/window_process, dropout, DropPath row scaling and
// operand packing. All are HBM-bound streaming kernels: 16-byte vector accesses, one warp per row.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {

namespace {

// Source row pointer for output row `r`, segment `seg` (PatchMerging: 4 segments of Cs channels, order
// (dy,dx) = (0,0),(1,0),(0,1),(1,1) as in swin_transformer_mtlora.py:462-466).
__device__ __forceinline__ long merge_src_row(long r, int seg, int H, int W) {
  const int H2 = H >> 1, W2 = W >> 1;
  const long b = r / (static_cast<long>(H2) * W2);
  const int rem = static_cast<int>(r - b * H2 * W2);
  const int i = rem / W2, j = rem - i * W2;
  return (b * H + 2 * i + (seg & 1)) * W + 2 * j + (seg >> 1);
}

struct LnFwdParams {
  const __nv_bfloat16* x;
  const float* gamma;
  const float* beta;
  __nv_bfloat16* y;
  __nv_bfloat16* y_drop;
  float* mean;
  float* rstd;
  long rows;
  long drop_rows;  // only rows < drop_rows get a dropped copy (stream 0 of a stream-stacked input)
  int C, merge, H, W;
  float eps, drop_p;
  uint64_t drop_seed;
};

__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const LnFwdParams p) {
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int nvec = p.C >> 3;
  const int Cs = p.merge ? p.C >> 2 : p.C;
  const int vec_per_seg = Cs >> 3;
  auto src_vec = [&](int v) -> const uint4* {
    if (!p.merge) return reinterpret_cast<const uint4*>(p.x + row * p.C) + v;
    const int seg = v / vec_per_seg;
    return reinterpret_cast<const uint4*>(p.x + merge_src_row(row, seg, p.H, p.W) * Cs) + (v - seg * vec_per_seg);
  };
  float sum = 0.f, sq = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = __ldg(src_vec(v));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = bf16lo_to_f32(w[e]), b = bf16hi_to_f32(w[e]);
      sum += a + b;
      sq += a * a + b * b;
    }
  }
  sum = warp_sum(sum);
  sq = warp_sum(sq);
  const float mean = sum / p.C;
  const float var = fmaxf(sq / p.C - mean * mean, 0.f);
  const float rstd = rsqrtf(var + p.eps);
  if (lane == 0) {
    if (p.mean) p.mean[row] = mean;
    if (p.rstd) p.rstd[row] = rstd;
  }
  const uint32_t thr = dropout_threshold(p.drop_p);
  const bool do_drop = p.y_drop != nullptr && row < p.drop_rows;
  const float keep_scale = do_drop ? 1.f / (1.f - p.drop_p) : 1.f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = __ldg(src_vec(v));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * v);
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.beta) + 2 * v + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      o[2 * e] = (bf16lo_to_f32(w[e]) - mean) * rstd * gg[2 * e] + bb[2 * e];
      o[2 * e + 1] = (bf16hi_to_f32(w[e]) - mean) * rstd * gg[2 * e + 1] + bb[2 * e + 1];
    }
    const size_t off = static_cast<size_t>(row) * p.C + v * 8;
    *reinterpret_cast<uint4*>(p.y + off) =
        make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    if (do_drop) {
      // the dropped copy is derived from the bf16-rounded value so that it equals D(y) exactly
      uint32_t d[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        d[e] = dropout_apply_pair(pack_bf16x2(o[2 * e], o[2 * e + 1]), p.drop_seed, off + 2 * e, thr, keep_scale);
      *reinterpret_cast<uint4*>(p.y_drop + off) = make_uint4(d[0], d[1], d[2], d[3]);
    }
  }
}

struct LnBwdParams {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* x;
  const float* gamma;
  const float* mean;
  const float* rstd;
  const __nv_bfloat16* dres;
  __nv_bfloat16* dx;
  float* dgamma;
  float* dbeta;
  long rows;
  int C, merge, H, W;
  int rows_per_cta;
};

// One warp per row. dgamma/dbeta: each lane owns the columns of vectors lane, lane+32, ... and keeps their
// partial sums in registers across all rows of the CTA strip (C <= 1024); wider rows (PatchMerging LN(4C) of the
// deep, short stages) fall back to shared-memory atomics. Per-CTA sums are flushed to the fp32 global
// accumulators with one atomicAdd per column per CTA.
template <bool REGS>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const LnBwdParams p) {
  extern __shared__ float red[];  // [2][C]
  float* dg_s = red;
  float* db_s = red + p.C;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int nvec = p.C >> 3;
  const int Cs = p.merge ? p.C >> 2 : p.C;
  const int vec_per_seg = Cs >> 3;
  const long row_begin = static_cast<long>(blockIdx.x) * p.rows_per_cta;
  long row_end = row_begin + p.rows_per_cta;
  if (row_end > p.rows) row_end = p.rows;
  constexpr int VPL = REGS ? 4 : 1;
  float dg_r[VPL][8], db_r[VPL][8];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) dg_r[k][e] = db_r[k][e] = 0.f;

  for (long row = row_begin + warp; row < row_end; row += nwarp) {
    auto src_off = [&](int v) -> size_t {
      if (!p.merge) return static_cast<size_t>(row) * p.C + v * 8;
      const int seg = v / vec_per_seg;
      return static_cast<size_t>(merge_src_row(row, seg, p.H, p.W)) * Cs + (v - seg * vec_per_seg) * 8;
    };
    const float mean = p.mean[row], rstd = p.rstd[row];
    float s1 = 0.f, s2 = 0.f;  // sum(g), sum(g * xhat), g = dy * gamma
    auto pass1 = [&](int v, int k) {
      const uint4 qx = __ldg(reinterpret_cast<const uint4*>(p.x + src_off(v)));
      const uint4 qd = __ldg(reinterpret_cast<const uint4*>(p.dy + static_cast<size_t>(row) * p.C + v * 8));
      const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = (e & 1) ? bf16hi_to_f32(wx[e >> 1]) : bf16lo_to_f32(wx[e >> 1]);
        const float dv = (e & 1) ? bf16hi_to_f32(wd[e >> 1]) : bf16lo_to_f32(wd[e >> 1]);
        const float xh = (xv - mean) * rstd;
        const float g = dv * gg[e];
        s1 += g;
        s2 += g * xh;
        if (REGS) {
          dg_r[k][e] += dv * xh;
          db_r[k][e] += dv;
        } else {
          atomicAdd(&dg_s[v * 8 + e], dv * xh);
          atomicAdd(&db_s[v * 8 + e], dv);
        }
      }
    };
    if (REGS) {
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int v = lane + 32 * k;
        if (v < nvec) pass1(v, k);
      }
    } else {
      for (int v = lane; v < nvec; v += 32) pass1(v, 0);
    }
    s1 = warp_sum(s1) / p.C;
    s2 = warp_sum(s2) / p.C;
    for (int v = lane; v < nvec; v += 32) {
      const size_t so = src_off(v);
      const uint4 qx = __ldg(reinterpret_cast<const uint4*>(p.x + so));
      const uint4 qd = __ldg(reinterpret_cast<const uint4*>(p.dy + static_cast<size_t>(row) * p.C + v * 8));
      const uint32_t wx[4] = {qx.x, qx.y, qx.z, qx.w}, wd[4] = {qd.x, qd.y, qd.z, qd.w};
      uint32_t wr[4] = {0, 0, 0, 0};
      if (p.dres) {
        const uint4 qr = __ldg(reinterpret_cast<const uint4*>(p.dres + so));
        wr[0] = qr.x; wr[1] = qr.y; wr[2] = qr.z; wr[3] = qr.w;
      }
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = (e & 1) ? bf16hi_to_f32(wx[e >> 1]) : bf16lo_to_f32(wx[e >> 1]);
        const float dv = (e & 1) ? bf16hi_to_f32(wd[e >> 1]) : bf16lo_to_f32(wd[e >> 1]);
        const float rv = (e & 1) ? bf16hi_to_f32(wr[e >> 1]) : bf16lo_to_f32(wr[e >> 1]);
        const float xh = (xv - mean) * rstd;
        o[e] = rstd * (dv * gg[e] - s1 - xh * s2) + rv;
      }
      *reinterpret_cast<uint4*>(p.dx + so) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  }
  if (REGS) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = lane + 32 * k;
      if (v < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(&dg_s[v * 8 + e], dg_r[k][e]);
          atomicAdd(&db_s[v * 8 + e], db_r[k][e]);
        }
      }
    }
  }
  __syncthreads();
  if (p.dgamma) {
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      atomicAdd(p.dgamma + i, dg_s[i]);
      atomicAdd(p.dbeta + i, db_s[i]);
    }
  }
}

// ---- fast LayerNorm for the Swin channel counts --------------------------------------------------------------------
// A row of C = LPR * VPT * 8 channels is owned by LPR lanes (a power of two <= 32) holding VPT 16-byte vectors each, so
// a warp normalises 32 / LPR rows at once with every lane busy (the one-warp-per-row kernels above leave 20 of 32 lanes
// idle at C = 96). The row lives in registers between the statistics and the output pass (one HBM read), gamma / beta
// are loaded once per thread, and the backward keeps the per-column dgamma / dbeta partial sums in registers over all
// the rows a lane visits.
template <int LPR>
__device__ __forceinline__ float subwarp_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR, int VPT>
__global__ void __launch_bounds__(256, VPT >= 6 ? 2 : 4) ln_fwd_fast_kernel(const LnFwdParams p) {
  // gamma / beta live in shared memory (broadcast 16-byte reads) and the row stays packed (bf16) in registers, so the
  // kernel fits 64 registers: 32 resident warps per SM keep ~48 KB of loads in flight (the kernel is latency-bound:
  // 16 warps x 48 B per lane measured 4.5 TB/s)
  extern __shared__ float gb_s[];   // [2][C]
  constexpr int RPW = 32 / LPR;
  for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
    gb_s[i] = p.gamma[i];
    gb_s[p.C + i] = p.beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rsel = lane / LPR;
  const long warp_g = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long n_warps = static_cast<long>(gridDim.x) * (blockDim.x >> 5);
  const int Cs = p.merge ? p.C >> 2 : p.C;
  const int vec_per_seg = Cs >> 3;
  const uint32_t thr = dropout_threshold(p.drop_p);
  const float inv_c = 1.f / p.C;
  const float keep_scale = 1.f / (1.f - p.drop_p);
  for (long rg = warp_g; rg * RPW < p.rows; rg += n_warps) {
    const long row = rg * RPW + rsel;
    const bool ok = row < p.rows;
    uint4 q[VPT];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = sub + k * LPR;
      q[k] = make_uint4(0, 0, 0, 0);
      if (ok) {
        const __nv_bfloat16* src;
        if (!p.merge) {
          src = p.x + row * p.C + v * 8;
        } else {
          const int seg = v / vec_per_seg;
          src = p.x + merge_src_row(row, seg, p.H, p.W) * Cs + (v - seg * vec_per_seg) * 8;
        }
        q[k] = __ldg(reinterpret_cast<const uint4*>(src));
      }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) sum += bf16lo_to_f32(w[e]) + bf16hi_to_f32(w[e]);
    }
    const float mean = subwarp_sum<LPR>(sum) * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = bf16lo_to_f32(w[e]) - mean, d1 = bf16hi_to_f32(w[e]) - mean;
        sq += d0 * d0 + d1 * d1;
      }
    }
    const float rstd = rsqrtf(subwarp_sum<LPR>(sq) * inv_c + p.eps);
    if (!ok) continue;
    if (sub == 0) {
      if (p.mean) p.mean[row] = mean;
      if (p.rstd) p.rstd[row] = rstd;
    }
    const bool do_drop = p.y_drop != nullptr && row < p.drop_rows;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = sub + k * LPR;
      const float4 g0 = *reinterpret_cast<const float4*>(gb_s + v * 8), g1 = *reinterpret_cast<const float4*>(gb_s + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(gb_s + p.C + v * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(gb_s + p.C + v * 8 + 4);
      const float gk[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bk[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        o[e] = pack_bf16x2((bf16lo_to_f32(w[e]) - mean) * rstd * gk[2 * e] + bk[2 * e],
                           (bf16hi_to_f32(w[e]) - mean) * rstd * gk[2 * e + 1] + bk[2 * e + 1]);
      const size_t off = static_cast<size_t>(row) * p.C + v * 8;
      *reinterpret_cast<uint4*>(p.y + off) = make_uint4(o[0], o[1], o[2], o[3]);
      if (do_drop) {
        uint32_t d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) d[e] = dropout_apply_pair(o[e], p.drop_seed, off + 2 * e, thr, keep_scale);
        *reinterpret_cast<uint4*>(p.y_drop + off) = make_uint4(d[0], d[1], d[2], d[3]);
      }
    }
  }
}

// VPT >= 6 (PatchMerging rows of 1536 / 2048 channels): gamma is read from shared memory instead of registers and one
// CTA per SM may use the whole register file (the row and the dgamma / dbeta partial sums stay in registers)
template <int LPR, int VPT, bool PARAM_GRADS>
__global__ void __launch_bounds__(256, VPT >= 6 ? 1 : 2) ln_bwd_fast_kernel(const LnBwdParams p) {
  extern __shared__ float red[];  // [2][C] (+ [C] gamma when GS)
  constexpr int RPW = 32 / LPR;
  constexpr bool GS = VPT >= 6;
  float* sgam = red + 2 * p.C;
  if (GS) {
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) sgam[i] = p.gamma[i];
  }
  if (PARAM_GRADS) {
    for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) red[i] = 0.f;
  }
  if (GS || PARAM_GRADS) __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rsel = lane / LPR;
  const long warp_g = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long n_warps = static_cast<long>(gridDim.x) * (blockDim.x >> 5);
  const int Cs = p.merge ? p.C >> 2 : p.C;
  const int vec_per_seg = Cs >> 3;
  float gam[GS ? 1 : VPT][8];
  float dg[PARAM_GRADS ? VPT : 1][8], db[PARAM_GRADS ? VPT : 1][8];
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int v = sub + k * LPR;
    if (!GS) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma) + 2 * v + 1);
      gam[k][0] = g0.x; gam[k][1] = g0.y; gam[k][2] = g0.z; gam[k][3] = g0.w;
      gam[k][4] = g1.x; gam[k][5] = g1.y; gam[k][6] = g1.z; gam[k][7] = g1.w;
    }
    if (PARAM_GRADS) {
#pragma unroll
      for (int e = 0; e < 8; ++e) dg[k][e] = db[k][e] = 0.f;
    }
  }
  const float inv_c = 1.f / p.C;
  for (long rg = warp_g; rg * RPW < p.rows; rg += n_warps) {
    const long row = rg * RPW + rsel;
    const bool ok = row < p.rows;
    const float mean = ok ? p.mean[row] : 0.f, rstd = ok ? p.rstd[row] : 0.f;
    uint4 qx[VPT], qd[VPT];   // the row stays packed in registers between the two passes
    size_t soff[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int v = sub + k * LPR;
      qx[k] = make_uint4(0, 0, 0, 0);
      qd[k] = make_uint4(0, 0, 0, 0);
      soff[k] = 0;
      if (ok) {
        if (!p.merge) {
          soff[k] = static_cast<size_t>(row) * p.C + v * 8;
        } else {
          const int seg = v / vec_per_seg;
          soff[k] = static_cast<size_t>(merge_src_row(row, seg, p.H, p.W)) * Cs + (v - seg * vec_per_seg) * 8;
        }
        qx[k] = __ldg(reinterpret_cast<const uint4*>(p.x + soff[k]));
        qd[k] = __ldg(reinterpret_cast<const uint4*>(p.dy + static_cast<size_t>(row) * p.C + v * 8));
      }
    }
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const uint32_t wx[4] = {qx[k].x, qx[k].y, qx[k].z, qx[k].w}, wd[4] = {qd[k].x, qd[k].y, qd[k].z, qd[k].w};
      float gk[8];
      if (GS) {
        const float4 g0 = *reinterpret_cast<const float4*>(sgam + (sub + k * LPR) * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgam + (sub + k * LPR) * 8 + 4);
        gk[0] = g0.x; gk[1] = g0.y; gk[2] = g0.z; gk[3] = g0.w; gk[4] = g1.x; gk[5] = g1.y; gk[6] = g1.z; gk[7] = g1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) gk[e] = gam[k][e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float x = (e & 1) ? bf16hi_to_f32(wx[e >> 1]) : bf16lo_to_f32(wx[e >> 1]);
        const float d = (e & 1) ? bf16hi_to_f32(wd[e >> 1]) : bf16lo_to_f32(wd[e >> 1]);
        const float h = (x - mean) * rstd;
        const float g = d * gk[e];
        s1 += g;
        s2 += g * h;
        if (PARAM_GRADS) {
          dg[k][e] += d * h;
          db[k][e] += d;
        }
      }
    }
    s1 = subwarp_sum<LPR>(s1) * inv_c;
    s2 = subwarp_sum<LPR>(s2) * inv_c;
    if (!ok) continue;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      uint32_t wr[4] = {0, 0, 0, 0};
      if (p.dres) {
        const uint4 qr = __ldg(reinterpret_cast<const uint4*>(p.dres + soff[k]));
        wr[0] = qr.x; wr[1] = qr.y; wr[2] = qr.z; wr[3] = qr.w;
      }
      const uint32_t wx[4] = {qx[k].x, qx[k].y, qx[k].z, qx[k].w}, wd[4] = {qd[k].x, qd[k].y, qd[k].z, qd[k].w};
      float gk[8];
      if (GS) {
        const float4 g0 = *reinterpret_cast<const float4*>(sgam + (sub + k * LPR) * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(sgam + (sub + k * LPR) * 8 + 4);
        gk[0] = g0.x; gk[1] = g0.y; gk[2] = g0.z; gk[3] = g0.w; gk[4] = g1.x; gk[5] = g1.y; gk[6] = g1.z; gk[7] = g1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) gk[e] = gam[k][e];
      }
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h0 = (bf16lo_to_f32(wx[e]) - mean) * rstd, h1 = (bf16hi_to_f32(wx[e]) - mean) * rstd;
        const float g0 = bf16lo_to_f32(wd[e]) * gk[2 * e], g1 = bf16hi_to_f32(wd[e]) * gk[2 * e + 1];
        o[e] = pack_bf16x2(rstd * (g0 - s1 - h0 * s2) + bf16lo_to_f32(wr[e]),
                           rstd * (g1 - s1 - h1 * s2) + bf16hi_to_f32(wr[e]));
      }
      *reinterpret_cast<uint4*>(p.dx + soff[k]) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (PARAM_GRADS) {
    // lanes with the same `sub` own the same columns: fold the RPW row slots of the warp, then warp -> CTA -> global
#pragma unroll
    for (int k = 0; k < VPT; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = dg[k][e], b = db[k][e];
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o);
          b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (rsel == 0) {
          atomicAdd(&red[(sub + k * LPR) * 8 + e], a);
          atomicAdd(&red[p.C + (sub + k * LPR) * 8 + e], b);
        }
      }
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) {
      atomicAdd(p.dgamma + i, red[i]);
      atomicAdd(p.dbeta + i, red[p.C + i]);
    }
  }
}

// (lanes per row, vectors per lane) for the channel counts of Swin-T/S (96 * 2^s, x4 for PatchMerging) and Swin-B
// (128 * 2^s); anything else (or very wide rows) takes the generic one-warp-per-row kernels.
static bool ln_fast_shape(int C, int* lpr, int* vpt) {
  switch (C) {
    case 96: *lpr = 4; *vpt = 3; return true;
    case 192: *lpr = 8; *vpt = 3; return true;
    case 384: *lpr = 16; *vpt = 3; return true;
    case 768: *lpr = 32; *vpt = 3; return true;
    case 128: *lpr = 8; *vpt = 2; return true;
    case 256: *lpr = 16; *vpt = 2; return true;
    case 512: *lpr = 32; *vpt = 2; return true;
    case 1024: *lpr = 32; *vpt = 4; return true;
    case 1536: *lpr = 32; *vpt = 6; return true;   // PatchMerging of stage 2 (Swin-T/S): 4 x 384
    case 2048: *lpr = 32; *vpt = 8; return true;   // ... Swin-B: 4 x 512
    default: return false;
  }
}

template <int LPR, int VPT>
static void ln_fwd_launch(const LnFwdParams& p, unsigned grid, cudaStream_t stream) {
  ln_fwd_fast_kernel<LPR, VPT><<<grid, 256, 2 * p.C * sizeof(float), stream>>>(p);
}
template <int LPR, int VPT>
static void ln_bwd_launch(const LnBwdParams& p, unsigned grid, cudaStream_t stream) {
  const size_t sm = (VPT >= 6 ? 3 : 2) * p.C * sizeof(float);
  if (p.dgamma) ln_bwd_fast_kernel<LPR, VPT, true><<<grid, 256, sm, stream>>>(p);
  else ln_bwd_fast_kernel<LPR, VPT, false><<<grid, 256, sm, stream>>>(p);
}
#define MTL_LN_CASES(FN, ...)                      \
  switch (lpr * 100 + vpt) {                       \
    case 403: FN<4, 3>(__VA_ARGS__); break;        \
    case 803: FN<8, 3>(__VA_ARGS__); break;        \
    case 1603: FN<16, 3>(__VA_ARGS__); break;      \
    case 3203: FN<32, 3>(__VA_ARGS__); break;      \
    case 802: FN<8, 2>(__VA_ARGS__); break;        \
    case 1602: FN<16, 2>(__VA_ARGS__); break;      \
    case 3202: FN<32, 2>(__VA_ARGS__); break;      \
    case 3204: FN<32, 4>(__VA_ARGS__); break;      \
    case 3206: FN<32, 6>(__VA_ARGS__); break;      \
    case 3208: FN<32, 8>(__VA_ARGS__); break;      \
    default: break;                                \
  }

// ---- kernels/window_process equivalents ----------------------------------------------------------
// partition: out[b*nWin + wy*nW + wx, iy, ix, :] = in[b, (wy*ws+iy+shift) % H, (wx*ws+ix+shift) % W, :]
// (torch.roll(x, (-shift,-shift)) + window_partition). `scatter` swaps source and destination, which is both
// the backward of partition and the forward of window_reverse + roll(+shift).
template <typename V>
__global__ void window_gather_kernel(const V* __restrict__ in, V* __restrict__ out, int B, int H, int W, int CV,
                                     int shift, int ws, int scatter) {
  const long total = static_cast<long>(B) * H * W * CV;
  const int nww = W / ws, nwh = H / ws;
  for (long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % CV);
    long t = idx / CV;  // window-ordered token index
    const int ix = static_cast<int>(t % ws); t /= ws;
    const int iy = static_cast<int>(t % ws); t /= ws;
    const int wx = static_cast<int>(t % nww); t /= nww;
    const int wy = static_cast<int>(t % nwh); t /= nwh;
    const long b = t;
    int r = wy * ws + iy + shift, q = wx * ws + ix + shift;
    r %= H; if (r < 0) r += H;
    q %= W; if (q < 0) q += W;
    const long img = ((b * H + r) * W + q) * CV + c;
    if (scatter) out[img] = in[idx];
    else out[idx] = in[img];
  }
}

__global__ void dropout_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long n8, float p,
                               uint64_t seed) {
  const uint32_t thr = dropout_threshold(p);
  const float ks = 1.f / (1.f - p);
  for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < n8;
       v += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x) + v);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = dropout_apply_pair(w[e], seed, static_cast<uint64_t>(v) * 8 + 2 * e, thr, ks);
    reinterpret_cast<uint4*>(y)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void scale_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                  __nv_bfloat16* __restrict__ y, int S, long M, int CV, int rows_per_sample,
                                  int n_samples) {
  const long total = static_cast<long>(S) * M * CV;
  for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
       v += static_cast<long>(gridDim.x) * blockDim.x) {
    const long rowg = v / CV;
    const int s = static_cast<int>(rowg / M);
    const long m = rowg - s * M;
    const float f = scale[s * n_samples + m / rows_per_sample];
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x) + v);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(bf16lo_to_f32(w[e]) * f, bf16hi_to_f32(w[e]) * f);
    reinterpret_cast<uint4*>(y)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// y[s] = x[s] * scale[s, sample] (scale == null: y[s] = x[s], not rewritten when y == x) and y[S] = sum_s y[s]
// (sum of the bf16-rounded scaled rows, fp32 accumulation): the pre-summed gradient stream of mtl_linear_bwd_input.
__global__ void scale_rows_sum_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                      __nv_bfloat16* __restrict__ y, int S, long M, int CV, int rows_per_sample,
                                      int n_samples) {
  const long per = M * CV;
  for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < per;
       v += static_cast<long>(gridDim.x) * blockDim.x) {
    const long m = v / CV;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < S; ++s) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x) + s * per + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t o[4];
      if (scale != nullptr) {
        const float f = scale[s * n_samples + m / rows_per_sample];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(bf16lo_to_f32(w[e]) * f, bf16hi_to_f32(w[e]) * f);
        reinterpret_cast<uint4*>(y)[s * per + v] = make_uint4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = w[e];
        if (y != x) reinterpret_cast<uint4*>(y)[s * per + v] = q;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += bf16lo_to_f32(o[e]);
        acc[2 * e + 1] += bf16hi_to_f32(o[e]);
      }
    }
    reinterpret_cast<uint4*>(y)[S * per + v] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                         pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

__global__ void add_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                           __nv_bfloat16* __restrict__ out, long n8) {
  for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < n8;
       v += static_cast<long>(gridDim.x) * blockDim.x) {
    const uint4 qa = __ldg(reinterpret_cast<const uint4*>(a) + v);
    const uint4 qb = __ldg(reinterpret_cast<const uint4*>(b) + v);
    const uint32_t wa[4] = {qa.x, qa.y, qa.z, qa.w}, wb[4] = {qb.x, qb.y, qb.z, qb.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = pack_bf16x2(bf16lo_to_f32(wa[e]) + bf16lo_to_f32(wb[e]), bf16hi_to_f32(wa[e]) + bf16hi_to_f32(wb[e]));
    reinterpret_cast<uint4*>(out)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// out[i] = sum_s x[s, i] (+ extra[i]); fp32 accumulation, bf16 in/out
__global__ void sum_streams_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ extra,
                                   __nv_bfloat16* __restrict__ out, int S, long n8) {
  for (long v = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; v < n8;
       v += static_cast<long>(gridDim.x) * blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s <= S; ++s) {
      if (s == S && extra == nullptr) break;
      const uint4 q = (s < S) ? __ldg(reinterpret_cast<const uint4*>(x) + s * n8 + v)
                              : __ldg(reinterpret_cast<const uint4*>(extra) + v);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] += bf16lo_to_f32(w[e]);
        acc[2 * e + 1] += bf16hi_to_f32(w[e]);
      }
    }
    reinterpret_cast<uint4*>(out)[v] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]),
                                                  pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

struct PackParams {
  const float* a[8];  // [r_i, K]
  const float* b[8];  // [N, r_i]
  int rank[8], off[8];
  int n, K, N, R_pad;
  __nv_bfloat16 *a_cat, *b_cat, *a_cat_t, *b_cat_t;
};

// a_cat [R_pad, K], b_cat [N, R_pad], a_cat_t [K, R_pad], b_cat_t [R_pad, N]; zero padded.
// Elements start, start + stride, ... of one layer's packed operands.
__device__ __forceinline__ void pack_adapters_body(const PackParams& p, long start, long stride) {
  const long nA = static_cast<long>(p.R_pad) * p.K, nB = static_cast<long>(p.N) * p.R_pad;
  for (long idx = start; idx < nA + nB; idx += stride) {
    if (idx < nA) {
      const int r = static_cast<int>(idx / p.K), k = static_cast<int>(idx - static_cast<long>(r) * p.K);
      float v = 0.f;
      for (int i = 0; i < p.n; ++i)
        if (r >= p.off[i] && r < p.off[i] + p.rank[i]) v = p.a[i][static_cast<long>(r - p.off[i]) * p.K + k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      if (p.a_cat) p.a_cat[idx] = h;
      if (p.a_cat_t) p.a_cat_t[static_cast<long>(k) * p.R_pad + r] = h;
    } else {
      const long j = idx - nA;
      const int n = static_cast<int>(j / p.R_pad), r = static_cast<int>(j - static_cast<long>(n) * p.R_pad);
      float v = 0.f;
      for (int i = 0; i < p.n; ++i)
        if (r >= p.off[i] && r < p.off[i] + p.rank[i]) v = p.b[i][static_cast<long>(n) * p.rank[i] + (r - p.off[i])];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      if (p.b_cat) p.b_cat[j] = h;
      if (p.b_cat_t) p.b_cat_t[static_cast<long>(r) * p.N + n] = h;
    }
  }
}

__global__ void pack_adapters_kernel(const PackParams p) {
  pack_adapters_body(p, static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<long>(gridDim.x) * blockDim.x);
}

// All layers of a model in ONE launch (the adapters change once per optimizer step: 48 launches of ~6 us otherwise).
// The job table travels as a kernel parameter (<= 32 KiB on sm_100); blocks [blk0[j], blk0[j + 1]) serve job j.
constexpr int kPackManyMax = 64;
struct PackMany {
  int n;
  int blk0[kPackManyMax + 1];
  PackParams job[kPackManyMax];
};
static_assert(sizeof(PackMany) <= 32 * 1024 - 64, "job table exceeds the kernel parameter space");

__global__ void pack_adapters_many_kernel(const __grid_constant__ PackMany pm) {
  int lo = 0, hi = pm.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pm.blk0[mid] <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid;
  }
  const int blk = blockIdx.x - pm.blk0[lo], n_blk = pm.blk0[lo + 1] - pm.blk0[lo];
  pack_adapters_body(pm.job[lo], static_cast<long>(blk) * blockDim.x + threadIdx.x, static_cast<long>(n_blk) * blockDim.x);
}

__global__ void cast_transpose_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb,
                                      __nv_bfloat16* __restrict__ wt, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = w[static_cast<long>(r) * cols + c];
      if (wb) wb[static_cast<long>(r) * cols + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (wt) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int c = c0 + i, r = r0 + threadIdx.x;
      if (r < rows && c < cols) wt[static_cast<long>(c) * rows + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
  }
}

int grid_for(long work, int block) {
  long g = (work + block - 1) / block;
  const long cap = sm_count() * 16L;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int launch_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, void* y_drop, long drop_rows,
                         float* mean, float* rstd, long rows, int C, float eps, int merge, int H, int W, float drop_p,
                         uint64_t drop_seed, cudaStream_t stream) {
  MTL_REQUIRE(rows > 0 && C > 0, "layernorm: empty input");
  MTL_REQUIRE(C % 8 == 0 && (!merge || C % 32 == 0), "layernorm: C=%d must be a multiple of 8 (32 when merging)", C);
  MTL_REQUIRE(!merge || (H % 2 == 0 && W % 2 == 0), "patch merging: x size (%d*%d) are not even.", H, W);
  LnFwdParams p{static_cast<const __nv_bfloat16*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y),
                static_cast<__nv_bfloat16*>(y_drop), mean, rstd, rows, drop_rows, C, merge, H, W, eps, drop_p, drop_seed};
  int lpr = 0, vpt = 0;
  if (ln_fast_shape(C, &lpr, &vpt)) {
    const long row_groups = (rows + (32 / lpr) - 1) / (32 / lpr);
    long grid = (row_groups + 7) / 8;
    if (grid > sm_count() * 16L) grid = sm_count() * 16L;
    const unsigned g = static_cast<unsigned>(grid);
    MTL_LN_CASES(ln_fwd_launch, p, g, stream)
  } else {
    const int wpb = 8;
    layernorm_fwd_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(p);
  }
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         const void* dres, void* dx, float* dgamma, float* dbeta, long rows, int C, int merge,
                         int H, int W, cudaStream_t stream) {
  MTL_REQUIRE(rows > 0 && C > 0, "layernorm bwd: empty input");
  MTL_REQUIRE(C % 8 == 0 && C <= 4096, "layernorm bwd: C=%d unsupported", C);
  MTL_REQUIRE(!(merge && dres), "layernorm bwd: residual gradient not supported together with merge");
  LnBwdParams p{static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), gamma, mean, rstd,
                static_cast<const __nv_bfloat16*>(dres), static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, rows, C,
                merge, H, W, 0};
  int lpr = 0, vpt = 0;
  if (ln_fast_shape(C, &lpr, &vpt)) {
    const long row_groups = (rows + (32 / lpr) - 1) / (32 / lpr);
    long grid = (row_groups + 7) / 8;
    if (grid > sm_count() * 4L) grid = sm_count() * 4L;
    const unsigned g = static_cast<unsigned>(grid);
    MTL_LN_CASES(ln_bwd_launch, p, g, stream)
    note_launch();
    MTL_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  long ctas = sm_count() * 4L;
  long rpc = (rows + ctas - 1) / ctas;
  if (rpc < 8) rpc = 8;
  p.rows_per_cta = static_cast<int>(rpc);
  const unsigned grid = static_cast<unsigned>((rows + rpc - 1) / rpc);
  if (C <= 1024) layernorm_bwd_kernel<true><<<grid, 256, 2 * C * sizeof(float), stream>>>(p);
  else layernorm_bwd_kernel<false><<<grid, 256, 2 * C * sizeof(float), stream>>>(p);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_window_gather(const void* in, void* out, int B, int H, int W, int C, int shift, int ws,
                                int elem_size, int scatter, cudaStream_t stream) {
  MTL_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "window_process: empty input");
  MTL_REQUIRE(ws > 0 && H % ws == 0 && W % ws == 0, "window_process: H,W (%d,%d) not divisible by window %d", H, W, ws);
  MTL_REQUIRE(elem_size == 2 || elem_size == 4, "window_process: element size %d unsupported", elem_size);
  const long row_bytes = static_cast<long>(C) * elem_size;
  const bool v16 = row_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(in) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (v16) {
    const int CV = static_cast<int>(row_bytes / 16);
    const long total = static_cast<long>(B) * H * W * CV;
    window_gather_kernel<uint4><<<grid_for(total, 256), 256, 0, stream>>>(
        static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, CV, shift, ws, scatter); note_launch();
  } else if (elem_size == 4) {
    const long total = static_cast<long>(B) * H * W * C;
    window_gather_kernel<uint32_t><<<grid_for(total, 256), 256, 0, stream>>>(
        static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), B, H, W, C, shift, ws, scatter); note_launch();
  } else {
    const long total = static_cast<long>(B) * H * W * C;
    window_gather_kernel<uint16_t><<<grid_for(total, 256), 256, 0, stream>>>(
        static_cast<const uint16_t*>(in), static_cast<uint16_t*>(out), B, H, W, C, shift, ws, scatter); note_launch();
  }
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// roll(-shift) + partition (inverse = its adjoint: un-partition + roll(+shift))
int launch_roll_partition(const void* in, void* out, int B, int H, int W, int C, int shift, int ws, int elem_size,
                          int inverse, cudaStream_t stream) {
  return launch_window_gather(in, out, B, H, W, C, shift, ws, elem_size, inverse, stream);
}
// window_reverse + roll(+shift) (inverse = its adjoint: roll(-shift) + partition)
int launch_merge_roll(const void* in, void* out, int B, int H, int W, int C, int shift, int ws, int elem_size,
                      int inverse, cudaStream_t stream) {
  return launch_window_gather(in, out, B, H, W, C, shift, ws, elem_size, !inverse, stream);
}

int launch_dropout(const void* x, void* y, long n, float p, uint64_t seed, cudaStream_t stream) {
  MTL_REQUIRE(n % 8 == 0, "dropout: element count %ld must be a multiple of 8", n);
  MTL_REQUIRE(p >= 0.f && p < 1.f, "dropout probability has to be in [0, 1), but got %f", p);
  if (n == 0) return 0;
  dropout_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                          static_cast<__nv_bfloat16*>(y), n / 8, p, seed); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_scale_rows(const void* x, const float* scale, void* y, int S, long M, int C, int rows_per_sample,
                      cudaStream_t stream) {
  MTL_REQUIRE(C % 8 == 0 && rows_per_sample > 0 && M % rows_per_sample == 0, "scale_rows: bad shape");
  const long total = static_cast<long>(S) * M * (C / 8);
  if (total == 0) return 0;
  scale_rows_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), scale,
                                                             static_cast<__nv_bfloat16*>(y), S, M, C / 8,
                                                             rows_per_sample, static_cast<int>(M / rows_per_sample)); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_scale_rows_sum(const void* x, const float* scale, void* y, int S, long M, int C, int rows_per_sample,
                          cudaStream_t stream) {
  MTL_REQUIRE(C % 8 == 0 && S >= 1, "scale_rows_sum: bad shape");
  MTL_REQUIRE(scale == nullptr || (rows_per_sample > 0 && M % rows_per_sample == 0), "scale_rows_sum: bad rows_per_sample");
  const long per = M * (C / 8);
  if (per == 0) return 0;
  const int rps = rows_per_sample > 0 ? rows_per_sample : 1;
  scale_rows_sum_kernel<<<grid_for(per, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), scale,
                                                               static_cast<__nv_bfloat16*>(y), S, M, C / 8, rps,
                                                               static_cast<int>(M / rps)); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_add(const void* a, const void* b, void* out, long n, cudaStream_t stream) {
  MTL_REQUIRE(n % 8 == 0, "add: element count %ld must be a multiple of 8", n);
  if (n == 0) return 0;
  add_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(a),
                                                      static_cast<const __nv_bfloat16*>(b),
                                                      static_cast<__nv_bfloat16*>(out), n / 8); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_sum_streams(const void* x, const void* extra, void* out, int S, long n, cudaStream_t stream) {
  MTL_REQUIRE(n % 8 == 0, "sum_streams: element count %ld must be a multiple of 8", n);
  if (n == 0) return 0;
  sum_streams_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                              static_cast<const __nv_bfloat16*>(extra),
                                                              static_cast<__nv_bfloat16*>(out), S, n / 8); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_adapters(const float* const* a_ptrs, const float* const* b_ptrs, const int* ranks, const int* offs,
                         int n_adapt, int K, int N, int R_pad, void* a_cat, void* b_cat, void* a_cat_t, void* b_cat_t,
                         cudaStream_t stream) {
  MTL_REQUIRE(n_adapt >= 1 && n_adapt <= 8, "pack: adapter count %d out of range", n_adapt);
  PackParams p;
  for (int i = 0; i < 8; ++i) {
    p.a[i] = i < n_adapt ? a_ptrs[i] : nullptr;
    p.b[i] = i < n_adapt ? b_ptrs[i] : nullptr;
    p.rank[i] = i < n_adapt ? ranks[i] : 0;
    p.off[i] = i < n_adapt ? offs[i] : 0;
    if (i < n_adapt) MTL_REQUIRE(offs[i] + ranks[i] <= R_pad, "pack: adapter %d exceeds R_pad", i);
  }
  p.n = n_adapt; p.K = K; p.N = N; p.R_pad = R_pad;
  p.a_cat = static_cast<__nv_bfloat16*>(a_cat);
  p.b_cat = static_cast<__nv_bfloat16*>(b_cat);
  p.a_cat_t = static_cast<__nv_bfloat16*>(a_cat_t);
  p.b_cat_t = static_cast<__nv_bfloat16*>(b_cat_t);
  const long total = static_cast<long>(R_pad) * (K + N);
  pack_adapters_kernel<<<grid_for(total, 256), 256, 0, stream>>>(p); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_adapters_many(const PackJobHost* jobs, int n_jobs, cudaStream_t stream) {
  for (int j0 = 0; j0 < n_jobs; j0 += kPackManyMax) {
    static thread_local PackMany pm;   // 15 KiB: keep it off the stack
    pm.n = n_jobs - j0 < kPackManyMax ? n_jobs - j0 : kPackManyMax;
    int blocks = 0;
    for (int j = 0; j < pm.n; ++j) {
      const PackJobHost& h = jobs[j0 + j];
      MTL_REQUIRE(h.n_adapt >= 1 && h.n_adapt <= 8, "pack_many: job %d: adapter count %d out of range", j0 + j, h.n_adapt);
      PackParams& p = pm.job[j];
      for (int i = 0; i < 8; ++i) {
        p.a[i] = i < h.n_adapt ? h.a[i] : nullptr;
        p.b[i] = i < h.n_adapt ? h.b[i] : nullptr;
        p.rank[i] = i < h.n_adapt ? h.rank[i] : 0;
        p.off[i] = i < h.n_adapt ? h.off[i] : 0;
        if (i < h.n_adapt) MTL_REQUIRE(h.off[i] + h.rank[i] <= h.R_pad, "pack_many: job %d: adapter %d exceeds R_pad", j0 + j, i);
      }
      p.n = h.n_adapt; p.K = h.K; p.N = h.N; p.R_pad = h.R_pad;
      p.a_cat = static_cast<__nv_bfloat16*>(h.a_cat);
      p.b_cat = static_cast<__nv_bfloat16*>(h.b_cat);
      p.a_cat_t = static_cast<__nv_bfloat16*>(h.a_cat_t);
      p.b_cat_t = static_cast<__nv_bfloat16*>(h.b_cat_t);
      const long total = static_cast<long>(h.R_pad) * (h.K + h.N);
      long nb = (total + 1023) / 1024;   // ~4 elements per thread
      if (nb > 128) nb = 128;
      pm.blk0[j] = blocks;
      blocks += static_cast<int>(nb);
    }
    pm.blk0[pm.n] = blocks;
    pack_adapters_many_kernel<<<blocks, 256, 0, stream>>>(pm); note_launch();
    MTL_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int launch_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int rows, int cols, cudaStream_t stream) {
  MTL_REQUIRE(rows > 0 && cols > 0, "cast_transpose: empty matrix");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  cast_transpose_kernel<<<grid, block, 0, stream>>>(w, static_cast<__nv_bfloat16*>(w_bf16),
                                                   static_cast<__nv_bfloat16*>(wt_bf16), rows, cols); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
