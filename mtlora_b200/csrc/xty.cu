// Parameter-gradient reduction C[a, b] += alpha * sum_m rs(m) * P[m, a] * Q[m, b] — the LEGACY mma.sync path. The hot
// path is xty_sm100.cu (tcgen05, MN-major operands straight from TMA boxes); this kernel remains for the cases it does
// not take: alpha != 1, GELU recomputation on load (x_gelu) and operands narrower than 72 columns.
//
// This is the only reduction over the (huge) token dimension M in the MTLoRA backward:
//   dB_s = dY_s^T U_s   (N x r),  dA_s = G_s^T X_s   (r x K),  dW_reduction = dY^T X   (PatchMerging, trainable)
// (reference: autograd of models/lora.py:260-265). The outputs are tiny and the inputs stream once, so the
// kernel is HBM-bound; mma.sync m16n8k16 tiles with both operands loaded transposed (ldmatrix.trans) from
// row-major [m][.] shared tiles, split over M across CTAs with fp32 atomics into C.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {

namespace {

constexpr int TA = 64, TB = 64, TM = 64;
constexpr int kChunks = TM * 8 / 128;   // 16-byte chunks of P (and of Q) each thread stages per step
constexpr int LDS = 72;  // smem row stride (bf16): 144 B, conflict-free for ldmatrix

struct XtyParams {
  const __nv_bfloat16* P;
  const __nv_bfloat16* Q;
  float* C;
  const float* rowscale;
  long ldp, ldq, ldc, M;
  int a, b, rows_per_sample, q_gelu;
  long m_per_cta;
  float alpha;
};

__device__ __forceinline__ uint4 load_chunk(const __nv_bfloat16* base, long ld, long m, long M, int col, int ncols) {
  if (m < M && col < ncols) return __ldg(reinterpret_cast<const uint4*>(base + m * ld + col));
  return make_uint4(0, 0, 0, 0);
}
__device__ __forceinline__ uint4 xform_chunk(uint4 q, float s, bool gelu) {
  uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float lo = bf16lo_to_f32(w[e]), hi = bf16hi_to_f32(w[e]);
    if (gelu) { lo = gelu_exact(lo); hi = gelu_exact(hi); }
    w[e] = pack_bf16x2(lo * s, hi * s);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(128) xty_kernel(const XtyParams p) {
  __shared__ __align__(16) __nv_bfloat16 Ps[TM * LDS];
  __shared__ __align__(16) __nv_bfloat16 Qs[TM * LDS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = blockIdx.x * TA, b0 = blockIdx.y * TB;
  const long m_begin = static_cast<long>(blockIdx.z) * p.m_per_cta;
  long m_end = m_begin + p.m_per_cta;
  if (m_end > p.M) m_end = p.M;

  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;

  // each thread stages kChunks chunks of P and of Q per step: chunk id = tid + 128*i -> row = id / 8, col8 = id % 8
  uint4 rp[kChunks], rq[kChunks];
  auto fetch = [&](long m0) {
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const int id = threadIdx.x + 128 * i;
      const int r = id >> 3, c8 = (id & 7) * 8;
      rp[i] = load_chunk(p.P, p.ldp, m0 + r, m_end, a0 + c8, p.a);
      rq[i] = load_chunk(p.Q, p.ldq, m0 + r, m_end, b0 + c8, p.b);
      if (p.rowscale != nullptr || p.q_gelu) {
        float s = 1.f;
        if (p.rowscale != nullptr && m0 + r < m_end) s = p.rowscale[(m0 + r) / p.rows_per_sample];
        rq[i] = xform_chunk(rq[i], s, p.q_gelu != 0);
      }
    }
  };
  if (m_begin < m_end) fetch(m_begin);
  for (long m0 = m_begin; m0 < m_end; m0 += TM) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const int id = threadIdx.x + 128 * i;
      const int r = id >> 3, c8 = (id & 7) * 8;
      *reinterpret_cast<uint4*>(Ps + r * LDS + c8) = rp[i];
      *reinterpret_cast<uint4*>(Qs + r * LDS + c8) = rq[i];
    }
    __syncthreads();
    if (m0 + TM < m_end) fetch(m0 + TM);
#pragma unroll
    for (int kk = 0; kk < TM / 16; ++kk) {
      uint32_t af[4];
      const int mi = lane >> 3;
      ldmatrix_x4_trans(af, smem_u32(Ps + (kk * 16 + (lane & 7) + (mi >> 1) * 8) * LDS + warp * 16 + (mi & 1) * 8));
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, smem_u32(Qs + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8));
        const uint32_t bA[2] = {bf[0], bf[1]}, bB[2] = {bf[2], bf[3]};
        mma_bf16_16816(acc[np * 2], af, bA);
        mma_bf16_16816(acc[np * 2 + 1], af, bB);
      }
    }
  }
  const int g4 = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ar = a0 + warp * 16 + g4 + (e >> 1) * 8;
      const int bc = b0 + nt * 8 + t4 * 2 + (e & 1);
      if (ar < p.a && bc < p.b && acc[nt][e] != 0.f) atomicAdd(p.C + static_cast<long>(ar) * p.ldc + bc, p.alpha * acc[nt][e]);
    }
  }
}

}  // namespace

int launch_xty(const void* P, long ldp, const void* Q, long ldq, float* C, long ldc, long M, int a, int b,
               const float* rowscale, int rows_per_sample, int q_gelu, float alpha, cudaStream_t stream) {
  MTL_REQUIRE(M > 0 && a > 0 && b > 0, "xty: empty problem");
  MTL_REQUIRE(a % 8 == 0 && b % 8 == 0 && ldp % 8 == 0 && ldq % 8 == 0, "xty: dims must be multiples of 8");
  MTL_REQUIRE((reinterpret_cast<uintptr_t>(P) % 16 == 0) && (reinterpret_cast<uintptr_t>(Q) % 16 == 0),
              "xty: operands must be 16-byte aligned");
  XtyParams p;
  p.P = static_cast<const __nv_bfloat16*>(P);
  p.Q = static_cast<const __nv_bfloat16*>(Q);
  p.C = C;
  p.rowscale = rowscale;
  p.ldp = ldp; p.ldq = ldq; p.ldc = ldc; p.M = M;
  p.a = a; p.b = b;
  p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : 1;
  p.q_gelu = q_gelu;
  p.alpha = alpha;
  const int at = (a + TA - 1) / TA, bt = (b + TB - 1) / TB;
  long splits = (sm_count() * 4L + at * bt - 1) / (at * bt);
  const long max_splits = (M + 4 * TM - 1) / (4 * TM);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long mpc = (M + splits - 1) / splits;
  mpc = (mpc + TM - 1) / TM * TM;
  p.m_per_cta = mpc;
  const unsigned gz = static_cast<unsigned>((M + mpc - 1) / mpc);
  xty_kernel<<<dim3(at, bt, gz), 128, 0, stream>>>(p); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
