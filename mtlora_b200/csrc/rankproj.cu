// Rank-space projection of a single-adapter MTLoRALinear as a launch of its own:
//     out[M, R] = scale * X[M, K] . Down[R, K]^T        (bf16 in / out, fp32 accumulation)
// forward  (reference models/lora.py:260, the `x @ A^T` half of the shared update): X = D(x), Down = A_cat
// backward (its adjoint on the output side):                                        X = dy,   Down = B_cat^T
// Used for the compute-bound layers of stages 2-3 (LinearSpec.pre_project): the main tcgen05 kernel then runs as one
// dense product over the concatenated contraction [x | U] . [W | B]^T and never waits for a TMEM -> smem conversion.
//
// The product is skinny (R <= 128 columns) and memory-bound: 19 MB of X for M = 25088, K = 384. A persistent tcgen05
// kernel spends most of such a launch in prologue / pipeline fill (measured 17-21 us through mtl_linear_kernel); here
// many small CTAs (64 rows x R columns, 4 warps, mma.sync m16n8k16, 3-stage cp.async ring, 3-4 CTAs per SM) keep the
// whole machine busy for the few microseconds the data needs to arrive.
#include "common.cuh"
#include "kernels.cuh"

namespace mtl {
namespace {

constexpr int kBK = 64;        // contraction elements per stage (one 128-byte smem row)
constexpr int kStages = 3;
constexpr int kThreads = 128;

// byte offset of 16-byte chunk `c` of row `r` in a [rows][64] bf16 tile with an XOR swizzle (conflict-free ldmatrix)
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

// ROWS = 64: warp w owns rows 16w..16w+15 and all R columns. ROWS = 32 (few row tiles: M = 25088 gives only 392 CTAs of
// 64 rows for 148 SMs): warp w owns rows 16(w & 1).. and the column half (w >> 1), twice as many CTAs in flight.
template <int R, int kRows>
__global__ void __launch_bounds__(kThreads, kRows == 64 ? 3 : 5)
rank_project_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ down,
                    __nv_bfloat16* __restrict__ out, int M, int K, float scale) {
  constexpr int kStageBytes = (kRows + R) * 128;
  constexpr int CW = kRows == 64 ? R : R / 2;        // columns per warp
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = (smem_u32(smem) + 127u) & ~127u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kRows;
  const int n_kb = (K + kBK - 1) / kBK;

  auto load_stage = [&](int kb, int s) {
    const uint32_t xs = sbase + s * kStageBytes, ds = xs + kRows * 128;
    const int k0 = kb * kBK;
    // X tile: 64 rows x 8 chunks; Down tile: R rows x 8 chunks
    for (int i = threadIdx.x; i < (kRows + R) * 8; i += kThreads) {
      const int r = i >> 3, c = i & 7;
      const bool kin = k0 + c * 8 < K;
      if (r < kRows) {
        const bool ok = kin && (m0 + r) < M;
        const __nv_bfloat16* src = x + static_cast<size_t>(ok ? m0 + r : 0) * K + (ok ? k0 + c * 8 : 0);
        cp_async_16_zfill(xs + tile_off(r, c), src, ok);
      } else {
        const int rr = r - kRows;
        const __nv_bfloat16* src = down + static_cast<size_t>(rr) * K + (kin ? k0 + c * 8 : 0);
        cp_async_16_zfill(ds + tile_off(rr, c), src, kin);
      }
    }
  };

  const int wrow = kRows == 64 ? warp * 16 : (warp & 1) * 16;
  const int wcol = kRows == 64 ? 0 : (warp >> 1) * CW;
  float acc[CW / 8][4];
#pragma unroll
  for (int j = 0; j < CW / 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

  for (int s = 0; s < kStages - 1; ++s) {
    if (s < n_kb) load_stage(s, s);
    cp_async_commit();
  }
  for (int kb = 0; kb < n_kb; ++kb) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    if (kb + kStages - 1 < n_kb) load_stage(kb + kStages - 1, (kb + kStages - 1) % kStages);
    cp_async_commit();
    const uint32_t xs = sbase + (kb % kStages) * kStageBytes, ds = xs + kRows * 128;
#pragma unroll
    for (int kk = 0; kk < kBK / 16; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, xs + tile_off(wrow + (lane & 15), kk * 2 + (lane >> 4)));
#pragma unroll
      for (int jp = 0; jp < CW / 16; ++jp) {
        uint32_t b[4];
        ldmatrix_x4(b, ds + tile_off(wcol + jp * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)));
        const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
        mma_bf16_16816(acc[2 * jp], a, b0);
        mma_bf16_16816(acc[2 * jp + 1], a, b1);
      }
    }
  }
  cp_async_wait<0>();
  // epilogue: scale, round, store (row g and g + 8 of the warp's 16 rows; columns 2t, 2t + 1 of every 8-column tile)
  const int g = lane >> 2, t = lane & 3;
  const int r0 = m0 + wrow + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < CW / 8; ++j) {
    const int col = wcol + j * 8 + 2 * t;
    if (r0 < M) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r0) * R + col) = pack_bf16x2(acc[j][0] * scale, acc[j][1] * scale);
    if (r1 < M) *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r1) * R + col) = pack_bf16x2(acc[j][2] * scale, acc[j][3] * scale);
  }
}

template <int R, int kRows>
int launch_rr(const void* x, const void* down, void* out, int M, int K, float scale, cudaStream_t stream) {
  constexpr int smem = kStages * (kRows + R) * 128 + 128;
  static bool attr_done[64] = {};
  MTL_CHECK_CUDA(ensure_max_dyn_smem(attr_done, rank_project_kernel<R, kRows>, smem));
  rank_project_kernel<R, kRows><<<(M + kRows - 1) / kRows, kThreads, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(down), static_cast<__nv_bfloat16*>(out), M,
      K, scale);
  note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int R>
int launch_r(const void* x, const void* down, void* out, int M, int K, float scale, cudaStream_t stream) {
  // (a 32-row variant — twice the CTAs, each warp half the columns — measured 5-7 % slower at M = 25088: the Down tile is
  // then re-read by twice as many CTAs and every warp issues fewer MMAs per ldmatrix)
  return launch_rr<R, 64>(x, down, out, M, K, scale, stream);
}

}  // namespace

bool rank_project_supported(int R, int K) { return (R == 16 || R == 32 || R == 64 || R == 128) && K % 8 == 0; }

int launch_rank_project(const void* x, const void* down, void* out, int M, int K, int R, float scale,
                        cudaStream_t stream) {
  switch (R) {
    case 16: return launch_r<16>(x, down, out, M, K, scale, stream);
    case 32: return launch_r<32>(x, down, out, M, K, scale, stream);
    case 64: return launch_r<64>(x, down, out, M, K, scale, stream);
    case 128: return launch_r<128>(x, down, out, M, K, scale, stream);
    default: set_error("rank_project: rank space %d unsupported by the skinny kernel", R); return 1;
  }
}

}  // namespace mtl
