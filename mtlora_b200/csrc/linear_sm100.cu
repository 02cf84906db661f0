// Fused MTLoRALinear GEMM for sm_100a: frozen dense product + (1+T) low-rank adapters in one kernel.
//
// Reference semantics: models/lora.py:253-284 (MTLoRALinear.forward) — the reference evaluates
//   F.linear(x, W, b), 2(1+T) skinny matmuls and (1+T) mul/add passes as separate kernels.
// Here each 128-row tile of X is staged by TMA, the rank-space products U = X.A_cat^T are formed by
// tcgen05.mma into TMEM, re-staged as a bf16 K-major operand, and replayed against B_cat so that
// every output stream is written exactly once; the frozen W tile is read once per tile and feeds all
// streams (one dense accumulator P shared by the stream epilogues).
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-7 = U converter + epilogue (each owns 32 TMEM lanes = 32 tile rows).
#include "linear_sm100.cuh"

#include <cudaTypedefs.h>
#include <mutex>

namespace mtl {

namespace {

constexpr int kThreads = 256;
constexpr int kStageABytes = LIN_BM * LIN_BK * 2;  // 16 KiB

struct SmemLayout {
  uint32_t stages;   // n_stages * (A + B)
  uint32_t usm;      // n_uatoms * 16 KiB
  uint32_t bars;     // mbarriers
  uint32_t total;
};

__host__ __device__ inline SmemLayout smem_layout(int n_stages, int stage_b_bytes, int r_pad) {
  SmemLayout l;
  const uint32_t n_uatoms = (r_pad + 63) / 64;
  l.stages = 0;
  l.usm = n_stages * (kStageABytes + stage_b_bytes);
  l.bars = l.usm + n_uatoms * kStageABytes;
  l.total = l.bars + 256;
  return l;
}

__device__ __forceinline__ bool col_in_out(const LinPlan& p, int j, int col) {
  return (col >= p.out_r0[j][0] && col < p.out_r0[j][0] + p.out_len[j][0]) ||
         (col >= p.out_r0[j][1] && col < p.out_r0[j][1] + p.out_len[j][1]);
}

__global__ void __launch_bounds__(kThreads)
mtl_linear_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_down,
                  const __grid_constant__ CUtensorMap tm_up, const __grid_constant__ LinPlan p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const SmemLayout L = smem_layout(p.n_stages, p.stage_b_bytes, p.R_pad);
  const uint32_t stage_bytes = kStageABytes + p.stage_b_bytes;
  const uint32_t usm_base = smem_base + L.usm;
  const uint32_t bar_base = smem_base + L.bars;
  // barrier slots (8 bytes each)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t u_full = bar_base + 8u * 16;
  const uint32_t u_ready = bar_base + 8u * 17;
  auto acc_full = [&](int b) { return bar_base + 8u * (18 + b); };
  auto acc_empty = [&](int b) { return bar_base + 8u * (20 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 22;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L.bars + 8u * 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tile = blockIdx.x / p.n_splits;
  const int split = blockIdx.x % p.n_splits;
  const int m0 = m_tile * LIN_BM;
  const int n_kb = (p.Kc + LIN_BK - 1) / LIN_BK;
  const int n_uatoms = (p.R_pad + 63) / 64;
  const int n_my_chunks = (p.n_chunks - split + p.n_splits - 1) / p.n_splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    if (p.R_pad > 0) {
      tma_prefetch_desc(&tm_down);
      tma_prefetch_desc(&tm_up);
    }
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(u_full, 1);
    mbar_init(u_ready, 128);
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 128);
    }
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

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      // phase 1: rank-space ("down") products
      for (int g = 0; g < p.n_groups; ++g) {
        const int len = p.grp_len[g];
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t a_dst = smem_base + stage * stage_bytes;
          const uint32_t b_dst = a_dst + kStageABytes;
          mbar_arrive_expect_tx(full_bar(stage), kStageABytes + len * 128);
          tma_load_3d(a_dst, &tm_x, full_bar(stage), kb * LIN_BK, m0, p.grp_in[g]);
          for (int i = 0; i < len / 16; ++i)
            tma_load_2d(b_dst + i * 2048, &tm_down, full_bar(stage), kb * LIN_BK, p.grp_r0[g] + 16 * i);
          advance();
        }
      }
      // phase 2: dense product + adapter replay per output chunk
      for (int ci = 0; ci < n_my_chunks; ++ci) {
        const int c = split + ci * p.n_splits;
        for (int i = 0; i < p.n_main; ++i) {
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            const uint32_t b_dst = a_dst + kStageABytes;
            mbar_arrive_expect_tx(full_bar(stage), kStageABytes + p.BN * 128);
            tma_load_3d(a_dst, &tm_x, full_bar(stage), kb * LIN_BK, m0, p.main_in[i]);
            tma_load_2d(b_dst, &tm_w, full_bar(stage), kb * LIN_BK, c * p.BN);
            advance();
          }
        }
        for (int a = 0; a < n_uatoms; ++a) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t b_dst = smem_base + stage * stage_bytes + kStageABytes;
          mbar_arrive_expect_tx(full_bar(stage), p.BN * 128);
          tma_load_2d(b_dst, &tm_up, full_bar(stage), a * 64, c * p.BN);
          advance();
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      };
      for (int g = 0; g < p.n_groups; ++g) {
        const uint32_t idesc = umma_idesc_bf16_m128(p.grp_len[g]);
        const uint32_t d_tmem = tmem_base + p.grp_r0[g];
        for (int kb = 0; kb < n_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_src = smem_base + stage * stage_bytes;
          const uint64_t adesc = umma_desc_sw128(a_src);
          const uint64_t bdesc = umma_desc_sw128(a_src + kStageABytes);
          for (int q = 0; q < 4; ++q) {
            if (kb * LIN_BK + q * 16 >= p.Kc) break;
            umma_bf16(d_tmem, adesc + 2 * q, bdesc + 2 * q, idesc, ((kb | q) != 0 || p.grp_acc[g]) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          advance();
        }
      }
      if (p.R_pad > 0) umma_commit(u_full);

      const uint32_t idesc_bn = umma_idesc_bf16_m128(p.BN);
      bool u_waited = false;
      for (int ci = 0; ci < n_my_chunks; ++ci) {
        const int b = ci % p.n_acc;
        const uint32_t use = ci / p.n_acc;
        mbar_wait(acc_empty(b), (use & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + p.acc_col0 + b * p.n_regions * p.BN;
        bool first = true;
        for (int i = 0; i < p.n_main; ++i) {
          for (int kb = 0; kb < n_kb; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_src = smem_base + stage * stage_bytes;
            const uint64_t adesc = umma_desc_sw128(a_src);
            const uint64_t bdesc = umma_desc_sw128(a_src + kStageABytes);
            for (int q = 0; q < 4; ++q) {
              if (kb * LIN_BK + q * 16 >= p.Kc) break;
              umma_bf16(acc0, adesc + 2 * q, bdesc + 2 * q, idesc_bn, first ? 0u : 1u);
              first = false;
            }
            umma_commit(empty_bar(stage));
            advance();
          }
        }
        uint32_t region_started = 0;  // bit r: region r already holds a partial sum
        if (!first) region_started |= 1u;
        for (int a = 0; a < n_uatoms; ++a) {
          if (!u_waited) {
            mbar_wait(u_ready, 0);
            u_waited = true;
          }
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(usm_base + a * kStageABytes);
          const uint64_t bdesc = umma_desc_sw128(smem_base + stage * stage_bytes + kStageABytes);
          for (int q = 0; q < 4; ++q) {
            const int col = a * 64 + q * 16;
            if (col >= p.R_pad) break;
            for (int j = 0; j < p.S_out; ++j) {
              if (!col_in_out(p, j, col)) continue;
              const int region = (p.n_regions == 1) ? 0 : 1 + j;
              umma_bf16(acc0 + region * p.BN, adesc + 2 * q, bdesc + 2 * q, idesc_bn,
                        (region_started >> region) & 1u);
              region_started |= 1u << region;
            }
          }
          umma_commit(empty_bar(stage));
          advance();
        }
        umma_commit(acc_full(b));
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =============================== U converter + epilogue ================================
    const int w = warp - 4;
    const int row = w * 32 + lane;
    const int grow = m0 + row;
    const bool row_ok = grow < p.M;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    const int sample = (p.rows_per_sample > 0) ? min(grow / p.rows_per_sample, p.n_samples - 1) : 0;

    if (p.R_pad > 0) {
      mbar_wait(u_full, 0);
      tc_fence_after();
      for (int gq = 0; gq < p.R_pad / 16; ++gq) {
        uint32_t r[16];
        tmem_ld16(t_lane + gq * 16, r);
        tmem_ld_wait();
        float s = p.gran_scale[gq];
        if (p.rowscale_in != nullptr) s *= p.rowscale_in[p.gran_in[gq] * p.n_samples + sample];
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]) * s, __uint_as_float(r[2 * i + 1]) * s);
        uint8_t* atom = smem_gen + L.usm + (gq >> 2) * kStageABytes;
        const uint32_t c0 = (gq & 3) * 16;
        *reinterpret_cast<uint4*>(atom + sw128_offset(row, c0)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(atom + sw128_offset(row, c0 + 8)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (p.u_save != nullptr && split == 0 && row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.u_save + static_cast<size_t>(grow) * p.R_pad + gq * 16);
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(u_ready);
    }

    const size_t stream_stride = static_cast<size_t>(p.M) * p.Nn;
    for (int ci = 0; ci < n_my_chunks; ++ci) {
      const int c = split + ci * p.n_splits;
      const int b = ci % p.n_acc;
      const uint32_t use = ci / p.n_acc;
      mbar_wait(acc_full(b), use & 1u);
      tc_fence_after();
      const uint32_t acc0 = t_lane + p.acc_col0 + b * p.n_regions * p.BN;
      for (int gq = 0; gq < p.BN / 16; ++gq) {
        const int n0 = c * p.BN + gq * 16;
        if (n0 >= p.Nn) break;
        uint32_t pr[16];
        bool have_p = false;
        if (p.n_regions > 1 && p.n_main > 0) {
          tmem_ld16(acc0 + gq * 16, pr);
          have_p = true;
        }
        float bias_v[16];
        if (p.bias != nullptr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + i);
            bias_v[4 * i + 0] = bv.x; bias_v[4 * i + 1] = bv.y;
            bias_v[4 * i + 2] = bv.z; bias_v[4 * i + 3] = bv.w;
          }
        }
        for (int j = 0; j < p.S_out; ++j) {
          float v[16];
          const bool has_delta = (p.out_len[j][0] + p.out_len[j][1]) > 0;
          if (p.n_regions == 1) {
            uint32_t dr[16];
            tmem_ld16(acc0 + gq * 16, dr);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(dr[i]);
          } else {
            uint32_t dr[16];
            if (has_delta) tmem_ld16(acc0 + (1 + j) * p.BN + gq * 16, dr);
            tmem_ld_wait();
            const bool mask_delta = (p.drop_mode == 2) && j == 0;
            const uint32_t thr = dropout_threshold(p.drop_p);
            const float keep_scale = 1.f / (1.f - p.drop_p);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float t = 0.f;
              if (have_p && p.out_useP[j]) t = __uint_as_float(pr[i]);
              if (has_delta) {
                float d = __uint_as_float(dr[i]);
                if (mask_delta)
                  d = dropout_hash(p.drop_seed, static_cast<uint64_t>(grow) * p.Nn + n0 + i) >= thr ? d * keep_scale : 0.f;
                t += d;
              }
              v[i] = t;
            }
          }
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += bias_v[i];
          }
          if (p.rowscale_out != nullptr) {
            const float rs = p.rowscale_out[j * p.n_samples + sample];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= rs;
          }
          if (row_ok) {
            const size_t off = j * stream_stride + static_cast<size_t>(grow) * p.Nn + n0;
            if (p.ep_mode == LIN_EP_GELU_BWD) {
              const uint4* ap = reinterpret_cast<const uint4*>(p.aux + off);
              const uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
              const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[2 * i] *= gelu_exact_grad(bf16lo_to_f32(aw[i]));
                v[2 * i + 1] *= gelu_exact_grad(bf16hi_to_f32(aw[i]));
              }
            }
            if (p.res != nullptr) {
              const size_t roff = (p.res_streams == 1 ? 0 : j * stream_stride) +
                                  static_cast<size_t>(grow) * p.Nn + n0;
              const uint4* rp = reinterpret_cast<const uint4*>(p.res + roff);
              const uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
              const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[2 * i] += bf16lo_to_f32(rw[i]);
                v[2 * i + 1] += bf16hi_to_f32(rw[i]);
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            uint4* dst = reinterpret_cast<uint4*>(p.y + off);
            dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (p.ep_mode == LIN_EP_GELU_DUAL) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                pk[i] = pack_bf16x2(gelu_exact(v[2 * i]), gelu_exact(v[2 * i + 1]));
              uint4* dst2 = reinterpret_cast<uint4*>(p.y2 + off);
              dst2[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst2[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              if (p.drop_mode == 1 && j == 0) {
                // D(m): derived from the bf16-rounded activation so it equals dropout(y2, seed + 1) exactly; the next
                // layer (fc2) is called with dropout_seed + 1 so that its mask is independent of this layer's input mask
                const uint32_t thr = dropout_threshold(p.drop_p);
                const float keep_scale = 1.f / (1.f - p.drop_p);
                const uint64_t e0 = static_cast<uint64_t>(grow) * p.Nn + n0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float lo = dropout_hash(p.drop_seed + 1, e0 + 2 * i) >= thr ? bf16lo_to_f32(pk[i]) * keep_scale : 0.f;
                  const float hi = dropout_hash(p.drop_seed + 1, e0 + 2 * i + 1) >= thr ? bf16hi_to_f32(pk[i]) * keep_scale : 0.f;
                  pk[i] = pack_bf16x2(lo, hi);
                }
                uint4* dst3 = reinterpret_cast<uint4*>(p.y2 + static_cast<size_t>(p.S_out) * stream_stride +
                                                       static_cast<size_t>(grow) * p.Nn + n0);
                dst3[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst3[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
            }
          }
        }
        if (have_p) tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(acc_empty(b));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
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

// bf16 row-major [d2][d1][d0] tensor (d0 contiguous), box = (b0, b1, 1), 128-byte swizzle.
int make_tmap(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0,
              uint32_t b1) {
  auto fn = get_encode_fn();
  MTL_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  MTL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer not 16-byte aligned");
  MTL_REQUIRE((d0 * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (inner dim %llu)",
              (unsigned long long)d0);
  const int rank = d2 > 0 ? 3 : 2;
  cuuint64_t dims[3] = {d0, d1, d2 > 0 ? d2 : 1};
  cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MTL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu,%llu,%llu box %u,%u)",
              (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, b0, b1);
  return 0;
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

int launch_linear(LinPlan p, const void* x, const void* wm, const void* down, const void* up,
                  cudaStream_t stream) {
  MTL_REQUIRE(p.M > 0 && p.Kc > 0 && p.Nn > 0, "linear: empty problem (M=%d K=%d N=%d)", p.M, p.Kc, p.Nn);
  MTL_REQUIRE(p.Kc % 16 == 0 && p.Nn % 16 == 0, "linear: K (%d) and N (%d) must be multiples of 16", p.Kc, p.Nn);
  MTL_REQUIRE(p.R_pad % 16 == 0 && p.R_pad <= 16 * LIN_MAX_GRAN, "linear: rank space %d unsupported (<= %d)",
              p.R_pad, 16 * LIN_MAX_GRAN);
  MTL_REQUIRE(p.S_in >= 1 && p.S_in <= LIN_MAX_STREAMS && p.S_out >= 1 && p.S_out <= LIN_MAX_STREAMS,
              "linear: stream counts out of range");
  MTL_REQUIRE(p.n_main > 0 || p.R_pad > 0, "linear: nothing to compute");

  // ---- tiling -------------------------------------------------------------------------------
  const int u_cols = round_up(p.R_pad, 32);
  p.n_regions = (p.S_out == 1 && !p.force_split && p.drop_mode != 2) ? 1 : 1 + p.S_out;
  if (p.R_pad == 0 || p.n_main == 0) p.n_regions = (p.S_out == 1) ? 1 : 1 + p.S_out;
  const int budget = 512 - u_cols;
  const int n_round = round_up(p.Nn, 32);
  auto fit_bn = [&](int n_acc) {
    int bn = budget / (p.n_regions * n_acc) / 32 * 32;
    if (bn > 128) bn = 128;
    if (bn > n_round) bn = n_round;
    return bn;
  };
  int bn = fit_bn(2);
  p.n_acc = 2;
  if (bn < 64 && fit_bn(1) > bn) {
    bn = fit_bn(1);
    p.n_acc = 1;
  }
  MTL_REQUIRE(bn >= 32, "linear: TMEM budget exceeded (R_pad=%d, S_out=%d)", p.R_pad, p.S_out);
  // even out the chunks: smallest multiple of 32 that keeps the chunk count
  const int chunks = (p.Nn + bn - 1) / bn;
  bn = round_up((p.Nn + chunks - 1) / chunks, 32);
  p.BN = bn;
  p.n_chunks = (p.Nn + bn - 1) / bn;
  if (p.n_chunks == 1) p.n_acc = 1;
  p.acc_col0 = u_cols;
  const int need_cols = u_cols + p.n_acc * p.n_regions * p.BN;
  p.tmem_cols = 32;
  while (p.tmem_cols < need_cols) p.tmem_cols *= 2;
  MTL_REQUIRE(p.tmem_cols <= 512, "linear: TMEM columns %d > 512", need_cols);

  int max_len = p.BN;
  for (int g = 0; g < p.n_groups; ++g) {
    MTL_REQUIRE(p.grp_len[g] % 16 == 0 && p.grp_len[g] >= 16 && p.grp_len[g] <= 128,
                "linear: down group length %d invalid", p.grp_len[g]);
    if (p.grp_len[g] > max_len) max_len = p.grp_len[g];
  }
  p.stage_b_bytes = round_up(max_len * 128, 1024);

  const int m_tiles = (p.M + LIN_BM - 1) / LIN_BM;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  p.n_splits = 1;
  if (m_tiles < 2 * n_sm) {
    p.n_splits = (2 * n_sm + m_tiles - 1) / m_tiles;
    if (p.n_splits > p.n_chunks) p.n_splits = p.n_chunks;
  }

  // ring depth: as many stages as fit; keep <= ~100 KiB when TMEM allows two CTAs per SM
  const int smem_cap = (p.tmem_cols <= 256) ? 110 * 1024 : 220 * 1024;
  p.n_stages = 8;
  while (p.n_stages > 2 && smem_layout(p.n_stages, p.stage_b_bytes, p.R_pad).total + 1024 > (uint32_t)smem_cap)
    --p.n_stages;
  const uint32_t smem_bytes = smem_layout(p.n_stages, p.stage_b_bytes, p.R_pad).total + 1024;
  MTL_REQUIRE(smem_bytes <= 227 * 1024, "linear: shared memory %u exceeds 227 KiB", smem_bytes);

  // ---- tensor maps ---------------------------------------------------------------------------
  CUtensorMap tm_x, tm_w, tm_down, tm_up;
  if (int e = make_tmap(&tm_x, x, p.Kc, p.M, p.S_in, LIN_BK, LIN_BM)) return e;
  if (int e = make_tmap(&tm_w, wm, p.Kc, p.Nn, 0, LIN_BK, p.BN)) return e;
  if (p.R_pad > 0) {
    if (int e = make_tmap(&tm_down, down, p.Kc, p.R_pad, 0, LIN_BK, 16)) return e;
    if (int e = make_tmap(&tm_up, up, p.R_pad, p.Nn, 0, LIN_BK, p.BN)) return e;
  } else {
    tm_down = tm_w;
    tm_up = tm_w;
  }

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, []() {
    attr_err = cudaFuncSetAttribute(mtl_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  MTL_CHECK_CUDA(attr_err);

  const dim3 grid(m_tiles * p.n_splits);
  mtl_linear_kernel<<<grid, kThreads, smem_bytes, stream>>>(tm_x, tm_w, tm_down, tm_up, p); note_launch();
  MTL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace mtl
