// Internal interface of the fused multi-stream low-rank-extended GEMM (tcgen05 / TMEM / TMA).
//
// One kernel serves MTLoRALinear forward (reference models/lora.py:253-284) and its input-gradient
// backward. In kernel terms (all operands bf16, K-major, fp32 accumulation in TMEM):
//
//   U[:, seg_a]  = scale_a * X[in(a)] . Down[seg_a, :]^T          "down" products (rank space)
//   P            = sum_{i in main} X[i] . Wm^T                       dense frozen product
//   Y[j]         = (useP_j ? P : 0) + U[:, ranges_j] . Up[:, ranges_j]^T   (+ epilogue)
//
// forward : X = activations, Wm = W (N,K), Down = A_cat (R,K), Up = B_cat (N,R)
// backward: X = dY streams,  Wm = W^T (K,N), Down = B_cat^T (R,N), Up = A_cat^T (K,R)
#pragma once
#include "common.cuh"

namespace mtl {

constexpr int LIN_BM = 128;        // rows per CTA tile (UMMA M)
constexpr int LIN_BK = 64;         // bf16 per 128-byte swizzle row
constexpr int LIN_MAX_STREAMS = 8; // 1 shared + up to 7 task streams
constexpr int LIN_MAX_GROUPS = 12; // phase-1 load groups (adapters, split to <= 128 rank rows)
constexpr int LIN_MAX_GRAN = 20;   // 16-column granules of the rank space (R_pad <= 320)

enum LinEpilogue : int {
  LIN_EP_NONE = 0,
  LIN_EP_GELU_DUAL = 1,  // y = v, y2 = gelu(v)                (fc1 forward, Mlp.forward :69-75)
  LIN_EP_GELU_BWD = 2,   // y = v * gelu'(aux)                 (fc2 input-gradient -> d(fc1 out))
  LIN_EP_GELU_DUAL_GRAD = 3,  // y = gelu'(v), y2 = gelu(v)     (fc1 forward keeping the derivative factor instead of v)
  LIN_EP_MUL_AUX = 4,    // y = v * aux                        (fc2 input-gradient, aux = gelu'(fc1 out) saved by mode 3)
};

struct LinPlan {
  int M, Kc, Nn;
  int S_in, S_out;
  int R_pad;     // rank-space width (multiple of 16; 0 = dense only)
  int BN;        // output columns per chunk (multiple of 32, <= 128)
  int n_chunks;  // ceil(Nn / BN)
  int n_splits;  // CTAs sharing one 128-row tile (each takes chunks split, split+n_splits, ...)
  int n_stages;  // TMA ring depth
  int stage_b_bytes;
  int n_regions;  // 1: dense + adapters merged in one accumulator per item; > 1: dense P + per-stream delta D
  int n_pbuf;     // dense accumulator buffers P (multi mode): 2, or 1 when the rank space leaves no room
  int n_dbuf;     // item accumulators per epilogue group (2, or 1 when TMEM is short)
  int d_shared;   // multi mode, one output stream: a single delta accumulator shared by both epilogue groups
  int n_work;     // work items = row tiles x column splits (walked by persistent CTAs)
  int n_slabs;    // per-epilogue-warp store slabs in shared memory
  int up_pack;    // Up tiles ([BN x 64] per rank atom) per ring stage: 1, or 2-3 when the rank space is wide
  int acc_col0;   // first accumulator column in TMEM
  int tmem_cols;  // allocation size (power of two >= 32)

  // phase 1 ("down") load groups: X[grp_in] . Down[grp_r0 : grp_r0+grp_len]^T
  int n_groups;
  int grp_in[LIN_MAX_GROUPS], grp_r0[LIN_MAX_GROUPS], grp_len[LIN_MAX_GROUPS];
  int grp_acc[LIN_MAX_GROUPS];  // 1: accumulate onto the columns an earlier group already produced
  // per 16-column granule of U: scale and which input stream it came from (for row scaling)
  float gran_scale[LIN_MAX_GRAN];
  int gran_in[LIN_MAX_GRAN];

  int n_main;
  int main_in[LIN_MAX_STREAMS];

  int out_useP[LIN_MAX_STREAMS];
  int out_r0[LIN_MAX_STREAMS][2], out_len[LIN_MAX_STREAMS][2];

  // epilogue
  int ep_mode;
  const float* bias;          // [Nn] or null
  __nv_bfloat16* y;           // [S_out, M, Nn]
  __nv_bfloat16* y2;          // [S_out, M, Nn] (GELU_DUAL)
  const __nv_bfloat16* aux;   // [S_out, M, Nn] (GELU_BWD)
  const __nv_bfloat16* res;   // [res_streams, M, Nn] residual added last, or null
  int res_streams;            // 1 (shared by all outputs) or S_out
  int in_streams;             // streams of the epilogue-input tensor (aux / res) the producer prefetches into L2
  const float* rowscale_out;  // [S_out, n_samples] multiplies the accumulator (DropPath), or null
  const float* rowscale_in;   // [S_in, n_samples] multiplies U rows, or null
  int rows_per_sample, n_samples;
  __nv_bfloat16* u_save;      // [M, R_pad] scaled U written once (split 0), or null

  // LoRA dropout (models/lora.py:258), counter-based mask over the flat index of the [M, Nn] matrix:
  //   drop_mode 1 (forward, GELU_DUAL): also write D(gelu(v)) of stream 0 as stream S_out of y2
  //   drop_mode 2 (backward): the adapter part of output 0 is multiplied by mask / (1 - p)
  int drop_mode;
  float drop_p;
  uint64_t drop_seed;
  unsigned long long* trace;  // debug timeline of CTA 0 (MTL_LINEAR_TRACE), else null: [role][1024] (time, code) pairs
  uint32_t wait_hint_ns;  // mbarrier.try_wait suspend-time hint used by every role's waits
  int force_split;  // keep dense and adapter accumulators in separate TMEM regions even if S_out == 1
  // u_in = 1: the rank-space operand U [M, R_pad] (bf16, already scaled) was produced by an earlier launch
  // (mtl_linear_rank_project) and lives at `u_save`: phase 1 and the TMEM -> smem conversion are skipped, the U atoms
  // travel through the TMA ring next to the Up tiles like one more k-block of the contraction. Single-adapter layers.
  int u_in;
  float out_scale;  // multiplies the accumulator (+ bias) before the residual; 0 is read as 1
};

// TMA descriptor of a bf16 row-major [d2][d1][d0] tensor (d0 contiguous, `pitch` elements between rows, 0 = dense),
// box = (b0, b1, 1). Shared by the linear and the adapter-gradient kernels.
int make_tmap(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
              CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint64_t pitch = 0);

// Host side, no CUDA calls: validates `plan` and fills its derived tiling fields (chunk width, TMEM / shared-memory
// budget, column splits for a grid of `n_sm` persistent CTAs). Returns 0 and the dynamic shared-memory size, or -1
// with mtl_last_error() set when the problem does not fit the kernel.
int plan_linear(LinPlan& plan, int n_sm, uint32_t* smem_bytes);

// Host side: plan_linear for the current device, tensor maps, launch.
// x: [S_in, M, Kc], wm: [Nn, Kc], down: [R_pad, Kc], up: [Nn, R_pad]; all bf16 row-major.
int launch_linear(LinPlan plan, const void* x, const void* wm, const void* down, const void* up,
                  cudaStream_t stream);

}  // namespace mtl
