// Internal launcher declarations (C++ linkage) shared by the C ABI layer in api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mtlora_b200.h"

namespace mtl {

// attention.cu ------------------------------------------------------------------------------------
int launch_win_attn_fwd(const void* qkv, const float* rpb, const float* mask, int n_mask, void* out, void* out_drop,
                        float* lse, int B, int H, int W, int C, int nH, int ws, int shift, float scale,
                        float drop_p, uint64_t drop_seed, cudaStream_t stream);
int launch_win_attn_bwd(const void* qkv, const void* dout, const float* rpb, const float* mask, int n_mask,
                        const float* lse, void* dqkv, float* drpb, int B, int H, int W, int C, int nH, int ws,
                        int shift, float scale, cudaStream_t stream);

// attention_sm100.cu ------------------------------------------------------------------------------
// tcgen05 / TMEM forward (two windows per 128-row tile, whole-row gathers); same contract as launch_win_attn_fwd without
// an explicit mask
bool win_attn_fwd_umma_supported(int C, int nH, int ws);
int launch_win_attn_fwd_umma(const void* qkv, const float* rpb, void* out, void* out_drop, float* lse, int B, int H,
                             int W, int C, int nH, int ws, int shift, float scale, float drop_p, uint64_t drop_seed,
                             cudaStream_t stream);

// rowwise.cu --------------------------------------------------------------------------------------
// LayerNorm over the last dim of [rows, C] (bf16 in/out, fp32 statistics and affine parameters).
// merge != 0: input is a (B, H, W, C/4) token grid and each output row is the PatchMerging 2x2 gather
// (reference swin_transformer_mtlora.py:462-466) of 4 source rows, so C = 4 * C_src.
int launch_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, void* y_drop, long drop_rows,
                         float* mean, float* rstd, long rows, int C, float eps, int merge, int H, int W, float drop_p,
                         uint64_t drop_seed, cudaStream_t stream);
// dx (= LN backward, optionally + dres) ; dgamma/dbeta accumulated with atomics into fp32 buffers.
int launch_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         const void* dres, void* dx, float* dgamma, float* dbeta, long rows, int C, int merge,
                         int H, int W, cudaStream_t stream);
// kernels/window_process equivalents (bitwise gathers), elem_size in {2, 4}
int launch_roll_partition(const void* in, void* out, int B, int H, int W, int C, int shift, int ws, int elem_size,
                          int inverse, cudaStream_t stream);
int launch_merge_roll(const void* in, void* out, int B, int H, int W, int C, int shift, int ws, int elem_size,
                      int inverse, cudaStream_t stream);
// y = x * mask / (1-p) with the counter-based mask shared by every kernel (index = flat element index)
int launch_dropout(const void* x, void* y, long n, float p, uint64_t seed, cudaStream_t stream);
// y[s, m, :] = x[s, m, :] * scale[s, m / rows_per_sample]   (DropPath gradient pre-scale)
int launch_scale_rows(const void* x, const float* scale, void* y, int S, long M, int C, int rows_per_sample,
                      cudaStream_t stream);
// y[s] = x[s] * scale[s, sample] for s < S (scale null: copy, skipped when y == x) and y[S] = sum_s y[s]
int launch_scale_rows_sum(const void* x, const float* scale, void* y, int S, long M, int C, int rows_per_sample,
                          cudaStream_t stream);
// out[i] = sum_{s<S} x[s, i] (+ extra[i] when extra != null), bf16 in/out, fp32 accumulation
int launch_sum_streams(const void* x, const void* extra, void* out, int S, long n, cudaStream_t stream);
// out = a + b (bf16)
int launch_add(const void* a, const void* b, void* out, long n, cudaStream_t stream);
// pack fp32 adapter parameters into the bf16 operand layouts of the fused linear kernel
int launch_pack_adapters(const float* const* a_ptrs, const float* const* b_ptrs, const int* ranks, const int* offs,
                         int n_adapt, int K, int N, int R_pad, void* a_cat, void* b_cat, void* a_cat_t, void* b_cat_t,
                         cudaStream_t stream);
// One launch for the packed operands of many layers (mtl_linear_pack_many).
struct PackJobHost {
  const float* a[8];
  const float* b[8];
  int rank[8], off[8];
  int n_adapt, K, N, R_pad;
  void *a_cat, *b_cat, *a_cat_t, *b_cat_t;
};
int launch_pack_adapters_many(const PackJobHost* jobs, int n_jobs, cudaStream_t stream);
// fp32 [rows, cols] -> bf16 copy and/or bf16 transposed copy
int launch_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int rows, int cols, cudaStream_t stream);

// patch_embed.cu ----------------------------------------------------------------------------------
// tokens = LayerNorm(Conv2d(k4, s4)(x) + bias): x fp32 [B, 3, H, W] -> proj (pre-norm, bf16), y (bf16), patches (im2col,
// bf16 [B*L, 48]), mean / rstd; proj / patches / mean / rstd / gamma may be null.
int launch_patch_embed_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                           void* proj, void* y, void* patches, float* mean, float* rstd, int B, int H, int W, int E,
                           float eps, cudaStream_t stream);

// xty.cu ------------------------------------------------------------------------------------------
// C[a, b] += sum_m rowscale(m) * P[m, a] * Q[m, b]   (fp32 atomics; P, Q bf16 row-major with leading dims)
// q_gelu != 0 applies exact GELU to Q elements on load (recomputing the fc2 input from the saved fc1 output).
int launch_xty(const void* P, long ldp, const void* Q, long ldq, float* C, long ldc, long M, int a, int b,
               const float* rowscale, int rows_per_sample, int q_gelu, float alpha, cudaStream_t stream);

// xty_sm100.cu ------------------------------------------------------------------------------------
// out[w, r] += sum_m Wide[m, w] * (rowscale(m) * Rank[m, r]) for a list of job groups in ONE launch (tcgen05, both
// operands consumed MN-major straight from their row-major layout). Operands are bf16 [streams][M][pitch].
struct XtyOperand {
  const void* base;   // null: unused slot
  int cols;           // valid columns
  long pitch;         // elements between rows (0: cols)
  int streams;        // wide operands: >= 1 (3rd TMA coordinate); rank operands: ignored
};
struct XtyJobGroup {
  int wide_op, wide_stream;    // which wide operand (0/1) and stream
  int rank_op, r0, rlen;       // which rank operand (0/1), its column range [r0, r0 + rlen)
  int out_rank_major, out_ld;  // 0: out[w * ld + r], 1: out[r * ld + w]   (r = absolute rank column)
  float* out;
  const float* rowscale;       // per-sample scale of the contraction rows (DropPath), or null
};
int launch_xty_groups(const XtyOperand* wide, const XtyOperand* rank, const XtyJobGroup* groups, int n_groups,
                      long M, int rows_per_sample, cudaStream_t stream);

// rankproj.cu -------------------------------------------------------------------------------------
// out[M, R] = scale * X[M, K] . Down[R, K]^T (bf16, fp32 accumulation) for R in {16, 32, 64, 128}
bool rank_project_supported(int R, int K);
int launch_rank_project(const void* x, const void* down, void* out, int M, int K, int R, float scale,
                        cudaStream_t stream);

// optim.cu ----------------------------------------------------------------------------------------
int opt_sqnorm(const mtl_opt_seg* segs, const int32_t* prefix, int n_segs, int n_chunks, float* out,
               cudaStream_t stream);
int opt_adamw(const mtl_opt_seg* segs, const int32_t* prefix, int n_segs, int n_chunks, float* flat_m, float* flat_v,
              float* state, const mtl_opt_group* groups, int n_groups, const float* grad_scale, const float* found_inf,
              const float* sqnorm, float max_norm, int adam_w, cudaStream_t stream);

}  // namespace mtl
