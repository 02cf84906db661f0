/*
 * mtlora_b200 — C ABI of the B200 (sm_100a) hot path of scale-lab/MTLoRA.
 *
 * The reference has no C ABI for this path: it is Python (torch.nn modules) plus one pybind/ATen extension
 * (kernels/window_process/swin_window_process.cpp:127-132). Each entry point below names the reference
 * interface it replaces (paths relative to the reference checkout). All pointers are DEVICE pointers owned by
 * the caller (torch caching allocator in the Python host layer), every call is asynchronous on `stream`, nothing
 * is allocated or freed on the device inside. Process-wide state, all of it behind a mutex or atomic: a thread-local
 * error string, a launch counter, per-device "shared-memory attribute set" flags and a cache of TMA descriptors keyed
 * by (base pointer, dims, box, swizzle) — a descriptor only encodes addresses and strides, so it cannot go stale.
 *
 * Conventions
 *   - activations are bf16, row-major, "stream-stacked": [S, M, C] where stream 0 is the task-shared stream and
 *     streams 1..T are the per-task streams in the module's task order; M = B*H*W tokens in (B, H, W) order.
 *   - parameters handed to compute calls are bf16 operand copies produced by the *_pack / mtl_cast_transpose
 *     calls from the fp32 master parameters; gradients are accumulated (+=) into fp32 buffers.
 *   - return value: 0 = ok, 1 = invalid argument / unsupported shape, 2 = CUDA error; mtl_last_error() explains.
 */
#ifndef MTLORA_B200_H_
#define MTLORA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTL_ABI_VERSION 3
#define MTL_MAX_TASKS 7

typedef void* mtl_stream_t; /* cudaStream_t */

/* MTLoRALinear shared_mode: 'matrix' (models/lora.py:259-266: task outputs = pretrained + task adapter) and 'matrixv2'
 * (:267-274: task outputs additionally carry the shared adapter's update). */
enum { MTL_MODE_MATRIX = 0, MTL_MODE_MATRIXV2 = 1 };
enum { MTL_ACT_NONE = 0, MTL_ACT_GELU = 1, MTL_ACT_GELU_GRAD = 2 };

int mtl_abi_version(void);
/* sizeof(mtl_linear_cfg) as compiled into the library: a binding checks its own struct mirror against it. */
int mtl_linear_cfg_size(void);
const char* mtl_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all threads); bench.py reports the delta over
 * its timed region as `gpu_launches`. */
uint64_t mtl_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * MTLoRALinear  — models/lora.py:159-284
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct mtl_linear_cfg {
  int64_t M;              /* rows (tokens) per stream */
  int32_t in_features;    /* K */
  int32_t out_features;   /* N */
  int32_t n_tasks;        /* T = len(tasks), 0 when the layer was built with tasks=None */
  int32_t x_tasks_given;  /* forward(x, x_tasks) got x_tasks: input is [1+T, M, K] (lora.py:263) */
  int32_t shared_mode;    /* MTL_MODE_MATRIX or MTL_MODE_MATRIXV2 */
  int32_t r_shared;       /* r['shared']; 0 = no adapters (lora.py:256-257, or CompatLinear) */
  int32_t r_task[MTL_MAX_TASKS];
  float scale_shared;     /* lora_shared_scale */
  float scale_task[MTL_MAX_TASKS];
  float dropout_p;        /* lora_dropout in training, else 0 */
  uint64_t dropout_seed;  /* counter-based mask seed; same value in forward and backward */
  int32_t rows_per_sample;/* L = H*W: rows of one image, for per-sample DropPath scales (0 = unused) */
  int32_t gelu_aux_is_grad;/* mtl_linear_bwd_input: gelu_aux already holds GELU'(pre-activation) (the producing layer ran
                            * with MTL_ACT_GELU_GRAD) -> dx *= gelu_aux instead of dx *= GELU'(gelu_aux) */
  int32_t dy_has_sum;     /* mtl_linear_bwd_input, layers with task streams: dy holds 1+T+1 streams, the last one being
                           * sum_j dy[j] (mtl_scale_rows_sum) — the frozen product then reads ONE stream per column
                           * chunk instead of re-summing the 1+T streams on the tensor cores */
  int32_t u_precomputed;  /* mtl_linear_fwd / mtl_linear_bwd_input on a layer WITHOUT task adapters: the rank-space
                           * activations (u_save resp. g_save) are an INPUT, produced by mtl_linear_rank_project; the
                           * kernel then runs as one dense product over the concatenated contraction [x | U].[W | B]^T */
} mtl_linear_cfg;

/* Width R of the packed rank space: every adapter (shared first, then the tasks in module order) starts on a
 * 16-column boundary and is zero-padded to a multiple of 16; R is their sum (<= 320). Returns -1 on error. */
int mtl_linear_rank_pad(const mtl_linear_cfg* cfg);
/* Column offset of adapter `idx` (0 = shared, 1+t = task t) inside the packed rank space. */
int mtl_linear_rank_offset(const mtl_linear_cfg* cfg, int idx);

/* Pack fp32 master adapters (lora_shared_A [r_s,K], lora_shared_B [N,r_s], lora_tasks_A[t] [r_t,K],
 * lora_tasks_B[t] [N,r_t]; lora.py:200-225) into the bf16 operands used by forward (a_cat [R,K], b_cat [N,R])
 * and backward (a_cat_t [K,R], b_cat_t [R,N]). Any output may be NULL. a_tasks/b_tasks: host arrays of T
 * device pointers. */
int mtl_linear_pack(const mtl_linear_cfg* cfg, const float* a_shared, const float* b_shared,
                    const float* const* a_tasks, const float* const* b_tasks, void* a_cat, void* b_cat,
                    void* a_cat_t, void* b_cat_t, mtl_stream_t stream);

/* mtl_linear_pack for many layers in ONE launch: the adapters of every layer change once per optimizer step, and ~50
 * launches of a few microseconds each add up (reference: none — models/lora.py:253-284 reads the fp32 parameters directly;
 * this is the operand staging of the bf16 kernels). Any of the four outputs of a job may be NULL. */
typedef struct mtl_pack_job {
  mtl_linear_cfg cfg;                       /* in_features, out_features, n_tasks, r_shared, r_task */
  const float* a_shared;                    /* [r_shared, in] */
  const float* b_shared;                    /* [out, r_shared] */
  const float* a_tasks[MTL_MAX_TASKS];      /* [r_task[t], in] */
  const float* b_tasks[MTL_MAX_TASKS];      /* [out, r_task[t]] */
  void* a_cat;                              /* bf16 [R, in] */
  void* b_cat;                              /* bf16 [out, R] */
  void* a_cat_t;                            /* bf16 [in, R] */
  void* b_cat_t;                            /* bf16 [R, out] */
} mtl_pack_job;
int mtl_pack_job_size(void);
int mtl_linear_pack_many(const mtl_pack_job* jobs, int32_t n_jobs, mtl_stream_t stream);

/* Rank-space projection of a layer without task adapters (lora.py:260: the `x @ A^T` half of the shared update) as a
 * launch of its own, for the compute-bound layers of stages 2-3:
 *   pass 0 (forward) : u[M, R] = s_sh * x[drop stream] . a_cat^T     x: [1 (+1 if dropout_p > 0), M, K], down = a_cat [R, K]
 *   pass 1 (backward): g[M, R] = s_sh * dy . b_cat_t^T               x: dy [1, M, N],                   down = b_cat_t [R, N]
 * The result is what mtl_linear_fwd would save as u_save (mtl_linear_bwd_input: g_save); hand it back to those calls
 * with cfg.u_precomputed = 1. */
int mtl_linear_rank_project(const mtl_linear_cfg* cfg, int32_t pass, const void* x, const void* down, void* u_out,
                            mtl_stream_t stream);

/* fp32 [rows, cols] -> bf16 copy (w_bf16, may be NULL) and bf16 transpose [cols, rows] (wt_bf16, may be NULL);
 * used once per frozen nn.Linear weight (lora.py:194). */
int mtl_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int32_t rows, int32_t cols, mtl_stream_t stream);

/* Forward (lora.py:253-284, mode 'matrix'):
 *   pre  = x[0] W^T + b
 *   y[0] = pre + s_sh * D(x[0]) A_sh^T B_sh^T
 *   y[t] = pre + s_t  * (x_tasks_given ? x[t] : D(x[0])) A_t^T B_t^T          t = 1..T
 *   (MTL_MODE_MATRIXV2: y[t] additionally + s_sh * D(x[0]) A_sh^T B_sh^T; in backward the shared adapter then receives
 *    the gradient of every stream: G_sh = s_sh (sum_j dy[j]) B_sh, dB_sh = (sum_j dy[j])^T U_sh)
 * x:     [S_in, M, K], S_in = 1 (+T if x_tasks_given) (+1 if dropout_p > 0: D(x[0]) appended as last stream,
 *        produced by the upstream kernel with the same seed)
 * y:     [1+T, M, N]  (T = 0 -> a single stream)
 * act == MTL_ACT_GELU (Mlp.forward swin_transformer_mtlora.py:69-75): y keeps the pre-activation (needed by
 *        backward), y_act [1+T (+1 if dropout_p > 0), M, N] receives GELU(y) and, last, D(GELU(y[0])) drawn with
 *        dropout_seed + 1 (the seed the consuming fc2 layer must be called with).
 * act == MTL_ACT_GELU_GRAD: as MTL_ACT_GELU, but y receives GELU'(pre-activation) = Phi(y) + y pdf(y) (evaluated on
 *        the fp32 accumulator) — all the backward of the consuming layer needs (cfg.gelu_aux_is_grad = 1), which
 *        turns its epilogue into one multiply per element.
 * residual/res_streams/path_scale (SwinTransformerBlock.forward :389-392,398-408): when residual != NULL,
 *        y[j] = residual[res_streams == 1 ? 0 : j] + path_scale[j, sample] * (above); path_scale may be NULL (=1),
 *        layout [1+T, M / rows_per_sample] fp32 (DropPath keep-mask / keep-prob, independent per stream).
 * u_save: [M, R] bf16, optional: scaled rank-space activations kept for mtl_linear_bwd_params. */
int mtl_linear_fwd(const mtl_linear_cfg* cfg, const void* x, const void* w_bf16, const float* bias,
                   const void* a_cat, const void* b_cat, int32_t act, void* y, void* y_act, const void* residual,
                   int32_t res_streams, const float* path_scale, void* u_save, mtl_stream_t stream);

/* Input gradient (autograd of lora.py:253-284; W is frozen so there is no dW):
 *   dPre  = sum_j ps[j] dy[j]
 *   dx[0] = dPre W + mask/keep * ( G_sh A_sh (+ sum_t G_t A_t if !x_tasks_given) ),  G_s = s_s * ps[s] dy[s] B_s
 *   dx[t] = G_t A_t                                      (only if x_tasks_given)
 * dy:       [1+T, M, N];  dx: [1 (+T if x_tasks_given), M, K]
 * gelu_aux: optional, same shape as dx: dx *= GELU'(gelu_aux)  (the producer of x was the fused GELU epilogue)
 * path_scale: optional per-sample scale of dy rows (DropPath backward); supported in-kernel only for a layer
 *        with a single output stream — with task streams pre-scale dy with mtl_scale_rows and pass NULL here and
 *        to mtl_linear_bwd_params.
 * g_save:   [M, R] bf16, optional: G (without path_scale) kept for mtl_linear_bwd_params. */
int mtl_linear_bwd_input(const mtl_linear_cfg* cfg, const void* dy, const void* wt_bf16, const void* a_cat_t,
                         const void* b_cat_t, void* dx, const void* gelu_aux, const float* path_scale,
                         void* g_save, mtl_stream_t stream);

/* Tiling the planner of the fused linear kernel picks for one launch (test / tuning aid; runs without a GPU).
 * pass 0 = mtl_linear_fwd (act = MTL_ACT_*, has_residual = a residual is added), pass 1 = mtl_linear_bwd_input
 * (act != MTL_ACT_NONE = a gelu_aux operand is given). n_sm = size of the persistent grid to balance for.
 * Returns -1 (mtl_last_error) when the configuration does not fit the kernel's TMEM / shared-memory budget. */
typedef struct mtl_linear_plan_info {
  int32_t bn;             /* output columns per chunk */
  int32_t n_chunks;       /* ceil(N / bn) */
  int32_t n_splits;       /* work items per 128-row tile */
  int32_t n_stages;       /* TMA ring depth */
  int32_t n_slabs;        /* store slabs per epilogue warp */
  int32_t n_regions;      /* 1 = dense + adapters in one accumulator; else 1 + output streams */
  int32_t n_pbuf, n_dbuf; /* dense / per-group item accumulators */
  int32_t d_shared;       /* one delta accumulator shared by both epilogue groups */
  int32_t tmem_cols;      /* TMEM allocation (power of two <= 512) */
  int32_t tmem_cols_used;
  int32_t smem_bytes;     /* dynamic shared memory of the launch (<= 227 KiB) */
  int32_t n_work;         /* work items walked by the persistent CTAs */
  int32_t n_groups;       /* rank-space load groups */
  int32_t s_in, s_out;    /* operand / result streams of the launch */
  int32_t up_pack;        /* Up tiles per ring stage (1; 2-3 when the packed rank space is wider than 128) */
} mtl_linear_plan_info;
int mtl_linear_plan(const mtl_linear_cfg* cfg, int32_t pass, int32_t act, int32_t has_residual, int32_t n_sm,
                    mtl_linear_plan_info* out);

/* Adapter gradients, accumulated into packed fp32 buffers da_cat [R, K] and db_cat [N, R]:
 *   dB_s += dy[s]^T (ps[s] U_s),   dA_s += G_s^T (ps[s] x_in(s))   (x_in(s): the stream adapter s consumed in forward;
 *   ps = path_scale [1+T, M / rows_per_sample] or NULL)
 * x is the forward input (same layout incl. the dropped copy); x_gelu != 0 applies GELU to x on load (x is then
 * the saved fc1 pre-activation, i.e. the fc2 input is recomputed instead of stored). */
int mtl_linear_bwd_params(const mtl_linear_cfg* cfg, const void* x, int32_t x_gelu, const void* dy,
                          const void* u_save, const void* g_save, const float* path_scale, float* da_cat,
                          float* db_cat, mtl_stream_t stream);

/* C[a, b] += alpha * sum_m P[m, a] Q[m, b] (bf16 in, fp32 accumulate): dW of a trainable dense weight, e.g.
 * PatchMerging.reduction when MTLORA.DOWNSAMPLER_ENABLED is False (lora.py:599-600). */
int mtl_xty(const void* p, int64_t ldp, const void* q, int64_t ldq, float* c, int64_t ldc, int64_t M, int32_t a,
            int32_t b, float alpha, mtl_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * WindowAttention core + window shift/partition/reverse — swin_transformer_mtlora.py:194-220, :338-377
 * qkv: [B, H, W, 3C] (output of the qkv MTLoRALinear in token order), out: [B, H, W, C].
 * rpb: relative_position_bias_table [(2ws-1)^2, nH] fp32 (:139-141). mask: optional explicit additive mask
 * [n_mask, ws*ws, ws*ws] fp32 (the reference's attn_mask buffer :297-322); when NULL and shift > 0 the same
 * 0 / -100 mask is evaluated analytically from the window position. lse: [B*nW, nH, 64] fp32, saved for backward.
 * out_drop (optional): D(out) for the LoRA dropout of the following proj layer.
 * ---------------------------------------------------------------------------------------------------------- */
int mtl_window_attention_fwd(const void* qkv, const float* rpb, const float* mask, int32_t n_mask, void* out,
                             void* out_drop, float* lse, int32_t B, int32_t H, int32_t W, int32_t C,
                             int32_t num_heads, int32_t window_size, int32_t shift_size, float scale,
                             float dropout_p, uint64_t dropout_seed, mtl_stream_t stream);
/* dqkv: [B, H, W, 3C]; drpb: fp32 [(2ws-1)^2, nH], accumulated (+=), may be NULL. */
int mtl_window_attention_bwd(const void* qkv, const void* dout, const float* rpb, const float* mask,
                             int32_t n_mask, const float* lse, void* dqkv, float* drpb, int32_t B, int32_t H,
                             int32_t W, int32_t C, int32_t num_heads, int32_t window_size, int32_t shift_size,
                             float scale, mtl_stream_t stream);

/* kernels/window_process (swin_window_process.cpp:64-125), same argument meaning incl. the sign of shift_size
 * (WindowProcess passes -shift, WindowProcessReverse +shift; swin_transformer_mtlora.py:344-345,374-375).
 * elem_size: 2 (fp16/bf16) or 4 (fp32); results are bitwise copies like the reference kernels. */
int mtl_roll_and_window_partition_forward(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                          int32_t shift_size, int32_t window_size, int32_t elem_size,
                                          mtl_stream_t stream);
int mtl_roll_and_window_partition_backward(const void* grad_in, void* grad_out, int32_t B, int32_t H, int32_t W,
                                           int32_t C, int32_t shift_size, int32_t window_size, int32_t elem_size,
                                           mtl_stream_t stream);
int mtl_window_merge_and_roll_forward(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                      int32_t shift_size, int32_t window_size, int32_t elem_size,
                                      mtl_stream_t stream);
int mtl_window_merge_and_roll_backward(const void* grad_in, void* grad_out, int32_t B, int32_t H, int32_t W,
                                       int32_t C, int32_t shift_size, int32_t window_size, int32_t elem_size,
                                       mtl_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * LayerNorm (norm1/norm2 swin_transformer_mtlora.py:332,396; PatchMerging.norm :469) and PatchMerging gather
 * (:462-467). merge != 0: x is a [*, H, W, C/4] token grid and each output row is the 2x2 gather of 4 source
 * rows in the reference channel order; rows counts OUTPUT rows. y_drop: optional [drop_rows, C] D(y) of the first
 * drop_rows rows (LoRA dropout copy of the shared stream of a stream-stacked input).
 * ---------------------------------------------------------------------------------------------------------- */
int mtl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, void* y_drop,
                      int64_t drop_rows, float* mean, float* rstd, int64_t rows, int32_t C, float eps, int32_t merge,
                      int32_t H, int32_t W, float dropout_p, uint64_t dropout_seed, mtl_stream_t stream);
/* dx = LN'(dy) (+ dres, the gradient arriving over the residual connection); dgamma/dbeta accumulated (+=). */
int mtl_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                      int32_t merge, int32_t H, int32_t W, mtl_stream_t stream);

/* PatchEmbed.forward (swin_transformer_mtlora.py:597-605) for patch_size 4, in_chans 3, embed_dim 96 / 128:
 * y = LayerNorm(Conv2d(3, E, k=4, s=4)(x).flatten(2).transpose(1, 2)) in one pass over the fp32 NCHW image.
 * x: [B, 3, H, W] fp32; w: [E, 3, 4, 4] fp32; bias, gamma, beta: [E] fp32 (gamma NULL: no norm).
 * y: [B*H/4*W/4, E] bf16. Optional outputs for backward: proj (the projection + bias, bf16, input of
 * mtl_layernorm_bwd), patches (im2col rows [B*L, 48] bf16, operand of dW = d_proj^T patches via mtl_xty), mean, rstd. */
int mtl_patch_embed_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                        void* proj, void* y, void* patches, float* mean, float* rstd, int32_t B, int32_t H, int32_t W,
                        int32_t E, float eps, mtl_stream_t stream);

/* Elementwise helpers of the block */
int mtl_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, mtl_stream_t stream);
int mtl_scale_rows(const void* x, const float* scale, void* y, int32_t S, int64_t M, int32_t C,
                   int32_t rows_per_sample, mtl_stream_t stream);
/* y[s] = x[s] * scale[s, sample] for s < S (scale NULL: plain copy, skipped when y == x) and y[S] = sum_s y[s];
 * y: [S+1, M, C]. DropPath backward pre-scale fused with the stream sum mtl_linear_bwd_input(dy_has_sum) consumes. */
int mtl_scale_rows_sum(const void* x, const float* scale, void* y, int32_t S, int64_t M, int32_t C,
                       int32_t rows_per_sample, mtl_stream_t stream);
int mtl_add(const void* a, const void* b, void* out, int64_t n, mtl_stream_t stream);
/* out[i] = sum_{s<S} x[s, i] (+ extra[i] if extra != NULL): gradient of a tensor consumed by S residual streams
 * (shortcut + drop_path(stream), swin_transformer_mtlora.py:389-392). */
int mtl_sum_streams(const void* x, const void* extra, void* out, int32_t S, int64_t n, mtl_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimizer step over the trainable tensors of the path — main.py:341-353 -> utils.py:348-369
 * (NativeScalerWithGradNormCount: scaler.unscale_ + clip_grad_norm_ + scaler.step(AdamW), optimizer.py:58-60)
 *
 * The trainable tensors (~200 adapters, LayerNorm affines, rel-pos tables, reductions) are described ONCE by a
 * device-side segment table; a step is two launches whatever their number. `segs[i].offset` addresses the two flat
 * fp32 moment buffers; `prefix[i]` = number of MTL_OPT_CHUNK-element chunks before segment i (prefix[n_segs] =
 * n_chunks), both in device memory. A segment whose `grad` is NULL (grad is None) is skipped entirely, like
 * torch.optim skips parameters without a gradient.
 * ---------------------------------------------------------------------------------------------------------- */
#define MTL_OPT_CHUNK 4096
#define MTL_OPT_MAX_GROUPS 8
typedef struct mtl_opt_seg {
  float* param;      /* fp32 master parameter (updated in place) */
  const float* grad; /* fp32 gradient, same shape, contiguous; NULL = no gradient this step */
  int64_t offset;    /* element offset of this tensor's moments in flat_m / flat_v */
  int64_t numel;
  int32_t group;     /* index into the groups array of mtl_opt_adamw */
  int32_t pad_;
} mtl_opt_seg;
typedef struct mtl_opt_group {
  float lr, beta1, beta2, eps, weight_decay;
} mtl_opt_group;
int mtl_opt_seg_size(void);
/* out_sq[0] = sum of squares of every gradient in the table (inf / nan propagate). */
int mtl_opt_sqnorm(const mtl_opt_seg* segs, const int32_t* prefix, int32_t n_segs, int32_t n_chunks, float* out_sq,
                   mtl_stream_t stream);
/* One AdamW (adam_w = 1) / Adam-with-L2 (adam_w = 0) step. state: device float[2 + n_segs] = {scratch, total grad
 * norm of this step (when sqnorm given), steps taken by segment 0, 1, ...} (one step counter per tensor, like
 * torch.optim: a tensor without a gradient does not advance), zero-initialised by the caller, owned by the optimizer.
 * grad_scale / found_inf: the GradScaler's device scalars or NULL (torch.amp contract of fused optimizers: gradients
 * are divided by *grad_scale, the whole step — including the step counter — is skipped when *found_inf != 0).
 * sqnorm: result of mtl_opt_sqnorm on the same (still scaled) gradients, needed when max_norm > 0:
 * coef = min(1, max_norm / (sqrt(sqnorm) / grad_scale + 1e-6)) (torch.nn.utils.clip_grad_norm_). */
int mtl_opt_adamw(const mtl_opt_seg* segs, const int32_t* prefix, int32_t n_segs, int32_t n_chunks, float* flat_m,
                  float* flat_v, float* state, const mtl_opt_group* groups, int32_t n_groups, const float* grad_scale,
                  const float* found_inf, const float* sqnorm, float max_norm, int32_t adam_w, mtl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MTLORA_B200_H_ */
