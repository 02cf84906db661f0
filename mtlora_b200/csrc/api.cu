// C ABI of libmtlora_b200.so (declared in include/mtlora_b200.h): argument checking, translation of the
// reference-level layer description (mtl_linear_cfg == the constructor arguments of MTLoRALinear,
// models/lora.py:161-176) into kernel plans, and the thread-local error string.
#include "../../include/mtlora_b200.h"

#include <stdarg.h>
#include <atomic>
#include <string.h>
#include <vector>

#include "kernels.cuh"
#include "linear_sm100.cuh"

namespace mtl {

static thread_local char g_err[768] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {

inline cudaStream_t S(mtl_stream_t s) { return static_cast<cudaStream_t>(s); }
inline int pad16(int r) { return (r + 15) / 16 * 16; }

// Packed rank space: adapter 0 = shared, adapter 1+t = task t; every adapter starts on a 16-column boundary.
struct RankLayout {
  int n;          // number of adapters (0 when r_shared == 0)
  int off[1 + MTL_MAX_TASKS];
  int len[1 + MTL_MAX_TASKS];  // padded length
  int rank[1 + MTL_MAX_TASKS];
  float scale[1 + MTL_MAX_TASKS];
  int R_pad;
};

int build_layout(const mtl_linear_cfg* c, RankLayout* L) {
  MTL_REQUIRE(c != nullptr, "linear: cfg is NULL");
  MTL_REQUIRE(c->n_tasks >= 0 && c->n_tasks <= MTL_MAX_TASKS, "linear: n_tasks=%d out of range [0, %d]", c->n_tasks,
              MTL_MAX_TASKS);
  MTL_REQUIRE(c->shared_mode == MTL_MODE_MATRIX || c->shared_mode == MTL_MODE_MATRIXV2,
              "linear: shared_mode %d not implemented ('matrix' and 'matrixv2' are)", c->shared_mode);
  MTL_REQUIRE(c->r_shared >= 0, "linear: negative rank");
  memset(L, 0, sizeof(*L));
  if (c->r_shared == 0) return 0;  // lora.py:256-257: r == 0 -> plain linear, no task outputs
  int off = 0;
  L->off[0] = 0;
  L->rank[0] = c->r_shared;
  L->len[0] = pad16(c->r_shared);
  L->scale[0] = c->scale_shared;
  off = L->len[0];
  L->n = 1;
  for (int t = 0; t < c->n_tasks; ++t) {
    MTL_REQUIRE(c->r_task[t] > 0, "linear: task %d has rank %d (must be > 0)", t, c->r_task[t]);
    L->off[1 + t] = off;
    L->rank[1 + t] = c->r_task[t];
    L->len[1 + t] = pad16(c->r_task[t]);
    L->scale[1 + t] = c->scale_task[t];
    off += L->len[1 + t];
    L->n = 2 + t;
  }
  L->R_pad = off;
  MTL_REQUIRE(L->R_pad <= 16 * LIN_MAX_GRAN, "linear: packed rank space %d exceeds the supported %d", L->R_pad,
              16 * LIN_MAX_GRAN);
  return 0;
}

// number of output streams of the layer: lora.py:262-266 — a task dict only when tasks is not None and r > 0
inline int out_streams(const mtl_linear_cfg* c) { return (c->r_shared > 0) ? 1 + c->n_tasks : 1; }

int add_group(LinPlan& p, int in, int r0, int len, int acc) {
  while (len > 0) {
    MTL_REQUIRE(p.n_groups < LIN_MAX_GROUPS, "linear: too many adapter groups");
    const int l = len > 128 ? 128 : len;
    p.grp_in[p.n_groups] = in;
    p.grp_r0[p.n_groups] = r0;
    p.grp_len[p.n_groups] = l;
    p.grp_acc[p.n_groups] = acc;
    ++p.n_groups;
    r0 += l;
    len -= l;
  }
  return 0;
}

void fill_granules(LinPlan& p, const RankLayout& L, const int* adapter_in) {
  for (int a = 0; a < L.n; ++a)
    for (int g = L.off[a] / 16; g < (L.off[a] + L.len[a]) / 16; ++g) {
      p.gran_scale[g] = L.scale[a];
      p.gran_in[g] = adapter_in[a];
    }
}

// mtl_linear_plan: while `probe` is set on this thread, the linear entry points stop after planning and report the
// tiling instead of launching (no CUDA call is made, so this works without a device).
struct PlanProbe {
  mtl_linear_plan_info* out;
  int n_sm;
};
thread_local PlanProbe* g_probe = nullptr;

int run_linear(LinPlan& p, const void* x, const void* wm, const void* down, const void* up, cudaStream_t stream) {
  if (g_probe == nullptr) return launch_linear(p, x, wm, down, up, stream);
  uint32_t smem = 0;
  if (int e = plan_linear(p, g_probe->n_sm, &smem)) return e;
  mtl_linear_plan_info* o = g_probe->out;
  memset(o, 0, sizeof(*o));
  o->bn = p.BN;
  o->n_chunks = p.n_chunks;
  o->n_splits = p.n_splits;
  o->n_stages = p.n_stages;
  o->n_slabs = p.n_slabs;
  o->n_regions = p.n_regions;
  o->n_pbuf = p.n_pbuf;
  o->n_dbuf = p.n_dbuf;
  o->d_shared = p.d_shared;
  o->tmem_cols = p.tmem_cols;
  o->tmem_cols_used = p.d_shared ? p.acc_col0 + (p.n_pbuf + 1) * p.BN
                                 : p.acc_col0 + (p.n_pbuf + 2 * p.n_dbuf) * p.BN;
  o->smem_bytes = static_cast<int32_t>(smem);
  o->n_work = p.n_work;
  o->n_groups = p.n_groups;
  o->s_in = p.S_in;
  o->s_out = p.S_out;
  o->up_pack = p.up_pack;
  return 0;
}

int check_ptr16(const void* p, const char* what) {
  MTL_REQUIRE(p != nullptr, "%s is NULL", what);
  MTL_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "%s must be 16-byte aligned", what);
  return 0;
}

}  // namespace
}  // namespace mtl

using namespace mtl;

extern "C" {

int mtl_abi_version(void) { return MTL_ABI_VERSION; }
int mtl_linear_cfg_size(void) { return static_cast<int>(sizeof(mtl_linear_cfg)); }
const char* mtl_last_error(void) { return g_err; }
uint64_t mtl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mtl_linear_rank_pad(const mtl_linear_cfg* cfg) {
  RankLayout L;
  if (build_layout(cfg, &L)) return -1;
  return L.R_pad;
}

int mtl_linear_rank_offset(const mtl_linear_cfg* cfg, int idx) {
  RankLayout L;
  if (build_layout(cfg, &L)) return -1;
  if (idx < 0 || idx >= L.n) {
    set_error("linear: adapter index %d out of range (have %d)", idx, L.n);
    return -1;
  }
  return L.off[idx];
}

int mtl_linear_pack(const mtl_linear_cfg* cfg, const float* a_shared, const float* b_shared,
                    const float* const* a_tasks, const float* const* b_tasks, void* a_cat, void* b_cat,
                    void* a_cat_t, void* b_cat_t, mtl_stream_t stream) {
  RankLayout L;
  if (int e = build_layout(cfg, &L)) return e;
  MTL_REQUIRE(L.n > 0, "linear_pack: layer has no adapters (r_shared == 0)");
  MTL_REQUIRE(a_shared != nullptr && b_shared != nullptr, "linear_pack: shared adapter pointers are NULL");
  MTL_REQUIRE(cfg->n_tasks == 0 || (a_tasks != nullptr && b_tasks != nullptr), "linear_pack: task adapter arrays are NULL");
  const float* ap[8];
  const float* bp[8];
  ap[0] = a_shared;
  bp[0] = b_shared;
  for (int t = 0; t < cfg->n_tasks; ++t) {
    MTL_REQUIRE(a_tasks[t] != nullptr && b_tasks[t] != nullptr, "linear_pack: task %d adapter pointer is NULL", t);
    ap[1 + t] = a_tasks[t];
    bp[1 + t] = b_tasks[t];
  }
  return launch_pack_adapters(ap, bp, L.rank, L.off, L.n, cfg->in_features, cfg->out_features, L.R_pad, a_cat, b_cat,
                              a_cat_t, b_cat_t, S(stream));
}

int mtl_pack_job_size(void) { return static_cast<int>(sizeof(mtl_pack_job)); }

int mtl_linear_pack_many(const mtl_pack_job* jobs, int32_t n_jobs, mtl_stream_t stream) {
  MTL_REQUIRE(n_jobs >= 0 && (n_jobs == 0 || jobs != nullptr), "linear_pack_many: bad job list");
  if (n_jobs == 0) return 0;
  std::vector<PackJobHost> hs(static_cast<size_t>(n_jobs));
  for (int j = 0; j < n_jobs; ++j) {
    const mtl_pack_job& jb = jobs[j];
    RankLayout L;
    if (int e = build_layout(&jb.cfg, &L)) return e;
    MTL_REQUIRE(L.n > 0, "linear_pack_many: job %d has no adapters (r_shared == 0)", j);
    MTL_REQUIRE(jb.a_shared != nullptr && jb.b_shared != nullptr, "linear_pack_many: job %d: shared adapter pointers are NULL", j);
    PackJobHost& h = hs[j];
    memset(&h, 0, sizeof(h));
    h.a[0] = jb.a_shared;
    h.b[0] = jb.b_shared;
    for (int t = 0; t < jb.cfg.n_tasks; ++t) {
      MTL_REQUIRE(jb.a_tasks[t] != nullptr && jb.b_tasks[t] != nullptr, "linear_pack_many: job %d: task %d adapter pointer is NULL", j, t);
      h.a[1 + t] = jb.a_tasks[t];
      h.b[1 + t] = jb.b_tasks[t];
    }
    for (int i = 0; i < L.n; ++i) {
      h.rank[i] = L.rank[i];
      h.off[i] = L.off[i];
    }
    h.n_adapt = L.n;
    h.K = jb.cfg.in_features;
    h.N = jb.cfg.out_features;
    h.R_pad = L.R_pad;
    h.a_cat = jb.a_cat; h.b_cat = jb.b_cat; h.a_cat_t = jb.a_cat_t; h.b_cat_t = jb.b_cat_t;
  }
  return launch_pack_adapters_many(hs.data(), n_jobs, S(stream));
}

int mtl_linear_rank_project(const mtl_linear_cfg* cfg, int32_t pass, const void* x, const void* down, void* u_out,
                            mtl_stream_t stream) {
  RankLayout L;
  if (int e = build_layout(cfg, &L)) return e;
  MTL_REQUIRE(L.n == 1, "linear_rank_project: needs a layer with a shared adapter and no task adapters");
  MTL_REQUIRE(pass == 0 || pass == 1, "linear_rank_project: pass %d (0 = forward, 1 = input gradient)", pass);
  if (int e = check_ptr16(x, "linear_rank_project: x")) return e;
  if (int e = check_ptr16(down, "linear_rank_project: down")) return e;
  if (int e = check_ptr16(u_out, "linear_rank_project: u_out")) return e;
  MTL_REQUIRE(cfg->M > 0 && cfg->M < (1ll << 31), "linear_rank_project: M=%lld out of range", (long long)cfg->M);
  const bool drop = pass == 0 && cfg->dropout_p > 0.f;
  const int Kc = pass == 0 ? cfg->in_features : cfg->out_features;
  if (g_probe == nullptr && rank_project_supported(L.R_pad, Kc)) {
    // forward: the adapters read D(x[0]), appended as the last stream (lora.py:258)
    const __nv_bfloat16* xs = static_cast<const __nv_bfloat16*>(x) + (drop ? static_cast<size_t>(cfg->M) * Kc : 0);
    return launch_rank_project(xs, down, u_out, static_cast<int>(cfg->M), Kc, L.R_pad, L.scale[0], S(stream));
  }
  LinPlan p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(cfg->M);
  p.Kc = Kc;
  p.Nn = L.R_pad;
  p.S_in = drop ? 2 : 1;
  p.S_out = 1;
  p.n_main = 1;
  p.main_in[0] = drop ? 1 : 0;   // forward: the adapters read D(x[0]), appended as the last stream (lora.py:258)
  p.out_useP[0] = 1;
  p.ep_mode = LIN_EP_NONE;
  p.y = static_cast<__nv_bfloat16*>(u_out);
  p.out_scale = L.scale[0];
  return run_linear(p, x, down, nullptr, nullptr, S(stream));
}

int mtl_cast_transpose(const float* w, void* w_bf16, void* wt_bf16, int32_t rows, int32_t cols, mtl_stream_t stream) {
  MTL_REQUIRE(w != nullptr, "cast_transpose: source is NULL");
  return launch_cast_transpose(w, w_bf16, wt_bf16, rows, cols, S(stream));
}

int mtl_linear_fwd(const mtl_linear_cfg* cfg, const void* x, const void* w_bf16, const float* bias,
                   const void* a_cat, const void* b_cat, int32_t act, void* y, void* y_act, const void* residual,
                   int32_t res_streams, const float* path_scale, void* u_save, mtl_stream_t stream) {
  RankLayout L;
  if (int e = build_layout(cfg, &L)) return e;
  if (int e = check_ptr16(x, "linear_fwd: x")) return e;
  if (int e = check_ptr16(w_bf16, "linear_fwd: w_bf16")) return e;
  if (int e = check_ptr16(y, "linear_fwd: y")) return e;
  MTL_REQUIRE(cfg->M > 0 && cfg->M < (1ll << 31), "linear_fwd: M=%lld out of range", (long long)cfg->M);
  MTL_REQUIRE(act == MTL_ACT_NONE || act == MTL_ACT_GELU || act == MTL_ACT_GELU_GRAD, "linear_fwd: unknown activation %d",
              act);
  MTL_REQUIRE(act == MTL_ACT_NONE || y_act != nullptr, "linear_fwd: y_act required with MTL_ACT_GELU / MTL_ACT_GELU_GRAD");
  MTL_REQUIRE(cfg->dropout_p >= 0.f && cfg->dropout_p < 1.f, "dropout probability has to be in [0, 1), but got %f",
              cfg->dropout_p);
  const int T = cfg->n_tasks;
  const bool lora = L.n > 0;
  const bool xt = lora && T > 0 && cfg->x_tasks_given;
  const bool drop = lora && cfg->dropout_p > 0.f;
  if (lora) {
    if (int e = check_ptr16(a_cat, "linear_fwd: a_cat")) return e;
    if (int e = check_ptr16(b_cat, "linear_fwd: b_cat")) return e;
  }

  LinPlan p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(cfg->M);
  p.Kc = cfg->in_features;
  p.Nn = cfg->out_features;
  p.S_out = out_streams(cfg);
  p.S_in = 1 + (xt ? T : 0) + (drop ? 1 : 0);
  p.R_pad = L.R_pad;
  const int drop_in = drop ? p.S_in - 1 : 0;  // stream holding D(x[0])
  int adapter_in[1 + MTL_MAX_TASKS] = {0};
  if (lora) {
    if (!xt) {
      for (int a = 0; a < L.n; ++a) adapter_in[a] = drop_in;
      if (int e = add_group(p, drop_in, 0, L.R_pad, 0)) return e;
    } else {
      adapter_in[0] = drop_in;
      if (int e = add_group(p, drop_in, L.off[0], L.len[0], 0)) return e;
      for (int t = 0; t < T; ++t) {
        adapter_in[1 + t] = 1 + t;  // x_tasks[t] is never dropped out (lora.py:263)
        if (int e = add_group(p, 1 + t, L.off[1 + t], L.len[1 + t], 0)) return e;
      }
    }
    fill_granules(p, L, adapter_in);
  }
  p.n_main = 1;
  p.main_in[0] = 0;
  for (int j = 0; j < p.S_out; ++j) {
    p.out_useP[j] = 1;
    if (lora) {
      p.out_r0[j][0] = L.off[j];
      p.out_len[j][0] = L.len[j];
      if (j > 0 && cfg->shared_mode == MTL_MODE_MATRIXV2) {   // lora.py:267-274: pretrained + lora (shared) + task adapter
        p.out_r0[j][1] = L.off[0];
        p.out_len[j][1] = L.len[0];
      }
    }
  }
  p.ep_mode = act == MTL_ACT_GELU ? LIN_EP_GELU_DUAL : act == MTL_ACT_GELU_GRAD ? LIN_EP_GELU_DUAL_GRAD : LIN_EP_NONE;
  p.bias = bias;
  p.y = static_cast<__nv_bfloat16*>(y);
  p.y2 = static_cast<__nv_bfloat16*>(y_act);
  p.res = static_cast<const __nv_bfloat16*>(residual);
  p.res_streams = res_streams;
  if (residual != nullptr)
    MTL_REQUIRE(res_streams == 1 || res_streams == p.S_out, "linear_fwd: res_streams=%d must be 1 or %d", res_streams,
                p.S_out);
  p.rowscale_out = path_scale;
  if (path_scale != nullptr || cfg->rows_per_sample > 0) {
    MTL_REQUIRE(cfg->rows_per_sample > 0 && cfg->M % cfg->rows_per_sample == 0,
                "linear_fwd: rows_per_sample=%d does not divide M=%lld", cfg->rows_per_sample, (long long)cfg->M);
    p.rows_per_sample = cfg->rows_per_sample;
    p.n_samples = static_cast<int>(cfg->M / cfg->rows_per_sample);
  }
  p.u_save = static_cast<__nv_bfloat16*>(u_save);
  if (cfg->u_precomputed && lora) {
    MTL_REQUIRE(L.n == 1 && u_save != nullptr, "linear_fwd: u_precomputed needs a layer without task adapters and U");
    p.u_in = 1;
  }
  if (drop && act != MTL_ACT_NONE) p.drop_mode = 1;
  p.drop_p = cfg->dropout_p;
  p.drop_seed = cfg->dropout_seed;
  return run_linear(p, x, w_bf16, a_cat, b_cat, S(stream));
}

int mtl_linear_bwd_input(const mtl_linear_cfg* cfg, const void* dy, const void* wt_bf16, const void* a_cat_t,
                         const void* b_cat_t, void* dx, const void* gelu_aux, const float* path_scale,
                         void* g_save, mtl_stream_t stream) {
  RankLayout L;
  if (int e = build_layout(cfg, &L)) return e;
  if (int e = check_ptr16(dy, "linear_bwd_input: dy")) return e;
  if (int e = check_ptr16(wt_bf16, "linear_bwd_input: wt_bf16")) return e;
  if (int e = check_ptr16(dx, "linear_bwd_input: dx")) return e;
  MTL_REQUIRE(cfg->M > 0 && cfg->M < (1ll << 31), "linear_bwd_input: M=%lld out of range", (long long)cfg->M);
  const int T = cfg->n_tasks;
  const bool lora = L.n > 0;
  const bool xt = lora && T > 0 && cfg->x_tasks_given;
  const bool drop = lora && cfg->dropout_p > 0.f;
  if (lora) {
    if (int e = check_ptr16(a_cat_t, "linear_bwd_input: a_cat_t")) return e;
    if (int e = check_ptr16(b_cat_t, "linear_bwd_input: b_cat_t")) return e;
  }

  LinPlan p;
  memset(&p, 0, sizeof(p));
  p.M = static_cast<int>(cfg->M);
  p.Kc = cfg->out_features;  // contraction runs over the layer's outputs
  p.Nn = cfg->in_features;
  const bool dy_sum = cfg->dy_has_sum != 0 && out_streams(cfg) > 1;
  p.S_in = out_streams(cfg) + (dy_sum ? 1 : 0);
  p.S_out = xt ? 1 + T : 1;
  p.R_pad = L.R_pad;
  int adapter_in[1 + MTL_MAX_TASKS] = {0};
  if (lora) {
    const bool v2 = cfg->shared_mode == MTL_MODE_MATRIXV2 && L.n > 1;
    for (int a = 0; a < L.n; ++a) {
      adapter_in[a] = a;  // adapter a back-propagates the gradient of output stream a
      if (a == 0 && v2) {
        // matrixv2: the shared adapter fed every output stream -> G_sh = s_sh (sum_j dy[j]) B_sh: one group on the
        // pre-summed stream, or one accumulating group per stream
        if (dy_sum) {
          adapter_in[0] = p.S_in - 1;
          if (int e = add_group(p, p.S_in - 1, L.off[0], L.len[0], 0)) return e;
        } else {
          for (int j = 0; j < out_streams(cfg); ++j)
            if (int e = add_group(p, j, L.off[0], L.len[0], j > 0 ? 1 : 0)) return e;
        }
        continue;
      }
      if (int e = add_group(p, a, L.off[a], L.len[a], 0)) return e;
    }
    fill_granules(p, L, adapter_in);
  }
  if (dy_sum) {   // the caller appended sum_j dy[j] as the last stream: the frozen product reads one stream, not 1+T
    p.n_main = 1;
    p.main_in[0] = p.S_in - 1;
  } else {
    p.n_main = p.S_in;  // dPre = sum_j dy[j]  (pretrained feeds every output stream, lora.py:262-266,284)
    for (int j = 0; j < p.S_in; ++j) p.main_in[j] = j;
  }
  if (!xt) {
    p.out_useP[0] = 1;
    p.out_r0[0][0] = 0;
    p.out_len[0][0] = L.R_pad;
  } else {
    p.out_useP[0] = 1;
    p.out_r0[0][0] = L.off[0];
    p.out_len[0][0] = L.len[0];
    for (int t = 0; t < T; ++t) {
      p.out_useP[1 + t] = 0;
      p.out_r0[1 + t][0] = L.off[1 + t];
      p.out_len[1 + t][0] = L.len[1 + t];
    }
  }
  p.ep_mode = gelu_aux == nullptr ? LIN_EP_NONE : (cfg->gelu_aux_is_grad ? LIN_EP_MUL_AUX : LIN_EP_GELU_BWD);
  p.aux = static_cast<const __nv_bfloat16*>(gelu_aux);
  p.y = static_cast<__nv_bfloat16*>(dx);
  if (path_scale != nullptr) {
    MTL_REQUIRE(out_streams(cfg) == 1, "linear_bwd_input: in-kernel path_scale needs a single dy stream; pre-scale dy instead");
    MTL_REQUIRE(cfg->rows_per_sample > 0 && cfg->M % cfg->rows_per_sample == 0,
                "linear_bwd_input: rows_per_sample=%d does not divide M=%lld", cfg->rows_per_sample, (long long)cfg->M);
    // a single stream: scaling the rows of dy == scaling the rows of the result. The saved G stays unscaled;
    // mtl_linear_bwd_params applies the same path_scale to its second operand.
    p.rowscale_out = path_scale;
    p.rows_per_sample = cfg->rows_per_sample;
    p.n_samples = static_cast<int>(cfg->M / cfg->rows_per_sample);
  }
  p.u_save = static_cast<__nv_bfloat16*>(g_save);
  if (cfg->u_precomputed && lora) {
    MTL_REQUIRE(L.n == 1 && g_save != nullptr, "linear_bwd_input: u_precomputed needs a layer without task adapters and G");
    p.u_in = 1;
  }
  if (drop) {
    p.drop_mode = 2;
    p.force_split = 1;
  }
  p.drop_p = cfg->dropout_p;
  p.drop_seed = cfg->dropout_seed;
  return run_linear(p, dy, wt_bf16, b_cat_t, a_cat_t, S(stream));
}

int mtl_linear_plan(const mtl_linear_cfg* cfg, int32_t pass, int32_t act, int32_t has_residual, int32_t n_sm,
                    mtl_linear_plan_info* out) {
  MTL_REQUIRE(cfg != nullptr && out != nullptr, "linear_plan: NULL argument");
  MTL_REQUIRE(pass == 0 || pass == 1, "linear_plan: pass %d (0 = forward, 1 = input gradient)", pass);
  MTL_REQUIRE(n_sm > 0, "linear_plan: n_sm=%d", n_sm);
  // the entry points only validate these pointers (non-NULL, 16-byte aligned) before planning; nothing is dereferenced
  void* const buf = reinterpret_cast<void*>(uintptr_t{4096});
  float* const fbuf = reinterpret_cast<float*>(uintptr_t{4096});
  PlanProbe probe{out, n_sm};
  g_probe = &probe;
  int rc;
  if (pass == 0)
    rc = mtl_linear_fwd(cfg, buf, buf, fbuf, buf, buf, act, buf, act != MTL_ACT_NONE ? buf : nullptr,
                        has_residual ? buf : nullptr, 1, nullptr, nullptr, nullptr);
  else
    rc = mtl_linear_bwd_input(cfg, buf, buf, buf, buf, buf, act != MTL_ACT_NONE ? buf : nullptr, nullptr, nullptr, nullptr);
  g_probe = nullptr;
  return rc;
}

int mtl_linear_bwd_params(const mtl_linear_cfg* cfg, const void* x, int32_t x_gelu, const void* dy,
                          const void* u_save, const void* g_save, const float* path_scale, float* da_cat,
                          float* db_cat, mtl_stream_t stream) {
  RankLayout L;
  if (int e = build_layout(cfg, &L)) return e;
  MTL_REQUIRE(L.n > 0, "linear_bwd_params: layer has no adapters");
  MTL_REQUIRE(x != nullptr && dy != nullptr && u_save != nullptr && g_save != nullptr && da_cat != nullptr &&
                  db_cat != nullptr,
              "linear_bwd_params: NULL argument");
  const int T = cfg->n_tasks;
  const bool xt = T > 0 && cfg->x_tasks_given;
  const bool drop = cfg->dropout_p > 0.f;
  MTL_REQUIRE(!(x_gelu && drop), "linear_bwd_params: x_gelu cannot be combined with LoRA dropout");
  const int S_in = 1 + (xt ? T : 0) + (drop ? 1 : 0);
  const int drop_in = drop ? S_in - 1 : 0;
  const int64_t M = cfg->M;
  const int K = cfg->in_features, N = cfg->out_features;
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* dyb = static_cast<const __nv_bfloat16*>(dy);
  const __nv_bfloat16* ub = static_cast<const __nv_bfloat16*>(u_save);
  const __nv_bfloat16* gb = static_cast<const __nv_bfloat16*>(g_save);
  if (path_scale != nullptr)
    MTL_REQUIRE(cfg->rows_per_sample > 0 && M % cfg->rows_per_sample == 0,
                "linear_bwd_params: rows_per_sample=%d does not divide M=%lld", cfg->rows_per_sample, (long long)M);
  const int n_samples = path_scale != nullptr ? static_cast<int>(M / cfg->rows_per_sample) : 0;
  MTL_REQUIRE(!((x_gelu || K < 72 || N < 72) && cfg->shared_mode == MTL_MODE_MATRIXV2 && L.n > 1),
              "linear_bwd_params: matrixv2 needs the tensor-core reduction (no x_gelu, K and N >= 72)");
  if (x_gelu || K < 72 || N < 72) {
    // legacy path (mma.sync): GELU recomputation on load, or operands too narrow for the 128-column UMMA tile
    for (int a = 0; a < L.n; ++a) {
      const int in = (a == 0 || !xt) ? drop_in : a;
      // dB_a [N, len] += dy[a]^T (ps[a] * U[:, off:off+len])
      if (int e = launch_xty(dyb + static_cast<size_t>(a) * M * N, N, ub + L.off[a], L.R_pad, db_cat + L.off[a], L.R_pad,
                             M, N, L.len[a], path_scale ? path_scale + static_cast<size_t>(a) * n_samples : nullptr,
                             cfg->rows_per_sample, 0, 1.f, S(stream)))
        return e;
      // dA_a [len, K] += G[:, off:off+len]^T (ps[a] * x_in(a))     (G carries the adapter scale, not the path scale)
      if (int e = launch_xty(gb + L.off[a], L.R_pad, xb + static_cast<size_t>(in) * M * K, K,
                             da_cat + static_cast<size_t>(L.off[a]) * K, K, M, L.len[a], K,
                             path_scale ? path_scale + static_cast<size_t>(a) * n_samples : nullptr,
                             cfg->rows_per_sample, x_gelu, 1.f, S(stream)))
        return e;
    }
    return 0;
  }
  // one tcgen05 launch for every adapter: wide operands dy (dB) and x (dA), rank operands U and G
  XtyOperand wide[2] = {{dyb, N, 0, out_streams(cfg)}, {xb, K, 0, S_in}};
  XtyOperand rank[2] = {{ub, L.R_pad, 0, 1}, {gb, L.R_pad, 0, 1}};
  XtyJobGroup grp[3 * (1 + MTL_MAX_TASKS)];
  int ng = 0;
  const bool v2 = cfg->shared_mode == MTL_MODE_MATRIXV2 && L.n > 1;
  const bool dy_sum = cfg->dy_has_sum != 0 && out_streams(cfg) > 1;
  if (dy_sum) wide[0].streams += 1;   // dy carries sum_j dy[j] as an extra stream (mtl_scale_rows_sum)
  for (int a = 0; a < L.n; ++a) {   // dB_a [N, len] += dy[a]^T (ps[a] * U[:, off:off+len])
    if (a == 0 && v2) {
      // matrixv2: dB_sh = (sum_j dy[j])^T U_sh — the pre-summed stream, or every stream accumulated by the reduction
      MTL_REQUIRE(path_scale == nullptr, "linear_bwd_params: matrixv2 with task streams expects pre-scaled dy");
      if (dy_sum) {
        grp[ng++] = XtyJobGroup{0, out_streams(cfg), 0, L.off[0], L.len[0], 0, L.R_pad, db_cat, nullptr};
      } else {
        for (int j = 0; j < out_streams(cfg); ++j)
          grp[ng++] = XtyJobGroup{0, j, 0, L.off[0], L.len[0], 0, L.R_pad, db_cat, nullptr};
      }
      continue;
    }
    grp[ng++] = XtyJobGroup{0, a, 0, L.off[a], L.len[a], 0, L.R_pad, db_cat,
                            path_scale ? path_scale + static_cast<size_t>(a) * n_samples : nullptr};
  }
  if (!xt && (path_scale == nullptr || L.n == 1)) {
    // every adapter consumed the same input stream: dA_cat [R, K] += G^T (ps * x) in one group
    grp[ng++] = XtyJobGroup{1, drop_in, 1, 0, L.R_pad, 1, K, da_cat, path_scale};
  } else {
    for (int a = 0; a < L.n; ++a) {   // dA_a [len, K] += (ps[a] * G[:, off:off+len])^T x_in(a)
      const int in = (a == 0 || !xt) ? drop_in : a;
      grp[ng++] = XtyJobGroup{1, in, 1, L.off[a], L.len[a], 1, K, da_cat,
                              path_scale ? path_scale + static_cast<size_t>(a) * n_samples : nullptr};
    }
  }
  return launch_xty_groups(wide, rank, grp, ng, M, cfg->rows_per_sample, S(stream));
}

int mtl_xty(const void* p, int64_t ldp, const void* q, int64_t ldq, float* c, int64_t ldc, int64_t M, int32_t a,
            int32_t b, float alpha, mtl_stream_t stream) {
  MTL_REQUIRE(p != nullptr && q != nullptr && c != nullptr, "xty: NULL argument");
  const bool q_wide = b >= a;
  const int wide_cols = q_wide ? b : a, rank_cols = q_wide ? a : b;
  if (alpha == 1.f && wide_cols >= 72 && rank_cols % 4 == 0 && ldp % 8 == 0 && ldq % 8 == 0 && M < (1ll << 31)) {
    // C[a, b]: wide = the larger dimension (128-column UMMA tiles), rank = the other (64-column chunks)
    XtyOperand wide[2] = {{q_wide ? q : p, wide_cols, q_wide ? ldq : ldp, 1}, {nullptr, 0, 0, 0}};
    XtyOperand rank[2] = {{q_wide ? p : q, rank_cols, q_wide ? ldp : ldq, 1}, {nullptr, 0, 0, 0}};
    XtyJobGroup g{0, 0, 0, 0, rank_cols, q_wide ? 1 : 0, static_cast<int>(ldc), c, nullptr};
    return launch_xty_groups(wide, rank, &g, 1, M, 0, S(stream));
  }
  return launch_xty(p, ldp, q, ldq, c, ldc, M, a, b, nullptr, 0, 0, alpha, S(stream));
}

int mtl_window_attention_fwd(const void* qkv, const float* rpb, const float* mask, int32_t n_mask, void* out,
                             void* out_drop, float* lse, int32_t B, int32_t H, int32_t W, int32_t C,
                             int32_t num_heads, int32_t window_size, int32_t shift_size, float scale,
                             float dropout_p, uint64_t dropout_seed, mtl_stream_t stream) {
  if (int e = check_ptr16(qkv, "window_attention_fwd: qkv")) return e;
  if (int e = check_ptr16(out, "window_attention_fwd: out")) return e;
  MTL_REQUIRE(rpb != nullptr, "window_attention_fwd: relative_position_bias_table is NULL");
  MTL_REQUIRE(mask == nullptr || n_mask > 0, "window_attention_fwd: mask given with n_mask=%d", n_mask);
  MTL_REQUIRE(out_drop == nullptr || (dropout_p > 0.f && dropout_p < 1.f), "window_attention_fwd: out_drop needs 0 < p < 1");
  return launch_win_attn_fwd(qkv, rpb, mask, n_mask, out, out_drop, lse, B, H, W, C, num_heads, window_size,
                             shift_size, scale, dropout_p, dropout_seed, S(stream));
}

int mtl_window_attention_bwd(const void* qkv, const void* dout, const float* rpb, const float* mask,
                             int32_t n_mask, const float* lse, void* dqkv, float* drpb, int32_t B, int32_t H,
                             int32_t W, int32_t C, int32_t num_heads, int32_t window_size, int32_t shift_size,
                             float scale, mtl_stream_t stream) {
  if (int e = check_ptr16(qkv, "window_attention_bwd: qkv")) return e;
  if (int e = check_ptr16(dout, "window_attention_bwd: dout")) return e;
  if (int e = check_ptr16(dqkv, "window_attention_bwd: dqkv")) return e;
  MTL_REQUIRE(rpb != nullptr && lse != nullptr, "window_attention_bwd: NULL argument");
  MTL_REQUIRE(mask == nullptr || n_mask > 0, "window_attention_bwd: mask given with n_mask=%d", n_mask);
  return launch_win_attn_bwd(qkv, dout, rpb, mask, n_mask, lse, dqkv, drpb, B, H, W, C, num_heads, window_size,
                             shift_size, scale, S(stream));
}

int mtl_roll_and_window_partition_forward(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                          int32_t shift_size, int32_t window_size, int32_t elem_size,
                                          mtl_stream_t stream) {
  MTL_REQUIRE(in != nullptr && out != nullptr, "window_process: NULL argument");
  // the reference passes shift_size = -shift (swin_transformer_mtlora.py:344-345): out reads in[.. + (-shift_size) ..]
  return launch_roll_partition(in, out, B, H, W, C, -shift_size, window_size, elem_size, 0, S(stream));
}
int mtl_roll_and_window_partition_backward(const void* grad_in, void* grad_out, int32_t B, int32_t H, int32_t W,
                                           int32_t C, int32_t shift_size, int32_t window_size, int32_t elem_size,
                                           mtl_stream_t stream) {
  MTL_REQUIRE(grad_in != nullptr && grad_out != nullptr, "window_process: NULL argument");
  return launch_roll_partition(grad_in, grad_out, B, H, W, C, -shift_size, window_size, elem_size, 1, S(stream));
}
int mtl_window_merge_and_roll_forward(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                                      int32_t shift_size, int32_t window_size, int32_t elem_size,
                                      mtl_stream_t stream) {
  MTL_REQUIRE(in != nullptr && out != nullptr, "window_process: NULL argument");
  return launch_merge_roll(in, out, B, H, W, C, shift_size, window_size, elem_size, 0, S(stream));
}
int mtl_window_merge_and_roll_backward(const void* grad_in, void* grad_out, int32_t B, int32_t H, int32_t W,
                                       int32_t C, int32_t shift_size, int32_t window_size, int32_t elem_size,
                                       mtl_stream_t stream) {
  MTL_REQUIRE(grad_in != nullptr && grad_out != nullptr, "window_process: NULL argument");
  return launch_merge_roll(grad_in, grad_out, B, H, W, C, shift_size, window_size, elem_size, 1, S(stream));
}

int mtl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, void* y_drop,
                      int64_t drop_rows, float* mean, float* rstd, int64_t rows, int32_t C, float eps, int32_t merge, int32_t H, int32_t W,
                      float dropout_p, uint64_t dropout_seed, mtl_stream_t stream) {
  if (int e = check_ptr16(x, "layernorm_fwd: x")) return e;
  if (int e = check_ptr16(y, "layernorm_fwd: y")) return e;
  if (int e = check_ptr16(gamma, "layernorm_fwd: gamma")) return e;
  if (int e = check_ptr16(beta, "layernorm_fwd: beta")) return e;
  MTL_REQUIRE(y_drop == nullptr || (dropout_p > 0.f && dropout_p < 1.f), "layernorm_fwd: y_drop needs 0 < p < 1");
  return launch_layernorm_fwd(x, gamma, beta, y, y_drop, drop_rows, mean, rstd, rows, C, eps, merge, H, W, dropout_p,
                              dropout_seed, S(stream));
}

int mtl_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dres, void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                      int32_t merge, int32_t H, int32_t W, mtl_stream_t stream) {
  if (int e = check_ptr16(dy, "layernorm_bwd: dy")) return e;
  if (int e = check_ptr16(x, "layernorm_bwd: x")) return e;
  if (int e = check_ptr16(dx, "layernorm_bwd: dx")) return e;
  if (int e = check_ptr16(gamma, "layernorm_bwd: gamma")) return e;
  MTL_REQUIRE(mean != nullptr && rstd != nullptr, "layernorm_bwd: saved statistics are NULL");
  MTL_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta must be given together");
  return launch_layernorm_bwd(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, C, merge, H, W, S(stream));
}

int mtl_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, mtl_stream_t stream) {
  MTL_REQUIRE(x != nullptr && y != nullptr, "dropout: NULL argument");
  return launch_dropout(x, y, n, p, seed, S(stream));
}
int mtl_scale_rows(const void* x, const float* scale, void* y, int32_t Sn, int64_t M, int32_t C,
                   int32_t rows_per_sample, mtl_stream_t stream) {
  MTL_REQUIRE(x != nullptr && y != nullptr && scale != nullptr, "scale_rows: NULL argument");
  return launch_scale_rows(x, scale, y, Sn, M, C, rows_per_sample, S(stream));
}
int mtl_patch_embed_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                        void* proj, void* y, void* patches, float* mean, float* rstd, int32_t B, int32_t H, int32_t W,
                        int32_t E, float eps, mtl_stream_t stream) {
  return launch_patch_embed_fwd(x, w, bias, gamma, beta, proj, y, patches, mean, rstd, B, H, W, E, eps, S(stream));
}
int mtl_scale_rows_sum(const void* x, const float* scale, void* y, int32_t Sn, int64_t M, int32_t C,
                       int32_t rows_per_sample, mtl_stream_t stream) {
  MTL_REQUIRE(x != nullptr && y != nullptr, "scale_rows_sum: NULL argument");
  return launch_scale_rows_sum(x, scale, y, Sn, M, C, rows_per_sample, S(stream));
}
int mtl_add(const void* a, const void* b, void* out, int64_t n, mtl_stream_t stream) {
  MTL_REQUIRE(a != nullptr && b != nullptr && out != nullptr, "add: NULL argument");
  return launch_add(a, b, out, n, S(stream));
}
int mtl_sum_streams(const void* x, const void* extra, void* out, int32_t Sn, int64_t n, mtl_stream_t stream) {
  MTL_REQUIRE(x != nullptr && out != nullptr && Sn >= 1, "sum_streams: bad argument");
  return launch_sum_streams(x, extra, out, Sn, n, S(stream));
}

int mtl_opt_seg_size(void) { return static_cast<int>(sizeof(mtl_opt_seg)); }

int mtl_opt_sqnorm(const mtl_opt_seg* segs, const int32_t* prefix, int32_t n_segs, int32_t n_chunks, float* out_sq,
                   mtl_stream_t stream) {
  MTL_REQUIRE(segs != nullptr && prefix != nullptr, "mtl_opt_sqnorm: NULL table");
  return opt_sqnorm(segs, prefix, n_segs, n_chunks, out_sq, S(stream));
}

int mtl_opt_adamw(const mtl_opt_seg* segs, const int32_t* prefix, int32_t n_segs, int32_t n_chunks, float* flat_m,
                  float* flat_v, float* state, const mtl_opt_group* groups, int32_t n_groups, const float* grad_scale,
                  const float* found_inf, const float* sqnorm, float max_norm, int32_t adam_w, mtl_stream_t stream) {
  MTL_REQUIRE(segs != nullptr && prefix != nullptr && flat_m != nullptr && flat_v != nullptr && state != nullptr &&
                  groups != nullptr, "mtl_opt_adamw: NULL argument");
  return opt_adamw(segs, prefix, n_segs, n_chunks, flat_m, flat_v, state, groups, n_groups, grad_scale, found_inf,
                   sqnorm, max_norm, adam_w, S(stream));
}

}  // extern "C"
