"""Tensor-level wrappers of the C ABI (include/mtlora_b200.h): allocate outputs with torch, pass raw device
pointers + sizes + the current CUDA stream to libmtlora_b200.so. No arithmetic happens in this file.

Layout conventions (see the header): activations are bf16 and stream-stacked [S, M, C] — stream 0 is the
task-shared stream, streams 1..T the per-task streams in module task order; parameters are fp32 masters whose
bf16 operand copies are produced by `pack_adapters` / `cast_transpose`; gradients come back in fp32.
"""
import ctypes
import os

import torch

from . import _native as N

BF16 = torch.bfloat16
PRE_PROJECT_MIN = int(os.environ.get("MTL_PRE_PROJECT_MIN", "256"))   # tuning aid; see LinearSpec.pre_project


class _ZeroArena:
    """Zero-initialised fp32 scratch for the gradient accumulators of one backward pass (dA / dB, LayerNorm affine,
    rel-pos tables): one cudaMemset-sized fill per 16 MB chunk instead of one fill kernel per accumulator (~100 per
    step). Slices are views: a chunk lives as long as any gradient carved from it."""
    CHUNK = 4 << 20   # floats

    def __init__(self):
        self.buf = {}

    def take(self, n, device):
        n_al = (n + 63) // 64 * 64          # 256-byte aligned slices
        if n_al > self.CHUNK // 4:
            return torch.zeros(n, dtype=torch.float32, device=device)
        key = (device.type, device.index)
        buf, off = self.buf.get(key, (None, 0))
        if buf is None or off + n_al > buf.numel():
            buf, off = torch.zeros(self.CHUNK, dtype=torch.float32, device=device), 0
        self.buf[key] = (buf, off + n_al)
        return buf[off:off + n]


_arena = _ZeroArena()


def zeros_f32(shape, device):
    n = 1
    for d in shape:
        n *= int(d)
    return _arena.take(n, device).view(*shape)


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor — mtlora_b200 has no CPU path")
    if t.device.index != N.current_device_index():
        raise RuntimeError(f"{name}: tensor lives on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()} (one process per GPU; wrap the call in torch.cuda.device)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")


class LinearSpec:
    """Static description of one MTLoRALinear (models/lora.py:161-233): sizes, ranks and scales.

    r_tasks / scale_tasks are lists in module task order; r_shared == 0 means a plain linear (lora.py:256-257 or
    CompatLinear, swin_transformer_mtlora.py:36-41), which has a single output stream.
    """

    def __init__(self, in_features, out_features, r_shared=0, r_tasks=(), scale_shared=1.0, scale_tasks=(),
                 shared_mode="matrix"):
        self.K, self.Nf = int(in_features), int(out_features)
        if shared_mode not in ("matrix", "matrixv2"):   # 'addition' is composed above the kernels (lora._AdditionTailFn)
            raise ValueError(f"mtlora_b200: LinearSpec shared_mode must be 'matrix' or 'matrixv2', got {shared_mode!r}")
        self.mode = N.MTL_MODE_MATRIXV2 if shared_mode == "matrixv2" else N.MTL_MODE_MATRIX
        self.r_shared = int(r_shared)
        self.r_tasks = [int(r) for r in r_tasks] if self.r_shared > 0 else []
        self.T = len(self.r_tasks)
        if self.T > N.MTL_MAX_TASKS:
            raise ValueError(f"at most {N.MTL_MAX_TASKS} tasks are supported, got {self.T}")
        self.scale_shared = float(scale_shared)
        self.scale_tasks = [float(s) for s in scale_tasks][: self.T]
        if len(self.scale_tasks) != self.T:
            raise ValueError("scale_tasks must have one entry per task")
        self.S_out = 1 + self.T if self.r_shared > 0 else 1
        c = self.cfg(1, False)
        lib = N.load()
        self.R_pad = lib.mtl_linear_rank_pad(ctypes.byref(c))
        if self.R_pad < 0:
            raise RuntimeError(lib.mtl_last_error().decode())
        self.offsets = [lib.mtl_linear_rank_offset(ctypes.byref(c), i) for i in range(self.S_out if self.r_shared else 0)]
        self.ranks = ([self.r_shared] + self.r_tasks) if self.r_shared else []

    def cfg(self, M, x_tasks_given, dropout_p=0.0, seed=0, rows_per_sample=0, gelu_aux_is_grad=False, dy_has_sum=False,
            u_precomputed=False):
        c = N.LinearCfg()
        c.M = int(M)
        c.in_features, c.out_features = self.K, self.Nf
        c.n_tasks = self.T
        c.x_tasks_given = 1 if (x_tasks_given and self.T > 0) else 0
        c.shared_mode = self.mode
        c.r_shared = self.r_shared
        for t in range(self.T):
            c.r_task[t] = self.r_tasks[t]
            c.scale_task[t] = self.scale_tasks[t]
        c.scale_shared = self.scale_shared
        c.dropout_p = float(dropout_p)
        c.dropout_seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        c.rows_per_sample = int(rows_per_sample)
        c.gelu_aux_is_grad = 1 if gelu_aux_is_grad else 0
        c.dy_has_sum = 1 if dy_has_sum else 0
        c.u_precomputed = 1 if u_precomputed else 0
        return c

    def pre_project(self, M):
        """Whether the rank-space projection runs as a launch of its own (mtl_linear_rank_project) and the main kernel as
        one dense product over the concatenated contraction: layers without task adapters in the compute-bound regime
        (stages 2-3: both sides >= PRE_PROJECT_MIN), where re-forming U inside every column split of a row tile costs more
        than one extra pass over x. Below that the layers are HBM-bound and keep U on chip."""
        return (self.r_shared > 0 and self.T == 0 and self.R_pad <= 128 and min(self.K, self.Nf) >= PRE_PROJECT_MIN
                and M >= 128)

    def n_in_streams(self, x_tasks_given, dropout_p):
        has_lora = self.r_shared > 0
        return 1 + (self.T if (x_tasks_given and has_lora) else 0) + (1 if (dropout_p > 0 and has_lora) else 0)


# ----------------------------------------------------------------------------------------------------------------
# parameter staging
# ----------------------------------------------------------------------------------------------------------------
def cast_transpose(w, want_w=True, want_wt=True):
    """fp32 [rows, cols] -> (bf16 copy [rows, cols], bf16 transpose [cols, rows])."""
    _chk(w, torch.float32, "w")
    rows, cols = w.shape
    wb = torch.empty((rows, cols), dtype=BF16, device=w.device) if want_w else None
    wt = torch.empty((cols, rows), dtype=BF16, device=w.device) if want_wt else None
    N.call("mtl_cast_transpose", N.ptr(w), N.ptr(wb), N.ptr(wt), rows, cols, N.stream())
    return wb, wt


def pack_adapters(spec, a_shared, b_shared, a_tasks=(), b_tasks=(), fwd=True, bwd=True):
    """fp32 adapters -> (a_cat [R,K], b_cat [N,R], a_cat_t [K,R], b_cat_t [R,N]) in bf16, zero padded."""
    dev = a_shared.device
    for i, t in enumerate([a_shared, b_shared, *a_tasks, *b_tasks]):
        _chk(t, torch.float32, f"adapter[{i}]")
    R, K, Nf = spec.R_pad, spec.K, spec.Nf
    a_cat = torch.empty((R, K), dtype=BF16, device=dev) if fwd else None
    b_cat = torch.empty((Nf, R), dtype=BF16, device=dev) if fwd else None
    a_cat_t = torch.empty((K, R), dtype=BF16, device=dev) if bwd else None
    b_cat_t = torch.empty((R, Nf), dtype=BF16, device=dev) if bwd else None
    T = spec.T
    arr_a = (ctypes.c_void_p * max(T, 1))(*[t.data_ptr() for t in a_tasks])
    arr_b = (ctypes.c_void_p * max(T, 1))(*[t.data_ptr() for t in b_tasks])
    c = spec.cfg(1, False)
    N.call("mtl_linear_pack", ctypes.byref(c), N.ptr(a_shared), N.ptr(b_shared), arr_a, arr_b, N.ptr(a_cat),
           N.ptr(b_cat), N.ptr(a_cat_t), N.ptr(b_cat_t), N.stream())
    return a_cat, b_cat, a_cat_t, b_cat_t


def pack_adapters_many(jobs):
    """pack_adapters for many layers in ONE launch (mtl_linear_pack_many). jobs: [(spec, a_shared, b_shared, a_tasks,
    b_tasks)] of fp32 CUDA tensors on one device -> [(a_cat, b_cat, a_cat_t, b_cat_t)] in job order; the packed operands
    are views of one freshly allocated bf16 buffer."""
    if not jobs:
        return []
    dev = jobs[0][1].device
    sizes = [2 * spec.R_pad * (spec.K + spec.Nf) for spec, *_ in jobs]
    offs, n = [], 0
    for sz in sizes:
        offs.append(n)
        n += (sz + 7) // 8 * 8          # every operand stays 16-byte aligned (R_pad * K and N * R_pad are multiples of 8)
    buf = torch.empty(n, dtype=BF16, device=dev)
    arr = (N.PackJob * len(jobs))()
    out = []
    for j, (spec, a_s, b_s, a_t, b_t) in enumerate(jobs):
        for i, t in enumerate([a_s, b_s, *a_t, *b_t]):
            _chk(t, torch.float32, f"job {j}: adapter[{i}]")
        if len(a_t) != spec.T or len(b_t) != spec.T:
            raise ValueError(f"pack_adapters_many: job {j} needs {spec.T} task adapters")
        R, K, Nf = spec.R_pad, spec.K, spec.Nf
        o = offs[j]
        a_cat = buf[o:o + R * K].view(R, K)
        b_cat = buf[o + R * K:o + R * (K + Nf)].view(Nf, R)
        a_cat_t = buf[o + R * (K + Nf):o + R * (2 * K + Nf)].view(K, R)
        b_cat_t = buf[o + R * (2 * K + Nf):o + 2 * R * (K + Nf)].view(R, Nf)
        jb = arr[j]
        jb.cfg = spec.cfg(1, False)
        jb.a_shared, jb.b_shared = a_s.data_ptr(), b_s.data_ptr()
        for t in range(spec.T):
            jb.a_tasks[t], jb.b_tasks[t] = a_t[t].data_ptr(), b_t[t].data_ptr()
        jb.a_cat, jb.b_cat, jb.a_cat_t, jb.b_cat_t = (a_cat.data_ptr(), b_cat.data_ptr(), a_cat_t.data_ptr(),
                                                      b_cat_t.data_ptr())
        out.append((a_cat, b_cat, a_cat_t, b_cat_t))
    N.call("mtl_linear_pack_many", arr, len(jobs), N.stream())
    return out


# ----------------------------------------------------------------------------------------------------------------
# MTLoRALinear
# ----------------------------------------------------------------------------------------------------------------
def linear_fwd(spec, x, w_bf16, bias, a_cat, b_cat, *, x_tasks_given=False, act_gelu=False, gelu_grad=False,
               residual=None, path_scale=None, rows_per_sample=0, dropout_p=0.0, seed=0, save_u=False):
    """x: [S_in, M, K] -> y [S_out, M, N], y_act (GELU) or None, u_save [M, R] or None.

    act_gelu: y = pre-activation, y_act = GELU(y); with gelu_grad=True y holds GELU'(pre-activation) instead (pass it as
    `gelu_aux` with aux_is_grad=True to the consuming layer's linear_bwd_input)."""
    _chk(x, BF16, "x"); _chk(w_bf16, BF16, "w_bf16"); _chk(bias, torch.float32, "bias")
    _chk(residual, BF16, "residual"); _chk(path_scale, torch.float32, "path_scale")
    S_in, M, K = x.shape
    if K != spec.K:
        raise ValueError(f"mat1 and mat2 shapes cannot be multiplied ({M}x{K} and {spec.K}x{spec.Nf})")
    want = spec.n_in_streams(x_tasks_given, dropout_p)
    if S_in != want:
        raise ValueError(f"linear_fwd: expected {want} input streams, got {S_in}")
    dev = x.device
    y = torch.empty((spec.S_out, M, spec.Nf), dtype=BF16, device=dev)
    drop = dropout_p > 0 and spec.r_shared > 0
    y_act = torch.empty((spec.S_out + (1 if drop else 0), M, spec.Nf), dtype=BF16, device=dev) if act_gelu else None
    pre = spec.pre_project(M)
    u = torch.empty((M, spec.R_pad), dtype=BF16, device=dev) if ((save_u or pre) and spec.r_shared > 0) else None
    res_streams = 0
    if residual is not None:
        res_streams = residual.shape[0]
    c = spec.cfg(M, x_tasks_given, dropout_p, seed, rows_per_sample, u_precomputed=pre)
    if pre:
        N.call("mtl_linear_rank_project", ctypes.byref(c), 0, N.ptr(x), N.ptr(a_cat), N.ptr(u), N.stream())
    N.call("mtl_linear_fwd", ctypes.byref(c), N.ptr(x), N.ptr(w_bf16), N.ptr(bias), N.ptr(a_cat), N.ptr(b_cat),
           (N.MTL_ACT_GELU_GRAD if gelu_grad else N.MTL_ACT_GELU) if act_gelu else N.MTL_ACT_NONE, N.ptr(y), N.ptr(y_act),
           N.ptr(residual), res_streams,
           N.ptr(path_scale), N.ptr(u), N.stream(),
           meta=("fwd", M, spec.K, spec.Nf, 1 + (spec.T if (x_tasks_given and spec.r_shared > 0) else 0), spec.S_out, spec.R_pad, sum(spec.ranks), bias is not None))
    return y, y_act, u


def linear_bwd_input(spec, dy, wt_bf16, a_cat_t, b_cat_t, *, x_tasks_given=False, gelu_aux=None, aux_is_grad=False,
                     dy_has_sum=False, path_scale=None, rows_per_sample=0, dropout_p=0.0, seed=0, save_g=False,
                     spare_stream=False):
    """dy: [S_out, M, N] (dy_has_sum: [S_out + 1, M, N], last stream = sum of the others, see scale_rows_sum)
    -> dx [1 (+T), M, K], g_save [M, R] or None. spare_stream: dx is returned as [1 (+T) + 1, M, K] with an unwritten
    last stream, for a consumer that wants to append the stream sum in place (LinearEngine.backward, dy_full)."""
    _chk(dy, BF16, "dy"); _chk(wt_bf16, BF16, "wt_bf16"); _chk(gelu_aux, BF16, "gelu_aux")
    S, M, Nf = dy.shape
    if dy_has_sum:
        S -= 1
    if S != spec.S_out or Nf != spec.Nf:
        raise ValueError(f"linear_bwd_input: dy shape {tuple(dy.shape)} does not match the layer ({spec.S_out}, M, {spec.Nf})")
    xt = x_tasks_given and spec.T > 0
    n_dx = 1 + (spec.T if xt else 0)
    dx_full = torch.empty((n_dx + (1 if spare_stream else 0), M, spec.K), dtype=BF16, device=dy.device)
    dx = dx_full[:n_dx]
    pre = spec.pre_project(M) and not dy_has_sum
    g = torch.empty((M, spec.R_pad), dtype=BF16, device=dy.device) if ((save_g or pre) and spec.r_shared > 0) else None
    c = spec.cfg(M, x_tasks_given, dropout_p, seed, rows_per_sample, gelu_aux_is_grad=aux_is_grad, dy_has_sum=dy_has_sum,
                 u_precomputed=pre)
    if pre:
        N.call("mtl_linear_rank_project", ctypes.byref(c), 1, N.ptr(dy), N.ptr(b_cat_t), N.ptr(g), N.stream())
    N.call("mtl_linear_bwd_input", ctypes.byref(c), N.ptr(dy), N.ptr(wt_bf16), N.ptr(a_cat_t), N.ptr(b_cat_t),
           N.ptr(dx), N.ptr(gelu_aux), N.ptr(path_scale), N.ptr(g), N.stream(),
           meta=("bwd_input", M, spec.K, spec.Nf, dx.shape[0], S, spec.R_pad, sum(spec.ranks), False))
    return dx_full, g


def linear_bwd_params(spec, x, dy, u_save, g_save, *, x_tasks_given=False, x_gelu=False, path_scale=None,
                      rows_per_sample=0, dropout_p=0.0, dy_has_sum=False):
    """-> (da_cat [R, K], db_cat [N, R]) fp32, packed like a_cat / b_cat. dy_has_sum: dy is [S_out + 1, M, N] with the
    stream sum last (scale_rows_sum); only the matrixv2 mode reads it here."""
    _chk(x, BF16, "x"); _chk(dy, BF16, "dy"); _chk(u_save, BF16, "u_save"); _chk(g_save, BF16, "g_save")
    M = dy.shape[1]
    if dy.shape[0] != spec.S_out + (1 if dy_has_sum else 0):
        raise ValueError(f"linear_bwd_params: dy has {dy.shape[0]} streams, expected {spec.S_out + (1 if dy_has_sum else 0)}")
    # one zero-fill for both accumulators
    buf = zeros_f32((spec.R_pad * (spec.K + spec.Nf),), dy.device)
    da = buf[:spec.R_pad * spec.K].view(spec.R_pad, spec.K)
    db = buf[spec.R_pad * spec.K:].view(spec.Nf, spec.R_pad)
    c = spec.cfg(M, x_tasks_given, dropout_p, 0, rows_per_sample, dy_has_sum=dy_has_sum)
    N.call("mtl_linear_bwd_params", ctypes.byref(c), N.ptr(x), 1 if x_gelu else 0, N.ptr(dy), N.ptr(u_save),
           N.ptr(g_save), N.ptr(path_scale), N.ptr(da), N.ptr(db), N.stream(),
           meta=("bwd_params", M, spec.K, spec.Nf, x.shape[0], dy.shape[0], spec.R_pad, sum(spec.ranks), False))
    return da, db


def xty(p, q, alpha=1.0, out=None):
    """C[a, b] (+)= alpha * sum_m P[m, a] Q[m, b]; P [M, a], Q [M, b] bf16 row-major -> fp32 [a, b]."""
    _chk(p, BF16, "p"); _chk(q, BF16, "q")
    M, a = p.shape
    b = q.shape[1]
    if out is None:
        out = zeros_f32((a, b), p.device)
    N.call("mtl_xty", N.ptr(p), a, N.ptr(q), b, N.ptr(out), b, M, a, b, float(alpha), N.stream())
    return out


# ----------------------------------------------------------------------------------------------------------------
# window attention / window process
# ----------------------------------------------------------------------------------------------------------------
def window_attention_fwd(qkv, rpb, num_heads, window_size, shift_size, scale, mask=None, dropout_p=0.0, seed=0):
    """qkv [B, H, W, 3C] -> out [1 or 2, B*H*W, C] (second stream = D(out) when dropout_p > 0), lse."""
    _chk(qkv, BF16, "qkv"); _chk(rpb, torch.float32, "relative_position_bias_table"); _chk(mask, torch.float32, "mask")
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    nW = (H // window_size) * (W // window_size)
    out = torch.empty((2 if dropout_p > 0 else 1, B * H * W, C), dtype=BF16, device=qkv.device)
    lse = torch.empty((B * nW, num_heads, 64), dtype=torch.float32, device=qkv.device)
    out_drop = out[1] if dropout_p > 0 else None
    N.call("mtl_window_attention_fwd", N.ptr(qkv), N.ptr(rpb), N.ptr(mask), 0 if mask is None else mask.shape[0],
           N.ptr(out), N.ptr(out_drop), N.ptr(lse), B, H, W, C, num_heads, window_size, shift_size, float(scale),
           float(dropout_p), int(seed) & 0xFFFFFFFFFFFFFFFF, N.stream())
    return out, lse


def window_attention_bwd(qkv, dout, rpb, lse, num_heads, window_size, shift_size, scale, mask=None, want_drpb=True):
    _chk(qkv, BF16, "qkv"); _chk(dout, BF16, "dout"); _chk(lse, torch.float32, "lse")
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    dqkv = torch.empty_like(qkv)
    drpb = zeros_f32(tuple(rpb.shape), rpb.device) if want_drpb else None
    N.call("mtl_window_attention_bwd", N.ptr(qkv), N.ptr(dout), N.ptr(rpb), N.ptr(mask),
           0 if mask is None else mask.shape[0], N.ptr(lse), N.ptr(dqkv), N.ptr(drpb), B, H, W, C, num_heads,
           window_size, shift_size, float(scale), N.stream())
    return dqkv, drpb


def _window_process(name, x, out_shape, B, H, W, C, shift_size, window_size):
    if not x.is_cuda:
        raise RuntimeError("window_process: expected a CUDA tensor — mtlora_b200 has no CPU path")
    if x.element_size() not in (2, 4):
        raise TypeError(f"window_process: unsupported dtype {x.dtype}")
    x = x.contiguous()
    out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
    N.call(name, N.ptr(x), N.ptr(out), B, H, W, C, shift_size, window_size, x.element_size(), N.stream())
    return out


def roll_and_window_partition_forward(x, B, H, W, C, shift_size, window_size):
    nW = (H // window_size) * (W // window_size)
    return _window_process("mtl_roll_and_window_partition_forward", x, (B * nW, window_size, window_size, C), B, H, W,
                           C, shift_size, window_size)


def roll_and_window_partition_backward(g, B, H, W, C, shift_size, window_size):
    return _window_process("mtl_roll_and_window_partition_backward", g, (B, H, W, C), B, H, W, C, shift_size,
                           window_size)


def window_merge_and_roll_forward(x, B, H, W, C, shift_size, window_size):
    return _window_process("mtl_window_merge_and_roll_forward", x, (B, H, W, C), B, H, W, C, shift_size, window_size)


def window_merge_and_roll_backward(g, B, H, W, C, shift_size, window_size):
    nW = (H // window_size) * (W // window_size)
    return _window_process("mtl_window_merge_and_roll_backward", g, (B * nW, window_size, window_size, C), B, H, W, C,
                           shift_size, window_size)


# ----------------------------------------------------------------------------------------------------------------
# LayerNorm (+ PatchMerging gather) and elementwise helpers
# ----------------------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps=1e-5, merge_hw=None, dropout_p=0.0, seed=0, drop_rows=0):
    """x [..., C] (or, merge_hw=(H, W): [n_img, H*W, C/4] token grids gathered 2x2 -> rows of C).

    Returns y [rows, C] — or [rows + drop_rows, C] with D(y[:drop_rows]) appended when dropout_p > 0 —, mean, rstd."""
    _chk(x, BF16, "x"); _chk(gamma, torch.float32, "gamma"); _chk(beta, torch.float32, "beta")
    C = gamma.numel()
    if merge_hw is None:
        rows = x.numel() // C
        H = W = 0
    else:
        H, W = merge_hw
        rows = x.numel() // C
    drop = dropout_p > 0 and drop_rows > 0
    y = torch.empty((rows + (drop_rows if drop else 0), C), dtype=BF16, device=x.device)
    mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
    y_drop = y[rows:] if drop else None
    N.call("mtl_layernorm_fwd", N.ptr(x), N.ptr(gamma), N.ptr(beta), N.ptr(y), N.ptr(y_drop), drop_rows if drop else 0,
           N.ptr(mean), N.ptr(rstd), rows, C, float(eps), 0 if merge_hw is None else 1, H, W, float(dropout_p),
           int(seed) & 0xFFFFFFFFFFFFFFFF, N.stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres=None, merge_hw=None, want_param_grads=True):
    """-> dx (same shape as x; + dres), dgamma, dbeta (fp32)."""
    _chk(dy, BF16, "dy"); _chk(x, BF16, "x"); _chk(dres, BF16, "dres")
    C = gamma.numel()
    rows = mean.numel()
    dx = torch.empty_like(x)
    if want_param_grads:
        gb = zeros_f32((2, C), gamma.device)
        dg, db = gb[0], gb[1]
    else:
        dg = db = None
    H, W = (0, 0) if merge_hw is None else merge_hw
    N.call("mtl_layernorm_bwd", N.ptr(dy), N.ptr(x), N.ptr(gamma), N.ptr(mean), N.ptr(rstd), N.ptr(dres), N.ptr(dx),
           N.ptr(dg), N.ptr(db), rows, C, 0 if merge_hw is None else 1, H, W, N.stream())
    return dx, dg, db


def patch_embed_fwd(x, w, bias, gamma, beta, eps=1e-5, save=True):
    """x [B, 3, H, W] fp32 -> y [B, L, E] bf16 (+ proj, patches, mean, rstd when save) — Conv2d(k4, s4) + bias + LayerNorm."""
    _chk(x, torch.float32, "x"); _chk(w, torch.float32, "weight"); _chk(bias, torch.float32, "bias")
    _chk(gamma, torch.float32, "gamma"); _chk(beta, torch.float32, "beta")
    B, C, H, W = x.shape
    E = w.shape[0]
    if C != 3 or tuple(w.shape[1:]) != (3, 4, 4):
        raise ValueError("patch_embed_fwd: only in_chans=3, patch_size=4 are supported")
    L = (H // 4) * (W // 4)
    dev = x.device
    y = torch.empty((B, L, E), dtype=BF16, device=dev)
    proj = torch.empty((B * L, E), dtype=BF16, device=dev) if save else None
    patches = torch.empty((B * L, 48), dtype=BF16, device=dev) if save else None
    mean = torch.empty((B * L,), dtype=torch.float32, device=dev) if save else None
    rstd = torch.empty((B * L,), dtype=torch.float32, device=dev) if save else None
    N.call("mtl_patch_embed_fwd", N.ptr(x), N.ptr(w), N.ptr(bias), N.ptr(gamma), N.ptr(beta), N.ptr(proj), N.ptr(y),
           N.ptr(patches), N.ptr(mean), N.ptr(rstd), B, H, W, E, float(eps), N.stream())
    return y, proj, patches, mean, rstd


def dropout(x, p, seed):
    _chk(x, BF16, "x")
    y = torch.empty_like(x)
    N.call("mtl_dropout", N.ptr(x), N.ptr(y), x.numel(), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, N.stream())
    return y


def scale_rows(x, scale, rows_per_sample):
    """x [S, M, C] * scale[S, M / rows_per_sample] (per stream, per sample)."""
    _chk(x, BF16, "x"); _chk(scale, torch.float32, "scale")
    S, M, C = x.shape
    y = torch.empty_like(x)
    N.call("mtl_scale_rows", N.ptr(x), N.ptr(scale), N.ptr(y), S, M, C, rows_per_sample, N.stream())
    return y


def scale_rows_sum(x, scale, rows_per_sample):
    """x [S, M, C] -> y [S + 1, M, C]: y[s] = x[s] * scale[s, sample] (scale None: copy), y[S] = sum_s y[s]."""
    _chk(x, BF16, "x"); _chk(scale, torch.float32, "scale")
    S, M, C = x.shape
    y = torch.empty((S + 1, M, C), dtype=BF16, device=x.device)
    N.call("mtl_scale_rows_sum", N.ptr(x), N.ptr(scale), N.ptr(y), S, M, C, int(rows_per_sample), N.stream())
    return y


def add(a, b):
    _chk(a, BF16, "a"); _chk(b, BF16, "b")
    out = torch.empty_like(a)
    N.call("mtl_add", N.ptr(a), N.ptr(b), N.ptr(out), a.numel(), N.stream())
    return out


def sum_streams(x, extra=None, out=None):
    """x [S, ...] -> sum over S (+ extra), into `out` when given."""
    _chk(x, BF16, "x"); _chk(extra, BF16, "extra"); _chk(out, BF16, "out")
    if out is None:
        out = torch.empty(x.shape[1:], dtype=BF16, device=x.device)
    elif out.shape != x.shape[1:]:
        raise ValueError(f"sum_streams: out has shape {tuple(out.shape)}, expected {tuple(x.shape[1:])}")
    N.call("mtl_sum_streams", N.ptr(x), N.ptr(extra), N.ptr(out), x.shape[0], out.numel(), N.stream())
    return out
