"""Host-side mirror of the reference's `models/swin_transformer_mtlora.py` (the Swin backbone whose every
QKV / proj / fc1 / fc2 linear is an MTLoRALinear) on top of libmtlora_b200.so.

Same classes, constructor signatures, parameter / buffer names and `(tensor, {task: tensor} | None)` return
conventions as the reference, so `models/build.py` + `main.py` + `utils.load_checkpoint` work by swapping one import
(INTEGRATION.md). The arithmetic runs in bf16 with fp32 accumulation / statistics inside hand-written sm_100a
kernels; activations are kept "stream-stacked" ([1+T, B*L, C]: shared stream first, then the task streams) so one
kernel launch serves all (1+T) streams of a last-of-stage block and of PatchMerging.

Per block (reference :326-408) the fused path issues 7 launches forward:
    LN1(+LoRA-dropout copy) -> qkv MTLoRALinear -> window attention (roll / partition / reverse folded into its
    loads and stores) -> proj MTLoRALinear (+shortcut, DropPath) -> LN2 over all streams -> fc1 (+GELU) -> fc2
    (+residual, DropPath)
and mirrors them in `_BlockFn.backward`, producing input, adapter, LayerNorm and relative-position-bias gradients.
There is no CPU / PyTorch fallback for this path: CPU tensors raise.
"""
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .lora import AdapterStager, LinearEngine, MTLoRALinear, _new_seed, run_linear_standalone

BF16 = torch.bfloat16


# ----------------------------------------------------------------------------------------------------------------
# tiny local equivalents of the three timm.models.layers symbols the reference imports (:19); timm is not a
# dependency of this package
# ----------------------------------------------------------------------------------------------------------------
def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


trunc_normal_ = nn.init.trunc_normal_


class DropPath(nn.Module):
    """Stochastic depth (timm 0.9.2 semantics): per-sample Bernoulli(keep) mask / keep, identity in eval mode.
    Inside the fused block the same draw is passed to the kernels as a per-sample row scale."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


# ----------------------------------------------------------------------------------------------------------------
# stream-stacking helpers
# ----------------------------------------------------------------------------------------------------------------
def _stacked(tensors, shape=None):
    """[t0, t1, ...] (same shape, bf16) -> one contiguous [S, *shape] tensor; zero-copy when the tensors are already
    consecutive slices of one buffer (the outputs of an upstream fused function), else one torch.stack."""
    t0 = tensors[0]
    n = t0.numel()
    ok = all(t.is_contiguous() and t.dtype == t0.dtype and t.numel() == n for t in tensors)
    if ok:
        es = t0.element_size()
        p0 = t0.data_ptr()
        ok = all(t.untyped_storage().data_ptr() == t0.untyped_storage().data_ptr() and t.data_ptr() == p0 + i * n * es
                 for i, t in enumerate(tensors))
    if ok:
        out = torch.as_strided(t0, (len(tensors),) + tuple(t0.shape), (n,) + tuple(t0.stride()), t0.storage_offset())
    else:
        out = torch.stack([t.to(t0.dtype) for t in tensors])
    if shape is not None:
        out = out.reshape((len(tensors),) + tuple(shape))
    return out


def _grad_stack(grads, like_shape, device):
    """Incoming per-stream gradients (some may be None) -> contiguous bf16 [S, *like_shape]."""
    if all(g is None for g in grads):
        return None
    fixed = []
    for g in grads:
        if g is None:
            fixed = None
            break
        fixed.append(g if g.dtype == BF16 else g.to(BF16))
    if fixed is not None:
        fixed = [g if g.is_contiguous() else g.contiguous() for g in fixed]
        return _stacked(fixed, like_shape)
    out = torch.zeros((len(grads),) + tuple(like_shape), dtype=BF16, device=device)
    for i, g in enumerate(grads):
        if g is not None:
            out[i].copy_(g.reshape(like_shape))
    return out


def _as_bf16_2d(x, C):
    x = x.reshape(-1, C)
    if x.dtype != BF16:
        x = x.to(BF16)
    return x if x.is_contiguous() else x.contiguous()


def _require_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError(f"mtlora_b200.{what}: expected CUDA tensors — this path has no CPU / PyTorch fallback "
                           "(build libmtlora_b200.so and run on the B200)")


def _autocast_cuda_dtype():
    if hasattr(torch, "get_autocast_dtype"):
        return torch.get_autocast_dtype("cuda")
    return torch.get_autocast_gpu_dtype()


def _out_dtype(x):
    """dtype handed back to the caller: the input dtype, or the autocast dtype — what the reference's linears produce
    under `torch.cuda.amp.autocast` (main.py:341: fp16 by default; bf16 in the BASELINE configs). The kernels compute
    in bf16 with fp32 accumulation either way; under fp16 autocast the stage outputs are cast once at the boundary."""
    if torch.is_autocast_enabled():
        return _autocast_cuda_dtype()
    return x.dtype


class _Slot:
    """Placeholder of a tensor that travels through ctx.save_for_backward (index into ctx.saved_tensors)."""
    __slots__ = ("i",)

    def __init__(self, i):
        self.i = i


def _save_state(ctx, state):
    """Store the (nested dict / list) backward state of a fused function: every tensor goes through
    ctx.save_for_backward — so autograd's version-counter check catches in-place writes by downstream code (e.g. a
    `relu_` on a stage output that a later stage saved as its input) and saved-tensor hooks / retain_graph work —
    everything else is kept on ctx as is."""
    tensors = []

    def walk(o):
        if torch.is_tensor(o):
            tensors.append(o)
            return _Slot(len(tensors) - 1)
        if isinstance(o, dict):
            return {k: walk(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(walk(v) for v in o)
        return o
    ctx.state_tree = walk(state)
    ctx.save_for_backward(*tensors)


def _load_state(ctx):
    tensors = ctx.saved_tensors

    def walk(o):
        if isinstance(o, _Slot):
            return tensors[o.i]
        if isinstance(o, dict):
            return {k: walk(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return type(o)(walk(v) for v in o)
        return o
    return walk(ctx.state_tree)


# ----------------------------------------------------------------------------------------------------------------
# CompatLinear — reference :36-41
# ----------------------------------------------------------------------------------------------------------------
class CompatLinear(nn.Linear):
    """nn.Linear returning `(y, None)`; used where a `*_ENABLED` flag of the mtlora config is False."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._engine = LinearEngine(None, self, ops.LinearSpec(self.in_features, self.out_features), None)
        self.lora_dropout_p = 0.0

    @property
    def engine(self):
        return self._engine

    def forward(self, input: Tensor, x_tasks: dict = None):
        y, _ = run_linear_standalone(self._engine, input, None, 0.0, self.training)
        return y, None


def _engine_of(layer):
    eng = getattr(layer, "engine", None)
    if eng is None:
        raise TypeError(f"{type(layer).__name__} is not an mtlora_b200 linear layer")
    return eng


def _layer_p(layer, training):
    """LoRA-dropout probability in effect for `layer` (models/lora.py:79-82, 258)."""
    eng = _engine_of(layer)
    if not training or eng.spec.r_shared == 0:
        return 0.0
    return float(getattr(layer, "lora_dropout_p", 0.0))


# ----------------------------------------------------------------------------------------------------------------
# Mlp — reference :44-81
# ----------------------------------------------------------------------------------------------------------------
class _GeluFn(torch.autograd.Function):
    """Exact-erf GELU on the stand-alone (unfused) path; torch evaluates it, the fused block never comes here."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.nn.functional.gelu(x)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        xf = x.float()
        cdf = 0.5 * (1.0 + torch.erf(xf * 0.7071067811865476))
        pdf = 0.3989422804014327 * torch.exp(-0.5 * xf * xf)
        return (dy.float() * (cdf + xf * pdf)).to(dy.dtype)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0., lora=False,
                 tasks=None, mtlora=None, layer_idx=0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        t = tasks if (lora or mtlora.INTERMEDIATE_SPECIALIZATION) else None
        kw = dict(r=mtlora.R_PER_TASK_LIST[layer_idx], lora_shared_scale=mtlora.SHARED_SCALE[layer_idx],
                  lora_task_scale=mtlora.SCALE_PER_TASK_LIST[layer_idx], lora_dropout=mtlora.DROPOUT[layer_idx], tasks=t,
                  trainable_scale_shared=mtlora.TRAINABLE_SCALE_SHARED,
                  trainable_scale_per_task=mtlora.TRAINABLE_SCALE_PER_TASK, shared_mode=mtlora.SHARED_MODE)
        self.fc1 = MTLoRALinear(in_features, hidden_features, **kw) if mtlora.FC1_ENABLED else \
            CompatLinear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = MTLoRALinear(hidden_features, out_features, **kw) if mtlora.FC2_ENABLED else \
            CompatLinear(hidden_features, out_features)
        self.tasks = tasks
        self.drop = nn.Dropout(drop)

    def forward(self, x, x_tasks=None):
        """Stand-alone path (the fused block calls the engines directly): fc1 -> act -> fc2 on every stream."""
        x, t1 = self.fc1(x, x_tasks)
        x = self.drop(self.act(x))
        if t1 is not None:
            for task in self.tasks:
                t1[task] = self.drop(self.act(t1[task]))
        x, t2 = self.fc2(x, t1)
        x = self.drop(x)
        if t2 is not None:
            for task in self.tasks:
                t2[task] = self.drop(t2[task])
        return x, t2


# ----------------------------------------------------------------------------------------------------------------
# window partition / reverse — reference :84-116 (views + one copy; the fused block never materialises them)
# ----------------------------------------------------------------------------------------------------------------
def window_partition(x, window_size):
    """(B, H, W, C) -> (num_windows*B, window_size, window_size, C)"""
    B, H, W, C = x.shape
    if x.is_cuda and x.element_size() in (2, 4):
        return ops.roll_and_window_partition_forward(x, B, H, W, C, 0, window_size)
    x = x.view(B, H // window_size, window_size, W // window_size, window_size, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, window_size, window_size, C)


def window_reverse(windows, window_size, H, W):
    """(num_windows*B, window_size, window_size, C) -> (B, H, W, C)"""
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    C = windows.shape[-1]
    if windows.is_cuda and windows.element_size() in (2, 4):
        return ops.window_merge_and_roll_forward(windows, B, H, W, C, 0, window_size)
    x = windows.view(B, H // window_size, W // window_size, window_size, window_size, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


# ----------------------------------------------------------------------------------------------------------------
# WindowAttention — reference :119-227
# ----------------------------------------------------------------------------------------------------------------
class _AttnCoreFn(torch.autograd.Function):
    """softmax(q k^T * scale + rpb + mask) v on qkv laid out as (B, H, W, 3C) (mtl_window_attention_fwd/bwd)."""

    @staticmethod
    def forward(ctx, qkv, rpb, mask, num_heads, ws, shift, scale):
        qkv = qkv if qkv.dtype == BF16 else qkv.to(BF16)
        qkv = qkv if qkv.is_contiguous() else qkv.contiguous()
        rpb32 = rpb.detach().float().contiguous()
        m = None if mask is None else mask.detach().float().contiguous()
        out, lse = ops.window_attention_fwd(qkv, rpb32, num_heads, ws, shift, scale, mask=m)
        ctx.save_for_backward(qkv, rpb32, lse, m)
        ctx.cfg = (num_heads, ws, shift, scale)
        B, H, W, C3 = qkv.shape
        return out[0].view(B, H, W, C3 // 3)

    @staticmethod
    def backward(ctx, dout):
        qkv, rpb32, lse, m = ctx.saved_tensors
        num_heads, ws, shift, scale = ctx.cfg
        C = qkv.shape[-1] // 3
        d = _as_bf16_2d(dout, C)
        dqkv, drpb = ops.window_attention_bwd(qkv, d, rpb32, lse, num_heads, ws, shift, scale, mask=m,
                                              want_drpb=ctx.needs_input_grad[1])
        return dqkv, drpb, None, None, None, None, None


class WindowAttention(nn.Module):
    r"""Window based multi-head self attention (W-MSA) with relative position bias; shifted and non-shifted."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0., lora=False,
                 tasks=None, mtlora=None, layer_idx=0):
        super().__init__()
        self.dim = dim
        self.window_size = window_size  # Wh, Ww
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * window_size[0] - 1) * (2 * window_size[1] - 1), num_heads))
        # (yi - yj + Wh-1) * (2 Ww - 1) + (xi - xj + Ww-1), reference :148-160
        idx = torch.arange(window_size[0] * window_size[1])
        yi, xi = idx // window_size[1], idx % window_size[1]
        rel = (yi[:, None] - yi[None, :] + window_size[0] - 1) * (2 * window_size[1] - 1) + \
              (xi[:, None] - xi[None, :] + window_size[1] - 1)
        self.register_buffer("relative_position_index", rel)
        kw = dict(r=mtlora.R_PER_TASK_LIST[layer_idx], lora_shared_scale=mtlora.SHARED_SCALE[layer_idx],
                  lora_task_scale=mtlora.SCALE_PER_TASK_LIST[layer_idx], lora_dropout=mtlora.DROPOUT[layer_idx],
                  trainable_scale_shared=mtlora.TRAINABLE_SCALE_SHARED,
                  trainable_scale_per_task=mtlora.TRAINABLE_SCALE_PER_TASK, shared_mode=mtlora.SHARED_MODE)
        self.qkv = MTLoRALinear(dim, dim * 3, tasks=None, bias=qkv_bias, **kw) if mtlora.QKV_ENABLED else \
            CompatLinear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = MTLoRALinear(dim, dim, tasks=(tasks if (lora or mtlora.INTERMEDIATE_SPECIALIZATION) else None),
                                 **kw) if mtlora.PROJ_ENABLED else CompatLinear(dim, dim)
        self.tasks = tasks
        self.proj_drop = nn.Dropout(proj_drop)
        trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x, mask=None):
        """Stand-alone path. x: (num_windows*B, N, C) already partitioned; mask: (num_windows, N, N) or None."""
        _require_cuda(x, "WindowAttention")
        if self.attn_drop.p > 0 and self.training:
            raise NotImplementedError("mtlora_b200: attn_drop > 0 is not implemented (0 in every shipped config)")
        B_, N, C = x.shape
        ws_h, ws_w = self.window_size
        if ws_h != ws_w or N != ws_h * ws_w:
            raise ValueError(f"WindowAttention: expected square windows of {ws_h}x{ws_w} tokens, got N={N}")
        qkv, _ = self.qkv(x)
        # every window is a (ws, ws) "image" with a single window: no shift, optional explicit mask
        a = _AttnCoreFn.apply(qkv.reshape(B_, ws_h, ws_w, 3 * C), self.relative_position_bias_table, mask,
                              self.num_heads, ws_h, 0, float(self.scale))
        x, t = self.proj(a.reshape(B_, N, C).to(qkv.dtype))
        x = self.proj_drop(x)
        if t is not None:
            for task in self.tasks:
                t[task] = self.proj_drop(t[task])
        return x, t

    def extra_repr(self) -> str:
        return f'dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}'

    def flops(self, N):
        flops = N * self.dim * 3 * self.dim
        flops += self.num_heads * N * (self.dim // self.num_heads) * N
        flops += self.num_heads * N * N * (self.dim // self.num_heads)
        flops += N * self.dim * self.dim
        return flops


# ----------------------------------------------------------------------------------------------------------------
# SwinTransformerBlock — reference :244-429
# ----------------------------------------------------------------------------------------------------------------
def _ln_params(norm):
    return norm.weight, norm.bias


class _BlockFn(torch.autograd.Function):
    """One whole SwinTransformerBlock forward / backward over the C ABI (7 launches forward).

    inputs : x (B, L, C); per-branch DropPath scales ps1 / ps2 ([S, B] fp32 or None); then every parameter of the
             block (so autograd routes their gradients).
    outputs: S = 1 + T tensors (B, L, C) — consecutive slices of one stream-stacked buffer.
    """

    @staticmethod
    def forward(ctx, blk, x, ps1, ps2, *params):
        B, L, C = x.shape
        H, W = blk.input_resolution
        M = B * L
        attn = blk.attn
        e_qkv, e_proj = _engine_of(attn.qkv), _engine_of(attn.proj)
        e_fc1, e_fc2 = _engine_of(blk.mlp.fc1), _engine_of(blk.mlp.fc2)
        tr = blk.training
        p_qkv, p_proj = _layer_p(attn.qkv, tr), _layer_p(attn.proj, tr)
        p_fc1, p_fc2 = _layer_p(blk.mlp.fc1, tr), _layer_p(blk.mlp.fc2, tr)
        s1, s2, s3 = (_new_seed(), _new_seed(), _new_seed()) if max(p_qkv, p_proj, p_fc1, p_fc2) > 0 else (0, 0, 0)
        need = any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)   # unused output streams arrive as None, not as zero tensors
        eps1, eps2 = blk.norm1.eps, blk.norm2.eps
        n1w, n1b = (t.detach().float() for t in _ln_params(blk.norm1))
        n2w, n2b = (t.detach().float() for t in _ln_params(blk.norm2))
        rpb = attn.relative_position_bias_table.detach().float().contiguous()

        xb = _as_bf16_2d(x, C)
        # LN1 (+ D(h) appended for the qkv adapters, lora.py:258)
        h, mean1, rstd1 = ops.layernorm_fwd(xb, n1w, n1b, eps1, dropout_p=p_qkv, seed=s1, drop_rows=M)
        qkv, _, sv_q = e_qkv.forward(h.view(-1, M, C), dropout_p=p_qkv, seed=s1, save=need)
        # window attention in token order: roll / partition / reverse are index math inside the kernel
        a, lse = ops.window_attention_fwd(qkv.view(B, H, W, 3 * C), rpb, attn.num_heads, blk.window_size,
                                          blk.shift_size, float(attn.scale), dropout_p=p_proj, seed=s2)
        # proj + shortcut + DropPath (:389-392): x1[j] = x + ps1[j] * proj_j(a)
        x1, _, sv_p = e_proj.forward(a, residual=xb.view(1, M, C), path_scale=ps1, rows_per_sample=L,
                                     dropout_p=p_proj, seed=s2, save=need)
        S = x1.shape[0]
        xt = S > 1
        # LN2 on every stream (:395-396) (+ D(h2[0]) appended)
        h2, mean2, rstd2 = ops.layernorm_fwd(x1.view(S * M, C), n2w, n2b, eps2, dropout_p=p_fc1, seed=s3, drop_rows=M)
        # g = GELU'(fc1 output) (all the fc2 backward needs of the pre-activation), m = GELU(fc1 output)
        g, m, sv_1 = e_fc1.forward(h2.view(-1, M, C), xt=xt, gelu=True, gelu_grad=True, dropout_p=p_fc1, seed=s3,
                                   save=need)
        y, _, sv_2 = e_fc2.forward(m, xt=xt, residual=x1, path_scale=ps2, rows_per_sample=L, dropout_p=p_fc2,
                                   seed=s3 + 1, save=need)
        if need:
            ctx.blk = blk
            ctx.params = params
            _save_state(ctx, dict(x=xb, mean1=mean1, rstd1=rstd1, sv_q=sv_q, qkv=qkv, lse=lse, rpb=rpb, sv_p=sv_p, x1=x1,
                                  mean2=mean2, rstd2=rstd2, sv_1=sv_1, g=g, sv_2=sv_2, n1w=n1w, n2w=n2w, S=S,
                                  dims=(B, L, C, H, W), in_dtype=x.dtype))
        return tuple(y[j].view(B, L, C) for j in range(S))

    @staticmethod
    def backward(ctx, *dys):
        blk, sv = ctx.blk, _load_state(ctx)
        B, L, C, H, W = sv["dims"]
        M, S = B * L, sv["S"]
        attn = blk.attn
        e_qkv, e_proj = _engine_of(attn.qkv), _engine_of(attn.proj)
        e_fc1, e_fc2 = _engine_of(blk.mlp.fc1), _engine_of(blk.mlp.fc2)
        dy = _grad_stack(dys, (M, C), sv["x"].device)
        grads = {}
        # fc2 -> d(fc1 pre-activation), GELU' fused into the epilogue
        # (multi-stream block: one spare stream behind d(fc1 out), where fc1's backward may append the stream sum)
        dg_full, g2 = e_fc2.backward(sv["sv_2"], dy, gelu_aux=sv["g"], aux_is_grad=True, dx_spare=S > 1)
        dg = dg_full[:S] if S > 1 else dg_full
        grads.update(g2)
        dead = []   # parameters that feed ONLY output streams nobody used: autograd reports None for them, not zeros
        if S > 1 and e_fc2.spec.r_shared > 0 and e_fc2.spec.mode == ops.N.MTL_MODE_MATRIX:
            fc2 = blk.mlp.fc2
            if dys[0] is None:
                # the shared output stream is unused downstream (last stage, reference quirk 8: its
                # fc2.lora_shared_{A,B} receive no gradient at all; every other shared adapter still feeds the frozen
                # product of the task outputs, lora.py:255,262-266)
                dead += [fc2.lora_shared_A, fc2.lora_shared_B]
            for j, t in enumerate(blk.tasks):
                if dys[1 + j] is None:
                    # INTERMEDIATE_SPECIALIZATION (:53,175,544-545): the task streams of every block but the last of a
                    # stage are computed and dropped; their adapters form a chain that reaches no other output
                    for lay in (attn.proj, blk.mlp.fc1, fc2):
                        if getattr(lay, "lora_tasks_A", None) is not None and t in lay.lora_tasks_A:
                            dead += [lay.lora_tasks_A[t], lay.lora_tasks_B[t]]
                        ts = getattr(lay, "lora_task_scale", None)
                        if isinstance(ts, nn.ParameterDict):
                            dead.append(ts[t])
        dh2, g1 = e_fc1.backward(sv["sv_1"], dg, dy_full=dg_full if S > 1 else None)
        grads.update(g1)
        del dg, dg_full
        n2 = blk.norm2
        want_n2 = n2.weight.requires_grad or n2.bias.requires_grad
        dx1, dw2, db2 = ops.layernorm_bwd(dh2.view(S * M, C), sv["x1"].view(S * M, C), sv["n2w"], sv["mean2"],
                                          sv["rstd2"], dres=dy.view(S * M, C), want_param_grads=want_n2)
        if want_n2:
            grads[n2.weight], grads[n2.bias] = dw2, db2
        del dh2, dy
        dx1 = dx1.view(S, M, C)
        da, gp = e_proj.backward(sv["sv_p"], dx1)
        grads.update(gp)
        rpb_p = attn.relative_position_bias_table
        dqkv, drpb = ops.window_attention_bwd(sv["qkv"].view(B, H, W, 3 * C), da[0], sv["rpb"], sv["lse"],
                                              attn.num_heads, blk.window_size, blk.shift_size, float(attn.scale),
                                              want_drpb=rpb_p.requires_grad)
        if rpb_p.requires_grad:
            grads[rpb_p] = drpb
        dh, gq = e_qkv.backward(sv["sv_q"], dqkv.view(1, M, 3 * C))
        grads.update(gq)
        dres = dx1[0] if S == 1 else ops.sum_streams(dx1)   # x feeds the shortcut of every stream (:389-392)
        n1 = blk.norm1
        want_n1 = n1.weight.requires_grad or n1.bias.requires_grad
        dx, dw1, db1 = ops.layernorm_bwd(dh[0], sv["x"], sv["n1w"], sv["mean1"], sv["rstd1"], dres=dres,
                                         want_param_grads=want_n1)
        if want_n1:
            grads[n1.weight], grads[n1.bias] = dw1, db1
        dx = dx.view(B, L, C)
        if sv["in_dtype"] != BF16:
            dx = dx.to(sv["in_dtype"])
        for p in dead:
            grads.pop(p, None)
        out = []
        for p in ctx.params:
            gr = grads.get(p) if p.requires_grad else None
            if gr is not None and gr.dtype != p.dtype:
                gr = gr.to(p.dtype)
            out.append(gr)
        return (None, dx, None, None) + tuple(out)


class SwinTransformerBlock(nn.Module):
    r"""Swin Transformer Block (reference :244-429); `lora=True` blocks also emit the per-task streams."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 fused_window_process=False, lora=False, tasks=None, mtlora=None, layer_idx=0):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.tasks = tasks
        self.lora = lora
        if min(self.input_resolution) <= self.window_size:
            # window larger than the feature map: a single unshifted window (:279-282)
            self.shift_size = 0
            self.window_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"

        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(self.window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop, lora=lora, tasks=tasks,
                                    mtlora=mtlora, layer_idx=layer_idx)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop, lora=lora,
                       tasks=tasks, mtlora=mtlora, layer_idx=layer_idx)

        if self.shift_size > 0:
            # SW-MSA mask (:297-319): 3x3 regions on the rolled grid; kept as a buffer for checkpoint / API parity —
            # the fused kernel evaluates the same 0 / -100 mask analytically from the window position
            H, W = self.input_resolution
            ws, s = self.window_size, self.shift_size

            def rid(n):
                r = torch.zeros(n)
                r[n - ws:n - s] = 1
                r[n - s:] = 2
                return r
            img_mask = (3 * rid(H)[:, None] + rid(W)[None, :]).view(1, H, W, 1)
            mw = img_mask.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
            attn_mask = mw.unsqueeze(1) - mw.unsqueeze(2)
            attn_mask = attn_mask.masked_fill(attn_mask != 0, float(-100.0)).masked_fill(attn_mask == 0, float(0.0))
        else:
            attn_mask = None
        self.register_buffer("attn_mask", attn_mask)
        self.fused_window_process = fused_window_process
        self._param_list = None

    # ---- fused path ----------------------------------------------------------------------------------------------
    def _fused_params(self):
        a, m = self.attn, self.mlp
        engs = (_engine_of(a.qkv), _engine_of(a.proj), _engine_of(m.fc1), _engine_of(m.fc2))
        key = tuple(id(e) for e in engs)      # MTLoRALinear.merge() swaps the engine in effect
        if self._param_list is None or self._param_list[0] != key:
            ps = [self.norm1.weight, self.norm1.bias, a.relative_position_bias_table]
            ps += engs[0].params() + engs[1].params()
            ps += [self.norm2.weight, self.norm2.bias]
            ps += engs[2].params() + engs[3].params()
            self._param_list = (key, ps)
        return self._param_list[1]

    def _fusable(self):
        a, m = self.attn, self.mlp
        if not (type(self.norm1) is nn.LayerNorm and type(self.norm2) is nn.LayerNorm and isinstance(m.act, nn.GELU)):
            return False
        if getattr(m.act, "approximate", "none") != "none":
            return False
        if self.training and (a.attn_drop.p > 0 or a.proj_drop.p > 0 or m.drop.p > 0):
            return False
        if self.dim != 32 * self.num_heads or self.window_size > 8:
            return False
        if any(_engine_of(l).addition for l in (a.qkv, a.proj, m.fc1, m.fc2)):
            return False    # 'addition' mode: LayerNorm of the task sum between the layers -> composed path
        es = [_engine_of(a.proj).spec, _engine_of(m.fc1).spec, _engine_of(m.fc2).spec]
        # all three task-bearing layers agree on the number of streams and on having adapters (every shipped YAML)
        if len({e.S_out for e in es}) != 1 or len({e.r_shared > 0 for e in es[1:]}) != 1:
            return False
        return True

    def path_scales(self, B, S, device):
        """Independent DropPath draws per residual branch and per stream (:389-392, :398-408): 2 x [S, B] fp32."""
        p = self.drop_path.drop_prob if isinstance(self.drop_path, DropPath) else 0.0
        if p == 0.0 or not self.training:
            return None, None
        keep = 1.0 - p
        m = torch.empty((2, S, B), dtype=torch.float32, device=device).bernoulli_(keep)
        if keep > 0.0:
            m.div_(keep)
        return m[0], m[1]

    def forward_streams(self, x, ps=None):
        """Fused path: x (B, L, C) -> tuple of 1+T tensors (B, L, C) (bf16, slices of one stacked buffer)."""
        S = _engine_of(self.attn.proj).spec.S_out
        ps1, ps2 = ps if ps is not None else self.path_scales(x.shape[0], S, x.device)
        return _BlockFn.apply(self, x, ps1, ps2, *self._fused_params())

    # ---- reference-compatible entry ------------------------------------------------------------------------------
    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        _require_cuda(x, "SwinTransformerBlock")
        if self._fusable():
            od = _out_dtype(x)
            ys = self.forward_streams(x)
            ys = [y if y.dtype == od else y.to(od) for y in ys]
            if len(ys) == 1:
                return ys[0], None
            return ys[0], {t: ys[1 + i] for i, t in enumerate(self.tasks)}
        return self._forward_composed(x)

    def _forward_composed(self, x):
        """General (unfused) composition of the sub-modules, same control flow as the reference :326-408; used for
        configurations the fused path does not cover (mixed *_ENABLED flags, non-GELU activations, proj_drop > 0)."""
        H, W = self.input_resolution
        B, L, C = x.shape
        ws, s = self.window_size, self.shift_size
        shortcut = x
        h = self.norm1(x).view(B, H, W, C)
        xw = _WindowProcessFn.apply(h, B, H, W, C, -s, ws)
        aw, aw_t = self.attn(xw.view(-1, ws * ws, C), mask=self.attn_mask)

        def unwin(t):
            t = t.reshape(-1, ws, ws, C)
            return _WindowProcessReverseFn.apply(t, B, H, W, C, s, ws).view(B, H * W, C)
        x_tasks = None
        if aw_t is not None:
            x_tasks = {t: shortcut + self.drop_path(unwin(aw_t[t])) for t in self.tasks}
        x = shortcut + self.drop_path(unwin(aw))
        m, m_t = self.mlp(self.norm2(x), None if x_tasks is None else {t: self.norm2(x_tasks[t]) for t in self.tasks})
        if m_t is None:
            return x + self.drop_path(m), None
        if x_tasks is None:
            return x + self.drop_path(m), {t: self.drop_path(m_t[t]) for t in self.tasks}
        return x + self.drop_path(m), {t: x_tasks[t] + self.drop_path(m_t[t]) for t in self.tasks}

    def extra_repr(self) -> str:
        return f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, " \
               f"window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}"

    def flops(self):
        H, W = self.input_resolution
        flops = self.dim * H * W
        nW = H * W / self.window_size / self.window_size
        flops += nW * self.attn.flops(self.window_size * self.window_size)
        flops += 2 * H * W * self.dim * self.dim * self.mlp_ratio
        flops += self.dim * H * W
        return flops


class _WindowProcessFn(torch.autograd.Function):
    """kernels/window_process WindowProcess (window_process.py:11-35): roll(shift) + window_partition."""

    @staticmethod
    def forward(ctx, x, B, H, W, C, shift_size, window_size):
        ctx.cfg = (B, H, W, C, shift_size, window_size)
        return ops.roll_and_window_partition_forward(x, B, H, W, C, shift_size, window_size)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C, shift_size, window_size = ctx.cfg
        return ops.roll_and_window_partition_backward(g, B, H, W, C, shift_size, window_size), None, None, None, None, None, None


class _WindowProcessReverseFn(torch.autograd.Function):
    """kernels/window_process WindowProcessReverse (window_process.py:38-63): window_reverse + roll(shift)."""

    @staticmethod
    def forward(ctx, x, B, H, W, C, shift_size, window_size):
        ctx.cfg = (B, H, W, C, shift_size, window_size)
        return ops.window_merge_and_roll_forward(x, B, H, W, C, shift_size, window_size)

    @staticmethod
    def backward(ctx, g):
        B, H, W, C, shift_size, window_size = ctx.cfg
        return ops.window_merge_and_roll_backward(g, B, H, W, C, shift_size, window_size), None, None, None, None, None, None


WindowProcess = _WindowProcessFn
WindowProcessReverse = _WindowProcessReverseFn


# ----------------------------------------------------------------------------------------------------------------
# PatchMerging — reference :432-483
# ----------------------------------------------------------------------------------------------------------------
class _PatchMergeFn(torch.autograd.Function):
    """2x2 gather + LayerNorm(4C) + reduction for S stacked streams in one pass (3 launches forward)."""

    @staticmethod
    def forward(ctx, pm, n_streams, *args):
        xs, params = args[:n_streams], args[n_streams:]
        H, W = pm.input_resolution
        B, L, C = xs[0].shape
        S = n_streams
        xin = [_as_bf16_2d(x, C) for x in xs]
        x = _stacked(xin, (B * L, C))
        eng = _engine_of(pm.reduction)
        p = _layer_p(pm.reduction, pm.training)
        seed = _new_seed() if p > 0 else 0
        rows = S * B * L // 4
        nw, nb = pm.norm.weight.detach().float(), pm.norm.bias.detach().float()
        need = any(ctx.needs_input_grad)
        ctx.set_materialize_grads(False)
        h, mean, rstd = ops.layernorm_fwd(x, nw, nb, pm.norm.eps, merge_hw=(H, W), dropout_p=p, seed=seed, drop_rows=rows)
        y, _, sv = eng.forward(h.view(-1, rows, 4 * C), dropout_p=p, seed=seed, save=need)
        if need:
            ctx.pm, ctx.params = pm, params
            _save_state(ctx, dict(x=x, mean=mean, rstd=rstd, nw=nw, sv=sv, dims=(S, B, L, C, H, W),
                                  in_dtypes=[t.dtype for t in xs]))
        y = y.view(S, B, L // 4, 2 * C)
        return tuple(y[j] for j in range(S))

    @staticmethod
    def backward(ctx, *dys):
        pm, sv = ctx.pm, _load_state(ctx)
        S, B, L, C, H, W = sv["dims"]
        rows = S * B * L // 4
        eng = _engine_of(pm.reduction)
        dy = _grad_stack(dys, (B * L // 4, 2 * C), sv["x"].device)
        dh, grads = eng.backward(sv["sv"], dy.view(1, rows, 2 * C))
        want_n = pm.norm.weight.requires_grad or pm.norm.bias.requires_grad
        dx, dw, db = ops.layernorm_bwd(dh[0], sv["x"], sv["nw"], sv["mean"], sv["rstd"], merge_hw=(H, W),
                                       want_param_grads=want_n)
        if want_n:
            grads[pm.norm.weight], grads[pm.norm.bias] = dw, db
        dx = dx.view(S, B, L, C)
        dxs = tuple(dx[j] if sv["in_dtypes"][j] == BF16 else dx[j].to(sv["in_dtypes"][j]) for j in range(S))
        out = []
        for p in ctx.params:
            gr = grads.get(p) if p.requires_grad else None
            if gr is not None and gr.dtype != p.dtype:
                gr = gr.to(p.dtype)
            out.append(gr)
        return (None, None) + dxs + tuple(out)


class PatchMerging(nn.Module):
    r"""Patch Merging Layer: 2x2 neighbourhood gather (channel order (0,0),(1,0),(0,1),(1,1)) -> LN(4C) -> 4C->2C."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm, layer_idx=0, mtlora=None):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        if mtlora.DOWNSAMPLER_ENABLED:
            self.reduction = MTLoRALinear(4 * dim, 2 * dim, r=mtlora.R_PER_TASK_LIST[layer_idx],
                                          lora_shared_scale=mtlora.SHARED_SCALE[layer_idx],
                                          lora_task_scale=mtlora.SCALE_PER_TASK_LIST[layer_idx],
                                          lora_dropout=mtlora.DROPOUT[layer_idx], tasks=None, bias=False,
                                          trainable_scale_shared=mtlora.TRAINABLE_SCALE_SHARED,
                                          trainable_scale_per_task=mtlora.TRAINABLE_SCALE_PER_TASK,
                                          shared_mode=mtlora.SHARED_MODE)
        else:
            self.reduction = CompatLinear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward_streams(self, xs):
        """xs: list of S tensors (B, L, C) -> tuple of S tensors (B, L/4, 2C), one pass over all streams."""
        H, W = self.input_resolution
        B, L, C = xs[0].shape
        assert L == H * W, "input feature has wrong size"
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        _require_cuda(xs[0], "PatchMerging")
        if type(self.norm) is not nn.LayerNorm:
            raise NotImplementedError("mtlora_b200.PatchMerging supports norm_layer=nn.LayerNorm only")
        params = [self.norm.weight, self.norm.bias] + _engine_of(self.reduction).params()
        return _PatchMergeFn.apply(self, len(xs), *xs, *params)

    def forward(self, x):
        """x: B, H*W, C"""
        od = _out_dtype(x)
        (y,) = self.forward_streams([x])
        return y if y.dtype == od else y.to(od)

    def extra_repr(self) -> str:
        return f"input_resolution={self.input_resolution}, dim={self.dim}"

    def flops(self):
        H, W = self.input_resolution
        flops = H * W * self.dim
        flops += (H // 2) * (W // 2) * 4 * self.dim * 2 * self.dim
        return flops


# ----------------------------------------------------------------------------------------------------------------
# BasicLayer — reference :486-565
# ----------------------------------------------------------------------------------------------------------------
class BasicLayer(nn.Module):
    """One Swin stage: `depth` blocks (only the last one has lora=True and emits task streams) + PatchMerging."""

    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 fused_window_process=False, tasks=None, mtlora=None, layer_idx=0):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.tasks = tasks
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads,
                                 window_size=window_size, shift_size=0 if (i % 2 == 0) else window_size // 2,
                                 mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                                 attn_drop=attn_drop,
                                 drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                 norm_layer=norm_layer, fused_window_process=fused_window_process,
                                 lora=(i == depth - 1), tasks=tasks, mtlora=mtlora, layer_idx=layer_idx)
            for i in range(depth)])
        if downsample is not None:
            self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer, layer_idx=layer_idx,
                                         mtlora=mtlora)
        else:
            self.downsample = None

    def forward_streams(self, x):
        """x (B, L, C) -> tuple of streams after the last block and the (batched) downsample; None when some block
        cannot take the fused path."""
        if not all(b._fusable() for b in self.blocks):
            return None
        ys = (x,)
        for blk in self.blocks:
            ys = blk.forward_streams(ys[0])       # only the shared stream feeds the next block (:544-545)
        if self.downsample is not None:
            ys = self.downsample.forward_streams(list(ys))
        return ys

    def forward(self, x):
        _require_cuda(x, "BasicLayer")
        od = _out_dtype(x)
        ys = self.forward_streams(x)
        if ys is not None:
            ys = [y if y.dtype == od else y.to(od) for y in ys]
            if len(ys) == 1:
                return ys[0], None
            return ys[0], {t: ys[1 + i] for i, t in enumerate(self.tasks)}
        tasks_lora = None
        for blk in self.blocks:
            x, tasks_lora = blk(x)
        if self.downsample is not None:
            if tasks_lora is not None:
                outs = self.downsample.forward_streams([x] + [tasks_lora[t] for t in self.tasks])
                outs = [o if o.dtype == od else o.to(od) for o in outs]
                x, tasks_lora = outs[0], {t: outs[1 + i] for i, t in enumerate(self.tasks)}
            else:
                x = self.downsample(x)
        return x, tasks_lora

    def extra_repr(self) -> str:
        return f"dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}"

    def flops(self):
        flops = 0
        for blk in self.blocks:
            flops += blk.flops()
        if self.downsample is not None:
            flops += self.downsample.flops()
        return flops


# ----------------------------------------------------------------------------------------------------------------
# PatchEmbed — reference :568-621 (stays PyTorch: Conv2d + LayerNorm, not on the replaced path)
# ----------------------------------------------------------------------------------------------------------------
class _LayerNormFn(torch.autograd.Function):
    """Stand-alone LayerNorm over the last dim through mtl_layernorm_fwd / _bwd (bf16 rows, fp32 statistics)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        C = x.shape[-1]
        x2 = x.reshape(-1, C)
        w, b = weight.detach().float(), bias.detach().float()
        y, mean, rstd = ops.layernorm_fwd(x2, w, b, eps)
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.want = weight.requires_grad or bias.requires_grad
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        dy2 = dy.reshape(x2.shape)
        if dy2.dtype != BF16 or not dy2.is_contiguous():
            dy2 = dy2.to(BF16).contiguous()
        dx, dw, db = ops.layernorm_bwd(dy2, x2, w, mean, rstd, want_param_grads=ctx.want)
        return dx.view(dy.shape), dw, db, None


class _PatchEmbedFn(torch.autograd.Function):
    """Conv2d(3, E, k4, s4) + bias + LayerNorm in one kernel (mtl_patch_embed_fwd); backward = mtl_layernorm_bwd on the
    saved projection, dW = d_proj^T patches and db = column sums of d_proj through mtl_xty. The image gets no gradient."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, eps):
        need = any(ctx.needs_input_grad[1:])
        y, proj, patches, mean, rstd = ops.patch_embed_fwd(x.contiguous(), w.detach().float().contiguous(),
                                                           b.detach().float(), gamma.detach().float(),
                                                           beta.detach().float(), eps, save=need)
        if need:
            ctx.save_for_backward(proj, patches, mean, rstd, gamma.detach().float())
            ctx.wshape = w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        proj, patches, mean, rstd, gamma = ctx.saved_tensors
        E = proj.shape[1]
        dy2 = dy.reshape(-1, E)
        if dy2.dtype != BF16 or not dy2.is_contiguous():
            dy2 = dy2.to(BF16).contiguous()
        dproj, dgam, dbet = ops.layernorm_bwd(dy2, proj, gamma, mean, rstd)
        dw = ops.xty(dproj, patches).view(ctx.wshape)                      # [E, 48] = d_proj^T patches
        ones = torch.ones((dproj.shape[0], 8), dtype=BF16, device=dproj.device)
        db = ops.xty(dproj, ones)[:, 0].contiguous()
        return None, dw, db, dgam, dbet, None


class PatchEmbed(nn.Module):
    r"""Image to Patch Embedding (reference :568-611: Conv2d k4 s4 + LayerNorm; the convolution stays PyTorch / cuDNN).

    Under bf16 autocast on CUDA the standard geometry (patch 4, 3 input channels, embed_dim 96 / 128) runs as ONE
    kernel (mtl_patch_embed_fwd: projection + bias + LayerNorm, bf16 tokens); other geometries run the convolution
    channels-last, so its output already is the (B, L, C) token matrix, with the LayerNorm through mtl_layernorm."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.img_size = img_size
        self.patch_size = patch_size
        self.patches_resolution = patches_resolution
        self.num_patches = patches_resolution[0] * patches_resolution[1]
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        # any autocast dtype: the kernels compute in bf16 with fp32 accumulation (the backbone casts its stage outputs to
        # the autocast dtype at its boundary, see _out_dtype)
        bf16_autocast = x.is_cuda and torch.is_autocast_enabled()
        if (bf16_autocast and type(self.norm) is nn.LayerNorm and x.dtype == torch.float32 and C == 3
                and not x.requires_grad   # the fused kernel produces no image gradient
                and tuple(self.patch_size) == (4, 4) and self.embed_dim in (96, 128) and self.proj.bias is not None):
            # one kernel: patch projection + bias + LayerNorm, bf16 tokens out
            return _PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, self.norm.weight, self.norm.bias,
                                       self.norm.eps)
        if bf16_autocast and type(self.norm) is nn.LayerNorm:
            y = self.proj(x.contiguous(memory_format=torch.channels_last))   # (B, C, Ph, Pw) bf16, NHWC strides
            y = y.permute(0, 2, 3, 1).contiguous()                           # no copy when the conv answered NHWC
            y = y.view(B, -1, self.embed_dim)
            if y.dtype != BF16:
                y = y.to(BF16)
            return _LayerNormFn.apply(y, self.norm.weight, self.norm.bias, self.norm.eps)
        x = self.proj(x).flatten(2).transpose(1, 2)  # B Ph*Pw C
        if self.norm is not None:
            x = self.norm(x)
        return x

    def flops(self):
        Ho, Wo = self.patches_resolution
        flops = Ho * Wo * self.embed_dim * self.in_chans * (self.patch_size[0] * self.patch_size[1])
        if self.norm is not None:
            flops += Ho * Wo * self.embed_dim
        return flops


# ----------------------------------------------------------------------------------------------------------------
# SwinTransformerMTLoRA — reference :624-780
# ----------------------------------------------------------------------------------------------------------------
class SwinTransformerMTLoRA(nn.Module):
    r"""Swin Transformer backbone with MTLoRA adapters; `forward(x, return_stages=True)` returns, per stage,
    `(x_s, {task: x_{s,t}})` for the dense-prediction heads (models/swin_mtl.py:224-231)."""

    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, fused_window_process=False, basic_layer=BasicLayer, tasks=None, mtlora=None,
                 **kwargs):
        super().__init__()
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio
        self.tasks = tasks
        self.mtlora = mtlora
        if mtlora is not None:
            print("\nMTLoRA params:")
            print(mtlora)
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        num_patches = self.patch_embed.num_patches
        patches_resolution = self.patch_embed.patches_resolution
        self.patches_resolution = patches_resolution
        if self.ape:
            self.absolute_pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
            trunc_normal_(self.absolute_pos_embed, std=.02)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]  # stochastic depth decay rule
        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            layer = basic_layer(dim=int(embed_dim * 2 ** i_layer),
                                input_resolution=(patches_resolution[0] // (2 ** i_layer),
                                                  patches_resolution[1] // (2 ** i_layer)),
                                depth=depths[i_layer], num_heads=num_heads[i_layer], window_size=window_size,
                                mlp_ratio=self.mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                                attn_drop=attn_drop_rate,
                                drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])], norm_layer=norm_layer,
                                downsample=PatchMerging if (i_layer < self.num_layers - 1) else None,
                                use_checkpoint=use_checkpoint, fused_window_process=fused_window_process, tasks=tasks,
                                mtlora=self.mtlora, layer_idx=i_layer)
            self.layers.append(layer)
        self.avgpool = nn.AdaptiveAvgPool1d(1)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def _attach_grad_sync(self):
        """Data-parallel runs (torch.distributed initialised, world size > 1): average the trainable gradients of the
        backbone over the ranks from autograd hooks, so `main.py`'s loop needs no change (mtlora_b200/dist.py). Cheap
        to call every forward: parameters already covered are skipped."""
        from . import dist as _dist
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()) or \
                torch.distributed.get_world_size() == 1 or not _dist.auto_sync_enabled():
            return
        # every step starts here with an idle GPU when the loop reads its loss back (main.py:359-361): walking the module
        # tree for uncovered parameters cost 0.4 ms of that. The parameter list is cached; a re-scan happens only when the
        # number of trainable parameters changes (mark_only_lora_as_trainable / requires_grad_ after the first forward).
        ps = self.__dict__.get("_gs_params")
        if ps is None:
            ps = self.__dict__["_gs_params"] = list(self.parameters())
            self.__dict__["_gs_trainable"] = -1
        n_train = sum([p.requires_grad for p in ps])
        if n_train == self.__dict__["_gs_trainable"]:
            return
        gs = _dist.sync_gradients(self)
        if gs is not None:
            self.__dict__.setdefault("_grad_syncs", []).append(gs)
        self.__dict__["_gs_trainable"] = n_train

    def _stage_adapters(self):
        """Re-pack the bf16 adapter operands of all layers in one launch when an optimizer step changed them."""
        st = self.__dict__.get("_adapter_stager")
        if st is None:
            st = AdapterStager(m for m in self.modules() if m is not self and hasattr(type(m), "engine"))
            self.__dict__["_adapter_stager"] = st
        return st.refresh()

    def forward_features(self, x, return_stages=False, flatten_ft=False):
        if self.training and torch.is_grad_enabled():
            self._attach_grad_sync()
        staging_due = x.is_cuda
        x = self.patch_embed(x)
        if staging_due:
            # after the patch embedding has been launched: a loop that reads its loss back every step starts each step
            # with an idle GPU, so the first launch should leave the host as early as possible (the patch embedding needs
            # no adapters). (Launching the packing on a side stream UNDER the patch embedding was measured 0.2 ms per step
            # slower: the two kernels slow each other down by more than the 90 us they hide.)
            self._stage_adapters()
        if self.ape:
            x = x + self.absolute_pos_embed
        x = self.pos_drop(x)
        _require_cuda(x, "SwinTransformerMTLoRA")
        od = _out_dtype(x)
        if x.dtype != BF16:
            x = x.to(BF16)   # the stages run in bf16; cast once here instead of once per block
        out = []
        for layer in self.layers:
            ys = layer.forward_streams(x) if hasattr(layer, "forward_streams") else None
            if ys is None:
                x, tasks_lora = layer(x)
            else:
                x = ys[0]
                tasks_lora = None if len(ys) == 1 else {t: ys[1 + i] for i, t in enumerate(self.tasks)}
            if tasks_lora is None:
                tasks_lora = {task: x for task in self.tasks}
            if return_stages:
                cast = (lambda t: t) if od == BF16 else (lambda t: t.to(od))
                out.append((cast(x), {k: cast(v) for k, v in tasks_lora.items()}))
        if return_stages:
            return out
        x = x if x.dtype == od else x.to(od)
        if flatten_ft:
            x = self.avgpool(x.transpose(1, 2))  # B C 1
            x = torch.flatten(x, 1)
        return x

    def forward(self, x, return_stages=False, flatten_ft=False):
        x = self.forward_features(x, return_stages, flatten_ft)
        x = self.head(x)
        return x

    def flops(self, images=None, logger=None, detailed=False):
        flops = self.patch_embed.flops()
        for layer in self.layers:
            flops += layer.flops()
        flops += self.num_features * self.patches_resolution[0] * self.patches_resolution[1] // (2 ** self.num_layers)
        flops += self.num_features * self.num_classes
        return flops
