"""Host-side mirror of the reference's `models/lora.py` for the hot path: `MTLoRALinear` (reference :159-284),
`LoRALayer` (:62-84), `mark_only_lora_as_trainable` (:580-630) and `map_old_state_dict_weights` (:644-668).

Same constructor arguments, parameter names / shapes / registration order (checkpoint surface), `(tensor,
dict | None)` return convention and error behaviour as the reference; the arithmetic runs in
libmtlora_b200.so (`mtl_linear_fwd` / `mtl_linear_bwd_input` / `mtl_linear_bwd_params`, include/mtlora_b200.h) in
bf16 with fp32 accumulation. There is no PyTorch fallback: CPU tensors raise.
"""
import math
import os
from typing import Any, Dict, Mapping, Optional, Union

import torch
import torch.nn as nn

from . import ops

BF16 = torch.bfloat16
# input-gradient width (in_features) above which a multi-stream layer takes the stream sum of dy appended in place by
# its caller (LinearEngine.backward, dy_full): 128 = one column chunk; tuning aid
SUM_IN_PLACE_MIN_K = int(os.environ.get("MTL_SUM_IN_PLACE_MIN_K", "128"))


class LoRALayer(nn.Module):
    """Reference models/lora.py:62-84: stores r, the scale ("lora_alpha") and the LoRA-branch dropout."""

    def __init__(self, r: int, lora_alpha: float, lora_dropout: float):
        super().__init__()
        assert r >= 0
        self.r = r
        self.lora_alpha = lora_alpha
        self.lora_dropout_p = float(lora_dropout)
        self.lora_dropout = nn.Dropout(p=lora_dropout) if lora_dropout > 0.0 else (lambda x: x)
        self.merged = False


def _new_seed():
    """Seed of the counter-based dropout mask, drawn from torch's CPU generator (so torch.manual_seed governs it)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class LinearEngine:
    """Per-layer staging + explicit forward / backward over the C ABI (not an autograd object).

    Used by `MTLoRALinear` / `CompatLinear` standalone (through `_LinearFn`) and by the fused block / patch-merging
    functions in swin_transformer_mtlora.py, which call `forward` / `backward` directly from their own backward.
    """

    def __init__(self, owner, linear: nn.Linear, spec: ops.LinearSpec, tasks):
        self.owner = owner          # the nn.Module holding lora_* parameters (or None for a plain linear)
        self.linear = linear
        self.spec = spec
        self.tasks = list(tasks) if tasks else []
        # shared_mode 'addition' (lora.py:275-282): no shared adapter; the kernels run with an all-zero rank-1 stand-in
        # so that stream 0 carries the frozen product, the LayerNorm-of-the-task-sum tail is `_AdditionTailFn`
        self.addition = owner is not None and getattr(owner, "shared_mode", "") == "addition" and spec.r_shared > 0
        self._zero = None
        self.invalidate()

    def invalidate(self):
        """Drop the staged bf16 operand copies (frozen W / W^T, packed adapters). They are keyed on (data_ptr, _version)
        of the fp32 masters, which in-place ops through `.data` (`p.data.copy_()`, weight surgery, EMA scripts) do NOT
        bump: call this after such writes (load_state_dict / optimizer steps — torch.optim or FlatAdamW — need nothing)."""
        self._wkey = None
        self._w = self._wt = None
        self._akey = None
        self._packed = (None, None, None, None)

    # ---- parameter staging -------------------------------------------------------------------------------------
    def adapters(self):
        o = self.owner
        if self.spec.r_shared == 0:
            return []
        if self.addition:
            w = self.linear.weight
            if self._zero is None or self._zero[0].device != w.device:
                self._zero = (torch.zeros((1, self.spec.K), device=w.device), torch.zeros((self.spec.Nf, 1), device=w.device))
            ps = [self._zero[0], self._zero[1]]
        else:
            ps = [o.lora_shared_A, o.lora_shared_B]
        ps += [o.lora_tasks_A[t] for t in self.tasks] + [o.lora_tasks_B[t] for t in self.tasks]
        return ps

    def scales(self):
        """Trainable LoRA scales (lora.py:210-216, 229-233) as (adapter index, Parameter) pairs: 0 = shared, 1+t = task."""
        o = self.owner
        if o is None or self.spec.r_shared == 0:
            return []
        out = []
        if isinstance(getattr(o, "lora_shared_scale", None), nn.Parameter) and not self.addition:
            out.append((0, o.lora_shared_scale))
        ts = getattr(o, "lora_task_scale", None)
        if isinstance(ts, nn.ParameterDict):
            out += [(1 + i, ts[t]) for i, t in enumerate(self.tasks)]
        return out

    def params(self):
        """Parameters in the order `backward` reports gradients: weight, bias?, shared A, B, task A..., task B..., scales."""
        lin = self.linear
        return ([lin.weight] + ([lin.bias] if lin.bias is not None else [])
                + [p for p in self.adapters() if isinstance(p, nn.Parameter)] + [p for _, p in self.scales()])

    def _adapter_key(self):
        ad = self.adapters()
        sc = self.scales()
        return ad, sc, tuple((p.data_ptr(), p._version) for p in ad + [q for _, q in sc])

    def stage(self):
        w = self.linear.weight
        if not w.is_cuda:
            raise RuntimeError("mtlora_b200: parameters must live on a CUDA device — there is no CPU path")
        key = (w.data_ptr(), w._version)
        if key != self._wkey:
            with torch.no_grad():
                self._w, self._wt = ops.cast_transpose(w.detach().float().contiguous())
            self._wkey = key
        if self.spec.r_shared > 0:
            ad = self.adapters()
            sc = self.scales()
            akey = tuple((p.data_ptr(), p._version) for p in ad + [q for _, q in sc])
            if akey != self._akey:
                T = len(self.tasks)
                with torch.no_grad():
                    d = [p.detach().float().contiguous() for p in ad]
                    # a trainable scale is folded into the packed B operand (the kernels then run with scale 1): no
                    # device -> host read of the parameter, forward / input gradient / dA exact as they stand
                    for i, q in sc:
                        j = 1 if i == 0 else 2 + T + (i - 1)
                        d[j] = d[j] * q.detach().float()
                    self._packed = ops.pack_adapters(self.spec, d[0], d[1], d[2:2 + T], d[2 + T:2 + 2 * T])
                self._akey = akey
        return self._w, self._wt, self._packed

    # ---- explicit forward / backward ---------------------------------------------------------------------------
    def forward(self, x, *, xt=False, gelu=False, gelu_grad=False, residual=None, path_scale=None, rows_per_sample=0,
                dropout_p=0.0, seed=0, save=True):
        """x [S_in, M, K] bf16 (incl. the appended D(x[0]) stream when dropout_p > 0) -> y, y_act, saved-dict."""
        staged = self.stage()
        w, _, (a_cat, b_cat, _, _) = staged
        bias = self.linear.bias
        bias = None if bias is None else bias.detach().float()
        y, y_act, u = ops.linear_fwd(self.spec, x, w, bias, a_cat, b_cat, x_tasks_given=xt, act_gelu=gelu, gelu_grad=gelu_grad,
                                     residual=residual, path_scale=path_scale, rows_per_sample=rows_per_sample,
                                     dropout_p=dropout_p, seed=seed, save_u=save)
        saved = None
        if save:
            # `staged`: the bf16 operand copies this forward ran with — backward uses the same ones (the masters cannot
            # have changed legitimately in between) instead of re-validating ~10 (data_ptr, _version) keys per layer
            saved = dict(x=x, u=u, xt=xt, dropout_p=dropout_p, seed=seed, path_scale=path_scale,
                         rows_per_sample=rows_per_sample, staged=staged)
        return y, y_act, saved

    def backward(self, saved, dy, *, gelu_aux=None, aux_is_grad=False, need_dx=True, dy_full=None, dx_spare=False):
        """dy [S_out, M, N] -> dx [1 (+T), M, K], {param: fp32 grad} for every parameter that requires grad.

        DropPath (`path_scale` given in forward): a single-stream layer scales inside the kernels, a multi-stream
        layer pre-scales dy once (header contract of mtl_linear_bwd_input).
        dx_spare: dx comes back as [1 (+T) + 1, M, K] with a free last stream. dy_full: such a buffer whose leading
        streams ARE `dy` — the stream sum the frozen product needs is then appended in place (one read of dy, one
        stream written) instead of re-streaming all 1+T operand tiles for every column chunk."""
        spec = self.spec
        _, wt, (_, _, a_cat_t, b_cat_t) = saved.get("staged") or self.stage()
        ps, rps = saved["path_scale"], saved["rows_per_sample"]
        dy_in, dy_sum = dy, False
        v2 = spec.mode == ops.N.MTL_MODE_MATRIXV2 and spec.S_out > 1   # shared adapter sees the gradient of every stream
        if 1 < spec.S_out < 8 and (ps is not None or spec.K >= 2 * spec.Nf or v2):
            # hand the kernel sum_j dy[j] as one extra stream so the frozen product streams one operand tile per column
            # chunk instead of 1+T: worth a pass of its own when there are many chunks (fc2: K = 4 N), free when the
            # DropPath pre-scale pass is needed anyway (it writes the sum along)
            dy_in = ops.scale_rows_sum(dy, ps, rps)
            dy, dy_sum, ps = dy_in[:spec.S_out], True, None
        elif (dy_full is not None and 1 < spec.S_out < 8 and ps is None and spec.K > SUM_IN_PLACE_MIN_K
              and dy_full.shape[0] == spec.S_out + 1 and dy_full.data_ptr() == dy.data_ptr()):
            # fc1 of a multi-stream block (K = C outputs of the input gradient in 2+ column chunks): see dy_full above
            ops.sum_streams(dy, out=dy_full[spec.S_out])
            dy_in, dy_sum = dy_full, True
        elif ps is not None and spec.S_out > 1:
            dy = dy_in = ops.scale_rows(dy, ps, rps)
            ps = None
        lin = self.linear
        ad = self.adapters()
        want_ad = any(p.requires_grad for p in ad)
        dx, g = ops.linear_bwd_input(spec, dy_in, wt, a_cat_t, b_cat_t, x_tasks_given=saved["xt"], gelu_aux=gelu_aux,
                                     aux_is_grad=aux_is_grad, dy_has_sum=dy_sum,
                                     path_scale=ps, rows_per_sample=rps if ps is not None else 0,
                                     dropout_p=saved["dropout_p"], seed=saved["seed"], save_g=want_ad,
                                     spare_stream=dx_spare)
        grads = {}
        if want_ad:
            da, db = ops.linear_bwd_params(spec, saved["x"], dy_in if (v2 and dy_sum) else dy, saved["u"], g,
                                           x_tasks_given=saved["xt"], path_scale=ps,
                                           rows_per_sample=rps if ps is not None else 0, dropout_p=saved["dropout_p"],
                                           dy_has_sum=v2 and dy_sum)
            T = len(self.tasks)
            tscale = dict(self.scales())
            # dB of the task adapters as ONE strided copy into a [T, N, r] block (autograd would otherwise clone each
            # non-contiguous column slice of db_cat on its own: 4 small copy kernels per layer)
            db_tasks = None
            if T > 1 and len(set(spec.ranks[1:])) == 1:
                r_t, step_t = spec.ranks[1], spec.offsets[2] - spec.offsets[1]
                if all(spec.offsets[1 + t] == spec.offsets[1] + t * step_t for t in range(T)):
                    db_tasks = db[:, spec.offsets[1]:spec.offsets[1] + T * step_t].view(spec.Nf, T, step_t)[:, :, :r_t] \
                        .permute(1, 0, 2).contiguous()
            for i in range(1 + T):
                off, r = spec.offsets[i], spec.ranks[i]
                pa = ad[0] if i == 0 else ad[2 + (i - 1)]
                pb = ad[1] if i == 0 else ad[2 + T + (i - 1)]
                if pa.requires_grad:
                    grads[pa] = da[off:off + r]
                dbi = db_tasks[i - 1] if (i > 0 and db_tasks is not None) else db[:, off:off + r]
                if i in tscale:
                    # the kernels ran with scale 1 on U = x A^T: db = dy^T U, so dB = s db and ds = <db, B>
                    q = tscale[i]
                    if q.requires_grad:
                        grads[q] = (dbi * pb.detach().float()).sum().reshape(1)
                    dbi = dbi * q.detach().float()
                if pb.requires_grad:
                    grads[pb] = dbi
        if lin.weight.requires_grad or (lin.bias is not None and lin.bias.requires_grad):
            # trainable dense weight (PatchMerging.reduction under plain MTLoRA, lora.py:599-600; or an unfrozen layer):
            # dW = dPre^T x[0], dbias = column sums of dPre, dPre = sum_j dy[j] (lora.py:255,262-266)
            M = dy.shape[1]
            dpre = dy[0] if dy.shape[0] == 1 else ops.sum_streams(dy)
            if ps is not None:
                dpre = ops.scale_rows(dpre.unsqueeze(0), ps[:1].contiguous(), rps)[0]
            if lin.weight.requires_grad:
                grads[lin.weight] = ops.xty(dpre, saved["x"][0])
            if lin.bias is not None and lin.bias.requires_grad:
                ones = torch.ones((M, 8), dtype=BF16, device=dy.device)
                grads[lin.bias] = ops.xty(dpre, ones)[:, 0]
        return dx, grads


class AdapterStager:
    """Re-packs the bf16 adapter operands of ALL layers of a model in one launch when an optimizer step (or any in-place
    update that bumps `_version`) changed their fp32 masters; `LinearEngine.stage` then finds them current. Job tables
    (mtl_pack_job per layer) and operand views are built once, for two alternating buffers: a refresh costs one pass over
    the parameters' version counters on the host, and the operands a graph was recorded with stay intact until the
    refresh after next (the fp32 masters of such a graph would be stale for autograd as well).
    Layers with a trainable LoRA scale (folded into B at staging time) keep staging themselves."""

    def __init__(self, modules):
        self.modules = list(modules)
        self._sig = None

    def _build(self, engines):
        self.items = []
        for e in engines:
            if e.spec.r_shared == 0 or e.scales():
                continue
            ad = e.adapters()
            if any((not p.is_cuda) or p.dtype != torch.float32 or not p.is_contiguous() for p in ad):
                continue
            self.items.append((e, ad))
        self.flat = [p for _, ad in self.items for p in ad]
        self.ptrs = [p.data_ptr() for p in self.flat]
        self.spans, k = [], 0
        for _, ad in self.items:
            self.spans.append((k, k + len(ad)))
            k += len(ad)
        self.versions = None
        self.side = 0
        self.arrs, self.packed = [], []
        if len(self.items) < 2 or any(p.device != self.flat[0].device for p in self.flat):
            self.items = []
            return
        offs, n = [], 0
        for e, _ in self.items:
            offs.append(n)
            n += (2 * e.spec.R_pad * (e.spec.K + e.spec.Nf) + 7) // 8 * 8     # keeps every operand 16-byte aligned
        for _ in range(2):
            buf = torch.empty(n, dtype=BF16, device=self.flat[0].device)
            arr = (ops.N.PackJob * len(self.items))()
            views = []
            for j, (e, ad) in enumerate(self.items):
                spec, T = e.spec, len(e.tasks)
                R, K, Nf, o = spec.R_pad, spec.K, spec.Nf, offs[j]
                pk = (buf[o:o + R * K].view(R, K), buf[o + R * K:o + R * (K + Nf)].view(Nf, R),
                      buf[o + R * (K + Nf):o + R * (2 * K + Nf)].view(K, R),
                      buf[o + R * (2 * K + Nf):o + 2 * R * (K + Nf)].view(R, Nf))
                jb = arr[j]
                jb.cfg = spec.cfg(1, False)
                jb.a_shared, jb.b_shared = ad[0].data_ptr(), ad[1].data_ptr()
                for t in range(T):
                    jb.a_tasks[t], jb.b_tasks[t] = ad[2 + t].data_ptr(), ad[2 + T + t].data_ptr()
                jb.a_cat, jb.b_cat, jb.a_cat_t, jb.b_cat_t = (t.data_ptr() for t in pk)
                views.append(pk)
            self.arrs.append(arr)
            self.packed.append(views)

    def refresh(self):
        """-> number of layers re-packed (0: everything was current)."""
        engines = [m.engine for m in self.modules]
        sig = tuple(map(id, engines))          # MTLoRALinear.merge() swaps the engine in effect
        if sig != self._sig:
            self._build(engines)
            self._sig = sig
        if not self.items:
            return 0                           # nothing to do, or a single layer: its own stage() handles it
        vers = [p._version for p in self.flat]
        if vers == self.versions and all(e._akey is not None for e, _ in self.items):
            return 0
        if [p.data_ptr() for p in self.flat] != self.ptrs:      # parameters moved (.to(), .cuda()): new job tables
            self._build(engines)
            if not self.items:
                return 0
            vers = [p._version for p in self.flat]
        self.side ^= 1
        with torch.cuda.device(self.flat[0].device):
            ops.N.call("mtl_linear_pack_many", self.arrs[self.side], len(self.items), ops.N.stream())
        keys = list(zip(self.ptrs, vers))
        for (e, _), pk, (k0, k1) in zip(self.items, self.packed[self.side], self.spans):
            e._packed = pk
            e._akey = tuple(keys[k0:k1])
        self.versions = vers
        return len(self.items)


class _LinearFn(torch.autograd.Function):
    """Autograd wrapper for a stand-alone MTLoRALinear / CompatLinear call (the fused block has its own)."""

    @staticmethod
    def forward(ctx, engine, xt, dropout_p, seed, x, *params):
        ctx.engine = engine
        need = any(ctx.needs_input_grad[4:])
        y, _, saved = engine.forward(x, xt=xt, dropout_p=dropout_p, seed=seed, save=need)
        if need:
            tens = {k: v for k, v in saved.items() if torch.is_tensor(v)}
            ctx.meta = {k: v for k, v in saved.items() if not torch.is_tensor(v)}
            ctx.keys = list(tens)
            ctx.save_for_backward(*tens.values())
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.engine
        saved = dict(ctx.meta)
        saved.update(zip(ctx.keys, ctx.saved_tensors))
        dx, grads = eng.backward(saved, dy.contiguous())
        if saved["dropout_p"] > 0 and eng.spec.r_shared > 0:
            # the appended D(x[0]) stream is produced from x[0] inside the wrapper; its gradient is folded into dx[0]
            pad = torch.zeros_like(dx[:1])
            dx = torch.cat([dx, pad]) if dx.shape[0] + 1 == saved["x"].shape[0] else dx
        return (None, None, None, None, dx) + tuple(grads.get(p) for p in ctx.params)


def run_linear_standalone(engine, x, x_tasks, dropout_p, training):
    """Common stand-alone call path: arbitrary leading shape (..., K) -> ((..., N), {task: (..., N)} | None)."""
    if not x.is_cuda:
        raise RuntimeError("mtlora_b200: expected CUDA tensors — there is no CPU path (build + run on the B200)")
    spec = engine.spec
    lead, K = x.shape[:-1], x.shape[-1]
    if K != spec.K:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({x.numel() // max(K, 1)}x{K} and {spec.K}x{spec.Nf})")
    in_dtype = x.dtype
    M = x.numel() // K
    streams = [x.reshape(M, K).to(BF16)]
    xt = x_tasks is not None and spec.T > 0
    if xt:
        streams += [x_tasks[t].reshape(M, K).to(BF16) for t in engine.tasks]
    p = dropout_p if (training and spec.r_shared > 0) else 0.0
    seed = _new_seed() if p > 0 else 0
    if p > 0:
        streams.append(_DropoutFn.apply(streams[0], p, seed))
    xs = torch.stack(streams)
    y = _LinearFn.apply(engine, xt, p, seed, xs, *engine.params())
    outs = [y[i].reshape(*lead, spec.Nf).to(in_dtype) for i in range(spec.S_out)]
    if engine.addition:
        ln = engine.owner.lora_norm
        outs[0] = _AdditionTailFn.apply(y, ln.weight, ln.bias, ln.eps).reshape(*lead, spec.Nf).to(in_dtype)
    if spec.S_out == 1 and not (spec.r_shared > 0 and engine.tasks):
        return outs[0], None
    return outs[0], {t: outs[1 + i] for i, t in enumerate(engine.tasks)}


class _AdditionTailFn(torch.autograd.Function):
    """shared_mode 'addition' (reference lora.py:279-282): y[0] + LayerNorm(sum_t y[1 + t]) over the stream-stacked output
    of the linear kernel (y[0] = frozen product, y[1 + t] = task outputs), through mtl_sum_streams / mtl_layernorm_* /
    mtl_add."""

    @staticmethod
    def forward(ctx, y, weight, bias, eps):
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        tot = ops.sum_streams(y[1:].contiguous())
        ln, mean, rstd = ops.layernorm_fwd(tot, w, b, eps)
        ctx.save_for_backward(tot, w, mean, rstd)
        ctx.n_tasks = y.shape[0] - 1
        ctx.want = weight.requires_grad or bias.requires_grad
        return ops.add(y[0].contiguous(), ln)

    @staticmethod
    def backward(ctx, dout):
        tot, w, mean, rstd = ctx.saved_tensors
        d = dout if (dout.dtype == BF16 and dout.is_contiguous()) else dout.to(BF16).contiguous()
        dtot, dw, db = ops.layernorm_bwd(d, tot, w, mean, rstd, want_param_grads=ctx.want)
        dy = torch.cat([d.unsqueeze(0), dtot.unsqueeze(0).expand(ctx.n_tasks, *dtot.shape)])
        return dy, dw, db, None


class _DropoutFn(torch.autograd.Function):
    """D(x) with the library's counter-based mask (mtl_dropout); backward re-applies the same mask."""

    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed = p, seed
        return ops.dropout(x.contiguous(), p, seed)

    @staticmethod
    def backward(ctx, dy):
        return ops.dropout(dy.contiguous(), ctx.p, ctx.seed), None, None


class MTLoRALinear(LoRALayer):
    """Frozen nn.Linear + one task-shared and T task-specific low-rank updates (reference models/lora.py:159-284).

    forward(x, x_tasks=None) -> (pretrained + shared_lora(x), {task: pretrained + task_lora(x or x_tasks[task])} | None)
    """

    def __init__(self, in_features: int, out_features: int, r: Union[int, Mapping[str, int]] = 0,
                 lora_shared_scale: float = 1.0, lora_task_scale: Union[float, Mapping[str, float]] = 1.0,
                 lora_dropout: float = 0.0, tasks=None, trainable_scale_shared=False, trainable_scale_per_task=False,
                 shared_mode: str = "matrix", **kwargs):
        assert shared_mode in ["matrix", "matrixv2", "add", "addition", "lora_only"]
        if shared_mode == "add":
            shared_mode = "addition"
        if shared_mode == "lora_only":
            tasks = None
        has_tasks = tasks is not None
        if not has_tasks and shared_mode not in ["matrix"]:
            shared_mode = "matrix"
        if isinstance(r, int):
            r = {"shared": r}
        super().__init__(r=r["shared"], lora_alpha=lora_shared_scale, lora_dropout=lora_dropout)
        self.linear = nn.Linear(in_features, out_features, **kwargs)
        self.tasks = tasks
        self.shared_mode = shared_mode
        r_tasks, s_tasks = [], []
        if r["shared"] > 0:
            if has_tasks:
                self.lora_tasks_A = nn.ParameterDict({
                    task: nn.Parameter(self.linear.weight.new_zeros((r[task], in_features))) for task in tasks})
                self.lora_tasks_B = nn.ParameterDict({
                    task: nn.Parameter(self.linear.weight.new_zeros((out_features, r[task]))) for task in tasks})
                if trainable_scale_per_task:
                    # reference :210-214 (it takes a float here; a per-task mapping is accepted too)
                    init = lora_task_scale if isinstance(lora_task_scale, Mapping) else {t: lora_task_scale for t in tasks}
                    self.lora_task_scale = nn.ParameterDict({
                        task: nn.Parameter(torch.FloatTensor([float(init[task])])) for task in tasks})
                    s_tasks = [1.0] * len(tasks)      # folded into the packed B operand (LinearEngine.stage)
                else:
                    self.lora_task_scale = {task: lora_task_scale[task] for task in tasks}
                    s_tasks = [float(self.lora_task_scale[t]) for t in tasks]
                r_tasks = [r[t] for t in tasks]
            if shared_mode == "addition":
                assert has_tasks
                self.lora_norm = nn.LayerNorm(out_features)          # reference :217-219
            else:
                self.lora_shared_A = nn.Parameter(self.linear.weight.new_zeros((r["shared"], in_features)))
                self.lora_shared_B = nn.Parameter(self.linear.weight.new_zeros((out_features, r["shared"])))
            if trainable_scale_shared:
                self.lora_shared_scale = nn.Parameter(torch.FloatTensor([float(lora_shared_scale)]))   # reference :229-231
            else:
                self.lora_shared_scale = lora_shared_scale
            self.reset_parameters()
        addition = shared_mode == "addition" and r["shared"] > 0
        spec = ops.LinearSpec(in_features, out_features, 1 if addition else r["shared"], r_tasks,
                              1.0 if (trainable_scale_shared and r["shared"] > 0) else float(lora_shared_scale), s_tasks,
                              shared_mode=shared_mode if (r_tasks and not addition) else "matrix")
        self._engine = LinearEngine(self, self.linear, spec, tasks if (has_tasks and r["shared"] > 0) else None)
        self._plain_engine = None

    def reset_parameters(self):
        """A ~ kaiming_uniform(a=sqrt(5)), B = 0 (reference :236-247): the layer starts equal to the frozen linear."""
        if hasattr(self, "lora_shared_A"):
            nn.init.kaiming_uniform_(self.lora_shared_A, a=math.sqrt(5))
            nn.init.zeros_(self.lora_shared_B)
        if hasattr(self, "lora_tasks_A"):
            for task in self.tasks:
                nn.init.kaiming_uniform_(self.lora_tasks_A[task], a=math.sqrt(5))
                nn.init.zeros_(self.lora_tasks_B[task])

    # ---- inference merge (the reference declares merge() and raises NotImplementedError, :249-251) ----
    def can_merge(self):
        """W <- W + scale * B_s A_s is exact when no output stream needs the un-merged product: layers without task
        adapters (qkv, every block but the last of a stage, MTLoRA+ reductions) and 'matrixv2' layers (whose task outputs
        carry the shared update too, :267-274). 'matrix' layers with tasks need both W and W + sBA: they stay as they are."""
        return self.r > 0 and hasattr(self, "lora_shared_A") and (self.tasks is None)

    @torch.no_grad()
    def merge(self):
        """Fold the shared adapter into the frozen weight (W = W + scale * B A) for evaluation; returns whether it did."""
        if self.merged or not self.can_merge():
            return False
        sc = self.lora_shared_scale
        sc = sc.detach().float() if isinstance(sc, torch.Tensor) else float(sc)
        self.linear.weight.add_((self.lora_shared_B.float() @ self.lora_shared_A.float()) * sc)
        self.merged = True
        return True

    @torch.no_grad()
    def unmerge(self):
        if not self.merged:
            return False
        sc = self.lora_shared_scale
        sc = sc.detach().float() if isinstance(sc, torch.Tensor) else float(sc)
        self.linear.weight.sub_((self.lora_shared_B.float() @ self.lora_shared_A.float()) * sc)
        self.merged = False
        return True

    def train(self, mode: bool = True):
        if mode and self.merged:
            self.unmerge()      # adapters must see their own gradients again (loralib convention)
        return super().train(mode)

    @property
    def engine(self):
        """The engine the fused block drives: the full one, or a plain dense one while the shared adapter is merged."""
        if self.merged:
            if self._plain_engine is None:
                self._plain_engine = LinearEngine(None, self.linear, ops.LinearSpec(self.linear.in_features,
                                                                                      self.linear.out_features), None)
            return self._plain_engine
        return self._engine

    def forward(self, x: torch.Tensor, x_tasks: Optional[Dict[str, torch.Tensor]] = None):
        return run_linear_standalone(self.engine, x, x_tasks, self.lora_dropout_p, self.training)


def mark_only_lora_as_trainable(model: nn.Module, bias: str = "none", freeze_patch_embed: bool = False,
                                freeze_norm: bool = False, free_relative_bias: bool = False,
                                freeze_downsample_reduction=False) -> None:
    """Freeze everything except LoRA parameters and the optional extras, by substring match on parameter names
    (reference models/lora.py:580-630; NB `free_relative_bias` means *freeze* there, kept for drop-in parity)."""
    def keep(name):
        return ("lora_" in name
                or (not freeze_patch_embed and "patch_embed" in name)
                or (not freeze_norm and "norm" in name)
                or (not freeze_downsample_reduction and "downsample.reduction" in name)
                or (not free_relative_bias and "relative_position_bias_table" in name))

    print(f"LoRA bias mode: {bias}")
    print(f"LoRA Freeze patch_embed: {freeze_patch_embed}")
    print(f"LoRA Freeze norm: {freeze_norm}")
    print(f"LoRA Freeze downsample_reduction: {freeze_downsample_reduction}")
    print(f"LoRA Freeze relative_position_bias: {free_relative_bias}")
    for n, p in model.named_parameters():
        if not keep(n):
            p.requires_grad = False
    if bias == "none":
        return
    if bias == "all":
        for n, p in model.named_parameters():
            if "bias" in n:
                p.requires_grad = True
    elif bias == "lora_only":
        for m in model.modules():
            if isinstance(m, LoRALayer) and hasattr(m, "bias") and m.bias is not None:
                m.bias.requires_grad = True
    else:
        raise NotImplementedError


def merge_lora(model: nn.Module) -> int:
    """Merge every mergeable MTLoRALinear of `model` (MTLoRALinear.merge) for evaluation; returns how many were merged.
    `model.train()` un-merges them again."""
    return sum(1 for m in model.modules() if isinstance(m, MTLoRALinear) and m.merge())


def lora_filter(key: str, value: Any) -> bool:
    return "lora_" in key


def map_old_state_dict_weights(state_dict: Dict, mapping: Mapping, prefix: str, split_qkv: bool = False) -> Dict:
    """Rename checkpoint keys (e.g. `attn.qkv.weight` -> `attn.qkv.linear.weight`), reference :644-668."""
    missing = []
    for old, new in mapping.items():
        src = prefix + old
        if src not in state_dict:
            missing.append(old)
            continue
        dst = prefix + new
        value = state_dict.pop(src)
        tail = ".".join(dst.split(".")[-4:])
        if split_qkv and tail in ("attn.qkv.linear.weight", "attn.qkv.linear.bias"):
            kind = tail.split(".")[-1]
            stem = ".".join(dst.split(".")[:-2])
            for part, chunk in zip("qkv", torch.chunk(value, chunks=3)):
                state_dict[f"{stem}.{part}.linear.{kind}"] = chunk
        else:
            state_dict[dst] = value
    if missing:
        print(f"WARNING: The following keys from the checkpoint were not mapped: {missing}")
    return state_dict
