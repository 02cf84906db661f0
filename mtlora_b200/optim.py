"""Flat multi-tensor optimizer step for the trainable tensors of the path.

Reference: main.py:341-353 -> utils.py:348-369 (`NativeScalerWithGradNormCount.__call__`: `scaler.unscale_(optimizer)`,
`clip_grad_norm_(parameters, clip_grad)`, `scaler.step(optimizer)`) with `optimizer = optim.AdamW(...)`
(optimizer.py:58-60). `FlatAdamW` is a drop-in `torch.optim.Optimizer` (same constructor arguments, param_groups, LR
schedulers, `state_dict()` layout of torch.optim.AdamW), whose `step()` is TWO launches of libmtlora_b200.so
(`mtl_opt_sqnorm`, `mtl_opt_adamw`) whatever the number of tensors:

  * the ~200 trainable tensors are described by a device-side segment table (parameter pointer, gradient pointer, offset
    into two flat fp32 moment buffers, param-group index), re-uploaded only when a gradient pointer changes;
  * `_step_supports_amp_scaling = True`: `GradScaler.step()` hands the optimizer its device-side `grad_scale` /
    `found_inf` scalars instead of synchronising with the host — the kernel unscales, and skips the whole step
    (including the step counter) when an inf / nan was found;
  * `max_grad_norm=5.0` folds `clip_grad_norm_` into the same step (the global norm is computed by `mtl_opt_sqnorm` over
    the same table; `last_grad_norm()` returns it as a device scalar, like the value utils.py:357 returns). When the
    caller's loop calls `clip_grad_norm_` itself (main.py unchanged), leave it at None.

There is no CPU path: parameters must live on a CUDA device.
"""
import ctypes

import torch

from . import _native as N


class FlatAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, max_grad_norm=None, decoupled=True):
        if amsgrad or maximize:
            raise NotImplementedError("mtlora_b200.FlatAdamW: amsgrad / maximize are not implemented")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("FlatAdamW: invalid hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False)
        super().__init__(params, defaults)
        if len(self.param_groups) > N.MTL_OPT_MAX_GROUPS:
            raise ValueError(f"FlatAdamW: at most {N.MTL_OPT_MAX_GROUPS} param groups, got {len(self.param_groups)}")
        self.max_grad_norm = max_grad_norm
        self.fused_clip = max_grad_norm is not None
        self.decoupled = decoupled
        self._built = False

    # ---- flat state --------------------------------------------------------------------------------------------
    def _build(self):
        ps, groups = [], []
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("mtlora_b200.FlatAdamW: parameters must live on a CUDA device — there is no CPU path")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("FlatAdamW: parameters must be contiguous fp32 tensors")
                ps.append(p)
                groups.append(gi)
        if not ps:
            raise ValueError("FlatAdamW: no trainable parameters")
        dev = ps[0].device
        if any(p.device != dev for p in ps):
            raise RuntimeError("mtlora_b200.FlatAdamW: every parameter must live on the same device (one process per GPU)")
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4          # 16-byte aligned moment slices
        self._ps, self._pgroup, self._offs = ps, groups, offs
        self._m = torch.zeros(n, dtype=torch.float32, device=dev)
        self._v = torch.zeros(n, dtype=torch.float32, device=dev)
        self._state = torch.zeros(2 + len(ps), dtype=torch.float32, device=dev)   # scratch, grad norm, per-tensor steps
        self._sq = torch.zeros(1, dtype=torch.float32, device=dev)
        prefix, c = [], 0
        for p in ps:
            prefix.append(c)
            c += (p.numel() + N.MTL_OPT_CHUNK - 1) // N.MTL_OPT_CHUNK
        prefix.append(c)
        self._n_chunks = c
        self._prefix = torch.tensor(prefix, dtype=torch.int32).to(dev)
        self._segs_host = (N.OptSeg * len(ps))()
        for i, p in enumerate(ps):
            s = self._segs_host[i]
            s.param, s.grad, s.offset, s.numel, s.group = p.data_ptr(), None, offs[i], p.numel(), groups[i]
        nbytes = ctypes.sizeof(self._segs_host)
        self._segs_dev = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        # the table travels through a small ring of pinned staging buffers with non-blocking copies: a blocking copy from
        # pageable memory would synchronise the stream (the whole step) every time a gradient pointer changes
        self._pin = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
        self._pin_ev = [None] * 4
        self._pin_k = 0
        self._grad_key = None
        assert N.load().mtl_opt_seg_size() == ctypes.sizeof(N.OptSeg)
        # torch.optim.AdamW-shaped per-parameter state (views into the flat buffers) for state_dict() / checkpoints
        for i, p in enumerate(ps):
            self.state[p] = {"step": self._state[2 + i], "exp_avg": self._m[offs[i]:offs[i] + p.numel()].view_as(p),
                             "exp_avg_sq": self._v[offs[i]:offs[i] + p.numel()].view_as(p)}
        self._built = True

    def add_param_group(self, param_group):
        """torch.optim API: parameters added later get fresh (zero) moments; the existing ones keep theirs."""
        old = None
        if getattr(self, "_built", False):
            old = {p: (self._m[o:o + p.numel()].clone(), self._v[o:o + p.numel()].clone(), self._state[2 + i].clone())
                   for i, (p, o) in enumerate(zip(self._ps, self._offs))}
            self._built = False
        super().add_param_group(param_group)
        if old is not None:
            if len(self.param_groups) > N.MTL_OPT_MAX_GROUPS:
                raise ValueError(f"FlatAdamW: at most {N.MTL_OPT_MAX_GROUPS} param groups")
            self._build()
            with torch.no_grad():
                for i, (p, o) in enumerate(zip(self._ps, self._offs)):
                    if p in old:
                        self._m[o:o + p.numel()].copy_(old[p][0])
                        self._v[o:o + p.numel()].copy_(old[p][1])
                        self._state[2 + i] = old[p][2]

    def load_state_dict(self, state_dict):
        if not self._built:
            self._build()
        super().load_state_dict(state_dict)
        # re-home the loaded moments into the flat buffers
        with torch.no_grad():
            for i, p in enumerate(self._ps):
                st = self.state.get(p, {})
                m = self._m[self._offs[i]:self._offs[i] + p.numel()].view_as(p)
                v = self._v[self._offs[i]:self._offs[i] + p.numel()].view_as(p)
                if "exp_avg" in st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    self._state[2 + i] = float(st.get("step", 0.0))
                self.state[p] = {"step": self._state[2 + i], "exp_avg": m, "exp_avg_sq": v}

    def last_grad_norm(self):
        """Total gradient norm of the last step (unscaled, before clipping) as a device scalar; needs max_grad_norm."""
        return self._state[1]

    # ---- step --------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self._built:
            self._build()
        ps = self._ps
        key = tuple(0 if p.grad is None else p.grad.data_ptr() for p in ps)
        if key != self._grad_key:
            for i, p in enumerate(ps):
                g = p.grad
                if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device):
                    raise TypeError("FlatAdamW: gradients must be contiguous fp32 tensors on the parameter's device")
                self._segs_host[i].grad = None if g is None else g.data_ptr()
                self._segs_host[i].param = p.data_ptr()
            k = self._pin_k
            self._pin_k = (k + 1) % len(self._pin)
            if self._pin_ev[k] is not None:
                self._pin_ev[k].synchronize()      # its previous upload (4 table changes ago) has long completed
            ctypes.memmove(self._pin[k].data_ptr(), ctypes.addressof(self._segs_host), ctypes.sizeof(self._segs_host))
            self._segs_dev.copy_(self._pin[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pin_ev[k] = ev
            self._grad_key = key
        if not any(key):
            return loss
        groups = (N.OptGroup * len(self.param_groups))()
        for i, g in enumerate(self.param_groups):
            groups[i].lr, groups[i].beta1, groups[i].beta2 = float(g["lr"]), float(g["betas"][0]), float(g["betas"][1])
            groups[i].eps, groups[i].weight_decay = float(g["eps"]), float(g["weight_decay"])
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        for t in (grad_scale, found_inf):
            if t is not None and (t.dtype != torch.float32 or not t.is_cuda):
                raise TypeError("FlatAdamW: grad_scale / found_inf must be fp32 CUDA scalars")
        clip = self.max_grad_norm is not None
        with torch.cuda.device(ps[0].device):
            st = N.stream()
            if clip:
                N.call("mtl_opt_sqnorm", self._segs_dev.data_ptr(), self._prefix.data_ptr(), len(ps), self._n_chunks,
                       self._sq.data_ptr(), st)
            N.call("mtl_opt_adamw", self._segs_dev.data_ptr(), self._prefix.data_ptr(), len(ps), self._n_chunks,
                   self._m.data_ptr(), self._v.data_ptr(), self._state.data_ptr(), groups, len(self.param_groups),
                   N.ptr(grad_scale), N.ptr(found_inf), self._sq.data_ptr() if clip else None,
                   float(self.max_grad_norm) if clip else 0.0, 1 if self.decoupled else 0, st)
        # the kernel wrote the parameters through raw pointers: bump their autograd version counters, so that saved-tensor
        # checks and the staged bf16 operand copies of the linears (keyed on (data_ptr, _version)) see the update
        torch.autograd.graph.increment_version([p for p, k in zip(ps, key) if k])
        return loss
