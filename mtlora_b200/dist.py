"""Data-parallel exchange step of the path: the all-reduce of the trainable gradients, driven from autograd hooks.

The reference launches one process per GPU but never synchronises gradients (main.py:167 has DDP commented out,
SURVEY.md §2a); BASELINE.json's north_star asks for pure data parallelism with an NCCL all-reduce of the (small)
adapter gradients — the frozen backbone has nothing to reduce — that "drops into main.py's torch.distributed loop
unchanged". So nobody calls anything: when `torch.distributed` is initialised with more than one rank,
`SwinTransformerMTLoRA` attaches a `GradSync` to its own trainable parameters at its first training forward
(`sync_gradients(module)` does the same for any other module, e.g. the decoder heads), and from then on every
`loss.backward()` ends with averaged gradients in `p.grad`:

  * the trainable parameters are split into a few buckets in reverse registration order (the order backward reaches
    them); each bucket is one persistent flat fp32 buffer;
  * a post-accumulate-grad hook per parameter counts the bucket down; a finished bucket is packed (one multi-tensor
    copy) and all-reduced asynchronously while backward continues (NCCL over NVLink / NVSwitch on the B200 box, gloo in
    the CPU tests). Buckets are launched strictly in index order on every rank;
  * a callback queued on the autograd engine runs when backward ends: it flushes the buckets that never filled up
    (parameters without a gradient contribute zeros), waits for the collectives and re-points `p.grad` at the averaged
    slice of the flat buffer (no copy back).

`grad is None` (the unused `layers.3.blocks.1.mlp.fc2.lora_shared_{A,B}`; a task whose loss is
missing on one rank): every bucket carries one presence slot per parameter through the same all-reduce.
  * `exact_presence=False` (default, no host synchronisation): a parameter without a local gradient keeps `grad = None`;
    the presence slots are checked on the device and read back one step late — ranks that disagree on which parameters
    received gradients raise a RuntimeError at the next step instead of silently diverging;
  * `exact_presence=True`: the presence slots are read right away (one small blocking copy per backward) and a
    parameter that is None locally but live on another rank receives the averaged gradient.
"""
import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

_SYNC_ATTR = "_mtlora_b200_grad_sync"


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class _Bucket:
    __slots__ = ("params", "offsets", "numel", "flat", "pending", "work", "seen", "launched", "div", "pres_cache")

    def __init__(self, params):
        self.params = params
        self.offsets = []
        n = 0
        for p in params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat = None
        self.pending = 0
        self.work = None
        self.seen = [False] * len(params)
        self.launched = False
        self.div = None
        self.pres_cache = {}


class GradSync:
    """Hook-driven, bucketed, overlapped averaging of the gradients of `params` over the process group."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, n_buckets: int = 4,
                 exact_presence: bool = False):
        ps: List[torch.nn.Parameter] = []
        for p in params:
            if p.requires_grad and getattr(p, _SYNC_ATTR, None) is None:
                ps.append(p)
        self.group = process_group
        self.exact_presence = exact_presence
        self.params = ps
        self.enabled = True
        # buckets over the reversed parameter list, balanced by element count
        rev = ps[::-1]
        total = sum(p.numel() for p in rev)
        target = max(1, -(-total // max(1, n_buckets)))
        self.buckets: List[_Bucket] = []
        cur, acc = [], 0
        for p in rev:
            cur.append(p)
            acc += p.numel()
            if acc >= target and len(self.buckets) < n_buckets - 1:
                self.buckets.append(_Bucket(cur))
                cur, acc = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        self._where = {}
        self._hooks = []
        for bi, b in enumerate(self.buckets):
            for pi, p in enumerate(b.params):
                self._where[id(p)] = (bi, pi)
                setattr(p, _SYNC_ATTR, self)
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._armed = False
        self._next = 0
        self._check = None        # (mismatch tensor on the host side, event) of the previous backward
        self.n_reductions = 0     # collectives issued so far (tests / bench)

    # ---- lifecycle ----------------------------------------------------------------------------------------------
    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for p in self.params:
            if getattr(p, _SYNC_ATTR, None) is self:
                delattr(p, _SYNC_ATTR)

    def world_size(self):
        return _world(self.group)

    # ---- hooks --------------------------------------------------------------------------------------------------
    def _on_grad(self, p):
        if not self.enabled or self.world_size() == 1:
            return
        if not self._armed:
            self._arm()
        bi, pi = self._where[id(p)]
        b = self.buckets[bi]
        if b.launched or b.seen[pi]:
            return   # a second accumulation into the same parameter after its bucket left: picked up by the next backward
        b.seen[pi] = True
        b.pending -= 1
        while self._next < len(self.buckets) and self.buckets[self._next].pending == 0:
            self._launch(self.buckets[self._next])
            self._next += 1

    def _arm(self):
        self._armed = True
        self._next = 0
        for b in self.buckets:
            b.pending = len(b.params)
            b.seen = [False] * len(b.params)
            b.launched = False
            b.work = None
        torch.autograd.Variable._execution_engine.queue_callback(self._finalize)

    @torch.no_grad()
    def _launch(self, b: _Bucket):
        dev = b.params[0].device
        n_p = len(b.params)
        if b.flat is None or b.flat.device != dev:
            b.flat = torch.zeros(b.numel + n_p, dtype=torch.float32, device=dev)
        flat = b.flat
        live = [i for i, p in enumerate(b.params) if p.grad is not None]
        if len(live) != n_p:
            flat.zero_()
            if live:
                key = tuple(live)
                pat = b.pres_cache.get(key)
                if pat is None:   # presence pattern of this bucket, uploaded once per distinct pattern
                    pat = torch.tensor([1.0 if p.grad is not None else 0.0 for p in b.params]).to(dev)
                    b.pres_cache[key] = pat
                flat[b.numel:].copy_(pat)
        else:
            flat[b.numel:].fill_(1.0)
        views, grads = [], []
        for i in live:
            p = b.params[i]
            v = flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
            if p.grad.data_ptr() != v.data_ptr():
                views.append(v)
                grads.append(p.grad)
        if views:
            torch._foreach_copy_(views, grads)
        ws = self.world_size()
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            b.work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        else:
            b.work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        b.launched = True
        self.n_reductions += 1
        b.div = None if backend == "nccl" else 1.0 / ws

    @torch.no_grad()
    def _finalize(self):
        self._armed = False
        # verdict of the previous backward's presence check (read back asynchronously, never blocks a healthy run)
        if self._check is not None:
            bad, ev = self._check
            self._check = None
            if ev is not None:
                ev.synchronize()
            if float(bad.item()) != 0.0:
                raise RuntimeError(
                    "mtlora_b200.GradSync: the ranks disagree on which parameters received a gradient in the previous "
                    "backward (a parameter had grad=None on some ranks only); construct the reducer with "
                    "exact_presence=True (sync_gradients(module, exact_presence=True)) for such workloads")
        for b in self.buckets[self._next:]:
            self._launch(b)
        self._next = len(self.buckets)
        mism = []
        for b in self.buckets:
            b.work.wait()
            if b.div is not None:
                b.flat.mul_(b.div)
            b.work = None
            pres = b.flat[b.numel:]     # fraction of ranks that had a gradient, per parameter
            if self.exact_presence:
                live = (pres > 0).tolist()          # small blocking read, by request
            else:
                live = [p.grad is not None for p in b.params]
                mism.append(((pres > 0) & (pres < 1)).any())
            for i, p in enumerate(b.params):
                if live[i]:
                    p.grad = b.flat[b.offsets[i]:b.offsets[i] + p.numel()].view_as(p)
        if mism:
            bad = torch.stack(mism).any().float()
            if bad.is_cuda:
                host = torch.empty((), dtype=torch.float32, pin_memory=True)
                host.copy_(bad, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                self._check = (host, ev)
            else:
                self._check = (bad, None)


def sync_gradients(module: torch.nn.Module, process_group=None, n_buckets: int = 4,
                   exact_presence: bool = False) -> Optional[GradSync]:
    """Attach a `GradSync` to every trainable parameter of `module` that no reducer covers yet. Returns it (None when
    there is nothing to do: a single process, or every parameter already covered)."""
    if _world(process_group) == 1:
        return None
    ps = [p for p in module.parameters() if p.requires_grad and getattr(p, _SYNC_ATTR, None) is None]
    if not ps:
        return None
    return GradSync(ps, process_group, n_buckets, exact_presence)


def auto_sync_enabled():
    """The backbone attaches its reducer by itself unless MTLORA_B200_GRAD_SYNC=0 (e.g. when the caller wraps the model
    in DistributedDataParallel or runs its own reducer)."""
    return os.environ.get("MTLORA_B200_GRAD_SYNC", "1") != "0"


class AdapterGradReducer:
    """Explicit variant (round-1 API, kept for callers that prefer one call after backward): packs the trainable gradients
    into one flat fp32 buffer, one blocking all-reduce, scatters the mean back. Parameters whose gradient is None on
    this rank contribute zeros; with `exact_presence=True` they receive the average when another rank produced one,
    otherwise disagreeing ranks raise."""

    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, exact_presence: bool = True):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.exact_presence = exact_presence
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat = None

    def world_size(self):
        return _world(self.group)

    @torch.no_grad()
    def reduce(self):
        ws = self.world_size()
        if ws == 1 or not self.params:
            return
        dev = self.params[0].device
        n_p = len(self.params)
        if self.flat is None or self.flat.device != dev:
            self.flat = torch.zeros(self.numel + n_p, dtype=torch.float32, device=dev)
        flat = self.flat
        have = [p.grad is not None for p in self.params]
        if not all(have):
            flat.zero_()
        flat[self.numel:].copy_(torch.tensor([1.0 if h else 0.0 for h in have]))
        live = [(p, off) for p, off, h in zip(self.params, self.offsets, have) if h]
        views = [flat[off:off + p.numel()].view_as(p) for p, off in live]
        if views:
            torch._foreach_copy_(views, [p.grad for p, _ in live])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(ws)
        pres = flat[self.numel:].tolist()
        if views:
            torch._foreach_copy_([p.grad for p, _ in live], views)
        for p, off, h, f in zip(self.params, self.offsets, have, pres):
            if not h and f > 0:
                if not self.exact_presence:
                    raise RuntimeError("AdapterGradReducer: ranks disagree on which parameters received a gradient")
                p.grad = flat[off:off + p.numel()].view_as(p).clone()
