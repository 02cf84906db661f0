"""Data-parallel exchange step of the path: one all-reduce per optimizer step over the trainable gradients only.

The reference launches one process per GPU but never synchronises gradients (main.py:167 has DDP commented out,
SURVEY.md §2a); BASELINE.json's north_star asks for pure data parallelism with an NCCL all-reduce of the (small)
adapter gradients — the frozen backbone has nothing to reduce. `AdapterGradReducer` packs every trainable gradient into
one flat fp32 buffer (≈ 6.4 M values = 26 MB for Swin-T / 4 tasks / r=64), issues a single `all_reduce(SUM)` on the
process group (NCCL over NVLink/NVSwitch on the B200 box, gloo in the CPU tests) and scatters the mean back.
Parameters whose gradient is None (the unused `layers.3.blocks.1.mlp.fc2.lora_shared_{A,B}`)
contribute zeros and keep `grad = None` unless another rank produced a gradient for them.
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


class AdapterGradReducer:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group=None, bucket_dtype=torch.float32):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = process_group
        self.dtype = bucket_dtype
        self.offsets = []
        n = 0
        for p in self.params:
            self.offsets.append(n)
            n += p.numel()
        self.numel = n
        self.flat = None

    def world_size(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def _buffer(self, device):
        if self.flat is None or self.flat.device != device:
            self.flat = torch.zeros(self.numel, dtype=self.dtype, device=device)
        return self.flat

    @torch.no_grad()
    def reduce(self):
        """Average gradients over the group in place (no host synchronisation). No-op for a single process.

        Every rank runs the same model, so `grad is None` holds for the same parameters everywhere; such a parameter
        contributes zeros to the bucket and keeps `grad = None`."""
        ws = self.world_size()
        if ws == 1 or not self.params:
            return
        flat = self._buffer(self.params[0].device)
        live = [(p, off) for p, off in zip(self.params, self.offsets) if p.grad is not None]
        if len(live) != len(self.params):
            flat.zero_()
        views = [flat[off:off + p.numel()].view_as(p) for p, off in live]
        grads = [p.grad for p, _ in live]
        torch._foreach_copy_(views, grads)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(ws)
        torch._foreach_copy_(grads, views)
