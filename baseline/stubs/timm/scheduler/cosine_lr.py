"""`timm.scheduler.cosine_lr.CosineLRScheduler` (single cycle, linear warm-up; the options lr_scheduler.py:37-47 passes)."""
import math

from .scheduler import Scheduler


class CosineLRScheduler(Scheduler):
    def __init__(self, optimizer, t_initial, lr_min=0.0, cycle_mul=1.0, cycle_decay=1.0, cycle_limit=1, warmup_t=0,
                 warmup_lr_init=0, warmup_prefix=False, t_in_epochs=True, initialize=True, **kwargs):
        super().__init__(optimizer, param_group_field="lr", initialize=initialize)
        self.t_initial, self.lr_min, self.warmup_t, self.warmup_lr_init = t_initial, lr_min, warmup_t, warmup_lr_init
        self.warmup_prefix, self.t_in_epochs = warmup_prefix, t_in_epochs
        self.warmup_steps = [(v - warmup_lr_init) / max(warmup_t, 1) for v in self.base_values]
        if warmup_t:
            super().update_groups(self.warmup_lr_init)

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        if self.warmup_prefix:
            t = t - self.warmup_t
        t = min(t, self.t_initial)
        return [self.lr_min + 0.5 * (v - self.lr_min) * (1 + math.cos(math.pi * t / max(self.t_initial, 1)))
                for v in self.base_values]

    def get_epoch_values(self, epoch):
        return self._get_lr(epoch) if self.t_in_epochs else None

    def get_update_values(self, num_updates):
        return self._get_lr(num_updates) if not self.t_in_epochs else None
