"""`timm.scheduler.step_lr.StepLRScheduler` (the options lr_scheduler.py of the reference passes)."""
from .scheduler import Scheduler


class StepLRScheduler(Scheduler):
    def __init__(self, optimizer, decay_t, decay_rate=1.0, warmup_t=0, warmup_lr_init=0, t_in_epochs=True,
                 initialize=True, **kwargs):
        super().__init__(optimizer, param_group_field="lr", initialize=initialize)
        self.decay_t, self.decay_rate, self.warmup_t, self.warmup_lr_init = decay_t, decay_rate, warmup_t, warmup_lr_init
        self.t_in_epochs = t_in_epochs
        self.warmup_steps = [(v - warmup_lr_init) / max(warmup_t, 1) for v in self.base_values]

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        return [v * (self.decay_rate ** (t // self.decay_t)) for v in self.base_values]

    def get_epoch_values(self, epoch):
        return self._get_lr(epoch) if self.t_in_epochs else None

    def get_update_values(self, num_updates):
        return self._get_lr(num_updates) if not self.t_in_epochs else None
