"""`timm.scheduler.scheduler.Scheduler` (timm 0.9.2 semantics, reduced to what lr_scheduler.py of the reference uses)."""


class Scheduler:
    def __init__(self, optimizer, param_group_field, noise_range_t=None, noise_type="normal", noise_pct=0.67,
                 noise_std=1.0, noise_seed=None, initialize=True):
        self.optimizer = optimizer
        self.param_group_field = param_group_field
        self._initial_param_group_field = f"initial_{param_group_field}"
        if initialize:
            for group in self.optimizer.param_groups:
                group.setdefault(self._initial_param_group_field, group[param_group_field])
        self.base_values = [group[self._initial_param_group_field] for group in self.optimizer.param_groups]
        self.metric = None
        self.noise_range_t = None

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, state_dict):
        self.__dict__.update(state_dict)

    def get_epoch_values(self, epoch):
        return None

    def get_update_values(self, num_updates):
        return None

    def step(self, epoch, metric=None):
        self.metric = metric
        values = self.get_epoch_values(epoch)
        if values is not None:
            self.update_groups(values)

    def step_update(self, num_updates, metric=None):
        self.metric = metric
        values = self.get_update_values(num_updates)
        if values is not None:
            self.update_groups(values)

    def update_groups(self, values):
        if not isinstance(values, (list, tuple)):
            values = [values] * len(self.optimizer.param_groups)
        for group, value in zip(self.optimizer.param_groups, values):
            if "lr_scale" in group:
                group[self.param_group_field] = value * group["lr_scale"]
            else:
                group[self.param_group_field] = value
