"""`DropPath`, `to_2tuple`, `trunc_normal_` with timm 0.9.2's published semantics."""
import torch


class DropPath(torch.nn.Module):
    """Per-sample stochastic depth: Bernoulli(keep) mask of shape (B, 1, ...) divided by keep; identity in eval mode."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


trunc_normal_ = torch.nn.init.trunc_normal_
