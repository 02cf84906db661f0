"""`timm.data` names the reference's data/ package imports at module level (ImageNet pipelines, never used on the
multi-task path the benchmarks and tests drive)."""
from .constants import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD  # noqa: F401
from . import transforms  # noqa: F401


class Mixup:
    def __init__(self, *a, **k):
        raise RuntimeError("timm.data.Mixup is not available in this image")


def create_transform(*a, **k):
    raise RuntimeError("timm.data.create_transform is not available in this image")
