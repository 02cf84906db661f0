"""Stand-in for imageio (utils.py:26 of the reference; only used when saving sample images)."""


def imwrite(*a, **k):
    raise RuntimeError("imageio is not installed in this image")


imsave = imwrite
