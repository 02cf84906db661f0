def __getattr__(name):
    raise RuntimeError("matplotlib is not installed in this image")
