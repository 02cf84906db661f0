from PIL import Image


def _pil_interp(method):
    return {"bicubic": Image.BICUBIC, "lanczos": Image.LANCZOS, "hamming": Image.HAMMING}.get(method, Image.BILINEAR)
