"""mtlora_b200 — B200-native (sm_100a) implementation of the MTLoRA Swin-backbone hot path.

The compute lives in libmtlora_b200.so (hand-written CUDA behind the C ABI of include/mtlora_b200.h); this package
is the PyTorch-facing host layer mirroring the reference's module API (models/lora.py, models/swin_transformer_mtlora.py).
"""
__version__ = "0.1.0"
