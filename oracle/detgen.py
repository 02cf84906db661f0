"""Deterministic, RNG-library-independent parameter / input generator shared by the golden-vector script, the
oracle tests, the GPU parity tests and bench.py (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Values come from a splitmix64 hash of (crc32(name), element index) evaluated with numpy uint64 arithmetic, so the
same name+shape gives bit-identical fp32 tensors on every machine and library version — the committed golden
vectors (tests/golden/) stay valid without committing multi-megabyte parameter files.
"""
import zlib
from collections import OrderedDict

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(name, shape, lo=-1.0, hi=1.0):
    """fp32 tensor of `shape`, i.i.d.-looking uniform in [lo, hi), a pure function of (name, shape)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(zlib.crc32(name.encode("utf-8")))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        z = _splitmix64(idx ^ _splitmix64(seed))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32)).reshape(tuple(shape))


def randint(name, shape, hi):
    return (uniform(name, shape, 0.0, 1.0) * hi).floor().clamp_(max=hi - 1)


def std_uniform(name, shape, std, mean=0.0):
    a = std * 3.0 ** 0.5
    return uniform(name, shape, mean - a, mean + a)


def backbone_param_shapes(cfg, ranks, downsampler_lora=False, qkv_bias=True, intermediate_specialization=False):
    """name -> shape for every parameter of SwinTransformerMTLoRA(num_classes=0), in the reference's
    named_parameters() naming (SURVEY.md §5 checkpoint row; validated against the real model by
    tests/golden/make_golden.py). `ranks[s]` = {'shared': r_s, task: r_t, ...} = mtlora.R_PER_TASK_LIST[s]."""
    E, ps = cfg.embed_dim, cfg.patch_size
    out = OrderedDict()
    out["patch_embed.proj.weight"] = (E, cfg.in_chans, ps, ps)
    out["patch_embed.proj.bias"] = (E,)
    out["patch_embed.norm.weight"] = (E,)
    out["patch_embed.norm.bias"] = (E,)

    def lin(prefix, K, N, r, tasks_on, bias=True):
        if r["shared"] > 0:
            out[prefix + "lora_shared_A"] = (r["shared"], K)
            out[prefix + "lora_shared_B"] = (N, r["shared"])
        out[prefix + "linear.weight"] = (N, K)
        if bias:
            out[prefix + "linear.bias"] = (N,)
        if r["shared"] > 0 and tasks_on:
            for t in sorted(cfg.tasks):
                out[prefix + "lora_tasks_A." + t] = (r[t], K)
            for t in sorted(cfg.tasks):
                out[prefix + "lora_tasks_B." + t] = (N, r[t])

    ws = cfg.window_size
    res = cfg.img_size // ps
    for s, depth in enumerate(cfg.depths):
        C = E * 2 ** s
        hid = int(C * cfg.mlp_ratio)
        w = min(ws, res // 2 ** s)
        for i in range(depth):
            b = f"layers.{s}.blocks.{i}."
            last = i == depth - 1 or intermediate_specialization
            out[b + "norm1.weight"] = (C,)
            out[b + "norm1.bias"] = (C,)
            out[b + "attn.relative_position_bias_table"] = ((2 * w - 1) ** 2, cfg.num_heads[s])
            lin(b + "attn.qkv.", C, 3 * C, ranks[s], False, qkv_bias)
            lin(b + "attn.proj.", C, C, ranks[s], last)
            out[b + "norm2.weight"] = (C,)
            out[b + "norm2.bias"] = (C,)
            lin(b + "mlp.fc1.", C, hid, ranks[s], last)
            lin(b + "mlp.fc2.", hid, C, ranks[s], last)
        if s < len(cfg.depths) - 1:
            d = f"layers.{s}.downsample."
            if downsampler_lora:
                lin(d + "reduction.", 4 * C, 2 * C, ranks[s], False, bias=False)
            else:
                out[d + "reduction.weight"] = (2 * C, 4 * C)
            out[d + "norm.weight"] = (4 * C,)
            out[d + "norm.bias"] = (4 * C,)
    return out


def param_value(name, shape):
    """Deterministic value for one parameter; magnitudes follow the reference's initialisers
    (swin_transformer_mtlora.py:715-724 trunc_normal .02; lora.py:236-247 kaiming A) except that lora B and the
    biases are non-zero so that every term of the forward contributes."""
    leaf = name.split(".")[-1]
    if "norm" in name:
        return std_uniform(name, shape, 0.1, 1.0) if leaf == "weight" else std_uniform(name, shape, 0.05)
    if "lora_shared_A" in name or "lora_tasks_A" in name:
        return std_uniform(name, shape, (1.0 / (3.0 * shape[1])) ** 0.5)
    if "lora_shared_B" in name or "lora_tasks_B" in name:
        return std_uniform(name, shape, 0.02)
    if "relative_position_bias_table" in name:
        return std_uniform(name, shape, 0.2)
    if name.startswith("patch_embed.proj"):
        return std_uniform(name, shape, 0.1 if leaf == "weight" else 0.02)
    if leaf == "bias":
        return std_uniform(name, shape, 0.02)
    return std_uniform(name, shape, 0.03)


def make_params(shapes, prefix="", device="cpu"):
    return {prefix + k: param_value(k, v).to(device) for k, v in shapes.items()}
