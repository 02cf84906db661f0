"""Loader of the unmodified reference installed under baseline/_ref (baseline/install_reference.py).

TEST / BENCH INFRASTRUCTURE ONLY: used by bench.py's reference arms and by the parity tests; `mtlora_b200/` never
imports this. `load()` returns a namespace with the reference's own modules:
    .lora (models/lora.py), .swin (models/swin_transformer_mtlora.py), .swin_mtl (models/swin_mtl.py),
    .losses (mtl_loss_schemes.py), .optimizer (optimizer.py)
"""
import contextlib
import importlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")

NUM_OUTPUT = {"semseg": 21, "human_parts": 7, "sal": 1, "normals": 3, "edge": 1, "depth": 1}   # data/mtl_ds.py:744-804
LOSS_WEIGHTS = {"depth": 1.0, "semseg": 1.0, "human_parts": 2.0, "sal": 5.0, "edge": 50.0, "normals": 10.0}  # main.py:192-199


def available():
    return os.path.isfile(os.path.join(REF, "models", "swin_transformer_mtlora.py"))


def load():
    if not available():
        raise FileNotFoundError("baseline/_ref is missing: run `python baseline/install_reference.py` in the build container")
    for p in (STUBS, REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    for name in ("timm", "termcolor", "ptflops"):
        try:
            importlib.import_module(name)
        except ImportError:   # a partially installed real package: fall back to the stand-in
            sys.modules.pop(name, None)
    with contextlib.redirect_stdout(io.StringIO()):
        ns = types.SimpleNamespace(
            lora=importlib.import_module("models.lora"),
            swin=importlib.import_module("models.swin_transformer_mtlora"),
            swin_mtl=importlib.import_module("models.swin_mtl"),
            losses=importlib.import_module("mtl_loss_schemes"),
            optimizer=importlib.import_module("optimizer"),
        )
    assert os.path.realpath(ns.swin.__file__).startswith(os.path.realpath(REF)), ns.swin.__file__
    return ns


def load_main():
    """The reference's training script and config machinery themselves (main.py, config.py, utils.py, lr_scheduler.py,
    unmodified): returns a namespace with .config (module), .main (module), .utils, .lr_scheduler, .build (models.build)."""
    load()
    with contextlib.redirect_stdout(io.StringIO()):
        ns = types.SimpleNamespace(config=importlib.import_module("config"), main=importlib.import_module("main"),
                                   utils=importlib.import_module("utils"),
                                   lr_scheduler=importlib.import_module("lr_scheduler"),
                                   build=importlib.import_module("models.build"))
    return ns


def yaml_path(name):
    """Path of a shipped YAML, e.g. 'mtlora/tiny_448/mtlora_tiny_448_r64_scale4_pertask.yaml'."""
    return os.path.join(REF, "configs", name)


def reference_config(yaml_name, tasks, opts=None, batch_size=None):
    """config.get_config(args) of the reference for a shipped YAML, exactly as main.py:parse_option builds it
    (`--cfg <yaml> --pascal <path> --tasks a,b,c [--opts ...]`)."""
    m = load_main()
    args = types.SimpleNamespace(cfg=yaml_path(yaml_name), opts=list(opts) if opts else None, tasks=",".join(tasks),
                                 pascal="/nonexistent/PASCAL_MT", local_rank=0, batch_size=batch_size)
    with contextlib.redirect_stdout(io.StringIO()):
        return m.config.get_config(args)


class _Node(dict):
    """Minimal attribute-dict standing in for the yacs CfgNode the reference passes around (config.py)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def mtlora_node(tasks, r_shared=64, r_task=4, n_stages=4, dropout=0.05, scale=4.0, **over):
    """config.MODEL.MTLORA as config.py:476-557 builds it from a `*_pertask.yaml` (R_PER_TASK_LIST etc.)."""
    d = _Node(ENABLED=True,
              R_PER_TASK_LIST=[dict({"shared": r_shared}, **{t: r_task for t in tasks}) for _ in range(n_stages)],
              SHARED_SCALE=[scale] * n_stages, SCALE_PER_TASK_LIST=[{t: scale for t in tasks} for _ in range(n_stages)],
              DROPOUT=[dropout] * n_stages, TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False,
              SHARED_MODE="matrix", INTERMEDIATE_SPECIALIZATION=False, QKV_ENABLED=True, PROJ_ENABLED=True,
              FC1_ENABLED=True, FC2_ENABLED=True, DOWNSAMPLER_ENABLED=False, FREEZE_PRETRAINED=True, BIAS="none")
    d.update(over)
    return d


def mtl_config(tasks, img_size, mtlora):
    """The slice of the reference's config that models/swin_mtl.py:138-221 and mtl_loss_schemes.get_loss read."""
    return _Node(
        TASKS=list(tasks), MTL=True,
        DATA=_Node(IMG_SIZE=img_size),
        TASKS_CONFIG=_Node(ALL_TASKS=_Node(NUM_OUTPUT={t: NUM_OUTPUT[t] for t in tasks}), edge_w=0.95),
        MODEL=_Node(DECODER_HEAD=_Node({t: "hrnet" for t in tasks}), DECODER_CHANNELS=[18, 36, 72, 144],
                    DECODER_DOWNSAMPLER=True, PER_TASK_DOWNSAMPLER=True, SEGFORMER_CHANNELS=256, MTLORA=mtlora))


def synthetic_targets(tasks, batch, img_size, generator=None, device="cpu"):
    """Targets of the shapes data/mtl_ds.py yields (B, C, H, W float): class maps for semseg / human_parts (255 = ignore
    is not drawn), unit-ish normals, binary saliency / edge maps, positive depth."""
    import torch
    out = {}
    for t in tasks:
        if t in ("semseg", "human_parts"):
            out[t] = torch.randint(0, NUM_OUTPUT[t], (batch, 1, img_size, img_size), generator=generator).float()
        elif t == "normals":
            n = torch.randn(batch, 3, img_size, img_size, generator=generator)
            out[t] = n / n.norm(dim=1, keepdim=True).clamp_min(1e-6)
        elif t in ("sal", "edge"):
            out[t] = (torch.rand(batch, 1, img_size, img_size, generator=generator) > 0.7).float()
        else:
            out[t] = torch.rand(batch, 1, img_size, img_size, generator=generator) * 5 + 0.5
    return {k: v.to(device) for k, v in out.items()}
