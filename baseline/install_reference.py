#!/usr/bin/env python
"""Install the UNMODIFIED reference (scale-lab/MTLoRA, a plain-Python repo without setup.py / pyproject.toml, so
`pip install /root/reference` is impossible) into the git-ignored `baseline/_ref/` so that it travels to the GPU box:

    python baseline/install_reference.py            # source: $MTLORA_REFERENCE or /root/reference

Copies `models/`, `kernels/`, `configs/`, `mtl_loss_schemes.py`, `optimizer.py`, `main.py`, `utils.py`, `config.py`,
`logger.py`, `lr_scheduler.py`, `data/`, `evaluation/` byte for byte (the hot path, its MultiTaskSwin caller, its train
loop and its YAML/config machinery) and records a manifest with the sha256 of every file. Nothing under `baseline/_ref/` is ever committed (see .gitignore) and the product (`mtlora_b200/`) never
imports it: it is the reference arm of bench.py (`--impl reference-gpu`) and the live comparator of the `-m gpu`
parity tests. Third-party imports of the reference that are absent here (timm==0.9.2, termcolor, ptflops) are served
by the tiny stand-ins under `baseline/stubs/`.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
ITEMS = ["models", "kernels", "configs", "mtl_loss_schemes.py", "optimizer.py", "LICENSE",
         # the training script itself, for the drop-in test that runs main.py's own train_one_epoch (its third-party
         # imports that this image lacks — yacs, timm, easydict, imageio, matplotlib, scikit-image — are served by
         # baseline/stubs/)
         "main.py", "utils.py", "config.py", "logger.py", "lr_scheduler.py", "data", "evaluation"]


def install(src=None, quiet=False):
    src = src or os.environ.get("MTLORA_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(src, "models")):
        raise FileNotFoundError(f"no reference checkout at {src}")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    manifest = {}
    for it in ITEMS:
        s, d = os.path.join(src, it), os.path.join(DST, it)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so", "build"))
        elif os.path.exists(s):
            shutil.copy2(s, d)
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            with open(p, "rb") as fh:
                manifest[os.path.relpath(p, DST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f"installed {len(manifest)} reference files from {src} into {DST}")
    return DST


if __name__ == "__main__":
    install(sys.argv[1] if len(sys.argv) > 1 else None)
