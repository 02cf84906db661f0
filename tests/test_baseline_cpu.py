"""CPU checks of the reference arm's plumbing (baseline/): the installed copy of the reference is byte-identical to its
manifest, and the reference's own `MultiTaskSwin` (models/swin_mtl.py:138-221) accepts this repo's backbone as a drop-in:
same attributes read at construction, same state_dict keys / shapes, checkpoints interchangeable in both directions.
Skipped when baseline/_ref has not been installed (python baseline/install_reference.py)."""
import contextlib
import hashlib
import io
import json
import os

import pytest
import torch

from baseline import refload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not refload.available(), reason="baseline/_ref not installed")


def test_installed_reference_matches_manifest():
    ref = os.path.join(ROOT, "baseline", "_ref")
    man = json.load(open(os.path.join(ref, "MANIFEST.json")))
    assert len(man["files"]) >= 25
    for rel, digest in man["files"].items():
        if rel == "MANIFEST.json":
            continue
        with open(os.path.join(ref, rel), "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == digest, rel
    src = man["source"]
    if os.path.isdir(os.path.join(src, "models")):      # in the build container: identical to the checkout itself
        for rel in ("models/lora.py", "models/swin_transformer_mtlora.py", "models/swin_mtl.py", "mtl_loss_schemes.py"):
            with open(os.path.join(src, rel), "rb") as a, open(os.path.join(ref, rel), "rb") as b:
                assert a.read() == b.read(), rel


def test_multitask_swin_accepts_this_backbone():
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable
    ref = refload.load()
    tasks = ["semseg", "normals", "sal", "human_parts"]
    ml = refload.mtlora_node(tasks, 64, 4)
    cfg = refload.mtl_config(tasks, 224, ml)
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(0)
        rnet = ref.swin_mtl.MultiTaskSwin(ref.swin.SwinTransformerMTLoRA(img_size=224, num_classes=0, tasks=tasks, mtlora=ml), cfg)
        net = ref.swin_mtl.MultiTaskSwin(S.SwinTransformerMTLoRA(img_size=224, num_classes=0, tasks=tasks, mtlora=ml), cfg)
    assert (net.dims, net.input_res, net.window_size, tuple(net.img_size)) == \
           (rnet.dims, rnet.input_res, rnet.window_size, tuple(rnet.img_size))
    sd_r, sd = rnet.state_dict(), net.state_dict()
    assert list(sd_r) == list(sd)
    assert all(sd_r[k].shape == sd[k].shape and sd_r[k].dtype == sd[k].dtype for k in sd)
    net.load_state_dict(sd_r)
    rnet.load_state_dict(net.state_dict())
    # the trainable set main.py:254-262 selects is the same on both models: 8,344,634 parameters for this YAML
    with contextlib.redirect_stdout(io.StringIO()):
        ref.lora.mark_only_lora_as_trainable(rnet.backbone, bias="none")
        mark_only_lora_as_trainable(net.backbone, bias="none")
    tr_r = [n for n, p in rnet.named_parameters() if p.requires_grad]
    tr = [n for n, p in net.named_parameters() if p.requires_grad]
    assert tr == tr_r
    assert sum(p.numel() for p in net.parameters() if p.requires_grad) == 8344634
    # the reference's optimizer grouping (optimizer.py:62-78) works on it unchanged
    groups = ref.optimizer.set_weight_decay(net, net.backbone.no_weight_decay(), net.backbone.no_weight_decay_keywords())
    assert len(groups) == 2 and sum(len(g["params"]) for g in groups) == len(tr)


SHIPPED_YAMLS = ["mtlora_plus_tiny_448_r16_scale4.yaml", "mtlora_plus_tiny_448_r16_scale4_pertask.yaml",
                 "mtlora_plus_tiny_448_r32_scale4_pertask.yaml", "mtlora_plus_tiny_448_r4_scale4.yaml",
                 "mtlora_plus_tiny_448_r64_scale4_pertask.yaml", "mtlora_plus_tiny_448_r8_scale4.yaml",
                 "mtlora_tiny_448_r16_scale4_pertask.yaml", "mtlora_tiny_448_r32_scale4_pertask.yaml",
                 "mtlora_tiny_448_r64_scale4_pertask.yaml"]


@pytest.mark.parametrize("name", SHIPPED_YAMLS)
def test_every_shipped_yaml_builds_through_the_references_own_factory(name):
    """YAML surface: the reference's own config.get_config + models.build.build_model / build_mtl_model
    (models/build.py:19-91), with ONE import swapped as INTEGRATION.md §1 says, construct this repo's backbone for every
    shipped YAML; parameter names / shapes equal the all-reference model's, and the planner accepts every layer."""
    from mtlora_b200 import swin_transformer_mtlora as S
    m = refload.load_main()
    tasks = ["semseg", "normals", "sal", "human_parts"]
    config = refload.reference_config("mtlora/tiny_448/" + name, tasks)
    assert config.MODEL.MTLORA.ENABLED and len(config.MODEL.MTLORA.R_PER_TASK_LIST) == 4
    with contextlib.redirect_stdout(io.StringIO()):
        rnet = m.build.build_mtl_model(m.build.build_model(config), config)
        orig = m.build.SwinTransformerMTLoRA
        m.build.SwinTransformerMTLoRA = S.SwinTransformerMTLoRA          # <- the one-line swap
        try:
            net = m.build.build_mtl_model(m.build.build_model(config), config)
        finally:
            m.build.SwinTransformerMTLoRA = orig
    assert type(net.backbone).__module__.startswith("mtlora_b200")
    a = {n: tuple(p.shape) for n, p in rnet.named_parameters()}
    b = {n: tuple(p.shape) for n, p in net.named_parameters()}
    assert list(a) == list(b) and a == b
    plus = config.MODEL.MTLORA.DOWNSAMPLER_ENABLED
    assert ("backbone.layers.0.downsample.reduction.lora_shared_A" in b) == bool(plus)
