"""Context measurement (not a parity test): the reference algorithm in PyTorch eager on the SAME GPU (the oracle is
device-agnostic plain torch, i.e. what the reference's nn.Modules execute: F.linear + 2(1+T) skinny matmuls + mul/add
per MTLoRALinear, roll / window_partition copies, bmm attention, separate LayerNorm / GELU / residual kernels) against
this repo's fused path, BASELINE config 2 (Swin-T 448, 4 tasks, r 64/4, batch 32; MTL_SPEED_BATCH overrides). BASELINE.json's
north_star target is >= 5x at 1 GPU; the measured ratio is printed and written to gpurun_out/eager_vs_fused.json, and
the test only asserts that the fused path is faster."""
import contextlib
import io
import json
import os
import types

import pytest
import torch

from oracle import detgen
from oracle import mtlora_oracle as O

pytestmark = pytest.mark.gpu
TASKS = ["semseg", "normals", "sal", "human_parts"]


def _time(fn, n_warm, n):
    for _ in range(n_warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def test_fused_beats_eager_reference_algorithm():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable
    B, img = int(os.environ.get("MTL_SPEED_BATCH", "32")), 448   # BASELINE configs[1] batch (README.md:28)
    ranks = [dict({"shared": 64}, **{t: 4 for t in TASKS})] * 4
    ns = types.SimpleNamespace(
        R_PER_TASK_LIST=ranks, SHARED_SCALE=[4.0] * 4, SCALE_PER_TASK_LIST=[{t: 4.0 for t in TASKS}] * 4,
        DROPOUT=[0.05] * 4, TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False, SHARED_MODE="matrix",
        INTERMEDIATE_SPECIALIZATION=False, QKV_ENABLED=True, PROJ_ENABLED=True, FC1_ENABLED=True, FC2_ENABLED=True,
        DOWNSAMPLER_ENABLED=False)
    with contextlib.redirect_stdout(io.StringIO()):
        net = S.SwinTransformerMTLoRA(img_size=img, num_classes=0, drop_path_rate=0.2, tasks=TASKS, mtlora=ns)
        mark_only_lora_as_trainable(net)
    net.cuda().train()
    x = torch.randn(B, 3, img, img, device="cuda")

    def fused():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            st = net(x, return_stages=True)
        loss = sum(v.float().pow(2).mean() for _, tl in st for v in tl.values())
        loss.backward()
        for p in net.parameters():
            p.grad = None

    # the reference algorithm, eager, same parameters (fp32 masters, bf16 autocast like the fused arm)
    p = {n: v.detach().clone().requires_grad_(v.requires_grad) for n, v in net.named_parameters()}
    cfg = O.OracleConfig(img_size=img, tasks=tuple(TASKS), dropout=(0.05,) * 4, drop_path_rate=0.2, training=True)

    def eager():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            st = O.backbone(p, x, cfg)
        loss = O.backbone_loss(st)
        loss.backward()
        for v in p.values():
            v.grad = None

    t_f = _time(fused, 3, 8)
    t_e = _time(eager, 2, 4)
    out = {"batch": B, "fused_ms": t_f, "eager_ms": t_e, "fused_img_s": B / t_f * 1e3, "eager_img_s": B / t_e * 1e3,
           "speedup": t_e / t_f, "what": "backbone fwd + loss + bwd, Swin-T 448, 4 tasks, r 64/4, bf16 autocast, train mode"}
    print(json.dumps(out))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "eager_vs_fused.json"), "w") as f:
        json.dump(out, f)
    assert t_f < t_e
