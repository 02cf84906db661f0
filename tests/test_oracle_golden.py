"""Pins oracle/mtlora_oracle.py against vectors produced by the unmodified reference (tests/golden/make_golden.py).
CPU only. Tolerances are fp32 round-off: both sides run the same math in a different op order."""
import numpy as np
import pytest
import torch

from oracle import detgen
from oracle import mtlora_oracle as O

TASKS = ["normals", "semseg"]
TSCALE = {"normals": 2.0, "semseg": 3.0}


def close(a, b, rtol=2e-5, atol=2e-6):
    a = a.detach().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max()
    assert err <= atol + rtol * scale, f"max abs err {err:.3e} vs scale {scale:.3e}"


def det_module_params(tag, names_shapes):
    return {n: detgen.param_value(tag + "." + n, s).requires_grad_() for n, s in names_shapes.items()}


def lin_shapes(K, N, r, tasks):
    s = {}
    if r["shared"] > 0:
        s["lora_shared_A"], s["lora_shared_B"] = (r["shared"], K), (N, r["shared"])
    s["linear.weight"], s["linear.bias"] = (N, K), (N,)
    if tasks and r["shared"] > 0:
        for t in tasks:
            s["lora_tasks_A." + t], s["lora_tasks_B." + t] = (r[t], K), (N, r[t])
    return s


@pytest.mark.parametrize("tag,K,N,r,use_tasks,xt,mode", [
    ("lin_shared", 96, 288, {"shared": 8}, False, False, "matrix"),
    ("lin_tasks", 96, 96, {"shared": 16, "normals": 4, "semseg": 4}, True, False, "matrix"),
    ("lin_xtasks", 96, 384, {"shared": 16, "normals": 4, "semseg": 8}, True, True, "matrix"),
    ("lin_r0", 64, 48, {"shared": 0}, False, False, "matrix"),
    ("lin_v2_tasks", 96, 96, {"shared": 16, "normals": 4, "semseg": 4}, True, False, "matrixv2"),
    ("lin_v2_xtasks", 96, 384, {"shared": 16, "normals": 4, "semseg": 8}, True, True, "matrixv2"),
    ("lin_tscale", 96, 192, {"shared": 16, "normals": 4, "semseg": 8}, True, True, "matrix+trainable_scales"),
])
def test_mtlora_linear(golden, tag, K, N, r, use_tasks, xt, mode):
    trainable = mode.endswith("+trainable_scales")
    mode = mode.split("+")[0]
    shapes = lin_shapes(K, N, r, TASKS if use_tasks else None)
    if trainable:   # the scales are parameters of the module (lora.py:210-216, 229-233), filled like every other one
        shapes["lora_shared_scale"] = (1,)
        for t in TASKS:
            shapes["lora_task_scale." + t] = (1,)
    p = det_module_params(tag, shapes)
    x = detgen.uniform(tag + ".x", (2, 49, K)).requires_grad_()
    x_tasks = {t: detgen.uniform(f"{tag}.x.{t}", (2, 49, K)).requires_grad_() for t in TASKS} if xt else None
    sc_sh = p["lora_shared_scale"] if trainable else 4.0
    sc_t = {t: p["lora_task_scale." + t] for t in TASKS} if trainable else TSCALE
    y, yt = O.mtlora_linear(p, "", x, x_tasks, TASKS if use_tasks else None, sc_sh, sc_t, mode=mode)
    loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
    if use_tasks:
        assert yt is not None and list(yt) == TASKS
        for t in TASKS:
            loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
    else:
        assert yt is None
    loss.backward()
    close(y, golden[tag + "/y"])
    if use_tasks:
        for t in TASKS:
            close(yt[t], golden[f"{tag}/y.{t}"])
    close(x.grad, golden[tag + "/dx"])
    if xt:
        for t in TASKS:
            close(x_tasks[t].grad, golden[f"{tag}/dx.{t}"])
    for n, v in p.items():
        if "lora" in n:
            close(v.grad, golden[f"{tag}/d.{n}"], rtol=5e-5)


@pytest.mark.parametrize("tag,B,H,W,C,shift,ws", [("win_s2", 2, 14, 14, 8, 2, 7), ("win_s3", 1, 28, 14, 4, 3, 7)])
def test_window_process_bitwise(golden, tag, B, H, W, C, shift, ws):
    """The reference's only unit test (kernels/window_process/unit_test.py) asserts bitwise equality."""
    x = detgen.uniform(tag + ".x", (B, H, W, C))
    part = O.roll_window_partition(x, shift, ws)
    assert np.array_equal(part.numpy(), golden[tag + "/partition"])
    assert torch.equal(O.window_merge_roll(part, shift, ws, H, W), x)
    merged = O.window_merge_roll(x.reshape(-1, ws, ws, C), shift, ws, H, W)
    assert np.array_equal(merged.numpy(), golden[tag + "/merge_of_x"])


def block_shapes(lora, ws):
    r = {"shared": 8, "normals": 4, "semseg": 4}
    s = {"norm1.weight": (96,), "norm1.bias": (96,), "attn.relative_position_bias_table": ((2 * ws - 1) ** 2, 3)}
    for pre, K, N, t in [("attn.qkv.", 96, 288, False), ("attn.proj.", 96, 96, lora)]:
        s.update({pre + k: v for k, v in lin_shapes(K, N, r, TASKS if t else None).items()})
    s.update({"norm2.weight": (96,), "norm2.bias": (96,)})
    for pre, K, N in [("mlp.fc1.", 96, 384), ("mlp.fc2.", 384, 96)]:
        s.update({pre + k: v for k, v in lin_shapes(K, N, r, TASKS if lora else None).items()})
    return s


@pytest.mark.parametrize("tag,H,shift,lora", [("blk_s0", 14, 0, False), ("blk_s3_lora", 14, 3, True),
                                             ("blk_s0_lora", 14, 0, True), ("blk_small", 7, 3, True)])
def test_swin_block(golden, tag, H, shift, lora):
    p = det_module_params(tag, block_shapes(lora, 7))
    cfg = O.OracleConfig(tasks=tuple(TASKS), shared_scale=(4.0,), task_scale=[{t: 4.0 for t in TASKS}], dropout=(0.0,))
    x = detgen.uniform(tag + ".x", (2, H * H, 96)).requires_grad_()
    y, yt = O.swin_block(p, "", x, H, H, 3, 7, shift, cfg, 0, lora)
    loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
    if lora:
        for t in TASKS:
            loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
    loss.backward()
    close(y, golden[tag + "/y"], rtol=5e-5)
    if lora:
        for t in TASKS:
            close(yt[t], golden[f"{tag}/y.{t}"], rtol=5e-5)
    close(x.grad, golden[tag + "/dx"], rtol=1e-4)
    for n, v in p.items():
        key = f"{tag}/d.{n}"
        if key in golden.files:
            close(v.grad, golden[key], rtol=2e-4)
    ws_eff = min(7, H)
    assert np.array_equal(O.relative_position_index(ws_eff).numpy(), golden[tag + "/relative_position_index"])
    if tag + "/attn_mask" in golden.files:
        assert np.array_equal(O.shift_attn_mask(H, H, 7, shift).numpy(), golden[tag + "/attn_mask"])


@pytest.mark.parametrize("tag,ds", [("pm_dense", False), ("pm_lora", True)])
def test_patch_merging(golden, tag, ds):
    if ds:
        s = {"reduction." + k: v for k, v in lin_shapes(384, 192, {"shared": 8}, None).items() if "bias" not in k}
    else:
        s = {"reduction.weight": (192, 384)}
    s.update({"norm.weight": (384,), "norm.bias": (384,)})
    p = det_module_params(tag, s)
    cfg = O.OracleConfig(shared_scale=(4.0,), dropout=(0.0,))
    x = detgen.uniform(tag + ".x", (2, 196, 96)).requires_grad_()
    y = O.patch_merging(p, "", x, 14, 14, cfg, 0)
    (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum().backward()
    close(y, golden[tag + "/y"])
    close(x.grad, golden[tag + "/dx"], rtol=5e-5)
    for n, v in p.items():
        close(v.grad, golden[f"{tag}/d.{n}"], rtol=1e-4)


def sample(t, stride=101):
    f = t.detach().reshape(-1).double()
    return np.concatenate([[f.sum().item(), f.abs().sum().item(), float(f.numel())], f[::stride].numpy()])


def test_backbone_config1(golden):
    """BASELINE.json configs[0]: Swin-T 224, 1 task (semseg), r = 4, batch 2, CPU fwd+bwd."""
    cfg = O.OracleConfig(img_size=224, tasks=("semseg",))
    ranks = [{"shared": 4, "semseg": 4}] * 4
    p = {k: v.requires_grad_() for k, v in detgen.make_params(detgen.backbone_param_shapes(cfg, ranks)).items()}
    img = detgen.uniform("c1.img", (2, 3, 224, 224), -2.0, 2.0)
    stages = O.backbone(p, img, cfg)
    loss = O.backbone_loss(stages)
    loss.backward()
    assert abs(loss.item() - golden["c1/loss"][0]) <= 1e-4 * abs(golden["c1/loss"][0])
    for s, (xs, tl) in enumerate(stages):
        for name, t in ((f"c1/stage{s}.x", xs), (f"c1/stage{s}.semseg", tl["semseg"])):
            g = golden[name]
            got = sample(t)
            assert got[2] == g[2]
            close(got[3:], g[3:], rtol=2e-4)
            assert abs(got[1] - g[1]) <= 1e-4 * g[1]
    none = sorted(n for n, v in p.items() if v.grad is None)
    assert none == sorted(golden["c1/none_grads"].tolist())
    checked = 0
    for n, v in p.items():
        key = f"c1/d.{n}"
        if key in golden.files:
            close(sample(v.grad, 53)[3:], golden[key][3:], rtol=1e-3)
            checked += 1
    assert checked > 150


# ---------------------------------------------------------------------------------------------------------------------
# tests/golden/headline_vectors.npz: the configurations the headline numbers are quoted on, 'addition' mode,
# INTERMEDIATE_SPECIALIZATION and MTLoRA+ (make_golden.headline)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,K,N,r,xt", [("lin_add_tasks", 96, 96, {"shared": 16, "normals": 4, "semseg": 4}, False),
                                          ("lin_add_xtasks", 96, 384, {"shared": 16, "normals": 4, "semseg": 8}, True)])
def test_mtlora_linear_addition(headline, tag, K, N, r, xt):
    """shared_mode='addition' (lora.py:275-282): no shared adapter; shared output = pretrained + LayerNorm(sum of the task
    outputs) with the layer's own `lora_norm` (:217-219)."""
    shapes = {"linear.weight": (N, K), "linear.bias": (N,)}
    for t in TASKS:
        shapes["lora_tasks_A." + t], shapes["lora_tasks_B." + t] = (r[t], K), (N, r[t])
    shapes["lora_norm.weight"], shapes["lora_norm.bias"] = (N,), (N,)
    assert sorted(shapes) == sorted(headline[tag + "/param_names"].tolist())
    p = det_module_params(tag, shapes)
    x = detgen.uniform(tag + ".x", (2, 49, K)).requires_grad_()
    x_tasks = {t: detgen.uniform(f"{tag}.x.{t}", (2, 49, K)).requires_grad_() for t in TASKS} if xt else None
    y, yt = O.mtlora_linear(p, "", x, x_tasks, TASKS, 4.0, TSCALE, mode="addition")
    loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
    for t in TASKS:
        loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
    loss.backward()
    close(y, headline[tag + "/y"])
    for t in TASKS:
        close(yt[t], headline[f"{tag}/y.{t}"])
    close(x.grad, headline[tag + "/dx"], rtol=5e-5)
    if xt:
        for t in TASKS:
            close(x_tasks[t].grad, headline[f"{tag}/dx.{t}"], rtol=5e-5)
    for n, v in p.items():
        if "lora" in n:
            close(v.grad, headline[f"{tag}/d.{n}"], rtol=1e-4)


HEADLINE_TASKS6 = ["semseg", "normals", "sal", "human_parts", "depth", "edge"]
HEADLINE_CASES = {   # mirrors tests/golden/make_golden.py::HEADLINE_CASES
    "h_t448": dict(img=448, embed_dim=96, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), n_tasks=4, r_s=64, r_t=4, B=1),
    "h_s448": dict(img=448, embed_dim=96, depths=(2, 2, 18, 2), heads=(3, 6, 12, 24), n_tasks=4, r_s=64, r_t=4, B=1),
    "h_b448": dict(img=448, embed_dim=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), n_tasks=6, r_s=32, r_t=4, B=1),
    "h_t224_r64all": dict(img=224, embed_dim=96, depths=(2, 2, 6, 2), heads=(3, 6, 12, 24), n_tasks=4, r_s=64, r_t=64, B=2),
    "h_t224_interm": dict(img=224, embed_dim=96, depths=(2, 2, 2, 2), heads=(3, 6, 12, 24), n_tasks=2, r_s=16, r_t=4, B=2,
                          interm=True),
    "h_t224_plus": dict(img=224, embed_dim=96, depths=(2, 2, 2, 2), heads=(3, 6, 12, 24), n_tasks=2, r_s=16, r_t=4, B=2,
                        downsampler=True),
}


# the two 18-block 448 cases take ~20 s each on 8 cores; they run on the GPU against the same vectors instead
@pytest.mark.parametrize("case", ["h_t448", "h_t224_r64all", "h_t224_interm", "h_t224_plus"])
def test_backbone_headline(headline, case):
    """The oracle on the configurations the numbers are quoted on (BASELINE.json configs[1], [4]) and on the
    INTERMEDIATE_SPECIALIZATION / MTLoRA+ variants, against the unmodified reference's own outputs and gradients."""
    c = HEADLINE_CASES[case]
    tasks = HEADLINE_TASKS6[:c["n_tasks"]]
    cfg = O.OracleConfig(img_size=c["img"], embed_dim=c["embed_dim"], depths=c["depths"], num_heads=c["heads"],
                         tasks=tuple(tasks), intermediate_specialization=c.get("interm", False))
    ranks = [dict({"shared": c["r_s"]}, **{t: c["r_t"] for t in tasks})] * 4
    shapes = detgen.backbone_param_shapes(cfg, ranks, downsampler_lora=c.get("downsampler", False),
                                          intermediate_specialization=c.get("interm", False))
    p = {k: v.requires_grad_() for k, v in detgen.make_params(shapes).items()}
    img = detgen.uniform(case + ".img", (c["B"], 3, c["img"], c["img"]), -2.0, 2.0)
    stages = O.backbone(p, img, cfg)
    loss = O.backbone_loss(stages)
    loss.backward()
    g0 = headline[f"{case}/loss"][0]
    assert abs(loss.item() - g0) <= 1e-4 * abs(g0)
    for s, (xs, tl) in enumerate(stages):
        for name, t in [(f"{case}/stage{s}.x", xs)] + [(f"{case}/stage{s}.{k}", tl[k]) for k in tasks]:
            f = t.detach().reshape(-1).double()
            stat = headline[name + ".stat"]
            assert stat[2] == f.numel()
            close(f[::101], headline[name], rtol=3e-4)
            assert abs(f.abs().sum().item() - stat[1]) <= 1e-4 * stat[1]
    none = sorted(n for n, v in p.items() if v.grad is None)
    assert none == sorted(headline[f"{case}/none_grads"].tolist())
    checked = 0
    for n, v in p.items():
        key = f"{case}/d.{n}"
        if key in headline.files:
            close(v.grad.reshape(-1)[::53], headline[key], rtol=2e-3, atol=1e-7)
            checked += 1
    assert checked > 150
