"""Generate tests/golden/reference_vectors.npz by running the UNMODIFIED reference (imported from /root/reference)
on the deterministic inputs of oracle/detgen.py. Run in the build container only (the GPU box has no reference):

    python tests/golden/make_golden.py

The reference needs `timm.models.layers.{DropPath,to_2tuple,trunc_normal_}` (timm==0.9.2 is not installed here);
a stub with timm's published semantics is injected before import. Nothing from the reference is copied into the
repo — only the numbers it produces.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("MTLORA_REFERENCE", "/root/reference")

from oracle import detgen  # noqa: E402
from oracle.mtlora_oracle import OracleConfig  # noqa: E402


def install_stubs():
    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath = DropPath
    layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    for name, attrs in (("termcolor", {"colored": lambda s, *a, **k: s}),
                        ("ptflops", {"get_model_complexity_info": lambda *a, **k: (0, 0)})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)


def mtlora_ns(ranks, tasks, dropout=0.0, downsampler=False, scale=4.0):
    n = len(ranks)
    return types.SimpleNamespace(
        R_PER_TASK_LIST=ranks, SHARED_SCALE=[scale] * n, SCALE_PER_TASK_LIST=[{t: scale for t in tasks} for _ in range(n)],
        DROPOUT=[dropout] * n, TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False, SHARED_MODE="matrix",
        INTERMEDIATE_SPECIALIZATION=False, QKV_ENABLED=True, PROJ_ENABLED=True, FC1_ENABLED=True, FC2_ENABLED=True,
        DOWNSAMPLER_ENABLED=downsampler)


def load_det(module, prefix_for_values=""):
    """Overwrite every parameter of `module` with detgen values keyed by its own parameter name."""
    with torch.no_grad():
        for name, prm in module.named_parameters():
            prm.copy_(detgen.param_value(prefix_for_values + name, tuple(prm.shape)))


def sample(t, stride=101):
    f = t.detach().reshape(-1).double()
    return np.concatenate([[f.sum().item(), f.abs().sum().item(), float(f.numel())], f[::stride].numpy()]).astype(np.float64)


HEADLINE_TASKS6 = ["semseg", "normals", "sal", "human_parts", "depth", "edge"]
# tag -> ctor arguments of the backbone cases the headline numbers are quoted on (BASELINE.json configs[1..4])
HEADLINE_CASES = {
    # configs[1]: Swin-T 448, 4 tasks, r 64 / 4 (configs/mtlora/tiny_448/mtlora_tiny_448_r64_scale4_pertask.yaml)
    "h_t448": dict(img=448, embed_dim=96, depths=[2, 2, 6, 2], heads=[3, 6, 12, 24], n_tasks=4, r_s=64, r_t=4, B=1),
    # configs[2]: Swin-S 448, 4 tasks, r 64 / 4
    "h_s448": dict(img=448, embed_dim=96, depths=[2, 2, 18, 2], heads=[3, 6, 12, 24], n_tasks=4, r_s=64, r_t=4, B=1),
    # configs[3]: Swin-B 448, 6 tasks, r 32 / 4
    "h_b448": dict(img=448, embed_dim=128, depths=[2, 2, 18, 2], heads=[4, 8, 16, 32], n_tasks=6, r_s=32, r_t=4, B=1),
    # configs[4] / non-pertask YAMLs: all-equal ranks of 64 (rank space 5 x 64 = 320 columns)
    "h_t224_r64all": dict(img=224, embed_dim=96, depths=[2, 2, 6, 2], heads=[3, 6, 12, 24], n_tasks=4, r_s=64, r_t=64, B=2),
    # INTERMEDIATE_SPECIALIZATION=True (swin_transformer_mtlora.py:53,175): every block carries the task adapters
    "h_t224_interm": dict(img=224, embed_dim=96, depths=[2, 2, 2, 2], heads=[3, 6, 12, 24], n_tasks=2, r_s=16, r_t=4, B=2,
                          interm=True),
    # MTLoRA+ (DOWNSAMPLER_ENABLED=True: the reductions are frozen and carry a shared adapter)
    "h_t224_plus": dict(img=224, embed_dim=96, depths=[2, 2, 2, 2], heads=[3, 6, 12, 24], n_tasks=2, r_s=16, r_t=4, B=2,
                        downsampler=True),
}


def headline_build(mod, case, ns_fn):
    """Backbone of `case` from module `mod` (the reference's or mtlora_b200's swin_transformer_mtlora)."""
    import contextlib
    import io
    c = HEADLINE_CASES[case]
    tasks = HEADLINE_TASKS6[:c["n_tasks"]]
    ranks = [dict({"shared": c["r_s"]}, **{t: c["r_t"] for t in tasks})] * 4
    ns = ns_fn(ranks, tasks, downsampler=c.get("downsampler", False))
    ns.INTERMEDIATE_SPECIALIZATION = c.get("interm", False)
    with contextlib.redirect_stdout(io.StringIO()):
        net = mod.SwinTransformerMTLoRA(img_size=c["img"], patch_size=4, in_chans=3, num_classes=0,
                                        embed_dim=c["embed_dim"], depths=c["depths"], num_heads=c["heads"],
                                        window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0,
                                        ape=False, patch_norm=True, tasks=tasks, mtlora=ns)
    return net, tasks


def sample32(t, stride):
    """(float64 [sum, sum |.|, numel], float32 strided sample) of a tensor."""
    f = t.detach().reshape(-1).double().cpu()
    return np.array([f.sum().item(), f.abs().sum().item(), float(f.numel())]), f[::stride].float().numpy()


HEADLINE_TRAINABLE = ("lora_", "norm", "relative_position_bias_table", "downsample.reduction", "patch_embed")


def headline(ref, MTLoRALinear):
    """Golden vectors of the configurations the headline numbers are quoted on -> tests/golden/headline_vectors.npz
    (samples: every 101st element of each stage tensor, every 53rd of each trainable gradient, plus sums)."""
    G = {}
    for case, c in HEADLINE_CASES.items():
        net, tasks = headline_build(ref, case, mtlora_ns)
        net.eval()
        load_det(net)
        img = detgen.uniform(case + ".img", (c["B"], 3, c["img"], c["img"]), -2.0, 2.0)
        stages = net(img, return_stages=True)
        loss = sum(v.pow(2).mean() for _, tl in stages for v in tl.values())
        loss.backward()
        G[f"{case}/loss"] = np.array([loss.item()])
        for s, (xs, tl) in enumerate(stages):
            G[f"{case}/stage{s}.x.stat"], G[f"{case}/stage{s}.x"] = sample32(xs, 101)
            for t in tasks:
                G[f"{case}/stage{s}.{t}.stat"], G[f"{case}/stage{s}.{t}"] = sample32(tl[t], 101)
        none_grads = []
        for name, prm in net.named_parameters():
            if prm.grad is None:
                none_grads.append(name)
            elif any(k in name for k in HEADLINE_TRAINABLE):
                G[f"{case}/d.{name}.stat"], G[f"{case}/d.{name}"] = sample32(prm.grad, 53)
        G[f"{case}/none_grads"] = np.array(none_grads)
        print(f"  {case}: loss {loss.item():.6f}, {sum(1 for k in G if k.startswith(case + '/d.')) // 2} gradient tensors")
        del net, stages, loss

    # shared_mode='addition' (lora.py:275-282 + lora_norm :217-219): layer level and inside a block
    tasks = ["normals", "semseg"]
    for tag, K, N, r, xt in [("lin_add_tasks", 96, 96, {"shared": 16, "normals": 4, "semseg": 4}, False),
                             ("lin_add_xtasks", 96, 384, {"shared": 16, "normals": 4, "semseg": 8}, True)]:
        m = MTLoRALinear(K, N, r=r, lora_shared_scale=4.0, lora_task_scale={t: 2.0 + i for i, t in enumerate(tasks)},
                         lora_dropout=0.0, tasks=tasks, shared_mode="addition")
        load_det(m, tag + ".")
        x = detgen.uniform(tag + ".x", (2, 49, K)).requires_grad_()
        x_tasks = {t: detgen.uniform(f"{tag}.x.{t}", (2, 49, K)).requires_grad_() for t in tasks} if xt else None
        y, yt = m(x, x_tasks)
        loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
        for t in tasks:
            loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
        loss.backward()
        G[tag + "/y"] = y.detach().numpy()
        for t in tasks:
            G[f"{tag}/y.{t}"] = yt[t].detach().numpy()
        G[tag + "/dx"] = x.grad.numpy()
        if xt:
            for t in tasks:
                G[f"{tag}/dx.{t}"] = x_tasks[t].grad.numpy()
        G[tag + "/param_names"] = np.array([n for n, _ in m.named_parameters()])
        for name, prm in m.named_parameters():
            if prm.grad is not None and "lora" in name:
                G[f"{tag}/d.{name}"] = prm.grad.numpy()
    ranks1 = [{"shared": 8, "normals": 4, "semseg": 4}]
    tag = "blk_add"
    import contextlib
    import io
    ns = mtlora_ns(ranks1, tasks)
    ns.SHARED_MODE = "addition"
    with contextlib.redirect_stdout(io.StringIO()):
        blk = ref.SwinTransformerBlock(dim=96, input_resolution=(14, 14), num_heads=3, window_size=7, shift_size=3,
                                       lora=True, tasks=tasks, mtlora=ns, layer_idx=0)
    blk.eval()
    load_det(blk, tag + ".")
    x = detgen.uniform(tag + ".x", (2, 196, 96)).requires_grad_()
    y, yt = blk(x)
    loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
    for t in tasks:
        loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
    loss.backward()
    G[tag + "/y"] = y.detach().numpy()
    for t in tasks:
        G[f"{tag}/y.{t}"] = yt[t].detach().numpy()
    G[tag + "/dx"] = x.grad.numpy()
    G[tag + "/param_names"] = np.array([n for n, _ in blk.named_parameters()])
    for name, prm in blk.named_parameters():
        if prm.grad is not None and ("lora" in name or "norm" in name or "relative_position" in name):
            G[f"{tag}/d.{name}"] = prm.grad.numpy()

    out = os.path.join(ROOT, "tests", "golden", "headline_vectors.npz")
    np.savez_compressed(out, **{k: (v.astype(np.float32) if v.dtype == np.float64 and not (k.endswith(".stat") or k.endswith("/loss")) else v)
                                for k, v in G.items()})
    print(f"wrote {out}: {len(G)} arrays, {os.path.getsize(out) / 1e6:.2f} MB")


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from models.lora import MTLoRALinear
        from models import swin_transformer_mtlora as ref
    G = {}
    torch.set_grad_enabled(True)

    # ---- 1. MTLoRALinear (models/lora.py:159-284) --------------------------------------------------------------
    tasks = ["normals", "semseg"]
    for tag, K, N, r, use_tasks, xt, mode in [
        ("lin_shared", 96, 288, {"shared": 8}, False, False, "matrix"),
        ("lin_tasks", 96, 96, {"shared": 16, "normals": 4, "semseg": 4}, True, False, "matrix"),
        ("lin_xtasks", 96, 384, {"shared": 16, "normals": 4, "semseg": 8}, True, True, "matrix"),
        ("lin_r0", 64, 48, {"shared": 0}, False, False, "matrix"),
        ("lin_v2_tasks", 96, 96, {"shared": 16, "normals": 4, "semseg": 4}, True, False, "matrixv2"),   # lora.py:267-274
        ("lin_v2_xtasks", 96, 384, {"shared": 16, "normals": 4, "semseg": 8}, True, True, "matrixv2"),
        # trainable scales (lora.py:210-216, 229-233): Parameters initialised to the given floats; the reference needs a
        # FLOAT lora_task_scale in this mode (torch.FloatTensor([lora_task_scale]))
        ("lin_tscale", 96, 192, {"shared": 16, "normals": 4, "semseg": 8}, True, True, "matrix+trainable_scales"),
    ]:
        trainable = mode.endswith("+trainable_scales")
        mode = mode.split("+")[0]
        m = MTLoRALinear(K, N, r=r, lora_shared_scale=4.0,
                         lora_task_scale=2.5 if trainable else {t: 2.0 + i for i, t in enumerate(tasks)},
                         lora_dropout=0.0, tasks=tasks if use_tasks else None, shared_mode=mode,
                         trainable_scale_shared=trainable, trainable_scale_per_task=trainable)
        load_det(m, tag + ".")
        x = detgen.uniform(tag + ".x", (2, 49, K)).requires_grad_()
        x_tasks = {t: detgen.uniform(f"{tag}.x.{t}", (2, 49, K)).requires_grad_() for t in tasks} if xt else None
        y, yt = m(x, x_tasks)
        loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
        if yt is not None:
            for t in tasks:
                loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
        loss.backward()
        G[tag + "/y"] = y.detach().numpy()
        if yt is not None:
            for t in tasks:
                G[f"{tag}/y.{t}"] = yt[t].detach().numpy()
        G[tag + "/dx"] = x.grad.numpy()
        if xt:
            for t in tasks:
                G[f"{tag}/dx.{t}"] = x_tasks[t].grad.numpy()
        for name, prm in m.named_parameters():
            if prm.grad is not None and "lora" in name:
                G[f"{tag}/d.{name}"] = prm.grad.numpy()

    # ---- 2. window index math: the reference unit test's oracle (kernels/window_process/unit_test.py:96-115) ---
    for tag, B, H, W, C, shift, ws in [("win_s2", 2, 14, 14, 8, 2, 7), ("win_s3", 1, 28, 14, 4, 3, 7)]:
        x = detgen.uniform(tag + ".x", (B, H, W, C))
        rolled = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
        part = ref.window_partition(rolled, ws)
        G[tag + "/partition"] = part.numpy()
        back = torch.roll(ref.window_reverse(part, ws, H, W), shifts=(shift, shift), dims=(1, 2))
        assert torch.equal(back, x)
        G[tag + "/merge_of_x"] = torch.roll(ref.window_reverse(x.reshape(-1, ws, ws, C), ws, H, W), shifts=(shift, shift),
                                            dims=(1, 2)).numpy()

    # ---- 3. WindowAttention / SwinTransformerBlock / PatchMerging ------------------------------------------------
    ranks1 = [{"shared": 8, "normals": 4, "semseg": 4}]
    for tag, H, shift, lora in [("blk_s0", 14, 0, False), ("blk_s3_lora", 14, 3, True), ("blk_s0_lora", 14, 0, True),
                                ("blk_small", 7, 3, True)]:
        with contextlib.redirect_stdout(io.StringIO()):
            blk = ref.SwinTransformerBlock(dim=96, input_resolution=(H, H), num_heads=3, window_size=7, shift_size=shift,
                                           lora=lora, tasks=tasks, mtlora=mtlora_ns(ranks1, tasks), layer_idx=0)
        blk.eval()
        load_det(blk, tag + ".")
        x = detgen.uniform(tag + ".x", (2, H * H, 96)).requires_grad_()
        y, yt = blk(x)
        loss = (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum()
        if yt is not None:
            for t in tasks:
                loss = loss + (yt[t] * detgen.uniform(f"{tag}.gy.{t}", tuple(y.shape))).sum()
        loss.backward()
        G[tag + "/y"] = y.detach().numpy()
        if yt is not None:
            for t in tasks:
                G[f"{tag}/y.{t}"] = yt[t].detach().numpy()
        G[tag + "/dx"] = x.grad.numpy()
        for name, prm in blk.named_parameters():
            if prm.grad is not None and ("lora" in name or "norm" in name or "relative_position" in name):
                G[f"{tag}/d.{name}"] = prm.grad.numpy()
        if blk.attn_mask is not None:
            G[tag + "/attn_mask"] = blk.attn_mask.numpy()
        G[tag + "/relative_position_index"] = blk.attn.relative_position_index.numpy()

    for tag, ds in [("pm_dense", False), ("pm_lora", True)]:
        pm = ref.PatchMerging((14, 14), 96, layer_idx=0, mtlora=mtlora_ns(ranks1, tasks, downsampler=ds))
        load_det(pm, tag + ".")
        x = detgen.uniform(tag + ".x", (2, 196, 96)).requires_grad_()
        y = pm(x)
        (y * detgen.uniform(tag + ".gy", tuple(y.shape))).sum().backward()
        G[tag + "/y"] = y.detach().numpy()
        G[tag + "/dx"] = x.grad.numpy()
        for name, prm in pm.named_parameters():
            G[f"{tag}/d.{name}"] = prm.grad.numpy()

    # ---- 4. full backbone, BASELINE config 1: Swin-T 224, tasks [semseg], r = 4/4, B = 2, fp32, eval-mode ---------
    cfg = OracleConfig(img_size=224, tasks=("semseg",))
    ranks = [{"shared": 4, "semseg": 4}] * 4
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref.SwinTransformerMTLoRA(img_size=224, patch_size=4, in_chans=3, num_classes=0, embed_dim=96,
                                        depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4.0,
                                        qkv_bias=True, drop_rate=0.0, drop_path_rate=0.0, ape=False, patch_norm=True,
                                        tasks=["semseg"], mtlora=mtlora_ns(ranks, ["semseg"]))
    net.eval()
    shapes = detgen.backbone_param_shapes(cfg, ranks)
    got = {n: tuple(p.shape) for n, p in net.named_parameters()}
    assert list(got.keys()) == list(shapes.keys()), "parameter names/order differ from detgen.backbone_param_shapes"
    assert got == dict(shapes), "parameter shapes differ from detgen.backbone_param_shapes"
    load_det(net)
    img = detgen.uniform("c1.img", (2, 3, 224, 224), -2.0, 2.0)
    stages = net(img, return_stages=True)
    loss = sum(v.pow(2).mean() for _, tl in stages for v in tl.values())
    loss.backward()
    G["c1/loss"] = np.array([loss.item()])
    for s, (xs, tl) in enumerate(stages):
        G[f"c1/stage{s}.x"] = sample(xs)
        G[f"c1/stage{s}.semseg"] = sample(tl["semseg"])
    none_grads = []
    for name, prm in net.named_parameters():
        if prm.grad is None:
            none_grads.append(name)
        elif any(k in name for k in ("lora_", "norm", "relative_position_bias_table", "downsample.reduction", "patch_embed")):
            G[f"c1/d.{name}"] = sample(prm.grad, 53)
    G["c1/none_grads"] = np.array(none_grads)

    # MTLoRA+ variant names (DOWNSAMPLER_ENABLED=True): only validate the parameter surface
    with contextlib.redirect_stdout(io.StringIO()):
        net2 = ref.SwinTransformerMTLoRA(img_size=224, num_classes=0, tasks=["semseg"],
                                         mtlora=mtlora_ns(ranks, ["semseg"], downsampler=True))
    shapes2 = detgen.backbone_param_shapes(cfg, ranks, downsampler_lora=True)
    got2 = {n: tuple(p.shape) for n, p in net2.named_parameters()}
    assert list(got2.keys()) == list(shapes2.keys()) and got2 == dict(shapes2)

    if os.environ.get("MTLORA_GOLDEN_SKIP_HEADLINE", "0") != "1":
        headline(ref, MTLoRALinear)

    out = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    np.savez_compressed(out, **{k: (v.astype(np.float32) if v.dtype == np.float64 and not k.startswith("c1/") else v)
                                for k, v in G.items()})
    print(f"wrote {out}: {len(G)} arrays, {os.path.getsize(out) / 1e6:.2f} MB; unused params: {none_grads}")


if __name__ == "__main__":
    main()
