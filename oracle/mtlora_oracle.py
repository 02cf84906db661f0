"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A functional, plain-PyTorch (fp32 by default, device-agnostic, autograd-differentiable) restatement of the
algorithm of the reference hot path: MTLoRALinear + windowed attention + SwinTransformerBlock + PatchMerging +
the SwinTransformerMTLoRA stage loop. Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module, and only as the checker / the timed CPU baseline — the product
(`mtlora_b200/`) never imports it and has no CPU fallback.

Parity status: PINNED against the reference itself. `tests/golden/make_golden.py` imports the unmodified reference
modules from /root/reference (with a 3-symbol timm stub), feeds them the deterministic parameters / inputs of
`oracle/detgen.py`, and stores their outputs and gradients under `tests/golden/`; `tests/test_oracle_golden.py`
checks this file against those vectors. The reference's own (only) test, kernels/window_process/unit_test.py, is
restated in `roll_window_partition` / `window_merge_roll` and covered by the same golden file.

Parameters are passed as a flat dict keyed by the reference's state_dict names (so a reference state_dict can be
dropped in unchanged); every function cites the reference lines it follows (paths relative to the reference
checkout). Third-party arithmetic the reference relies on: torch (unpinned, README.md:23 "PyTorch>=1.12") for
linear / softmax / LayerNorm(eps=1e-5) / exact-erf GELU / conv2d / dropout, timm==0.9.2 (requirements.txt:17)
for DropPath — restated in `drop_path` below from its published definition (per-sample Bernoulli(keep) mask of
shape (B,1,...) divided by keep, identity in eval mode).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class OracleConfig:
    """The subset of the reference's ctor arguments / `mtlora` CfgNode the path reads
    (models/build.py:39-58, swin_transformer_mtlora.py:50-62,164-178,442-447, config.py:476-557)."""
    img_size: int = 224
    patch_size: int = 4
    in_chans: int = 3
    embed_dim: int = 96
    depths: Sequence[int] = (2, 2, 6, 2)
    num_heads: Sequence[int] = (3, 6, 12, 24)
    window_size: int = 7
    mlp_ratio: float = 4.0
    tasks: Sequence[str] = ("semseg",)
    shared_scale: Sequence[float] = (4.0, 4.0, 4.0, 4.0)          # mtlora.SHARED_SCALE[stage]
    task_scale: Optional[List[Dict[str, float]]] = None           # mtlora.SCALE_PER_TASK_LIST[stage]; None -> shared_scale
    dropout: Sequence[float] = (0.0, 0.0, 0.0, 0.0)               # mtlora.DROPOUT[stage]
    drop_path_rate: float = 0.0
    shared_mode: str = "matrix"
    intermediate_specialization: bool = False                     # mtlora.INTERMEDIATE_SPECIALIZATION (:53,61,175)
    training: bool = False
    ln_eps: float = 1e-5

    def stage_task_scale(self, s, task):
        if self.task_scale is None:
            return self.shared_scale[s]
        return self.task_scale[s][task]


# ---------------------------------------------------------------------------------------------------------------
# stochastic pieces (only active with cfg.training and p > 0; parity tests run them off)
# ---------------------------------------------------------------------------------------------------------------
def drop_path(x, p, training):
    """timm 0.9.2 DropPath (call sites swin_transformer_mtlora.py:290-291,390,392,399,403,407)."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    shape = (x.shape[0],) + (1,) * (x.dim() - 1)
    mask = x.new_empty(shape).bernoulli_(keep)
    if keep > 0.0:
        mask.div_(keep)
    return x * mask


# ---------------------------------------------------------------------------------------------------------------
# MTLoRALinear — models/lora.py:253-284
# ---------------------------------------------------------------------------------------------------------------
def mtlora_linear(p, prefix, x, x_tasks=None, tasks=None, scale_shared=1.0, scale_tasks=None, dropout=0.0,
                  training=False, mode="matrix"):
    """Returns (shared_out, {task: out} | None).

    `p[prefix + 'linear.weight']` (+ bias) is the frozen layer; adapters are looked up by the reference's
    parameter names. A layer without `lora_shared_A` is the r == 0 / CompatLinear case (lora.py:256-257,
    swin_transformer_mtlora.py:36-41)."""
    W = p[prefix + "linear.weight"] if (prefix + "linear.weight") in p else p[prefix + "weight"]
    b = p.get(prefix + "linear.bias", p.get(prefix + "bias"))
    pretrained = F.linear(x, W, b)                                             # lora.py:255
    has_tasks = tasks is not None and (prefix + "lora_tasks_A." + tasks[0]) in p
    if (prefix + "lora_shared_A") not in p and not has_tasks:
        return pretrained, None
    xd = F.dropout(x, dropout, training) if dropout > 0 else x                 # lora.py:258
    if (prefix + "lora_norm.weight") in p:                                     # shared_mode 'addition', lora.py:275-282
        lora_tasks = {}
        for t in tasks:
            xin = xd if x_tasks is None else x_tasks[t]
            At, Bt = p[prefix + "lora_tasks_A." + t], p[prefix + "lora_tasks_B." + t]
            lora_tasks[t] = pretrained + (xin @ At.t() @ Bt.t()) * scale_tasks[t]
        tot = torch.stack(list(lora_tasks.values()), 0).sum(0)
        lora = F.layer_norm(tot, (tot.shape[-1],), p[prefix + "lora_norm.weight"], p[prefix + "lora_norm.bias"], 1e-5)
        return pretrained + lora, lora_tasks
    A, B = p[prefix + "lora_shared_A"], p[prefix + "lora_shared_B"]
    lora = (xd @ A.t() @ B.t()) * scale_shared                                 # lora.py:260-261
    lora_tasks = None
    if has_tasks:
        lora_tasks = {}
        for t in tasks:                                                        # lora.py:262-266 / 267-274
            xin = xd if x_tasks is None else x_tasks[t]
            At, Bt = p[prefix + "lora_tasks_A." + t], p[prefix + "lora_tasks_B." + t]
            delta = (xin @ At.t() @ Bt.t()) * scale_tasks[t]
            if mode == "matrix":
                lora_tasks[t] = pretrained + delta
            elif mode == "matrixv2":
                lora_tasks[t] = pretrained + lora + delta
            else:
                raise NotImplementedError(mode)
    return pretrained + lora, lora_tasks                                       # lora.py:284


# ---------------------------------------------------------------------------------------------------------------
# window index math — swin_transformer_mtlora.py:84-116, 148-160, 297-319; kernels/window_process
# ---------------------------------------------------------------------------------------------------------------
def window_partition(x, ws):
    """(B, H, W, C) -> (B*nW, ws, ws, C), windows ordered batch, window-row, window-col (:84-98)."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def window_reverse(windows, ws, H, W):
    """(B*nW, ws, ws, C) -> (B, H, W, C) (:101-116)."""
    B = windows.shape[0] // ((H // ws) * (W // ws))
    x = windows.reshape(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def roll_window_partition(x, shift, ws):
    """torch.roll(-shift) + window_partition (:338-342) == WindowProcess.apply(x, ..., -shift, ws)
    (kernels/window_process/unit_test.py:96-103 `pyt_forward`)."""
    if shift > 0:
        x = torch.roll(x, shifts=(-shift, -shift), dims=(1, 2))
    return window_partition(x, ws)


def window_merge_roll(windows, shift, ws, H, W):
    """window_reverse + torch.roll(+shift) (:365-377) == WindowProcessReverse (unit_test.py:106-115)."""
    x = window_reverse(windows, ws, H, W)
    if shift > 0:
        x = torch.roll(x, shifts=(shift, shift), dims=(1, 2))
    return x


def relative_position_index(ws):
    """(ws*ws, ws*ws) int64: (yi - yj + ws-1) * (2ws-1) + (xi - xj + ws-1) (:148-160)."""
    idx = torch.arange(ws * ws)
    yi, xi = idx // ws, idx % ws
    return (yi[:, None] - yi[None, :] + ws - 1) * (2 * ws - 1) + (xi[:, None] - xi[None, :] + ws - 1)


def shift_attn_mask(H, W, ws, shift):
    """(nW, N, N) of {0, -100}: 3x3 region ids on the rolled grid, differing regions masked (:297-319)."""
    def rid(size):
        r = torch.zeros(size, dtype=torch.long)
        r[size - ws:size - shift] = 1
        r[size - shift:] = 2
        return r
    reg = (3 * rid(H)[:, None] + rid(W)[None, :]).reshape(1, H, W, 1).float()
    mw = window_partition(reg, ws).reshape(-1, ws * ws)
    diff = mw[:, None, :] - mw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


# ---------------------------------------------------------------------------------------------------------------
# WindowAttention.forward — swin_transformer_mtlora.py:186-227
# ---------------------------------------------------------------------------------------------------------------
def window_attention_core(qkv, rpb_table, num_heads, ws, mask=None, scale=None):
    """qkv (B_, N, 3C) -> (B_, N, C): q*scale, q k^T + bias + mask, softmax, @ v (:194-220)."""
    B_, N, C3 = qkv.shape
    C = C3 // 3
    hd = C // num_heads
    scale = hd ** -0.5 if scale is None else scale
    qkv = qkv.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    idx = relative_position_index(ws).to(rpb_table.device)
    bias = rpb_table[idx.reshape(-1)].reshape(N, N, num_heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.reshape(B_ // nW, nW, num_heads, N, N) + mask.to(attn.dtype)[None, :, None]
        attn = attn.reshape(-1, num_heads, N, N)
    attn = attn.softmax(dim=-1)
    return (attn @ v).transpose(1, 2).reshape(B_, N, C)


def window_attention(p, prefix, xw, num_heads, ws, mask, cfg, stage, tasks_on):
    """Full module: qkv MTLoRALinear (tasks=None, :164-171) -> core -> proj MTLoRALinear (:173-180)."""
    sc, dp = cfg.shared_scale[stage], cfg.dropout[stage]
    qkv, _ = mtlora_linear(p, prefix + "qkv.", xw, scale_shared=sc, dropout=dp, training=cfg.training)
    a = window_attention_core(qkv, p[prefix + "relative_position_bias_table"], num_heads, ws, mask)
    ts = {t: cfg.stage_task_scale(stage, t) for t in cfg.tasks}
    return mtlora_linear(p, prefix + "proj.", a, None, list(cfg.tasks) if tasks_on else None, sc, ts, dp,
                         cfg.training, cfg.shared_mode)


# ---------------------------------------------------------------------------------------------------------------
# Mlp / SwinTransformerBlock / PatchMerging / BasicLayer / PatchEmbed / backbone
# ---------------------------------------------------------------------------------------------------------------
def mlp(p, prefix, x, x_tasks, cfg, stage, tasks_on):
    """Mlp.forward (:68-81): fc1 -> exact GELU on every stream -> fc2 (task inputs = task hidden streams)."""
    sc, dp = cfg.shared_scale[stage], cfg.dropout[stage]
    ts = {t: cfg.stage_task_scale(stage, t) for t in cfg.tasks}
    tl = list(cfg.tasks) if tasks_on else None
    h, h_tasks = mtlora_linear(p, prefix + "fc1.", x, x_tasks, tl, sc, ts, dp, cfg.training, cfg.shared_mode)
    h = F.gelu(h)
    if h_tasks is not None:
        h_tasks = {t: F.gelu(v) for t, v in h_tasks.items()}
    return mtlora_linear(p, prefix + "fc2.", h, h_tasks, tl, sc, ts, dp, cfg.training, cfg.shared_mode)


def swin_block(p, prefix, x, H, W, num_heads, ws, shift, cfg, stage, tasks_on, drop_path_p=0.0):
    """SwinTransformerBlock.forward (:326-408). x (B, H*W, C) -> (x_out, {task: x_t} | None)."""
    B, L, C = x.shape
    if min(H, W) <= ws:                                                        # :279-282
        shift, ws = 0, min(H, W)
    shortcut = x
    h = F.layer_norm(x, (C,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], cfg.ln_eps)
    xw = roll_window_partition(h.reshape(B, H, W, C), shift, ws).reshape(-1, ws * ws, C)
    mask = shift_attn_mask(H, W, ws, shift).to(x.device) if shift > 0 else None
    aw, aw_tasks = window_attention(p, prefix + "attn.", xw, num_heads, ws, mask, cfg, stage, tasks_on)
    unwin = lambda t: window_merge_roll(t.reshape(-1, ws, ws, C), shift, ws, H, W).reshape(B, L, C)
    x_tasks = None
    if aw_tasks is not None:                                                   # :378-390
        x_tasks = {t: shortcut + drop_path(unwin(v), drop_path_p, cfg.training) for t, v in aw_tasks.items()}
    x = shortcut + drop_path(unwin(aw), drop_path_p, cfg.training)            # :391-392
    ln2 = lambda t: F.layer_norm(t, (C,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], cfg.ln_eps)
    m, m_tasks = mlp(p, prefix + "mlp.", ln2(x), None if x_tasks is None else {t: ln2(v) for t, v in x_tasks.items()},
                     cfg, stage, tasks_on)
    out = x + drop_path(m, drop_path_p, cfg.training)                          # :398-399
    if m_tasks is None:
        return out, None
    if x_tasks is None:                                                        # :401-403
        return out, {t: drop_path(v, drop_path_p, cfg.training) for t, v in m_tasks.items()}
    return out, {t: x_tasks[t] + drop_path(m_tasks[t], drop_path_p, cfg.training) for t in m_tasks}   # :405-408


def patch_merging(p, prefix, x, H, W, cfg, stage):
    """PatchMerging.forward (:451-472): 2x2 gather in channel order (0,0),(1,0),(0,1),(1,1) -> LN(4C) -> reduction."""
    B, L, C = x.shape
    assert L == H * W, "input feature has wrong size"
    assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
    g = x.reshape(B, H, W, C)
    g = torch.cat([g[:, 0::2, 0::2], g[:, 1::2, 0::2], g[:, 0::2, 1::2], g[:, 1::2, 1::2]], -1).reshape(B, -1, 4 * C)
    g = F.layer_norm(g, (4 * C,), p[prefix + "norm.weight"], p[prefix + "norm.bias"], cfg.ln_eps)
    y, _ = mtlora_linear(p, prefix + "reduction.", g, scale_shared=cfg.shared_scale[stage],
                         dropout=cfg.dropout[stage], training=cfg.training)
    return y


def basic_layer(p, prefix, x, H, W, depth, num_heads, cfg, stage, dpr, has_downsample):
    """BasicLayer.forward (:543-551): only the last block (lora=True, :530) produces task streams."""
    tasks_lora = None
    for i in range(depth):
        x, tasks_lora = swin_block(p, f"{prefix}blocks.{i}.", x, H, W, num_heads, cfg.window_size,
                                   0 if i % 2 == 0 else cfg.window_size // 2, cfg, stage,
                                   i == depth - 1 or cfg.intermediate_specialization, dpr[i])
    if has_downsample:
        x = patch_merging(p, prefix + "downsample.", x, H, W, cfg, stage)
        if tasks_lora is not None:
            tasks_lora = {t: patch_merging(p, prefix + "downsample.", v, H, W, cfg, stage) for t, v in tasks_lora.items()}
    return x, tasks_lora


def patch_embed(p, prefix, img, cfg):
    """PatchEmbed.forward (:597-605): Conv2d(k=s=patch) -> flatten -> LayerNorm."""
    x = F.conv2d(img, p[prefix + "proj.weight"], p[prefix + "proj.bias"], stride=cfg.patch_size)
    x = x.flatten(2).transpose(1, 2)
    if (prefix + "norm.weight") in p:
        x = F.layer_norm(x, (x.shape[-1],), p[prefix + "norm.weight"], p[prefix + "norm.bias"], cfg.ln_eps)
    return x


def backbone(p, img, cfg, prefix=""):
    """SwinTransformerMTLoRA.forward(x, return_stages=True) (:734-761) -> [(x_s, {task: x_{s,t}})] per stage."""
    x = patch_embed(p, prefix + "patch_embed.", img, cfg)
    res = cfg.img_size // cfg.patch_size
    n_blocks = sum(cfg.depths)
    dpr = [cfg.drop_path_rate * i / max(n_blocks - 1, 1) for i in range(n_blocks)]   # torch.linspace(0, rate, n) :683
    out = []
    for s, depth in enumerate(cfg.depths):
        H = W = res // (2 ** s)
        lo = sum(cfg.depths[:s])
        x, tl = basic_layer(p, f"{prefix}layers.{s}.", x, H, W, depth, cfg.num_heads[s], cfg, s, dpr[lo:lo + depth],
                            s < len(cfg.depths) - 1)
        if tl is None:
            tl = {t: x for t in cfg.tasks}                                     # :744-745
        out.append((x, tl))
    return out


def backbone_loss(stages):
    """Scalar used by the parity tests: sum over stages and tasks of mean(x_{s,t}^2)."""
    return sum(v.float().pow(2).mean() for _, tl in stages for v in tl.values())
