"""CPU checks of the fused linear kernel's planner (mtl_linear_plan, include/mtlora_b200.h): every MTLoRALinear the
reference's Swin variants and MTLoRA configurations can build (models/swin_transformer_mtlora.py:622-700 — qkv, proj,
fc1, fc2 at four stages; configs/mtlora/**: shared rank 16/32/64, task rank 4, up to 6 tasks; BASELINE.json configs[4]:
rank sweep up to 128) must get a tiling that fits the SM's 512 TMEM columns and 227 KiB of shared memory, for the
forward and the input-gradient launch, with and without LoRA dropout / x_tasks. No GPU is touched: the planner is pure
host code and the probe stops before any CUDA call."""
import ctypes
import itertools

import pytest

from mtlora_b200 import _native as N

MTL_ACT_NONE, MTL_ACT_GELU, MTL_ACT_GELU_GRAD = 0, 1, 2
SMEM_MAX = 227 * 1024


@pytest.fixture(scope="module")
def nat():
    return N.load()


def plan(nat, K, Nf, M, r_shared, r_task, xt=1, pass_=0, act=0, res=0, drop=0.0, v2=0, dy_sum=0, n_sm=148):
    c = N.LinearCfg()
    c.M, c.in_features, c.out_features, c.n_tasks, c.r_shared = M, K, Nf, len(r_task), r_shared
    c.x_tasks_given, c.shared_mode, c.dropout_p, c.dy_has_sum, c.scale_shared = xt, v2, drop, dy_sum, 4.0
    for t, r in enumerate(r_task):
        c.r_task[t], c.scale_task[t] = r, 4.0
    o = N.LinearPlanInfo()
    rc = nat.mtl_linear_plan(ctypes.byref(c), pass_, act, res, n_sm, ctypes.byref(o))
    return rc, o, nat.mtl_last_error().decode()


def layer_shapes(embed, img=448, batch=2):
    """(name, K, N, M, forward activation, residual) of every MTLoRALinear of a Swin backbone."""
    out = []
    for s in range(4):
        C = embed << s
        M = batch * (img // 4 >> s) ** 2
        out += [(f"s{s}.qkv", C, 3 * C, M, MTL_ACT_NONE, 0), (f"s{s}.proj", C, C, M, MTL_ACT_NONE, 1),
                (f"s{s}.fc1", C, 4 * C, M, MTL_ACT_GELU_GRAD, 0), (f"s{s}.fc2", 4 * C, C, M, MTL_ACT_NONE, 1)]
        if s < 3:
            out.append((f"s{s}.reduction", 4 * C, 2 * C, M // 4, MTL_ACT_NONE, 0))
    return out


def check_info(o, K, Nf, M, what):
    assert 0 < o.smem_bytes <= SMEM_MAX, what
    assert o.tmem_cols_used <= o.tmem_cols <= 512 and o.tmem_cols & (o.tmem_cols - 1) == 0, what
    assert o.bn in (64, 128, 192) and o.n_chunks == -(-Nf // o.bn), what
    assert 1 <= o.n_splits <= o.n_chunks and o.n_work == -(-M // 128) * o.n_splits, what
    assert o.n_slabs in (2, 3) and o.n_pbuf in (0, 1, 2) and o.n_dbuf in (1, 2), what
    assert 2 <= o.n_stages <= 8 and 1 <= o.up_pack <= 3, what
    if o.d_shared:
        assert o.s_out == 1 and o.n_pbuf == 2 and o.n_dbuf == 1 and o.bn == 128, what


RANK_CONFIGS = [   # (shared rank, task ranks)
    (16, [4] * 4), (32, [4] * 4), (64, [4] * 4),           # configs/mtlora/tiny_448/*.yaml on PASCAL (4 tasks)
    (64, [4] * 6), (32, [4] * 6),                          # BASELINE.json configs[3]: 6 tasks
    (4, [4]), (8, [4] * 2), (64, []),                      # one / two tasks, shared adapter only
    (128, [4] * 4),                                        # rank sweep, r = 128 -> packed rank space 192
    (64, [64] * 4), (32, [32] * 6), (128, [16] * 4),       # equal ranks: rank space 320 / 224 / 192
]


@pytest.mark.parametrize("embed", [96, 128, 192])          # Swin-T/S, Swin-B, Swin-L
def test_every_backbone_layer_gets_a_plan(nat, embed):
    n = 0
    for (name, K, Nf, M, act, res), (rs, rt) in itertools.product(layer_shapes(embed), RANK_CONFIGS):
        for xt, drop in itertools.product((0, 1), (0.0, 0.05)):
            if not rt and xt:
                continue
            what = f"{name} embed={embed} r={rs}/{rt} xt={xt} drop={drop}"
            # forward: operand streams = x, the T task inputs when given, and D(x) when LoRA dropout is on
            rc, o, err = plan(nat, K, Nf, M, rs, rt, xt, 0, act, res, drop)
            assert rc == 0, f"{what} fwd: {err}"
            check_info(o, K, Nf, M, what + " fwd")
            assert o.s_out == 1 + len(rt) and o.s_in == 1 + (len(rt) if xt else 0) + (1 if drop else 0)
            # input gradient (fc2's carries the GELU' factor of fc1); with task streams also the pre-summed variant
            aux = MTL_ACT_GELU if name.endswith("fc2") else MTL_ACT_NONE
            for dy_sum in ((0, 1) if rt else (0,)):
                rc, o, err = plan(nat, K, Nf, M, rs, rt, xt, 1, aux, 0, drop, 0, dy_sum)
                assert rc == 0, f"{what} bwd dy_sum={dy_sum}: {err}"
                check_info(o, Nf, K, M, what + " bwd")
                assert o.s_out == (1 + len(rt) if xt else 1) and o.s_in == 1 + len(rt) + dy_sum
            n += 1
    assert n > 400


def test_wide_rank_space_packs_up_tiles(nat):
    """R_pad <= 128: one Up tile per ring stage (the layout every GPU profile of this round was taken with).
    R_pad > 128: the tiles share stages so that U operand + store slabs + ring still fit 227 KiB."""
    _, o, _ = plan(nat, 96, 288, 401408, 64, [4] * 4)
    assert o.up_pack == 1 and o.n_stages >= 4
    for rs, rt, r_pad in [(128, [4] * 4, 192), (64, [64] * 4, 320)]:
        rc, o, err = plan(nat, 96, 288, 401408, rs, rt)
        assert rc == 0, err
        n_atoms = -(-r_pad // 64)
        assert o.up_pack == 3 and o.bn == 64 and o.n_stages >= -(-n_atoms // 3) + 1
        assert o.smem_bytes <= SMEM_MAX


def test_persistent_grid_balance(nat):
    """Few row tiles (late stages, small batches): the chunks of a tile are split over several work items so that the
    148 persistent CTAs all get work; many row tiles: no split."""
    _, o, _ = plan(nat, 768, 3072, 6272, 64, [4] * 4, act=MTL_ACT_GELU_GRAD)      # 49 row tiles
    assert o.n_splits > 1 and o.n_work >= 98
    _, o, _ = plan(nat, 96, 384, 401408, 64, [4] * 4, act=MTL_ACT_GELU_GRAD)      # 3136 row tiles
    assert o.n_splits == 1
    _, o1, _ = plan(nat, 768, 3072, 6272, 64, [4] * 4, n_sm=49)                   # one tile per CTA: nothing to balance
    assert o1.n_splits == 1


def test_unplannable_configurations_fail_loudly(nat):
    rc, _, err = plan(nat, 96, 288, 1024, 64, [64] * 5)                          # rank space 384 > 320
    assert rc != 0 and "rank space" in err
    rc, _, err = plan(nat, 100, 288, 1024, 16, [4])                              # K not a multiple of 16
    assert rc != 0 and "multiples of 16" in err
    rc, _, err = plan(nat, 96, 288, 0, 16, [4])
    assert rc != 0 and "M=" in err
    c = N.LinearCfg()
    assert nat.mtl_linear_plan(ctypes.byref(c), 2, 0, 0, 148, ctypes.byref(N.LinearPlanInfo())) != 0
    assert "pass" in nat.mtl_last_error().decode()
