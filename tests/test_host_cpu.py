"""CPU tests of the host-side logic: the module / checkpoint surface of the reference-facing classes, the
stream-stacking helpers, the error behaviour without a GPU, and the data-parallel gradient reducer on a
world_size-2 gloo group."""
import contextlib
import io
import os
import socket
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import detgen
from oracle.mtlora_oracle import OracleConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mtlora_ns(ranks, tasks, dropout=0.0, downsampler=False, scale=4.0, **over):
    n = len(ranks)
    d = dict(R_PER_TASK_LIST=ranks, SHARED_SCALE=[scale] * n, SCALE_PER_TASK_LIST=[{t: scale for t in tasks} for _ in range(n)],
             DROPOUT=[dropout] * n, TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False, SHARED_MODE="matrix",
             INTERMEDIATE_SPECIALIZATION=False, QKV_ENABLED=True, PROJ_ENABLED=True, FC1_ENABLED=True, FC2_ENABLED=True,
             DOWNSAMPLER_ENABLED=downsampler)
    d.update(over)
    return types.SimpleNamespace(**d)


def build(tasks, ranks, **kw):
    from mtlora_b200 import swin_transformer_mtlora as S
    with contextlib.redirect_stdout(io.StringIO()):
        return S.SwinTransformerMTLoRA(img_size=224, num_classes=0, tasks=tasks, mtlora=mtlora_ns(ranks, tasks, **kw))


@pytest.mark.parametrize("ds", [False, True])
def test_parameter_surface_matches_reference(ds):
    """Names, order and shapes of named_parameters() == the reference's (validated against the real reference model
    by tests/golden/make_golden.py through detgen.backbone_param_shapes) — the checkpoint / optimizer-state surface."""
    cfg = OracleConfig(img_size=224, tasks=("semseg",))
    ranks = [{"shared": 4, "semseg": 4}] * 4
    net = build(["semseg"], ranks, downsampler=ds)
    shapes = detgen.backbone_param_shapes(cfg, ranks, downsampler_lora=ds)
    got = {n: tuple(p.shape) for n, p in net.named_parameters()}
    assert list(got.keys()) == list(shapes.keys())
    assert got == dict(shapes)
    sd = net.state_dict()
    assert "layers.0.blocks.1.attn_mask" in sd and "layers.0.blocks.0.attn_mask" not in sd
    assert sd["layers.0.blocks.0.attn.relative_position_index"].dtype == torch.int64


def test_task_order_is_sorted_like_parameterdict():
    """nn.ParameterDict sorts the keys of a plain dict, so the checkpoint order of the task adapters is sorted(tasks)
    (SURVEY.md §5); the output dict / stream order follows the module's task list."""
    tasks = ["semseg", "normals", "sal", "human_parts"]
    ranks = [dict({"shared": 8}, **{t: 4 for t in tasks})] * 4
    net = build(tasks, ranks)
    names = [n for n, _ in net.layers[0].blocks[1].attn.proj.named_parameters()]
    assert names[:2] == ["lora_shared_A", "lora_shared_B"]
    assert names[2:4] == ["linear.weight", "linear.bias"]
    assert [n.split(".")[-1] for n in names[4:8]] == sorted(tasks)
    assert net.layers[0].blocks[0].attn.proj.tasks is None and net.layers[0].blocks[1].attn.proj.tasks == tasks


def test_init_matches_reference_rules():
    """_init_weights (reference :715-724): Linear weights trunc-normal(.02) with zero bias, LayerNorm (1, 0); LoRA B = 0
    so the layer starts equal to the frozen linear (lora.py:236-247)."""
    net = build(["semseg"], [{"shared": 4, "semseg": 4}] * 4)
    blk = net.layers[1].blocks[1]
    assert float(blk.attn.qkv.linear.bias.abs().max()) == 0.0
    assert 0.015 < float(blk.attn.qkv.linear.weight.std()) < 0.025
    assert float(blk.attn.qkv.lora_shared_B.abs().max()) == 0.0
    assert float(blk.mlp.fc1.lora_tasks_B["semseg"].abs().max()) == 0.0
    assert float(blk.attn.qkv.lora_shared_A.abs().max()) > 0.0
    assert float(blk.norm1.weight.min()) == 1.0 and float(blk.norm1.bias.abs().max()) == 0.0


def test_attn_mask_and_index_buffers_match_oracle():
    from oracle import mtlora_oracle as O
    net = build(["semseg"], [{"shared": 4, "semseg": 4}] * 4)
    blk = net.layers[0].blocks[1]
    assert torch.equal(blk.attn_mask, O.shift_attn_mask(56, 56, 7, 3))
    assert torch.equal(blk.attn.relative_position_index, O.relative_position_index(7))
    last = net.layers[3].blocks[1]       # 7x7 feature map: window = whole map, no shift (:279-282)
    assert last.shift_size == 0 and last.window_size == 7 and last.attn_mask is None


def test_cpu_inputs_raise_loudly():
    net = build(["semseg"], [{"shared": 4, "semseg": 4}] * 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        net(torch.zeros(1, 3, 224, 224), return_stages=True)
    with pytest.raises(RuntimeError):
        net.layers[0].blocks[0].attn.qkv(torch.zeros(2, 49, 96))


def test_unsupported_modes_raise():
    from mtlora_b200.lora import MTLoRALinear
    # 'addition' (lora.py:275-282): task adapters + the layer's own LayerNorm, no shared adapter; same surface as the
    # reference (checked name by name against its golden parameter list in tests/test_headline_gpu.py)
    ad = MTLoRALinear(8, 8, r={"shared": 4, "a": 4}, tasks=["a"], lora_task_scale={"a": 1.0}, shared_mode="add")
    assert ad.shared_mode == "addition" and not hasattr(ad, "lora_shared_A")
    assert [n for n, _ in ad.named_parameters()] == ["linear.weight", "linear.bias", "lora_tasks_A.a", "lora_tasks_B.a",
                                                      "lora_norm.weight", "lora_norm.bias"]
    assert ad.engine.addition and ad.engine.params() == [ad.linear.weight, ad.linear.bias, ad.lora_tasks_A["a"],
                                                          ad.lora_tasks_B["a"]]
    # merge(): layers without task adapters fold the shared update into W
    mg = MTLoRALinear(8, 8, r=4, lora_shared_scale=2.0)
    torch.nn.init.normal_(mg.lora_shared_B)
    w0 = mg.linear.weight.detach().clone()
    assert mg.can_merge() and mg.merge() and mg.merged and not mg.merge()
    assert torch.allclose(mg.linear.weight, w0 + 2.0 * mg.lora_shared_B @ mg.lora_shared_A)
    assert mg.engine.spec.r_shared == 0 and mg.engine is not mg._engine
    mg.train()
    assert not mg.merged and torch.allclose(mg.linear.weight, w0, atol=1e-6) and mg.engine is mg._engine
    assert not MTLoRALinear(8, 8, r={"shared": 4, "a": 4}, tasks=["a"], lora_task_scale={"a": 1.0}).can_merge()
    ts = MTLoRALinear(8, 8, r={"shared": 4, "a": 4}, tasks=["a"], lora_task_scale=2.5, lora_shared_scale=4.0,
                      trainable_scale_shared=True, trainable_scale_per_task=True)
    # same registration order as the reference (direct parameters first, then linear, the task dicts, the scale dict)
    assert [n for n, _ in ts.named_parameters()] == ["lora_shared_A", "lora_shared_B", "lora_shared_scale", "linear.weight",
                                                      "linear.bias", "lora_tasks_A.a", "lora_tasks_B.a", "lora_task_scale.a"]
    assert float(ts.lora_shared_scale) == 4.0 and float(ts.lora_task_scale["a"]) == 2.5
    assert ts.engine.spec.scale_shared == 1.0 and ts.engine.spec.scale_tasks == [1.0]   # folded into the packed B
    m = MTLoRALinear(8, 8, r={"shared": 4, "a": 4}, tasks=["a"], lora_task_scale={"a": 1.0}, shared_mode="matrixv2")
    assert m.shared_mode == "matrixv2" and m.engine.spec.cfg(1, False).shared_mode == 1   # MTL_MODE_MATRIXV2
    # without tasks the reference falls back to 'matrix' (lora.py:183-186)
    assert MTLoRALinear(8, 8, r=4, shared_mode="matrixv2").shared_mode == "matrix"
    with pytest.raises(AssertionError):
        MTLoRALinear(8, 8, r=4, shared_mode="bogus")


def test_stacked_is_zero_copy_for_adjacent_slices():
    from mtlora_b200.swin_transformer_mtlora import _grad_stack, _stacked
    buf = torch.arange(3 * 4 * 5, dtype=torch.float32).reshape(3, 4, 5)
    parts = [buf[i] for i in range(3)]
    st = _stacked(parts)
    assert st.data_ptr() == buf.data_ptr() and torch.equal(st, buf)
    st2 = _stacked([buf[0], buf[2]])
    assert st2.data_ptr() != buf.data_ptr() and torch.equal(st2, torch.stack([buf[0], buf[2]]))
    g = _grad_stack([None, torch.ones(4, 5, dtype=torch.bfloat16)], (4, 5), torch.device("cpu"))
    assert g.shape == (2, 4, 5) and float(g[0].abs().max()) == 0 and float(g[1].min()) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reducer_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtlora_b200.dist import AdapterGradReducer
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)),
              torch.nn.Parameter(torch.zeros(2, 2), requires_grad=False), torch.nn.Parameter(torch.zeros(7))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5, dtype=torch.float32) * (rank + 1)
    # params[3].grad stays None on every rank (the unused fc2.lora_shared_* of the last stage)
    red = AdapterGradReducer(params)
    assert len(red.params) == 3 and red.numel == 12 + 5 + 7
    red.reduce()
    mean = sum(range(1, world + 1)) / world
    ok = (torch.allclose(params[0].grad, torch.full((3, 4), mean))
          and torch.allclose(params[1].grad, torch.arange(5, dtype=torch.float32) * mean)
          and params[3].grad is None and params[2].grad is None)
    # a second step reuses the bucket
    params[0].grad = torch.full((3, 4), 2.0 * (rank + 1))
    red.reduce()
    ok = ok and torch.allclose(params[0].grad, torch.full((3, 4), 2.0 * mean))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_adapter_grad_reducer_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reducer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_reducer_single_process_is_noop():
    from mtlora_b200.dist import AdapterGradReducer
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    AdapterGradReducer([p]).reduce()
    assert torch.equal(p.grad, torch.ones(3))


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm) on a tiny configuration: one JSON line with the contract keys."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--img", "224", "--tasks",
                        "1", "--r-shared", "4", "--steps", "1", "--warmup", "0", "--cpu-batch", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1


def _gradsync_worker(rank, world, port, out):
    """An UNMODIFIED training loop (forward, loss.backward(), optimizer.step()) on two ranks: gradients arrive averaged
    in p.grad without any call into the reducer (mtlora_b200/dist.py, hook-driven)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtlora_b200.dist import GradSync, sync_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 4), torch.nn.Tanh(),
                              torch.nn.Linear(4, 3))
    unused = torch.nn.Parameter(torch.ones(3))      # never reached by backward on any rank (quirk 8)
    net.register_parameter("unused", unused)
    net[2].bias.requires_grad_(False)
    ref = [p.detach().clone() for p in net.parameters()]
    gs = sync_gradients(net, n_buckets=3)
    assert isinstance(gs, GradSync) and len(gs.buckets) == 3 and sync_gradients(net) is None   # covered only once
    ok = True
    for it in range(3):
        xs = [torch.randn(7, 6, generator=torch.Generator().manual_seed(100 * it + r)) for r in range(world)]
        loss = net(xs[rank]).pow(2).mean()
        net.zero_grad(set_to_none=True)
        loss.backward()                               # <- nothing else: the hooks do the exchange
        # expected: mean over ranks of the per-rank gradients, computed locally on a copy
        exp = None
        for r in range(world):
            import copy
            m2 = copy.deepcopy(net)
            for p in m2.parameters():
                p.grad = None
            m2(xs[r]).pow(2).mean().backward()
            g = [None if p.grad is None else p.grad.clone() for p in m2.parameters()]
            exp = g if exp is None else [None if a is None else a + b for a, b in zip(exp, g)]
        for p, e in zip(net.parameters(), exp):
            if e is None:
                ok = ok and p.grad is None
            else:
                ok = ok and p.grad is not None and torch.allclose(p.grad, e / world, atol=1e-6)
    ok = ok and gs.n_reductions == 9 and unused.grad is None
    del ref
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _attach_worker(rank, world, port, out):
    """SwinTransformerMTLoRA._attach_grad_sync (called at every training forward): covers every trainable parameter once,
    is a no-op afterwards, and re-scans when the set of trainable parameters grows (cached parameter list)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import contextlib
    import io
    import types
    from mtlora_b200 import dist as D
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable
    tasks = ["normals", "semseg"]
    ranks = [{"shared": 8, "normals": 4, "semseg": 4}] * 4
    ns = types.SimpleNamespace(ENABLED=True, R_PER_TASK_LIST=ranks, SHARED_SCALE=[4.0] * 4,
                               SCALE_PER_TASK_LIST=[{t: 4.0 for t in tasks}] * 4, DROPOUT=[0.0] * 4,
                               TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False, SHARED_MODE="matrix",
                               INTERMEDIATE_SPECIALIZATION=False, DOWNSAMPLER_ENABLED=False, QKV_ENABLED=True,
                               PROJ_ENABLED=True, FC1_ENABLED=True, FC2_ENABLED=True, FREEZE_PRETRAINED=True,
                               R=ranks, R_PER_TASK={t: 4 for t in tasks}, SCALE_PER_TASK={t: 4.0 for t in tasks})
    with contextlib.redirect_stdout(io.StringIO()):
        net = S.SwinTransformerMTLoRA(img_size=224, num_classes=0, depths=[2, 2, 2, 2], tasks=tasks, mtlora=ns)
        mark_only_lora_as_trainable(net, bias="none", freeze_patch_embed=True, freeze_norm=True,
                                    free_relative_bias=False, freeze_downsample_reduction=True)
    covered = lambda: sum(getattr(p, D._SYNC_ATTR, None) is not None for p in net.parameters())
    n_train = sum(p.requires_grad for p in net.parameters())
    ok = 0 < n_train < sum(1 for _ in net.parameters()) and covered() == 0
    net._attach_grad_sync()
    ok = ok and covered() == n_train and len(net.__dict__["_grad_syncs"]) == 1
    net._attach_grad_sync()                         # nothing changed: no second reducer, no re-scan
    ok = ok and len(net.__dict__["_grad_syncs"]) == 1
    for p in net.patch_embed.parameters():          # unfreeze more parameters after the first "forward"
        p.requires_grad_(True)
    net._attach_grad_sync()
    n2 = sum(p.requires_grad for p in net.parameters())
    ok = ok and n2 > n_train and covered() == n2 and len(net.__dict__["_grad_syncs"]) == 2
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_backbone_attaches_grad_sync_once_and_rescans_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_attach_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_grad_sync_hooks_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gradsync_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def _gradsync_asym_worker(rank, world, port, out, exact):
    """Asymmetric None gradients: rank 1 never produces a gradient for `b`. exact_presence=True hands rank 1 the
    average; the default detects the disagreement (one step late) and raises instead of letting the replicas diverge."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtlora_b200.dist import GradSync
    a = torch.nn.Parameter(torch.ones(4))
    b = torch.nn.Parameter(torch.ones(3))
    gs = GradSync([a, b], n_buckets=1, exact_presence=exact)
    res = "ok"
    try:
        for it in range(2):
            a.grad = b.grad = None
            loss = (a * (rank + 1)).sum()
            if rank == 0:
                loss = loss + (b * 2.0).sum()
            loss.backward()
            if exact:
                good = torch.allclose(a.grad, torch.full((4,), 1.5)) and b.grad is not None and \
                    torch.allclose(b.grad, torch.full((3,), 1.0))      # (2 + 0) / 2
                res = res if good else "wrong values"
    except RuntimeError as e:
        res = "raised" if "disagree" in str(e) else f"other error: {e}"
    out[rank] = res
    del gs
    dist.destroy_process_group()


@pytest.mark.parametrize("exact", [True, False])
def test_grad_sync_asymmetric_none_gloo_world2(exact):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gradsync_asym_worker, args=(world, _free_port(), out, exact), nprocs=world, join=True)
    assert dict(out) == ({0: "ok", 1: "ok"} if exact else {0: "raised", 1: "raised"})


def _reducer_asym_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtlora_b200.dist import AdapterGradReducer
    a, b = torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(3))
    a.grad = torch.full((4,), float(rank + 1))
    if rank == 0:
        b.grad = torch.full((3,), 6.0)
    AdapterGradReducer([a, b]).reduce()
    out[rank] = bool(torch.allclose(a.grad, torch.full((4,), 1.5)) and b.grad is not None
                     and torch.allclose(b.grad, torch.full((3,), 3.0)))
    dist.destroy_process_group()


def test_adapter_grad_reducer_asymmetric_none_gloo_world2():
    """a parameter that is None on one rank only must still receive the averaged gradient there."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_reducer_asym_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
