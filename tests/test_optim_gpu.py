"""GPU parity of the flat optimizer step (mtlora_b200/optim.py: mtl_opt_sqnorm + mtl_opt_adamw) against the chain the
reference's train step runs (main.py:341-353 -> utils.py:348-369): GradScaler.unscale_ + clip_grad_norm_ +
torch.optim.AdamW (optimizer.py:58-60). fp32 arithmetic on both sides: tolerance 2e-6 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda")


SHAPES = [(64, 96), (96, 64), (4, 96), (96, 4), (96,), (169, 3), (192, 384), (1,), (5000,), (3, 7, 11)]


def make_params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in SHAPES]


def set_grads(ps, it, scale=1.0, none_idx=()):
    g = torch.Generator().manual_seed(100 + it)
    for i, p in enumerate(ps):
        gr = torch.randn(p.shape, generator=g).to(p.device) * (0.5 + i) * scale
        p.grad = None if i in none_idx else gr


def groups_of(ps):
    return [{"params": [p for p in ps if p.ndim > 1], "weight_decay": 0.05},
            {"params": [p for p in ps if p.ndim <= 1], "weight_decay": 0.0}]


@pytest.mark.parametrize("clip", [None, 5.0, 0.01])
def test_flat_adamw_matches_torch(cuda, clip):
    from mtlora_b200.optim import FlatAdamW
    a, b = make_params(cuda), make_params(cuda)
    oa = FlatAdamW(groups_of(a), lr=3e-3, betas=(0.9, 0.95), eps=1e-8, max_grad_norm=clip)
    ob = torch.optim.AdamW(groups_of(b), lr=3e-3, betas=(0.9, 0.95), eps=1e-8)
    for it in range(6):
        none = (3, 7) if it % 2 else (7,)       # parameters without a gradient are skipped, like torch.optim does
        set_grads(a, it, none_idx=none)
        set_grads(b, it, none_idx=none)
        if it == 3:                              # an LR scheduler changes param_group["lr"] between steps
            for o in (oa, ob):
                o.param_groups[0]["lr"] = 1e-3
        if clip is not None:
            norm = torch.nn.utils.clip_grad_norm_(b, clip)
        oa.step()
        ob.step()
        if clip is not None:
            assert torch.allclose(oa.last_grad_norm(), norm, rtol=1e-5)
        for i, (x, y) in enumerate(zip(a, b)):
            assert torch.allclose(x, y, rtol=2e-6, atol=1e-7), (it, i, (x - y).abs().max().item())
    sd = oa.state_dict()
    assert len(sd["state"]) == len(SHAPES)
    # per-parameter step counters (state indices follow the param_groups: 7 matrices first, then the 1-D tensors):
    # SHAPES[7] = (1,) never had a gradient, SHAPES[3] = (96, 4) skipped every other step
    assert float(sd["state"][8]["step"]) == 0.0 and float(sd["state"][3]["step"]) == 3.0
    for k, st in sd["state"].items():
        assert set(st) == {"step", "exp_avg", "exp_avg_sq"}


def test_flat_adamw_grad_scaler_contract(cuda):
    """GradScaler.step() hands `grad_scale` / `found_inf` to an optimizer with _step_supports_amp_scaling (no host
    sync); a step with an inf gradient is skipped entirely and does not advance the bias-correction counter."""
    from mtlora_b200.optim import FlatAdamW
    a, b = make_params(cuda, 1), make_params(cuda, 1)
    oa = FlatAdamW(groups_of(a), lr=1e-2, max_grad_norm=5.0)
    ob = torch.optim.AdamW(groups_of(b), lr=1e-2)
    sa = torch.amp.GradScaler("cuda", init_scale=1024.0)
    sb = torch.amp.GradScaler("cuda", init_scale=1024.0)
    for it in range(5):
        set_grads(a, it, scale=1024.0)
        set_grads(b, it, scale=1024.0)
        if it == 2:
            a[0].grad[0, 0] = float("inf")
            b[0].grad[0, 0] = float("inf")
        # ours: no unscale_ / clip call — both are fused into the step
        sa._lazy_init_scale_growth_tracker(cuda) if sa._scale is None else None
        sa.step(oa)
        sa.update()
        # reference chain (utils.py:352-366)
        sb._lazy_init_scale_growth_tracker(cuda) if sb._scale is None else None
        sb.unscale_(ob)
        torch.nn.utils.clip_grad_norm_(b, 5.0)
        sb.step(ob)
        sb.update()
        assert sa.get_scale() == sb.get_scale()
        for i, (x, y) in enumerate(zip(a, b)):
            assert torch.allclose(x, y, rtol=3e-6, atol=1e-7), (it, i, (x - y).abs().max().item())
    assert float(oa._state[2]) == 4.0            # one of the five steps was skipped


def test_flat_adamw_state_dict_round_trip(cuda):
    from mtlora_b200.optim import FlatAdamW
    a = make_params(cuda, 2)
    oa = FlatAdamW(groups_of(a), lr=1e-2)
    for it in range(3):
        set_grads(a, it)
        oa.step()
    sd = oa.state_dict()
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    ob = FlatAdamW(groups_of(b), lr=1e-2)
    ob.load_state_dict(sd)
    set_grads(a, 9)
    set_grads(b, 9)
    oa.step()
    ob.step()
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_flat_adamw_add_param_group(cuda):
    """add_param_group after the first step (e.g. unfreezing the decoder heads later): old moments survive."""
    from mtlora_b200.optim import FlatAdamW
    a, b = make_params(cuda, 3), make_params(cuda, 3)
    oa = FlatAdamW(a[:6], lr=1e-2, weight_decay=0.0)
    ob = torch.optim.AdamW(b[:6], lr=1e-2, weight_decay=0.0)
    for it in range(2):
        set_grads(a, it)
        set_grads(b, it)
        oa.step()
        ob.step()
    oa.add_param_group({"params": a[6:], "weight_decay": 0.0})
    ob.add_param_group({"params": b[6:], "weight_decay": 0.0})
    for it in range(2, 4):
        set_grads(a, it)
        set_grads(b, it)
        oa.step()
        ob.step()
    for i, (x, y) in enumerate(zip(a, b)):
        assert torch.allclose(x, y, rtol=3e-6, atol=1e-7), (i, (x - y).abs().max().item())


def test_flat_adamw_rejects_cpu():
    from mtlora_b200.optim import FlatAdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        FlatAdamW([p]).step()
