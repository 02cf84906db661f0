#!/usr/bin/env python
"""Benchmark of the MTLoRA Swin-backbone hot path (BASELINE.json: "images/sec Swin-T 448 4-task r=64").

    python bench.py --gpus N --steps K --warmup W                  # this repo's sm_100a path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...        # the reference's own modules on the host CPU cores
    python bench.py --impl reference-gpu --steps K ...             # the reference's own modules, PyTorch eager, same GPU

One step (`--scope backbone`, the headline `value`) = one training pass of the backbone over one synthetic batch:
forward of SwinTransformerMTLoRA under autocast (every stage through libmtlora_b200.so), the backbone loss of
SURVEY.md §8d (sum over stages and tasks of mean(x^2)), backward (adapter / LayerNorm / rel-pos-bias / reduction /
patch_embed gradients), the data-parallel all-reduce of the trainable gradients (N > 1) and an AdamW step over them.
`--scope full` = the reference's whole train step (main.py:341-353): the reference's own `MultiTaskSwin` (hrnet heads,
models/swin_mtl.py:138-246, unmodified from baseline/_ref) around the backbone, `MultiTaskLoss` with the weights of
main.py:192-199, backward, clip_grad_norm_(5.0), AdamW; with `--amp fp16` through a GradScaler exactly like
utils.py:348-369. Our line reports the full step next to the backbone step (`full_step`) and, at N = 1, the unmodified
reference in PyTorch eager on the same GPU for both scopes (`reference_gpu`).

Workload at N = 1: BASELINE.json configs[1] — Swin-T, 448x448, tasks semseg/normals/sal/human_parts, r_shared = 64,
r_task = 4 (configs/mtlora/tiny_448/mtlora_tiny_448_r64_scale4_pertask.yaml), LoRA dropout 0.05, DropPath 0.2,
batch 32 per GPU (README.md:28). Weak scaling: every rank processes its own batch.

Prints ONE JSON line (rank 0). `value` = images/s with the batch resident in HBM; `e2e` = images/s through the public
module API with the batch in pinned host memory: every step's 77 MB batch is copied H2D inside the timed region (on a
copy stream, one step ahead, like a prefetching loader) and the loss is read back (D2H) every step.
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = {
    "swin_t": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]),
    "swin_s": dict(embed_dim=96, depths=[2, 2, 18, 2], num_heads=[3, 6, 12, 24]),
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32]),
}
TASKS6 = ["semseg", "normals", "sal", "human_parts", "depth", "edge"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--scope", default="backbone", choices=["backbone", "full"],
                    help="what `value` times: the backbone train step (hot path) or the reference's whole train step")
    ap.add_argument("--amp", default="bf16", choices=["bf16", "fp16"],
                    help="autocast dtype; fp16 adds the reference's GradScaler (main.py:341, utils.py:352)")
    ap.add_argument("--model", default="swin_t", choices=list(MODELS))
    ap.add_argument("--img", type=int, default=448)
    ap.add_argument("--tasks", type=int, default=4)
    ap.add_argument("--r-shared", type=int, default=64)
    ap.add_argument("--r-task", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--dropout", type=float, default=0.05)
    ap.add_argument("--drop-path", type=float, default=0.2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gc-freeze", action="store_true",
                    help="do not gc.freeze() the warmed-up process before the timed loops (all arms freeze by default: a "
                         "full collection over the ~1e6 live objects of torch + the model pauses the host for ~50 ms, "
                         "longer than the launch queue covers)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the full_step and reference_gpu blocks of our line (profiling / sweeps)")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--torch-optimizer", action="store_true",
                    help="use torch.optim.AdamW(fused=True) instead of the library's flat multi-tensor step")
    ap.add_argument("--profile-ops", default="", help="write the per-C-ABI-call CUDA-event profile to this JSON file")
    ap.add_argument("--ncu-range", action="store_true",
                    help="profiling aid: after the warm-up run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`); prints no bench line")
    return ap.parse_args()


def mtlora_ns(n_stages, tasks, r_shared, r_task, dropout):
    ranks = [dict({"shared": r_shared}, **{t: r_task for t in tasks}) for _ in range(n_stages)]
    return types.SimpleNamespace(
        ENABLED=True, R_PER_TASK_LIST=ranks, SHARED_SCALE=[4.0] * n_stages,
        SCALE_PER_TASK_LIST=[{t: 4.0 for t in tasks} for _ in range(n_stages)], DROPOUT=[dropout] * n_stages,
        TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False, SHARED_MODE="matrix",
        INTERMEDIATE_SPECIALIZATION=False, QKV_ENABLED=True, PROJ_ENABLED=True, FC1_ENABLED=True, FC2_ENABLED=True,
        DOWNSAMPLER_ENABLED=False)


def workload_name(a, scope=None):
    scope = scope or a.scope
    tail = ("train fwd+bwd+allreduce+adamw" if scope == "backbone" else
            "full train step: MultiTaskSwin hrnet heads + MultiTaskLoss + clip5 + adamw")
    return (f"{a.model} img{a.img} tasks{a.tasks} r_shared{a.r_shared} r_task{a.r_task} batch{a.batch}/gpu "
            f"lora_dropout{a.dropout} drop_path{a.drop_path} {tail}")


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (profiling recipe's clocks line)
# ----------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        if os.environ.get("MTL_BENCH_NO_CLOCKS"):      # diagnosis only: a run without the sampler has no clocks evidence
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# algorithmic bytes / flops of one fused-linear launch (SURVEY.md §8d / BASELINE.md §3.4), bf16 activations
# ----------------------------------------------------------------------------------------------------------------
def linear_alg_bytes(meta):
    kind, M, K, N, s_a, s_b, R_pad, r_sum, has_bias = meta
    s = 2
    if kind == "fwd":        # s_a = input streams, s_b = output streams
        return s * (M * K * s_a + N * K + r_sum * (K + N) + M * N * s_b) + (4 * N if has_bias else 0)
    if kind == "bwd_input":  # s_a = dx streams, s_b = dy streams
        return s * (M * N * s_b + N * K + r_sum * (K + N) + M * K * s_a)
    # bwd_params: re-reads x and dy plus the saved rank-space activations, writes fp32 dA / dB
    return s * (M * K * s_a + M * N * s_b + 2 * M * R_pad) + 4 * r_sum * (K + N)


def linear_alg_flops(meta):
    """Dense frozen product + the true-rank adapter products of one fwd / bwd_input launch."""
    kind, M, K, N, s_a, s_b, R_pad, r_sum, has_bias = meta
    return 2.0 * M * K * N + 2.0 * M * r_sum * (K + N)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return (float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", 1380.8)),
                "measured (MEASURED_PEAKS.json: hbm_gbs copy kernel; bf16_tflops_sustained for kernels inside a long step)")
    except Exception:
        return 6650.0, 1380.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
# model builders
# ----------------------------------------------------------------------------------------------------------------
def tasks_of(a):
    return TASKS6[:a.tasks]


def build_backbone(a, module):
    """`module` = mtlora_b200.swin_transformer_mtlora or the reference's models.swin_transformer_mtlora: same ctor."""
    import torch
    tasks = tasks_of(a)
    m = MODELS[a.model]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = module.SwinTransformerMTLoRA(img_size=a.img, patch_size=4, in_chans=3, num_classes=0,
                                           embed_dim=m["embed_dim"], depths=m["depths"], num_heads=m["num_heads"],
                                           window_size=7, mlp_ratio=4.0, qkv_bias=True, drop_rate=0.0,
                                           drop_path_rate=a.drop_path, ape=False, patch_norm=True, tasks=tasks,
                                           mtlora=mtlora_ns(4, tasks, a.r_shared, a.r_task, a.dropout))
        gen = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for n, p in net.named_parameters():
                if "lora_shared_B" in n or "lora_tasks_B" in n:
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.02)   # non-zero adapters
    return net


def wrap_full(a, backbone):
    """The reference's own MultiTaskSwin + MultiTaskLoss (unmodified, baseline/_ref) around `backbone`."""
    import torch
    from baseline import refload
    ref = refload.load()
    tasks = tasks_of(a)
    cfg = refload.mtl_config(tasks, a.img, refload.mtlora_node(tasks, a.r_shared, a.r_task, dropout=a.dropout))
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        net = ref.swin_mtl.MultiTaskSwin(backbone, cfg)
    crit = ref.losses.MultiTaskLoss(
        tasks, torch.nn.ModuleDict({t: ref.losses.get_loss(cfg.TASKS_CONFIG, t, cfg) for t in tasks}),
        {t: refload.LOSS_WEIGHTS[t] for t in tasks})
    return net, crit


def mark_trainable(mark_fn, backbone):
    with contextlib.redirect_stdout(io.StringIO()):
        mark_fn(backbone, bias="none", freeze_patch_embed=False, freeze_norm=False, free_relative_bias=False,
                freeze_downsample_reduction=False)


def make_mean_square():
    import torch

    class MeanSquare(torch.autograd.Function):
        """mean(x^2) with fp32 accumulation: one reduction kernel forward, one elementwise kernel backward."""

        @staticmethod
        def forward(ctx, x):
            ctx.save_for_backward(x)
            return torch.linalg.vector_norm(x, dtype=torch.float32).square() / x.numel()

        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            return x * (g * (2.0 / x.numel())).to(x.dtype)
    return MeanSquare


def make_step(a, net, crit, opt, scope, amp, reducer=None):
    """One train step, written after main.py:341-353 + utils.py:352-366 (NativeScalerWithGradNormCount)."""
    import torch
    amp_dtype = torch.bfloat16 if amp == "bf16" else torch.float16
    scaler = torch.amp.GradScaler("cuda") if amp == "fp16" else None
    MeanSquare = make_mean_square()
    params = [p for p in net.parameters() if p.requires_grad]

    phase_log = [] if os.environ.get("MTL_BENCH_PHASE_TRACE") else None   # diagnosis: host time per phase
    make_step.phase_log = phase_log

    def mark(tag):
        if phase_log is not None:
            phase_log.append((tag, time.perf_counter()))

    def step(img, targets):
        mark("enter")
        with torch.autocast("cuda", dtype=amp_dtype):
            if scope == "full":
                loss, _ = crit(net(img), targets)
            else:
                stages = net(img, return_stages=True)
                mark("backbone")
                loss = sum(MeanSquare.apply(v) for _, tl in stages for v in tl.values())
        mark("forward")
        fused_clip = getattr(opt, "fused_clip", False)   # FlatAdamW(max_grad_norm=...): unscale + clip inside step()
        if scaler is not None:
            scaler.scale(loss).backward()
            if reducer is not None:
                reducer.reduce()
            if not fused_clip:
                scaler.unscale_(opt)
                if scope == "full":
                    torch.nn.utils.clip_grad_norm_(params, 5.0)
            scaler.step(opt)
            scaler.update()
        else:
            loss.backward()
            mark("backward")
            if reducer is not None:
                reducer.reduce()
            if scope == "full" and not fused_clip:
                torch.nn.utils.clip_grad_norm_(params, 5.0)
            opt.step()
            mark("opt")
        opt.zero_grad(set_to_none=True)
        mark("zero_grad")
        return loss
    return step


class GcWatch:
    """Counts the Python garbage collections that fall into the timed loops and their longest pause (gc.callbacks)."""

    def __init__(self):
        self.n, self.max_ms, self._t0 = 0, 0.0, 0.0

    def _cb(self, phase, info):
        if phase == "start":
            self._t0 = time.perf_counter()
        else:
            self.n += 1
            self.max_ms = max(self.max_ms, 1e3 * (time.perf_counter() - self._t0))

    def __enter__(self):
        import gc
        gc.callbacks.append(self._cb)
        return self

    def __exit__(self, *exc):
        import gc
        gc.callbacks.remove(self._cb)

    def report(self, frozen):
        return {"collections": self.n, "max_pause_ms": round(self.max_ms, 2), "frozen_after_warmup": bool(frozen)}


def gc_freeze(a):
    """After the warm-up: collect once and move every live object to the permanent generation, so that the collections
    inside the timed loops only walk what the steps themselves allocate."""
    import gc
    if not a.no_gc_freeze:
        gc.collect()
        gc.freeze()
    return not a.no_gc_freeze


def time_steps(step, batches, n_steps, warmup, a=None):
    """CUDA-event time of n_steps back-to-back steps over alternating resident batches -> (ms total, last loss)."""
    import torch
    for i in range(warmup):
        step(*batches[i % 2])
    torch.cuda.synchronize()
    if a is not None:
        gc_freeze(a)      # same treatment for every arm (reference modules on the GPU included)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(n_steps):
        last = step(*batches[i % 2]).item()      # main.py:359-361: the loss is read back every step (see run_ours)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), float(last)


# ----------------------------------------------------------------------------------------------------------------
# the CPU leg: the reference's own modules (baseline/_ref, unmodified) — or, when that install is absent, the oracle
# port of the reference algorithm (oracle/mtlora_oracle.py, pinned to the reference's golden vectors) — in fp32
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_fn(a, batch):
    import torch
    from baseline import refload
    if refload.available():
        ref = refload.load()
        net = build_backbone(a, ref.swin)
        mark_trainable(ref.lora.mark_only_lora_as_trainable, net)
        net.train()
        opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.05)
        img = torch.randn(batch, 3, a.img, a.img, generator=torch.Generator().manual_seed(2))

        def step():
            stages = net(img, return_stages=True)
            loss = sum(v.pow(2).mean() for _, tl in stages for v in tl.values())
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            return float(loss.detach())
        return step, "reference"

    from oracle import detgen
    from oracle import mtlora_oracle as O
    tasks = tasks_of(a)
    m = MODELS[a.model]
    cfg = O.OracleConfig(img_size=a.img, embed_dim=m["embed_dim"], depths=tuple(m["depths"]),
                         num_heads=tuple(m["num_heads"]), tasks=tuple(tasks), dropout=(a.dropout,) * 4,
                         drop_path_rate=a.drop_path, training=True)
    ranks = [dict({"shared": a.r_shared}, **{t: a.r_task for t in tasks}) for _ in range(4)]
    shapes = detgen.backbone_param_shapes(cfg, ranks)
    p = detgen.make_params(shapes)
    train = [k for k in p if ("lora_" in k or "norm" in k or "patch_embed" in k or "downsample.reduction" in k
                              or "relative_position_bias_table" in k)]
    for k in train:
        p[k].requires_grad_()
    opt = torch.optim.AdamW([p[k] for k in train], lr=1e-4)
    img = detgen.uniform("bench.cpu.img", (batch, 3, a.img, a.img), -2.0, 2.0)

    def step():
        loss = O.backbone_loss(O.backbone(p, img, cfg))
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss.detach())
    return step, "port"


def time_cpu(a, batch, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = cpu_step_fn(a, batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts, torch.get_num_threads(), kind


CPU_KIND_NOTE = {
    "reference": "the reference's own modules (baseline/_ref, unmodified models/swin_transformer_mtlora.py + models/lora.py)",
    "port": "oracle port of the reference algorithm (pinned to the reference's golden vectors; baseline/_ref absent)",
}


def run_reference(a):
    """`--impl reference` (tier contract): the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.steps + a.warmup
    batch = a.cpu_batch if n <= 40 else 1
    ts, cores, kind = time_cpu(a, batch, a.steps, a.warmup)
    total = sum(ts)
    v = batch * len(ts) / total
    sample = (f"{len(ts)} timed steps (+{a.warmup} warm-up) of batch {batch} (bounded sample of the batch-{a.batch} "
              f"workload), fp32, CPU, train mode, {CPU_KIND_NOTE[kind]}")
    line = {
        "impl": "reference", "metric": "images/sec", "value": v, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, "backbone"), "arm": f"host CPU, fp32, {cores} threads, batch {batch} per "
                   "step (bounded sample of the workload's batch); one process, rank 0 only",
                   "batch_timed": batch, "device": "cpu", "same_batch_as_workload": batch == a.batch},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# the reference in PyTorch eager on the same GPU (the denominator of the >= 5x target)
# ----------------------------------------------------------------------------------------------------------------
def reference_gpu_numbers(a, dev, steps, warmup, scopes=("backbone", "full"), amps=("bf16", "fp16")):
    """images/s of the UNMODIFIED reference modules (baseline/_ref) in eager mode on `dev`, same config / batch /
    optimizer family (optimizer.py:58-60: optim.AdamW over set_weight_decay's groups), batch resident in HBM."""
    import torch
    from baseline import refload
    if not refload.available():
        return {"unavailable": "baseline/_ref is not installed (python baseline/install_reference.py)"}
    ref = refload.load()
    tasks = tasks_of(a)
    out = {"modules": "baseline/_ref models/swin_transformer_mtlora.py, models/lora.py, models/swin_mtl.py (unmodified), "
                      "fused_window_process=False, PyTorch eager", "batch": a.batch, "steps": steps, "warmup": warmup}
    gen = torch.Generator().manual_seed(2)
    imgs = [torch.randn(a.batch, 3, a.img, a.img, generator=gen).to(dev) for _ in range(2)]
    for scope in scopes:
        targets = [refload.synthetic_targets(tasks, a.batch, a.img, gen, dev) if scope == "full" else None
                   for _ in range(2)]
        for amp in amps:
            bb = build_backbone(a, ref.swin)
            net, crit = wrap_full(a, bb) if scope == "full" else (bb, None)
            mark_trainable(ref.lora.mark_only_lora_as_trainable, bb)
            net.to(dev).train()
            skip_kw = bb.no_weight_decay_keywords()
            groups = ref.optimizer.set_weight_decay(net, bb.no_weight_decay(), skip_kw)
            opt = torch.optim.AdamW(groups, eps=1e-8, betas=(0.9, 0.999), lr=1e-4, weight_decay=0.05)
            step = make_step(a, net, crit, opt, scope, amp)
            torch.manual_seed(1234)
            key = f"{scope}_{amp}"
            try:
                ms, loss = time_steps(step, list(zip(imgs, targets)), steps, warmup, a)
                out[key] = {"images_per_s": a.batch * steps / (ms * 1e-3), "ms_per_step": ms / steps, "loss": loss,
                            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
            except torch.OutOfMemoryError as e:
                out[key] = {"unavailable": f"out of memory: {str(e)[:80]}"}
            del net, crit, opt, step, bb, groups
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
    return out


def run_reference_gpu(a):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    assert torch.cuda.is_available(), "--impl reference-gpu needs a CUDA device"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    clocks = Clocks(dev.index)
    clocks.start()
    r = reference_gpu_numbers(a, dev, a.steps, a.warmup, scopes=(a.scope,), amps=(a.amp,))
    clk = clocks.stop()
    if "unavailable" in r:
        print(json.dumps({"impl": "reference-gpu", "unavailable": r["unavailable"]}), flush=True)
        return
    k = r[f"{a.scope}_{a.amp}"]
    if "unavailable" in k:
        print(json.dumps({"impl": "reference-gpu", "unavailable": k["unavailable"]}), flush=True)
        return
    line = {
        "impl": "reference-gpu", "metric": "images/sec", "value": k["images_per_s"], "unit": "images/s", "n_gpus": 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": k["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": a.amp, "data": "synthetic",
        "config": {"workload": workload_name(a), "arm": r["modules"], "amp": a.amp, "loss": k["loss"],
                   "peak_mem_gb": k["peak_mem_gb"]},
        "gpu_launches": 0, "clocks": clk,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from baseline import refload
    from mtlora_b200 import _native
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable

    tasks = tasks_of(a)
    B = a.batch
    gen2 = torch.Generator().manual_seed(2 + rank)
    host = [torch.randn(B, 3, a.img, a.img, generator=gen2).pin_memory() for _ in range(2)]
    resident = [h.to(dev) for h in host]
    h2d_bytes = host[0].numel() * host[0].element_size()

    def make_optimizer(params, scope):
        if a.torch_optimizer:
            return torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True)
        from mtlora_b200.optim import FlatAdamW
        return FlatAdamW(params, lr=1e-4, weight_decay=0.05, max_grad_norm=5.0 if scope == "full" else None)

    def build(scope):
        bb = build_backbone(a, S)
        net, crit = (bb, None)
        if scope == "full":
            net, crit = wrap_full(a, bb)
        mark_trainable(mark_only_lora_as_trainable, bb)
        net.to(dev).train()
        params = [p for p in net.parameters() if p.requires_grad]
        opt = make_optimizer(params, scope)
        # the backbone synchronises its own trainable gradients from autograd hooks when torch.distributed is
        # initialised (mtlora_b200/dist.py); the decoder heads of the full step are registered with the same reducer
        if world > 1 and scope == "full":
            from mtlora_b200.dist import sync_gradients
            sync_gradients(net)
        return net, crit, opt, params

    scope = a.scope
    if scope == "full" and not refload.available():
        raise SystemExit("--scope full needs the reference's decoder heads: run python baseline/install_reference.py")
    net, crit, opt, trainable = build(scope)
    n_train = sum(p.numel() for p in trainable)
    step_fn = make_step(a, net, crit, opt, scope, a.amp)
    torch.manual_seed(1234 + rank)   # per-rank stochastic masks and data (main.py:570-575: seed + rank)
    targets = [refload.synthetic_targets(tasks, B, a.img, gen2, dev) if scope == "full" else None for _ in range(2)]

    def step(img, i=0):
        return step_fn(img, targets[i % 2])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # end-to-end input pipeline: every step's batch is copied from pinned host memory inside the timed region, on a copy
    # stream and one step ahead (the usual prefetching loader), into one of two device buffers
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [torch.empty_like(resident[0]) for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])   # the buffer's previous consumer (step i - 2) has finished
            staged[i % 2].copy_(host[i % 2], non_blocking=True)
            landed[i % 2].record(copy_stream)

    marks = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    free_run = bool(os.environ.get("MTL_BENCH_FREE_RUN"))   # diagnosis: let the host run ahead in the resident loop

    def timed(n_steps, e2e):
        barrier()
        k0 = _native.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        stamps = [time.perf_counter()]
        if e2e:
            prefetch(0)
        for i in range(n_steps):
            if e2e:
                torch.cuda.current_stream().wait_event(landed[i % 2])
                loss = step(staged[i % 2], i)
                consumed[i % 2].record()
                if i + 1 < n_steps:
                    prefetch(i + 1)                # H2D of the next batch overlaps this step's kernels
                last = loss.item()                 # D2H read of the step's result
                stamps.append(time.perf_counter())
            else:
                last = step(resident[i % 2], i)
                marks[i].record()                  # per-step device time
                # the reference's loop reads the loss back every step (main.py:359-361 `loss_meter.update(loss.item())`);
                # so does this one. It also keeps the host from running steps ahead of the GPU: left alone it fills the
                # driver's launch queue, and on some bench hosts one step of such a loop (always the 4th or 5th after the
                # bracketing synchronize) then lost 60-500 ms on the HOST side with no garbage collection, cudaMalloc or
                # clock event to blame (profiles/r02_host_stalls.txt) — a loop that waits for each step's result never did.
                if not free_run:
                    last = last.item()
                stamps.append(time.perf_counter())
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if not e2e and getattr(make_step, "phase_log", None):
            log = make_step.phase_log
            # host seconds spent in each phase of the slowest step of this loop
            enters = [k for k, (tag, _) in enumerate(log) if tag == "enter"][-n_steps:]
            worst = max(range(len(enters)), key=lambda j: (log[enters[j + 1]][1] if j + 1 < len(enters) else
                                                           stamps[-1]) - log[enters[j]][1])
            k0_, k1_ = enters[worst], (enters[worst + 1] if worst + 1 < len(enters) else len(log))
            seq = log[k0_:k1_]
            sys.stderr.write("slowest resident step %d, host ms per phase: %s\n" % (
                worst, ", ".join("%s %.1f" % (b[0], 1e3 * (b[1] - a_[1])) for a_, b in zip(seq, seq[1:]))))
        if not e2e:
            ev = [e0] + marks[:n_steps]
            timed.resident_gpu = [x.elapsed_time(y) for x, y in zip(ev, ev[1:])]
            timed.resident_host = [1e3 * (b - a_) for a_, b in zip(stamps, stamps[1:])]
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if e2e:
            timed.per_step = [1e3 * (b - a_) for a_, b in zip(stamps, stamps[1:])]
        return t.item(), _native.kernel_launches() - k0, float(last.detach() if hasattr(last, "detach") else last)

    # the clocks sampler (nvidia-smi -lms) is started BEFORE the warm-up: its start-up (NVML initialisation, ~1-2 s of
    # driver queries) must not fall into the timed region it is there to observe
    clocks = Clocks(local)
    if rank == 0 and not a.ncu_range:
        clocks.start()
    # warm-up (also stages the bf16 copies of the frozen weights)
    for i in range(a.warmup):
        step(resident[i % 2], i)
    if a.ncu_range:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    # host-side issue time of one step (queue empty at the start, so the launches never block on the GPU): the step is
    # GPU-bound as long as this stays below ms_per_step
    host_ms = []
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(resident[i % 2], i)
        host_ms.append(1e3 * (time.perf_counter() - t0))
    torch.cuda.synchronize()
    frozen = gc_freeze(a)
    ms0 = torch.cuda.memory_stats(dev)
    with GcWatch() as gcw:
        ms_dev, launches, loss_dev = timed(a.steps, e2e=False)
        ms1 = torch.cuda.memory_stats(dev)
        timed(min(a.warmup, 3), e2e=True)              # warm the end-to-end loop (copy stream, staging buffers)
        ms_e2e, _, loss_e2e = timed(a.steps, e2e=True)
    e2e_steps = sorted(getattr(timed, "per_step", []) or [0.0])
    res_gpu = getattr(timed, "resident_gpu", None) or [0.0]
    res_host = getattr(timed, "resident_host", None) or [0.0]
    clk = clocks.stop() if rank == 0 else None

    # per-call CUDA-event profile of 2 more steps: time share per C-ABI entry point + roofline of the fused linear
    _native.profile = []
    for i in range(2):
        step(resident[i % 2], i)
    torch.cuda.synchronize()
    prof, _native.profile = _native.profile, None
    by = {}
    lin_bytes = lin_ms = 0.0
    lin_n = 0
    heavy_flops = heavy_ms = 0.0
    for name, meta, s0, s1 in prof:
        ms = s0.elapsed_time(s1)
        d = by.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ms
        if name in ("mtl_linear_fwd", "mtl_linear_bwd_input") and meta is not None:
            lin_bytes += linear_alg_bytes(meta)
            lin_ms += ms
            lin_n += 1
            if min(meta[2], meta[3]) >= 384:   # stage 2 / 3 layers: the tensor-pipe regime
                heavy_flops += linear_alg_flops(meta)
                heavy_ms += ms
    tot_ms = sum(v[1] for v in by.values())
    peak, peak_tf, peak_src = peaks()
    achieved = lin_bytes / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else 0.0
    traffic, traffic_note = None, "not measured in this run"
    tr_path = os.path.join(ROOT, "profiles", "linear_traffic.json")
    if os.path.exists(tr_path):
        with open(tr_path) as f:
            tj = json.load(f)
        traffic = tj.get("dram_bytes_per_launch")
        traffic_note = ("from the committed ncu capture " + str(tj.get("source", "profiles/linear_traffic.json")) +
                        " (ncu cannot run inside a bench run; re-captured per round by tools/evidence.sh)")
    roof = {"kernel": "mtl_linear_kernel (mtl_linear_fwd + mtl_linear_bwd_input launches of one step)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "launches_per_step": lin_n // 2,
            "alg_bytes_per_launch": lin_bytes / max(lin_n, 1), "avg_launch_ms": lin_ms / max(lin_n, 1),
            "share_of_step_kernel_time": lin_ms / tot_ms if tot_ms else None}
    tf = heavy_flops / (heavy_ms * 1e-3) / 1e12 if heavy_ms > 0 else 0.0
    roof_tensor = {"kernel": "mtl_linear_kernel, stage-2/3 launches (min(K, N) >= 384)", "bound": "tensor",
                   "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": None,
                   "ms_per_step": heavy_ms / 2}
    breakdown = {k: {"calls_per_step": v[0] // 2, "ms_per_step": v[1] / 2} for k, v in
                 sorted(by.items(), key=lambda kv: -kv[1][1])}
    if a.profile_ops and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(a.profile_ops)), exist_ok=True)
        rows = [{"name": n, "meta": list(mt) if mt else None, "ms": s0.elapsed_time(s1)} for n, mt, s0, s1 in prof]
        with open(a.profile_ops, "w") as f:
            json.dump({"workload": workload_name(a), "steps_profiled": 2, "breakdown": breakdown, "calls": rows}, f)

    # ---- extras: the other scope through the same harness, and the reference in eager mode on this GPU -------------
    full_step = ref_gpu = None
    if not a.no_extras and scope == "backbone" and refload.available():
        del step_fn, opt, net, crit, trainable
        torch.cuda.empty_cache()
        n_x = min(a.steps, 20)
        fnet, fcrit, fopt, fparams = build("full")
        ftargets = [refload.synthetic_targets(tasks, B, a.img, gen2, dev) for _ in range(2)]
        out = {}
        for amp in ("bf16", "fp16"):
            fstep = make_step(a, fnet, fcrit, fopt, "full", amp)
            barrier()
            # 20 untimed steps first: the decoder heads' allocation pattern takes longer than the backbone's to settle in
            # the caching allocator (cudaMalloc synchronises the device)
            ms, loss = time_steps(fstep, list(zip(resident, ftargets)), n_x, 20, a)
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out[amp] = {"images_per_s": B * world * n_x / (t.item() * 1e-3), "ms_per_step": t.item() / n_x, "loss": loss}
        full_step = {"workload": workload_name(a, "full"), "steps": n_x,
                     "trainable_params": sum(p.numel() for p in fparams), **out}
        del fstep, fopt, fnet, fcrit, fparams, ftargets
        torch.cuda.empty_cache()
        if world == 1:
            ref_gpu = reference_gpu_numbers(a, dev, n_x, 5)
            for k in ("backbone_bf16", "backbone_fp16"):
                if k in ref_gpu and "images_per_s" in ref_gpu[k]:
                    ref_gpu[k]["ours_over_reference"] = (B * a.steps / (ms_dev * 1e-3)) / ref_gpu[k]["images_per_s"]
            for amp in ("bf16", "fp16"):
                k = f"full_{amp}"
                if k in ref_gpu and "images_per_s" in ref_gpu[k]:
                    ref_gpu[k]["ours_over_reference"] = full_step[amp]["images_per_s"] / ref_gpu[k]["images_per_s"]

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        ts, cores, kind = time_cpu(a, a.cpu_batch, 3, 1)
        cpu = {"value": a.cpu_batch * len(ts) / sum(ts), "unit": "images/s", "cores": cores, "kind": kind,
               "sample": f"3 timed steps (+1 warm-up) of batch {a.cpu_batch}, same model/config, fp32, "
                         f"{CPU_KIND_NOTE[kind]}"}
    imgs = B * world * a.steps
    line = {
        "metric": "images/sec", "value": imgs / (ms_dev * 1e-3), "unit": "images/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.amp, "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": B * world, "parallelism": f"dp{world}",
                   "trainable_params": n_train, "l2": "per-step working set (GBs of activations) >> 126 MB L2; two "
                   "alternating input batches", "loss_dev": loss_dev, "loss_e2e": loss_e2e,
                   "optimizer": "torch fused AdamW" if a.torch_optimizer else "mtlora_b200 FlatAdamW (one launch)"},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                # host-clock spread of the individual steps (rank 0): a slow host or a contended PCIe link shows here
                "step_ms_min_median_max": [round(e2e_steps[0], 2), round(e2e_steps[len(e2e_steps) // 2], 2),
                                           round(e2e_steps[-1], 2)]},
        "gpu_launches": launches,
        "host_issue_ms_per_step": round(statistics.median(host_ms), 3),
        "clocks": clk,
        "host_gc": gcw.report(frozen),
        # caching-allocator activity inside the resident timed loop: a cudaMalloc / cudaFree there synchronises the device
        "allocator": {"cuda_mallocs_in_loop": ms1.get("num_device_alloc", 0) - ms0.get("num_device_alloc", 0),
                      "cuda_frees_in_loop": ms1.get("num_device_free", 0) - ms0.get("num_device_free", 0),
                      "alloc_retries_in_loop": ms1.get("num_alloc_retries", 0) - ms0.get("num_alloc_retries", 0),
                      "reserved_gb_peak": round(ms1.get("reserved_bytes.all.peak", 0) / 2 ** 30, 1),
                      "allocated_gb_peak": round(ms1.get("allocated_bytes.all.peak", 0) / 2 ** 30, 1)},
        # per-step device time of the resident loop (events between the steps) and the longest host-side issue of a step:
        # a step whose device time stands out while the host time does too was starved by the host, not slow on the GPU
        "resident_steps": {"gpu_ms_min_median_max": [round(min(res_gpu), 2), round(statistics.median(res_gpu), 2),
                                                     round(max(res_gpu), 2)],
                           "host_ms_median_max": [round(statistics.median(res_host), 2), round(max(res_host), 2)],
                           "slowest_step": res_host.index(max(res_host))},
        "roofline": roof,
        "roofline_tensor": roof_tensor,
        "cpu_baseline": cpu,
        "full_step": full_step,
        "reference_gpu": ref_gpu,
        "breakdown_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in list(breakdown.items())[:10]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-gpu":
        run_reference_gpu(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
