#!/usr/bin/env python
"""Benchmark of the MTLoRA Swin-backbone hot path (BASELINE.json: "images/sec Swin-T 448 4-task r=64").

    python bench.py --gpus N --steps K --warmup W              # this repo's sm_100a path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...    # the reference algorithm on the host CPU cores

One step = one training pass of the backbone over one synthetic batch: forward of SwinTransformerMTLoRA under bf16
autocast (patch_embed in PyTorch, every stage through libmtlora_b200.so), the backbone loss of SURVEY.md §8d
(sum over stages and tasks of mean(x^2)), backward (adapter / LayerNorm / rel-pos-bias / reduction / patch_embed
gradients), the data-parallel all-reduce of the trainable gradients (N > 1) and a fused AdamW step over them.
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="swin_t", choices=list(MODELS))
    ap.add_argument("--img", type=int, default=448)
    ap.add_argument("--tasks", type=int, default=4)
    ap.add_argument("--r-shared", type=int, default=64)
    ap.add_argument("--r-task", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--dropout", type=float, default=0.05)
    ap.add_argument("--drop-path", type=float, default=0.2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--profile-ops", default="", help="write the per-C-ABI-call CUDA-event profile to this JSON file")
    ap.add_argument("--ncu-range", action="store_true",
                    help="profiling aid: after the warm-up run ONE step between cudaProfilerStart/Stop and exit "
                         "(use with `ncu --profile-from-start off`); prints no bench line")
    return ap.parse_args()


def mtlora_ns(n_stages, tasks, r_shared, r_task, dropout):
    ranks = [dict({"shared": r_shared}, **{t: r_task for t in tasks}) for _ in range(n_stages)]
    return types.SimpleNamespace(
        R_PER_TASK_LIST=ranks, SHARED_SCALE=[4.0] * n_stages,
        SCALE_PER_TASK_LIST=[{t: 4.0 for t in tasks} for _ in range(n_stages)], DROPOUT=[dropout] * n_stages,
        TRAINABLE_SCALE_SHARED=False, TRAINABLE_SCALE_PER_TASK=False, SHARED_MODE="matrix",
        INTERMEDIATE_SPECIALIZATION=False, QKV_ENABLED=True, PROJ_ENABLED=True, FC1_ENABLED=True, FC2_ENABLED=True,
        DOWNSAMPLER_ENABLED=False)


def workload_name(a):
    return (f"{a.model} img{a.img} tasks{a.tasks} r_shared{a.r_shared} r_task{a.r_task} batch{a.batch}/gpu "
            f"lora_dropout{a.dropout} drop_path{a.drop_path} train fwd+bwd+allreduce+adamw")


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
# algorithmic bytes of one fused-linear launch (SURVEY.md §8d / BASELINE.md §3.4), bf16 activations
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


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
# the CPU leg: the reference algorithm (oracle/mtlora_oracle.py, pinned to the reference's golden vectors) in fp32
# ----------------------------------------------------------------------------------------------------------------
def cpu_step_fn(a, batch):
    import torch
    from oracle import detgen
    from oracle import mtlora_oracle as O
    tasks = TASKS6[:a.tasks]
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
    return step


def time_cpu(a, batch, steps, warmup):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn(a, batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.steps + a.warmup
    batch = a.cpu_batch if n <= 40 else 1
    ts, cores = time_cpu(a, batch, a.steps, a.warmup)
    total = sum(ts)
    v = batch * len(ts) / total
    sample = f"{len(ts)} timed steps (+{a.warmup} warm-up) of batch {batch} on the same model/config, fp32, train mode"
    line = {
        "impl": "reference", "metric": "images/sec", "value": v, "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "note": "reference algorithm (oracle port, pinned to the reference's "
                   "golden vectors) on the host CPU; one process, all host threads; rank 0 only"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
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

    from mtlora_b200 import _native
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.dist import AdapterGradReducer
    from mtlora_b200.lora import mark_only_lora_as_trainable

    tasks = TASKS6[:a.tasks]
    m = MODELS[a.model]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = S.SwinTransformerMTLoRA(img_size=a.img, patch_size=4, in_chans=3, num_classes=0, embed_dim=m["embed_dim"],
                                      depths=m["depths"], num_heads=m["num_heads"], window_size=7, mlp_ratio=4.0,
                                      qkv_bias=True, drop_rate=0.0, drop_path_rate=a.drop_path, ape=False,
                                      patch_norm=True, tasks=tasks,
                                      mtlora=mtlora_ns(4, tasks, a.r_shared, a.r_task, a.dropout))
        gen = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for n, p in net.named_parameters():
                if "lora_shared_B" in n or "lora_tasks_B" in n:
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.02)   # non-zero adapters
        mark_only_lora_as_trainable(net, bias="none", freeze_patch_embed=False, freeze_norm=False,
                                    free_relative_bias=False, freeze_downsample_reduction=False)
    net.to(dev).train()
    trainable = [p for p in net.parameters() if p.requires_grad]
    n_train = sum(p.numel() for p in trainable)
    opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=0.05, fused=True)
    reducer = AdapterGradReducer(trainable)
    torch.manual_seed(1234 + rank)   # per-rank stochastic masks and data (main.py:570-575: seed + rank)

    B = a.batch
    gen2 = torch.Generator().manual_seed(2 + rank)
    host = [torch.randn(B, 3, a.img, a.img, generator=gen2).pin_memory() for _ in range(2)]
    resident = [h.to(dev) for h in host]
    h2d_bytes = host[0].numel() * host[0].element_size()

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

    def step(img):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            stages = net(img, return_stages=True)
        # backbone loss of SURVEY.md §8d: sum over stages and tasks of mean(x^2)
        loss = sum(MeanSquare.apply(v) for _, tl in stages for v in tl.values())
        loss.backward()
        reducer.reduce()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

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

    def timed(n_steps, e2e):
        barrier()
        k0 = _native.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        if e2e:
            prefetch(0)
        for i in range(n_steps):
            if e2e:
                torch.cuda.current_stream().wait_event(landed[i % 2])
                loss = step(staged[i % 2])
                consumed[i % 2].record()
                if i + 1 < n_steps:
                    prefetch(i + 1)                # H2D of the next batch overlaps this step's kernels
                last = loss.item()                 # D2H read of the step's result
            else:
                last = step(resident[i % 2])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), _native.kernel_launches() - k0, float(last)

    # warm-up (also stages the bf16 copies of the frozen weights)
    for i in range(a.warmup):
        step(resident[i % 2])
    if a.ncu_range:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    ms_dev, launches, loss_dev = timed(a.steps, e2e=False)
    timed(min(a.warmup, 3), e2e=True)              # warm the end-to-end loop (copy stream, staging buffers)
    ms_e2e, _, loss_e2e = timed(a.steps, e2e=True)
    clk = clocks.stop() if rank == 0 else None

    # per-call CUDA-event profile of 2 more steps: time share per C-ABI entry point + roofline of the fused linear
    _native.profile = []
    for i in range(2):
        step(resident[i % 2])
    torch.cuda.synchronize()
    prof, _native.profile = _native.profile, None
    by = {}
    lin_bytes = lin_ms = 0.0
    lin_n = 0
    for name, meta, s0, s1 in prof:
        ms = s0.elapsed_time(s1)
        d = by.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ms
        if name in ("mtl_linear_fwd", "mtl_linear_bwd_input") and meta is not None:
            lin_bytes += linear_alg_bytes(meta)
            lin_ms += ms
            lin_n += 1
    tot_ms = sum(v[1] for v in by.values())
    peak, peak_src = peaks()
    achieved = lin_bytes / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else 0.0
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "linear_traffic.json")
    if os.path.exists(tr_path):
        with open(tr_path) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roof = {"kernel": "mtl_linear_kernel (mtl_linear_fwd + mtl_linear_bwd_input launches of one step)",
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "launches_per_step": lin_n // 2,
            "alg_bytes_per_launch": lin_bytes / max(lin_n, 1), "avg_launch_ms": lin_ms / max(lin_n, 1),
            "share_of_step_kernel_time": lin_ms / tot_ms if tot_ms else None}
    breakdown = {k: {"calls_per_step": v[0] // 2, "ms_per_step": v[1] / 2} for k, v in
                 sorted(by.items(), key=lambda kv: -kv[1][1])}
    if a.profile_ops and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(a.profile_ops)), exist_ok=True)
        rows = [{"name": n, "meta": list(mt) if mt else None, "ms": s0.elapsed_time(s1)} for n, mt, s0, s1 in prof]
        with open(a.profile_ops, "w") as f:
            json.dump({"workload": workload_name(a), "steps_profiled": 2, "breakdown": breakdown, "calls": rows}, f)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        ts, cores = time_cpu(a, a.cpu_batch, 3, 1)
        cpu = {"value": a.cpu_batch * len(ts) / sum(ts), "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"3 timed steps (+1 warm-up) of batch {a.cpu_batch}, same model/config, fp32, oracle port of the "
                         "reference algorithm (pinned to the reference's golden vectors)"}
    imgs = B * world * a.steps
    line = {
        "metric": "images/sec", "value": imgs / (ms_dev * 1e-3), "unit": "images/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": B * world, "parallelism": f"dp{world}",
                   "trainable_params": n_train, "l2": "per-step working set (GBs of activations) >> 126 MB L2; two "
                   "alternating input batches", "loss_dev": loss_dev, "loss_e2e": loss_e2e},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
        "cpu_baseline": cpu,
        "breakdown_ms_per_step": {k: round(v["ms_per_step"], 3) for k, v in list(breakdown.items())[:8]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
