"""Micro-benchmark of the fused MTLoRALinear kernel through the C ABI (ops layer) on one shape.

    python tools/bench_linear.py --case fc1_fwd [--iters 20]       # CUDA-event timing, algorithmic GB/s
Used under ncu for the per-kernel captures committed in profiles/ (never quote a time measured under ncu)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import linear_alg_bytes  # noqa: E402
from mtlora_b200 import ops  # noqa: E402

BF = torch.bfloat16
# name: (M, K, N, r_s, r_t list, kind)   — BASELINE config 2 (Swin-T 448, batch 32), stage-0 / stage-2 layers
CASES = {
    "fc1_fwd": (401408, 96, 384, 64, [4] * 4, "fc1"),
    "fc2_fwd": (401408, 384, 96, 64, [4] * 4, "fc2"),
    "proj_fwd": (401408, 96, 96, 64, [4] * 4, "proj"),
    "qkv_fwd": (401408, 96, 288, 64, [], "qkv"),
    "fc2_bwd": (401408, 384, 96, 64, [4] * 4, "fc2_bwd"),
    "fc1_bwd": (401408, 96, 384, 64, [4] * 4, "fc1_bwd"),
    "proj_bwd": (401408, 96, 96, 64, [4] * 4, "proj_bwd"),
    "qkv_bwd": (401408, 96, 288, 64, [], "qkv_bwd"),
    "s1_fc2_bwd": (100352, 768, 192, 64, [4] * 4, "fc2_bwd"),
    "s1_fc1_fwd": (100352, 192, 768, 64, [4] * 4, "fc1"),
    "s0_fc2_bwd": (401408, 384, 96, 64, [], "fc2_bwd_single"),
    "s0_fc1_fwd": (401408, 96, 384, 64, [], "fc1_single"),
    "s0_fc2_fwd": (401408, 384, 96, 64, [], "fc2_single"),
    "s0_fc1_bwd": (401408, 96, 384, 64, [], "fc1_bwd_single"),
    "s0_proj_bwd": (401408, 96, 96, 64, [], "proj_bwd_single"),
    "s2_qkv_fwd": (25088, 384, 1152, 64, [], "qkv"),
    "s2_fc2_fwd": (25088, 1536, 384, 64, [], "fc2_single"),
    "s2_qkv_bwd": (25088, 384, 1152, 64, [], "qkv_bwd"),
    "s2_fc1_fwd": (25088, 384, 1536, 64, [], "fc1_single"),
    "s2_fc2_bwd": (25088, 1536, 384, 64, [], "fc2_bwd_single"),
    "s3_fc2_bwd": (6272, 3072, 768, 64, [4] * 4, "fc2_bwd"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="fc1_fwd")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--dropout", type=float, default=0.05)
    a = ap.parse_args()
    M, K, N, r_s, r_t, kind = CASES[a.case]
    T = len(r_t)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    spec = ops.LinearSpec(K, N, r_s, r_t, 4.0, [4.0] * T)
    W = torch.randn(N, K, device=dev, generator=g) * 0.03
    bias = torch.randn(N, device=dev, generator=g) * 0.02
    As, Bs = torch.randn(r_s, K, device=dev, generator=g) * 0.05, torch.randn(N, r_s, device=dev, generator=g) * 0.02
    At = [torch.randn(r, K, device=dev, generator=g) * 0.05 for r in r_t]
    Bt = [torch.randn(N, r, device=dev, generator=g) * 0.02 for r in r_t]
    wb, wt = ops.cast_transpose(W)
    a_cat, b_cat, a_cat_t, b_cat_t = ops.pack_adapters(spec, As, Bs, At, Bt)
    p = a.dropout
    xt = kind in ("fc1", "fc2", "fc2_bwd", "fc1_bwd") and T > 0
    S_in = 1 + (T if xt else 0) + (1 if p > 0 else 0)
    x = (torch.randn(S_in, M, K, device=dev, generator=g)).to(BF)
    S_out = spec.S_out

    if kind in ("fc1", "fc1_single"):
        fn = lambda: ops.linear_fwd(spec, x, wb, bias, a_cat, b_cat, x_tasks_given=xt, act_gelu=True, gelu_grad=True, dropout_p=p, seed=1, save_u=True)
        meta = ("fwd", M, K, N, 1 + (T if xt else 0), S_out, spec.R_pad, sum(spec.ranks), True)
    elif kind in ("fc2", "proj", "qkv", "fc2_single"):
        res = torch.randn(S_out if kind.startswith("fc2") else 1, M, N, device=dev, generator=g).to(BF) if kind != "qkv" else None
        ps = torch.ones(S_out, 32, device=dev) if kind != "qkv" else None
        fn = lambda: ops.linear_fwd(spec, x, wb, bias, a_cat, b_cat, x_tasks_given=xt, residual=res, path_scale=ps,
                                    rows_per_sample=M // 32 if ps is not None else 0, dropout_p=p, seed=1, save_u=True)
        meta = ("fwd", M, K, N, 1 + (T if xt else 0), S_out, spec.R_pad, sum(spec.ranks), True)
    else:
        dy = torch.randn(S_out, M, N, device=dev, generator=g).to(BF)
        n_dx = 1 + (T if xt else 0)
        aux = torch.randn(n_dx, M, K, device=dev, generator=g).to(BF) if kind.startswith("fc2_bwd") else None
        # what LinearEngine.backward does for fc2-shaped layers with task streams and for DropPath layers (proj)
        presum = S_out > 1 and (K >= 2 * N or kind == "proj_bwd")
        if presum:
            dy = ops.scale_rows_sum(dy, None, 0)
        fn = lambda: ops.linear_bwd_input(spec, dy, wt, a_cat_t, b_cat_t, x_tasks_given=xt, gelu_aux=aux, aux_is_grad=True,
                                          dy_has_sum=presum, dropout_p=p, seed=1, save_g=True)
        meta = ("bwd_input", M, K, N, n_dx, S_out, spec.R_pad, sum(spec.ranks), False)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    b = linear_alg_bytes(meta)
    print(json.dumps({"case": a.case, "meta": meta, "ms": ms, "alg_bytes": b, "alg_GBps": b / ms / 1e6}))


if __name__ == "__main__":
    main()
