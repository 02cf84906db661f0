"""Micro-benchmark of mtl_linear_bwd_params (adapter-gradient reductions) through the C ABI on one layer shape.
    python tools/bench_xty.py --case fc1_last [--iters 20]     # CUDA-event timing, algorithmic GB/s"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import linear_alg_bytes  # noqa: E402
from mtlora_b200 import ops  # noqa: E402

BF = torch.bfloat16
# name: (M, K, N, r_s, r_t list, x_tasks_given, x_gelu, path_scale)
CASES = {
    "fc1_last": (401408, 96, 384, 64, [4] * 4, True, False, False),
    "fc2_last": (401408, 384, 96, 64, [4] * 4, True, False, False),
    "proj_last": (401408, 96, 96, 64, [4] * 4, False, False, False),
    "qkv": (401408, 96, 288, 64, [], False, False, False),
    "fc2": (401408, 384, 96, 64, [], False, False, True),
    "s2_fc1": (25088, 384, 1536, 64, [], False, False, False),
    "s2_fc2": (25088, 1536, 384, 64, [], False, False, True),
    "s3_fc1_last": (6272, 768, 3072, 64, [4] * 4, True, False, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="fc1_last")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    M, K, N, r_s, r_t, xt, x_gelu, use_ps = CASES[a.case]
    T = len(r_t)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    spec = ops.LinearSpec(K, N, r_s, r_t, 4.0, [4.0] * T)
    p = 0.0 if x_gelu else 0.05
    S_in = 1 + (T if xt else 0) + (1 if p > 0 else 0)
    x = torch.randn(S_in, M, K, device=dev, generator=g).to(BF)
    dy = torch.randn(spec.S_out, M, N, device=dev, generator=g).to(BF)
    u = torch.randn(M, spec.R_pad, device=dev, generator=g).to(BF)
    gs = torch.randn(M, spec.R_pad, device=dev, generator=g).to(BF)
    ps = torch.ones(1, 32, device=dev) if use_ps else None
    fn = lambda: ops.linear_bwd_params(spec, x, dy, u, gs, x_tasks_given=xt, x_gelu=x_gelu, path_scale=ps,
                                       rows_per_sample=M // 32 if use_ps else 0, dropout_p=p)
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
    meta = ("bwd_params", M, K, N, 1 + (T if xt else 0), spec.S_out, spec.R_pad, sum(spec.ranks), False)
    b = linear_alg_bytes(meta)
    print(json.dumps({"case": a.case, "ms": ms, "alg_bytes": b, "alg_GBps": b / ms / 1e6}))


if __name__ == "__main__":
    main()
