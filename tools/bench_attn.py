"""Micro-benchmark of the window-attention forward through the C ABI (ops layer), BASELINE config 2 shapes.
    python tools/bench_attn.py [--iters 20]      # MTL_ATTN_UMMA=1 selects the tcgen05 kernel"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mtlora_b200 import ops  # noqa: E402

CASES = {"s0": (32, 112, 96, 3), "s1": (32, 56, 192, 6), "s2": (32, 28, 384, 12), "s3": (32, 14, 768, 24)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    for name, (B, H, C, nH) in CASES.items():
        for shift in (0, 3):
            g = torch.Generator(device="cuda").manual_seed(0)
            qkv = torch.randn(B, H, H, 3 * C, device="cuda", generator=g).to(torch.bfloat16)
            rpb = torch.randn(169, nH, device="cuda", generator=g) * 0.1
            fn = lambda: ops.window_attention_fwd(qkv, rpb, nH, 7, shift, 32 ** -0.5, dropout_p=0.05, seed=3)
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
            by = B * H * H * C * 2 * (3 + 2)     # qkv read, out + dropped copy written
            print(json.dumps({"case": name, "shift": shift, "umma": os.environ.get("MTL_ATTN_UMMA") == "1", "ms": round(ms, 4),
                              "alg_GBps": round(by / ms / 1e6, 1)}))


if __name__ == "__main__":
    main()
