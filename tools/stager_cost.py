"""Host cost of SwinTransformerMTLoRA._stage_adapters (one-launch adapter staging) on the bench workload."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    a = bench.parse()
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable
    net = bench.build_backbone(a, S)
    bench.mark_trainable(mark_only_lora_as_trainable, net)
    net.cuda().train()
    ps = [p for n, p in net.named_parameters() if "lora_" in n]
    net._stage_adapters()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        net._stage_adapters()
    t1 = time.perf_counter()
    print(f"current (no launch): {(t1 - t0) / 200 * 1e6:.1f} us per call")
    n = 0
    t = 0.0
    for _ in range(50):
        with torch.no_grad():
            torch._foreach_add_(ps, 0.0)
        t0 = time.perf_counter()
        n += net._stage_adapters()
        t += time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"stale (one launch for {n // 50} layers): {t / 50 * 1e6:.1f} us per call")


if __name__ == "__main__":
    main()
