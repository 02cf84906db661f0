"""Does the training step churn the CUDA caching allocator (cudaMalloc / cudaFree per step)? Host cost of torch.empty."""
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
    from mtlora_b200.optim import FlatAdamW
    print("PYTORCH_CUDA_ALLOC_CONF =", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
    dev = torch.device("cuda", 0)
    net = bench.build_backbone(a, S)
    bench.mark_trainable(mark_only_lora_as_trainable, net)
    net.to(dev).train()
    params = [p for p in net.parameters() if p.requires_grad]
    opt = FlatAdamW(params, lr=1e-4, weight_decay=0.05)
    step = bench.make_step(a, net, None, opt, "backbone", "bf16")
    img = torch.randn(a.batch, 3, a.img, a.img, device=dev)
    for _ in range(5):
        step(img, None)
    torch.cuda.synchronize()
    keys = ["num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated", "segment.all.allocated"]
    for it in range(4):
        s0 = torch.cuda.memory_stats()
        t0 = time.perf_counter()
        step(img, None)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        s1 = torch.cuda.memory_stats()
        print(f"step {it}: host issue {1e3 * (t1 - t0):.2f} ms;", {k: s1.get(k, 0) - s0.get(k, 0) for k in keys},
              f"reserved {s1['reserved_bytes.all.current'] / 2**30:.1f} GiB, peak allocated {s1['allocated_bytes.all.peak'] / 2**30:.1f} GiB")
    # raw cost of torch.empty for a few sizes (cache warm)
    for n in (1 << 10, 1 << 20, 1 << 26, 1 << 29):
        xs = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(3)]
        del xs
        t0 = time.perf_counter()
        for _ in range(200):
            x = torch.empty(n, dtype=torch.uint8, device=dev)
            del x
        print(f"torch.empty({n} B): {1e6 * (time.perf_counter() - t0) / 200:.1f} us")


if __name__ == "__main__":
    main()
