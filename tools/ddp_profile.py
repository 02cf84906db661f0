"""torchrun tool: GPU timeline summary of a few training steps of the bench workload on every rank's GPU (rank 0 prints):
busy time of the compute stream, idle gaps > 20 us and what follows them, time of the non-library kernels (NCCL, torch).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_profile.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sys.argv = sys.argv[:1]
    a = bench.parse()
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable
    from mtlora_b200.optim import FlatAdamW
    net = bench.build_backbone(a, S)
    bench.mark_trainable(mark_only_lora_as_trainable, net)
    net.to(dev).train()
    opt = FlatAdamW([p for p in net.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.05)
    step = bench.make_step(a, net, None, opt, "backbone", "bf16")
    torch.manual_seed(1234 + rank)
    img = torch.randn(a.batch, 3, a.img, a.img, device=dev)
    for _ in range(8):
        step(img, None).item()
    torch.cuda.synchronize()
    n = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            step(img, None).item()
        torch.cuda.synchronize()
    if rank != 0:
        return
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    by = collections.defaultdict(lambda: [0, 0.0])
    for e in evs:
        k = e.name[:60]
        by[k][0] += 1
        by[k][1] += e.time_range.end - e.time_range.start
    print(f"world {world}: span {(t1 - t0) / n / 1e3:.3f} ms/step over {n} steps, {len(evs) / n:.0f} device events/step")
    lib = sum(v[1] for k, v in by.items() if "mtl::" in k)
    print(f"library kernels {lib / n / 1e3:.3f} ms/step")
    for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        if "mtl::" not in k and t / n > 5:
            print(f"  {t / n:9.1f} us/step {c / n:6.1f} x  {k}")
    # idle gaps of the union of all device activity
    cur_end, gaps = evs[0].time_range.end, []
    for e in evs[1:]:
        if e.time_range.start > cur_end:
            gaps.append((e.time_range.start - cur_end, e.name[:50]))
        cur_end = max(cur_end, e.time_range.end)
    tot = sum(g for g, _ in gaps)
    print(f"idle {tot / n / 1e3:.3f} ms/step in {len(gaps) / n:.0f} gaps/step; gaps > 20 us:")
    for g, name in sorted(gaps, reverse=True)[:25]:
        print(f"  {g:8.1f} us before {name}")


if __name__ == "__main__":
    main()
