"""Which Python lines launch the small PyTorch kernels (copies, fills, casts, random draws) of one training step of the
bench workload: torch.profiler with stacks, grouped by the innermost frames that lie inside this repository."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    a = bench.parse()
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.lora import mark_only_lora_as_trainable
    from mtlora_b200.optim import FlatAdamW
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
    n = 2
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True,
                 experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
        for _ in range(n):
            step(img, None)
        torch.cuda.synchronize()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows = []
    for ev in prof.key_averages(group_by_stack_n=16):
        dt = getattr(ev, "device_time_total", 0)
        if not dt or not ev.key.startswith("aten::"):
            continue
        frames = [f for f in (ev.stack or []) if "mtlora_b200/" in f or "bench.py" in f]
        where = " <- ".join(os.path.basename(f.split("(")[0]) + ":" + f.split("(")[1].split(")")[0] + " " + f.split(": ")[-1]
                            for f in frames[:3] if "(" in f) or "(no repo frame)"
        rows.append((dt / n, ev.count / n, ev.key, where))
    rows.sort(reverse=True)
    tot = 0.0
    for dt, cnt, name, where in rows[:70]:
        print(f"{dt:9.1f} us/step {cnt:6.1f} calls/step  {name:28s} {where}")
        tot += dt
    print(f"total of the listed aten ops: {tot:.1f} us/step")


if __name__ == "__main__":
    main()
