"""torch.profiler view of one bench step (our arm): every CUDA kernel, ours and torch's, with its share of the step.
    python tools/profile_step.py [--batch 32] > gpurun_out/step_kernels.txt"""
import contextlib
import io
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mtlora_b200 import swin_transformer_mtlora as S  # noqa: E402
from mtlora_b200.lora import mark_only_lora_as_trainable  # noqa: E402


def main():
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    a = bench.parse()
    tasks = bench.TASKS6[:a.tasks]
    m = bench.MODELS[a.model]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = S.SwinTransformerMTLoRA(img_size=a.img, num_classes=0, embed_dim=m["embed_dim"], depths=m["depths"],
                                      num_heads=m["num_heads"], drop_path_rate=a.drop_path, tasks=tasks,
                                      mtlora=bench.mtlora_ns(4, tasks, a.r_shared, a.r_task, a.dropout))
        mark_only_lora_as_trainable(net)
    net.cuda().train()
    tr = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(tr, lr=1e-4, fused=True)
    img = torch.randn(a.batch, 3, a.img, a.img, device="cuda")

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            st = net(img, return_stages=True)
        loss = sum(v.float().pow(2).mean() for _, tl in st for v in tl.values())
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    if ev:
        t0 = min(e.time_range.start for e in ev)
        t1 = max(e.time_range.end for e in ev)
        busy = sum(e.time_range.end - e.time_range.start for e in ev)
        print(f"GPU span {1e-3 * (t1 - t0) / 2:.2f} ms/step, kernel-busy {1e-3 * busy / 2:.2f} ms/step over {len(ev) // 2} kernels/step")


if __name__ == "__main__":
    main()
