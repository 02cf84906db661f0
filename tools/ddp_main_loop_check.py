"""torchrun check: the reference's own main.train_one_epoch (baseline/_ref/main.py, unmodified) on N ranks with DIFFERENT
data per rank; the backbone's trainable parameters stay bit-identical across ranks because mtlora_b200 averages their
gradients from autograd hooks (mtlora_b200/dist.py) — nothing in the loop calls a reducer. The decoder heads are covered
by the one optional line `sync_gradients(model)`.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_main_loop_check.py
"""
import contextlib
import io
import json
import logging
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import refload  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mtlora_b200 import swin_transformer_mtlora as S
    from mtlora_b200.dist import sync_gradients
    m = refload.load_main()
    tasks = ["semseg", "normals", "sal", "human_parts"]
    config = refload.reference_config("mtlora/tiny_448/mtlora_tiny_448_r64_scale4_pertask.yaml", tasks,
                                      opts=["DATA.IMG_SIZE", 224, "TRAIN.EPOCHS", 1, "TRAIN.WARMUP_EPOCHS", 0, "PRINT_FREQ", 100])
    m.build.SwinTransformerMTLoRA = S.SwinTransformerMTLoRA          # the one-line swap of INTEGRATION.md §1
    torch.manual_seed(0)                                               # same initial weights on every rank
    with contextlib.redirect_stdout(io.StringIO()):
        model = m.build.build_mtl_model(m.build.build_model(config), config)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "lora_shared_B" in n or "lora_tasks_B" in n:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        model.cuda()
        m.main.mark_only_lora_as_trainable(model.backbone, bias="none")
    sync_gradients(model)                                              # optional line: also average the decoder heads
    optimizer = m.main.build_optimizer(config, model)
    loss_scaler = m.main.NativeScalerWithGradNormCount()
    gen = torch.Generator().manual_seed(100 + rank)                    # DIFFERENT data on every rank
    loader = []
    for _ in range(4):
        batch = {"image": torch.randn(2, 3, 224, 224, generator=gen)}
        batch.update(refload.synthetic_targets(tasks, 2, 224, gen))
        loader.append(batch)
    lr_scheduler = m.main.build_scheduler(config, optimizer, len(loader))
    loss_ft = torch.nn.ModuleDict({t: m.main.get_loss(config["TASKS_CONFIG"], t, config) for t in tasks})
    criterion = m.main.MultiTaskLoss(tasks, loss_ft, {t: refload.LOSS_WEIGHTS[t] for t in tasks})
    m.main.logger = logging.getLogger("ddp_check")
    m.main.wandb_available = False
    torch.manual_seed(1234 + rank)                                     # per-rank dropout / DropPath draws (main.py:570-575)
    m.main.train_one_epoch(config, model, criterion, loader, optimizer, 0, None, lr_scheduler, loss_scaler)
    torch.cuda.synchronize()
    # compare every trainable parameter with rank 0's
    worst, n_tr, moved = 0.0, 0, 0
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        worst = max(worst, float((p.detach() - ref).abs().max()))
        n_tr += 1
    ok = worst == 0.0
    if rank == 0:
        print(json.dumps({"world": world, "trainable_tensors": n_tr, "max_abs_difference_across_ranks": worst,
                          "identical": ok, "loop": "baseline/_ref/main.py::train_one_epoch (unmodified), 4 steps, fp16 AMP"}))
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
