"""cProfile of the host side of a few training steps of the bench workload (which Python functions cost launch time)."""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    sys.argv = [sys.argv[0]] + sys.argv[1:]
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
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        step(img, None)
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
    print(s.getvalue()[:6000])


if __name__ == "__main__":
    main()
