#!/bin/bash
# BASELINE.json configs[2] (Swin-S 448, 4 tasks, r 64/4, 8 x B200) and configs[3] (Swin-B 448, 6 tasks, r 32/4, 8 x B200):
# one torchrun per config, adapter-only gradient exchange from autograd hooks (mtlora_b200/dist.py).
out=gpurun_out
run() {  # name, extra args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-8} --master-addr 127.0.0.1 --master-port 29521 bench.py \
    --gpus ${NGPU:-8} --steps 20 --warmup 5 --no-extras "$@" > $out/r02_$name.json 2> $out/r02_$name.err
  echo "== $name rc=$?"; tail -2 $out/r02_$name.err | cut -c1-300; head -c 700 $out/r02_$name.json; echo
}
run swin_s_${NGPU:-8}gpu --model swin_s
run swin_b_6task_${NGPU:-8}gpu --model swin_b --tasks 6 --r-shared 32 || true
if ! grep -q '"value"' $out/r02_swin_b_6task_${NGPU:-8}gpu.json; then run swin_b_6task_${NGPU:-8}gpu_b16 --model swin_b --tasks 6 --r-shared 32 --batch 16; fi
