#!/bin/bash
# scaling run on an N-GPU lease: bench.py at N (and the reference arm once) — usage: tools/gpu_scale.sh <tag> <N>
tag=${1:-sc}; n=${2:-8}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $n --steps 20 --warmup 5 \
  > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo "bench rc=$?" >> $out/${tag}_bench_n$n.err
