#!/bin/bash
# multi-GPU session (gpurun --gpus N): the GPU test suite with no multi-GPU test skipped, then bench.py under torchrun
# usage: tools/gpu_multi.sh <tag> <N>
tag=${1:-m}; n=${2:-2}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
nvidia-smi topo -m >> $out/${tag}_smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -rs > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 20 --warmup 5 \
  > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo "bench rc=$?" >> $out/${tag}_bench_n$n.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo done > $out/${tag}_done.txt
