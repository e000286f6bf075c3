#!/bin/bash
# A/B of an environment switch under torchrun: tools/gpu_scale_ab.sh <tag> <N> "VAR=a" "VAR=b" ...
tag=$1; n=$2; shift 2
out=gpurun_out; mkdir -p $out
i=0
for e in "$@"; do
  env $e timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29560 + i)) bench.py --gpus $n --steps 20 --warmup 5 --no-extras \
    > $out/${tag}_$i.json 2> $out/${tag}_$i.err
  echo "[$e] $(python -c "import json,sys; d=json.loads(open('$out/${tag}_$i.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['per_rank_render_ms_and_kernel_sum'][0])")"
  i=$((i+1))
done > $out/${tag}_ab.log 2>&1
