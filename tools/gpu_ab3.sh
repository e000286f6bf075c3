#!/bin/bash
tag=${1:-ab3}
out=gpurun_out
mkdir -p $out
tools/ab_variants.sh base wide4 anybin c9 a9 c7 > $out/${tag}_variants.log 2>&1
tools/ab_env.sh "" "YRT_LEAF_BLAS=2" "YRT_LEAF_BLAS=4" "YRT_ROTATE_BLAS=4" "YRT_ROTATE_BLAS=2" "YRT_BLOCKS_PER_SM=6" "YRT_BLOCKS_PER_SM=7" > $out/${tag}_env.log 2>&1
echo done > $out/${tag}_done.txt
