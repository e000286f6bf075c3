#!/bin/bash
tag=${1:-ab4}
out=gpurun_out
mkdir -p $out
CONFIGS="simple basic refl lines instance_real instance" tools/ab_variants.sh base wide4 > $out/${tag}_variants.log 2>&1
echo done > $out/${tag}_done.txt
