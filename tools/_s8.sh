#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/s8_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
tools/ab_env.sh "" "YRT_STREAMS=2" "YRT_STREAMS=3" > gpurun_out/s8_env.log 2>&1
