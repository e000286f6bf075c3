#!/bin/bash
# quick A/B on the GPU box: tools/gpu_ab.sh <tag> variants...   (+ YRT_BUILD_TRACE of one scene creation)
tag=${1:-ab}; shift
mkdir -p gpurun_out
tools/ab_variants.sh "$@" > gpurun_out/${tag}_ab.log 2>&1
YRT_BUILD_TRACE=1 timeout 120 python tools/profile_frame.py --frames 1 > gpurun_out/${tag}_buildtrace.log 2>&1
YRT_BUILD_TRACE=1 timeout 120 python tools/profile_frame.py --frames 1 --config lines >> gpurun_out/${tag}_buildtrace.log 2>&1
