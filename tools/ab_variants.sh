#!/bin/bash
# A/B of library build variants on the GPU box: tools/ab_variants.sh name1 name2 ... ("base" = the in-tree library)
# prints the last frame line of tools/profile_frame.py for each (CUDA-event times, no profiler)
cd "$(dirname "$0")/.."
for v in "$@"; do
  if [ "$v" = base ]; then unset YRT_B200_LIB; else export YRT_B200_LIB=$PWD/build/variants/$v.so; fi
  printf "%-12s " "$v"; timeout 120 python tools/profile_frame.py --frames 4 2>&1 | tail -1
done
