#!/bin/bash
# A/B of library build variants on the GPU box: tools/ab_variants.sh name1 name2 ... ("base" = the in-tree library)
# prints the last frame line of tools/profile_frame.py for each (CUDA-event times, no profiler); CONFIGS="instance refl ..." to sweep configs
cd "$(dirname "$0")/.."
for c in ${CONFIGS:-instance}; do
for v in "$@"; do
  if [ "$v" = base ]; then unset YRT_B200_LIB; else export YRT_B200_LIB=$PWD/build/variants/$v.so; fi
  printf "%-9s %-10s " "$c" "$v"; timeout 300 python tools/profile_frame.py --frames 4 --config $c 2>&1 | tail -1
done
done
