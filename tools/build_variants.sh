#!/bin/bash
# build A/B variants of the library: tools/build_variants.sh name1 "flags1" name2 "flags2" ... -> build/variants/<name>.so
# run one with  YRT_B200_LIB=build/variants/<name>.so python tools/profile_frame.py
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  rm -rf build/v_$name; mkdir -p build/v_$name
  for f in yrt_host yrt_build yrt_render yrt_api yrt_png; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC $flags -c yocto_raytracing_b200/csrc/$f.cu -o build/v_$name/$f.o &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so build/v_$name/*.o -lz
  echo "built build/variants/$name.so ($flags)"
done
