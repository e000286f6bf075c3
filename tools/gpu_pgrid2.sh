#!/bin/bash
# second GPU session for the apex grids: build-phase trace, ncu launch lists (build + frame), full capture of the traversal kernels, counters
tag=${1:-pg2}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/build_trace.py > $out/${tag}_build.log 2>&1
timeout 300 python tools/build_trace.py --config lines >> $out/${tag}_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_build_launches.csv python tools/build_trace.py --builds 2 > $out/${tag}_ncu_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_launch.log 2>&1
YRT_B200_LIB=$PWD/yocto_raytracing_b200/libyrt_b200_counters.so timeout 300 python tools/frame_counters.py > $out/${tag}_counters.json 2> $out/${tag}_counters.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_trace --launch-skip 2 -c 2 -f -o $out/${tag}_trace python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_full.log 2>&1
echo done > $out/${tag}_done.txt
