#!/bin/bash
# GPU session for the apex grids: parity suite, then A/B of the grid switches on the headline frame and the other configs.
# usage (on the box, from the repo root): tools/gpu_pgrid.sh <tag>
tag=${1:-pg}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tools/ab_env.sh "" "YRT_PGRID=0" "YRT_CAM_GRID=0" "YRT_LIGHT_GRID_R=0" "YRT_CAM_CELL_SHIFT=2" "YRT_CAM_CELL_SHIFT=4" \
    "YRT_LIGHT_GRID_R=64" "YRT_LIGHT_GRID_R=256" > $out/${tag}_ab.log 2>&1
for c in simple basic refl lines instance_real; do
  for e in "" "YRT_PGRID=0"; do
    printf "%-14s %-12s " "$c" "[$e]"; env $e timeout 120 python tools/profile_frame.py --frames 4 --config $c 2>&1 | tail -1
  done
done > $out/${tag}_configs.log 2>&1
timeout 300 python tools/build_trace.py > $out/${tag}_build.log 2>&1
YRT_B200_LIB=$PWD/yocto_raytracing_b200/libyrt_b200_counters.so timeout 300 python tools/frame_counters.py > $out/${tag}_counters.json 2> $out/${tag}_counters.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 150 -c 60 --csv --log-file $out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_launch.log 2>&1
echo done > $out/${tag}_done.txt
