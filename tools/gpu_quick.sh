#!/bin/bash
# quick check: GPU tests, the headline frame with / without the grids, launch list of one frame
tag=${1:-q}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tools/ab_env.sh "" "YRT_PGRID=0" "YRT_CAM_CELL_SHIFT=4" "YRT_CAM_CELL_SHIFT=2" > $out/${tag}_ab.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 100 -c 60 --csv --log-file $out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_launch.log 2>&1
echo done > $out/${tag}_done.txt
