#!/bin/bash
# one GPU-box session: tests, bench, A/B of library variants, device counters, ncu launch list + full capture.
# usage (on the box, from the repo root): tools/gpu_session.sh <tag> [variants...]
tag=${1:-s}; shift
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err
if [ $# -gt 0 ]; then tools/ab_variants.sh "$@" > $out/${tag}_ab.log 2>&1; fi
if [ -f build/variants/counters.so ]; then
  YRT_B200_LIB=$PWD/build/variants/counters.so timeout 300 python tools/frame_counters.py > $out/${tag}_counters.json 2> $out/${tag}_counters.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_trace\|k_shade\|k_resolve --launch-skip 4 -c 4 -f -o $out/${tag}_trace python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_full.log 2>&1
timeout 300 python tools/gpu_dump_frames.py > $out/${tag}_frames.log 2>&1
echo done > $out/${tag}_done.txt
