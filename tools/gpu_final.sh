#!/bin/bash
# final 1-GPU session of a round: parity suite, bench (both arms), launch lists, full ncu capture, counters, build trace, CLI phases
# usage (on the box, from the repo root): tools/gpu_final.sh <tag>
tag=${1:-fin}
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rs > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?" >> $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_n1.json 2> $out/${tag}_bench_reference_n1.err
timeout 300 python tools/build_trace.py > $out/${tag}_build.log 2>&1
YRT_B200_LIB=$PWD/yocto_raytracing_b200/libyrt_b200_counters.so timeout 300 python tools/frame_counters.py > $out/${tag}_counters.json 2> $out/${tag}_counters.err
timeout 300 python tools/profile_frame.py --frames 4 > $out/${tag}_frame.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_trace\|k_shade\|k_resolve --launch-skip 4 -c 4 -f -o $out/${tag}_trace python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_full.log 2>&1
timeout 300 python tools/cli_phases.py > $out/${tag}_cli_phases.log 2>&1
echo done > $out/${tag}_done.txt
