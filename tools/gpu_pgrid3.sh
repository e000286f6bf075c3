#!/bin/bash
# tests + bench + ncu full capture of the traversal kernels (chain-node grids)
tag=${1:-pg4}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 100 -c 80 --csv --log-file $out/${tag}_launches.csv python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_trace\|k_shade --launch-skip 3 -c 3 -f -o $out/${tag}_trace python tools/profile_frame.py --frames 2 > $out/${tag}_ncu_full.log 2>&1
echo done > $out/${tag}_done.txt
