#!/bin/bash
# GPU tests + bench at N = 1 (short)
tag=${1:-chk}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?" >> $out/${tag}_bench_n1.err
timeout 300 python tools/profile_frame.py --frames 4 > $out/${tag}_frame.log 2>&1
timeout 300 python tools/diag_streams.py > $out/${tag}_streams.log 2>&1
echo done > $out/${tag}_done.txt
