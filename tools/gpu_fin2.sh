#!/bin/bash
# short final pass: GPU tests, both bench arms, frame log
tag=${1:-fin2}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -rs > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench rc=$?" >> $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_n1.json 2> $out/${tag}_bench_reference_n1.err
timeout 300 python tools/profile_frame.py --frames 4 > $out/${tag}_frame.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1
echo done > $out/${tag}_done.txt
