#!/bin/bash
tag=${1:-ab2}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q -k "apex or row_tiles or determinism or batch_size" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tools/ab_env.sh "" "YRT_TILE_SHIFT=3" "YRT_TILE_SHIFT=3 YRT_CHUNK=128" "YRT_TILE_SHIFT=3 YRT_CHUNK=256" "YRT_TILE_SHIFT=3 YRT_CHUNK=512" "YRT_TILE_SHIFT=3 YRT_CHUNK=1024" \
   "YRT_TILE_SHIFT=4 YRT_CHUNK=1024" "YRT_TILE_SHIFT=4 YRT_CHUNK=4096" "YRT_TILE_SHIFT=2 YRT_CHUNK=256" "YRT_CHUNK=256" "YRT_TILE_SHIFT=3 YRT_CHUNK=1024 YRT_PGRID=0" > $out/${tag}_ab.log 2>&1
echo done > $out/${tag}_done.txt
