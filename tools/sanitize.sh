#!/bin/bash
# compute-sanitizer passes over the whole path on small scenes: LBVH build (lock-free bottom-up refit / rotation /
# stack-need walks with arrival counters), one frame of a reflective scene (wave queues, recursion planes), one frame of
# the instance scene, the generic ray queries and the row-tile entry points.   usage (GPU box): tools/sanitize.sh [outdir]
out=${1:-gpurun_out}
mkdir -p $out
cd "$(dirname "$0")/.."
cat > /tmp/yrt_sanitize_job.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import synth
y.init(1)
for name, sc, (w, h, s) in (("mixed (mirrors, lines, points, textures)", synth.mixed_scene(7), (96, 54, 2)),
                            ("instance grid 12x12", synth.instance_grid_scene(12, seed=2), (96, 54, 2)),
                            ("hair 256", synth.hair_scene(256), (64, 36, 1))):
    flat = sc.flat()
    with y.Scene(flat) as scn:
        img, st = scn.render(w, h, s, 0.1)
        ids, dist, uv = scn.trace_primary(w, h, 1)
        ldr, _ = scn.render_ldr(w, h, s, 0.1)
        rng = np.random.RandomState(0)
        rays = np.concatenate([rng.uniform(-3, 3, (500, 3)), rng.normal(size=(500, 3)), np.full((500, 1), 1e-4), np.full((500, 1), 50.0)], 1).astype(np.float32)
        scn.intersect_first(rays); scn.intersect_any(rays)
        print(name, st.total_rays, "rays, depth", st.max_depth, "finite", bool(np.isfinite(img).all()))
print("SANITIZE_JOB_OK")
PY
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/yrt_sanitize_job.py > $out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" >> $out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_JOB_OK|rc=" $out/sanitize_$tool.log | tail -4
done
