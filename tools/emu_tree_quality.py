"""Per-ray work of the LBVH on the host emulation (no GPU): box tests / element tests / instance entries per primary ray
for different build settings, e.g.  YRT_ROTATE_BLAS=2 YRT_ROTATE_TLAS=1 python tools/emu_tree_quality.py"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _emu  # noqa: E402
from yocto_raytracing_b200 import synth  # noqa: E402

flat = synth.instance_grid_scene(100).flat()
w, h = 640, 360
t0 = time.time()
es = _emu.EmuScene(flat)
t1 = time.time()
ids, dist, uv, c = es.trace_primary(w, h, 1)
n = w * h
print(f"rotate blas={os.environ.get('YRT_ROTATE_BLAS', 'default')} tlas={os.environ.get('YRT_ROTATE_TLAS', 'default')}: "
      f"box tests/ray {c[0] / n:.2f} (tlas {c[6] / n:.2f}, blas {(c[0] - c[6]) / n:.2f}), element tests {c[1] / n:.2f}, "
      f"instance entries {c[2] / n:.2f}, max stack {c[3]}, false rejects {c[4]}, depth blas/tlas {es.info()[2]}/{es.info()[3]}, "
      f"build {t1 - t0:.2f} s, hits {(ids[:, 0] >= 0).mean():.4f}, checksum {int(ids.astype('int64').sum())} {float(dist[ids[:,0]>=0].astype('float64').sum()):.6f}")
