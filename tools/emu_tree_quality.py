"""Per-ray work of the traversal tree on the host emulation (no GPU): node visits / box tests / element tests / instance
entries per primary ray and per shadow ray for a library variant and build settings, e.g.
    python tools/emu_tree_quality.py            # default build (YRT_WIDE=4)
    python tools/emu_tree_quality.py bin        # -DYRT_WIDE=2
    YRT_ROTATE_BLAS=2 python tools/emu_tree_quality.py pack"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _emu  # noqa: E402
from yocto_raytracing_b200 import synth  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else ""
flat = synth.instance_grid_scene(100).flat()
w, h = 640, 360
t0 = time.time()
es = _emu.EmuScene(flat, variant=variant)
t1 = time.time()
ids, dist, uv, c = es.trace_primary(w, h, 1)
n = w * h
print(f"variant '{variant or 'default'}' primary: visits/ray {c[7] / n:.2f}, box tests/ray {c[0] / n:.2f} (tlas {c[6] / n:.2f}, blas {(c[0] - c[6]) / n:.2f}), "
      f"element tests {c[1] / n:.2f}, instance entries {c[2] / n:.2f}, max stack {c[3]}, false rejects {c[4]}, depth blas/tlas {es.info()[2]}/{es.info()[3]}, "
      f"build {t1 - t0:.2f} s, hits {(ids[:, 0] >= 0).mean():.4f}, checksum {int(ids.astype('int64').sum())} {float(dist[ids[:,0]>=0].astype('float64').sum()):.6f}")
img, rc = es.render(w // 2, h // 2, 1)
ns = max(rc[2], 1)
print(f"  shadow rays: visits/ray {rc[8] / ns:.2f}, box tests/ray {rc[3] / ns:.2f} (tlas {rc[4] / ns:.2f}), element tests {rc[5] / ns:.2f}, "
      f"instance entries {rc[6] / ns:.2f}, occluded {rc[7] / ns:.3f}")
