"""Create the scene of a bench config a few times (YRT_BUILD_TRACE=1 prints the phases of every build) — the command profiled
under ncu for the per-kernel times of the scene build (LBVH + light grids).  --config as in tools/profile_frame.py."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="instance")
ap.add_argument("--builds", type=int, default=3)
a = ap.parse_args()
os.environ.setdefault("YRT_BUILD_TRACE", "1")
import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import configs

y.init(1)
flat, res, smp, name = configs.load(a.config)
for b in range(a.builds):
    t0 = time.perf_counter()
    scn = y.Scene(flat)
    t1 = time.perf_counter()
    print(f"build {b}: yrt_scene_create {1e3 * (t1 - t0):.3f} ms wall, {scn.info()}")
    scn.close() if hasattr(scn, "close") else None
