"""Render a few frames of the headline config; the command profiled under ncu (see profiles/README.md)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--resolution", type=int, default=1080)
ap.add_argument("--samples", type=int, default=4)
ap.add_argument("--n-side", type=int, default=100)
ap.add_argument("--scene", default="grid", choices=["grid", "hair", "mixed"])
a = ap.parse_args()

import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import synth

y.init(1)
sc = {"grid": lambda: synth.instance_grid_scene(a.n_side), "hair": lambda: synth.hair_scene(16384), "mixed": lambda: synth.mixed_scene(7)}[a.scene]()
flat = sc.flat()
w = flat.image_width(a.resolution)
with y.Scene(flat) as scn:
    print(scn.info())
    for f in range(a.frames):
        img, st = scn.render(w, a.resolution, a.samples, 0.1)
        print(f"frame {f}: {st.total_rays} rays {st.ms_total:.3f} ms -> {st.total_rays / st.ms_total / 1e3:.1f} Mrays/s | closest {st.ms_trace_closest:.3f} any {st.ms_trace_any:.3f} "
              f"shade {st.ms_shade:.3f} other {st.ms_other:.3f} launches {st.launches}")
