"""Render a few frames of a bench config (default: the headline one) and print the per-kernel CUDA-event times of each; the
command profiled under ncu (see profiles/README.md).  --config: simple | basic | refl | lines | instance | instance_real."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2)
ap.add_argument("--config", default="instance")
ap.add_argument("--resolution", type=int, default=0)
ap.add_argument("--samples", type=int, default=0)
a = ap.parse_args()

import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import configs

y.init(1)
flat, res, smp, name = configs.load(a.config)
res, smp = a.resolution or res, a.samples or smp
w = flat.image_width(res)
with y.Scene(flat) as scn:
    print(a.config, scn.info())
    for f in range(a.frames):
        img, st = scn.render(w, res, smp, 0.1)
        print(f"frame {f}: {st.total_rays} rays {st.ms_total:.3f} ms -> {st.total_rays / st.ms_total / 1e3:.1f} Mrays/s | closest {st.ms_trace_closest:.3f} any {st.ms_trace_any:.3f} "
              f"shade {st.ms_shade:.3f} other {st.ms_other:.3f} launches {st.launches} depth {st.max_depth} truncated {st.truncated_paths}")
