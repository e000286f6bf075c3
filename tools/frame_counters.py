"""Per-ray work of the shipped traversal kernels, counted ON THE DEVICE by a -DYRT_COUNTERS=1 build of the library
(tools/build_variants.sh counters "-DYRT_COUNTERS=1"; selected with YRT_B200_LIB): renders one frame of a bench config and
prints one JSON object with, per kernel class, rays / node visits / box tests / element tests / instance entries per ray
and the flops and bytes they stand for (same unit costs as SURVEY 8d: 26 flop per box test, 54 per triangle test, 43 per
instance entry; node record bytes as laid out in csrc/yrt_scene.cuh).  Never the timed build: bench.py runs this in a
separate process after its timed region."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

FLOP_BOX, FLOP_ELEM, FLOP_INST = 26.0, 54.0, 43.0       # SURVEY 8d unit costs (reference arithmetic)
BYTES_NODE = {2: 56.0, 4: 112.0}                          # bytes a lane reads per node visit: binary record (3 quads + 8 B) / 4-wide record (7 quads)
BYTES_ELEM, BYTES_INST = 48.0, 64.0                       # ... per element test / instance entry
ARITY = {"camera_rays": 2, "mirror_rays": 2, "shadow_rays": 4}   # default build: YRT_WIDE_CLOSEST = 2, YRT_WIDE_ANY = 4 (csrc/yrt_scene.cuh)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="instance")
    ap.add_argument("--resolution", type=int, default=0)
    ap.add_argument("--samples", type=int, default=0)
    a = ap.parse_args()
    import yocto_raytracing_b200 as y
    from yocto_raytracing_b200 import _lib, configs
    y.init(1)
    flat, res, smp, name = configs.load(a.config)
    res, smp = a.resolution or res, a.samples or smp
    w = flat.image_width(res)
    out = (C.c_uint64 * 24)()
    with y.Scene(flat) as scn:
        lib = _lib.load()
        st = lib.yrt_counters_read(scn._h, out)
        if st != 0:
            print(json.dumps({"error": lib.yrt_last_error().decode()}))
            return 1
        img, stats = scn.render(w, res, smp, 0.1)
        _lib.check(lib.yrt_counters_read(scn._h, out))
    v = list(out)
    res_ = {"config": name, "width": w, "height": res, "spp": smp * smp, "library": os.environ.get("YRT_B200_LIB", "default"),
            "rays_by_reference_semantics": stats.total_rays}
    for k, cls in enumerate(("camera_rays", "mirror_rays", "shadow_rays")):
        rays, visits, box, tbox, elem, inst = v[8 * k: 8 * k + 6]
        if not rays:
            res_[cls] = None
            continue
        r = float(rays)
        flops = (box * FLOP_BOX + elem * FLOP_ELEM + inst * FLOP_INST) / r
        byts = (visits * BYTES_NODE[ARITY[cls]] + elem * BYTES_ELEM + inst * BYTES_INST) / r
        res_[cls] = {"rays": rays, "node_arity": ARITY[cls], "node_visits_per_ray": visits / r, "box_tests_per_ray": box / r, "tlas_box_tests_per_ray": tbox / r,
                     "element_tests_per_ray": elem / r, "instance_entries_per_ray": inst / r, "flops_per_ray": flops, "l1_bytes_per_ray": byts}
    print(json.dumps(res_))
    return 0


if __name__ == "__main__":
    sys.exit(main())
