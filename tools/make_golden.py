"""Generate tests/golden/ from the UNMODIFIED reference compiled here (oracle/_ref, see Makefile `ref`).

Run in a container that has /root/reference:   python tools/make_golden.py
Writes, per reference scene (in/*_pointlight/*.obj) and per synthetic scene (OBJ dialect, and one glTF written by tools/make_gltf.py):
  tests/golden/<name>.scene.npz   flattened scene as the reference's loader produced it (bin/yrt_flatten)
  tests/golden/<name>.ref.npz     outputs of the reference itself (oracle/_ref/ref_probe):
                                  image   float32 H x W x 4  raytrace() before tonemap (-r R -s S -a 0.1)
                                  ids     int32 n x 3        (instance, shape, element) per primary ray, 1 spp
                                  dist    float32 n          intersection3f::dist
                                  + the parameters used
The GPU box has no /root/reference; tests read only these files.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_probe  # noqa: E402
from yocto_raytracing_b200 import synth  # noqa: E402
from yocto_raytracing_b200.scene import FlatScene  # noqa: E402

REF = os.environ.get("YRT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
FLATTEN = os.path.join(ROOT, "bin", "yrt_flatten")

# name -> (obj path, image resolution, image samples, ids resolution)
CASES = {
    "simple": (f"{REF}/in/simple_pointlight/simple_pointlight.obj", 90, 2, 180),
    "basic": (f"{REF}/in/basic_pointlight/basic_pointlight.obj", 90, 2, 180),
    "refl": (f"{REF}/in/refl_pointlight/refl_pointlight.obj", 90, 2, 180),
    "instance10000": (f"{REF}/in/instance10000_pointlight/instance10000_pointlight.obj", 90, 2, 180),
}


def one(name, obj, res, smp, ids_res):
    with tempfile.TemporaryDirectory() as td:
        y = os.path.join(td, "s.yrts")
        subprocess.run([FLATTEN, os.path.basename(obj), y], check=True, cwd=os.path.dirname(obj), stdout=subprocess.DEVNULL)
        flat = FlatScene.load(y)
    flat.save_npz(os.path.join(OUT, name + ".scene.npz"))
    img, info = ref_probe.image(obj, res, smp, 0.1)
    w, h, rec = ref_probe.ids(obj, ids_res, 1)
    np.savez_compressed(os.path.join(OUT, name + ".ref.npz"), image=img, image_resolution=res, image_samples=smp, ambient=np.float32(0.1),
                        ids=np.stack([rec["inst"], rec["shape"], rec["ei"]], 1).astype(np.int32), dist=rec["dist"].astype(np.float32),
                        uv=np.stack([rec["w1"], rec["w2"]], 1).astype(np.float32), ids_width=w, ids_height=h)
    print(name, "image", img.shape, "ids", rec.shape, info)


def main():
    """python tools/make_golden.py [case ...]   (no arguments: every case)"""
    only = set(sys.argv[1:])
    want = lambda name: not only or name in only
    os.makedirs(OUT, exist_ok=True)
    for name, (obj, res, smp, ids_res) in CASES.items():
        if want(name):
            one(name, obj, res, smp, ids_res)
    # synthetic scenes written in the reference's OBJ dialect and rendered by the reference
    # (nonrigid31: scaled + sheared `i` lines — the reference's result depends on its own instance tree there)
    with tempfile.TemporaryDirectory() as td:
        for sc, res, smp, ids_res in ((synth.hair_scene(1024), 90, 2, 180), (synth.mixed_scene(7), 90, 2, 180), (synth.nonrigid_scene(31, 5, 3), 90, 2, 180)):
            if want(sc.name):
                obj = sc.write_obj(os.path.join(td, sc.name))
                one(sc.name, obj, res, smp, ids_res)
        # glTF input: node hierarchy with translation + rotation quaternions, KHR_materials_pbrSpecularGlossiness
        # materials, POINTS primitives as lights, a camera node — loaded by the reference's own glTF loader
        import make_gltf
        if want("gltf7"):
            one("gltf7", make_gltf.gltf_scene(os.path.join(td, "gltf7")), 90, 2, 180)
        # ... and the same with node scales on every third object (children inherit them): non-rigid instance frames
        if want("gltf23s"):
            one("gltf23s", make_gltf.gltf_scene(os.path.join(td, "gltf23s"), seed=23, n_objects=14, name="gltf23s", scales=True), 90, 2, 180)


if __name__ == "__main__":
    main()
