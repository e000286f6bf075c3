"""Full-size goldens from the UNMODIFIED reference compiled here (oracle/_ref, `make ref`) — the BASELINE target config on the
REAL scene, and the lines config at its specified size.  Run in a container that has /root/reference (takes ~15 min on one core:
the reference is single-threaded):

    python tools/make_golden_full.py [instance] [lines]

instance:  /root/reference/in/instance10000_pointlight at 1920x1080
  tests/golden/instance10000_1080p_s4.png      the reference CLI's own output (raytrace_ref -r 1080 -s 4: raytrace() + tonemap +
                                               stb PNG), sha256 of the decoded RGBA bytes starts 22bc1ac0e6f98ba9
  tests/golden/instance10000_1080p.ref.npz     ref_probe ids at 1 spp: inst, ei (int16/int32), dist (float32) per primary ray
lines:     synth.lines_config4() (2 x 65 536 hairs x 8 segments) at 1280x720
  tests/golden/lines_config4.ref.npz           ref_probe ids at 1 spp (the scene itself is regenerated from its seed, not stored)
  tests/golden/lines_config4_720p_s3.png       raytrace_ref -r 720 -s 3
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_probe  # noqa: E402

REF = os.environ.get("YRT_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "oracle", "_ref", "raytrace_ref")


def small(a):
    return a.astype(np.int16) if a.min() >= -32768 and a.max() < 32768 else a.astype(np.int32)


def cli_png(obj, res, smp, out_png):
    subprocess.run([CLI, "-r", str(res), "-s", str(smp), "-o", out_png, os.path.basename(obj)], check=True, cwd=os.path.dirname(obj),
                   stdout=subprocess.DEVNULL)
    from PIL import Image
    a = np.array(Image.open(out_png))
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def ids_npz(obj, res, out_npz, **extra):
    w, h, rec = ref_probe.ids(obj, res, 1)
    np.savez_compressed(out_npz, inst=small(rec["inst"]), ei=small(rec["ei"]), dist=rec["dist"].astype(np.float32), ids_width=w, ids_height=h, **extra)
    return w, h


def main():
    what = sys.argv[1:] or ["instance", "lines"]
    if "instance" in what:
        obj = f"{REF}/in/instance10000_pointlight/instance10000_pointlight.obj"
        print("ids", ids_npz(obj, 1080, os.path.join(OUT, "instance10000_1080p.ref.npz")))
        print("png sha", cli_png(obj, 1080, 4, os.path.join(OUT, "instance10000_1080p_s4.png")))
    if "lines" in what:
        from yocto_raytracing_b200 import synth
        sc = synth.lines_config4()
        with tempfile.TemporaryDirectory() as td:
            obj = sc.write_obj(td)
            print("ids", ids_npz(obj, 720, os.path.join(OUT, "lines_config4.ref.npz")))
            print("png sha", cli_png(obj, 720, 3, os.path.join(OUT, "lines_config4_720p_s3.png")))


if __name__ == "__main__":
    main()
