"""On the GPU box: render the reference's four scenes with the CUDA path at run.sh's settings (-r 720 -s 3), apply the reference
tonemap (C oracle's restatement) and save them as PNGs (gpurun_out/frames/<scene>.png), so that tools/report_vs_author.py can
compare them with the reference repo's out/*.png and check/*.png in the container that has /root/reference."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from PIL import Image  # noqa: E402

import yocto_raytracing_b200 as y  # noqa: E402
from oracle import oracle  # noqa: E402
from yocto_raytracing_b200.scene import FlatScene  # noqa: E402

oracle.build()
y.init(1)
out = os.path.join(ROOT, "gpurun_out", "frames")
os.makedirs(out, exist_ok=True)
for short, gold in (("simple", "simple"), ("basic", "basic"), ("refl", "refl"), ("instance", "instance10000")):
    flat = FlatScene.load(os.path.join(ROOT, "tests", "golden", gold + ".scene.npz"))
    w = flat.image_width(720)
    with y.Scene(flat) as scn:
        img, st = scn.render(w, 720, 3, 0.1)
    Image.fromarray(oracle.tonemap(img), "RGBA").save(os.path.join(out, short + ".png"))
    print(short, st.total_rays, "rays", round(st.ms_total, 3), "ms, truncated", st.truncated_paths)
