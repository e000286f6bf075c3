"""Report-only statistics against the images the reference repo ships (SURVEY §4 plan item 4, north_star "match the committed
check/*.png"): out/*.png (the author's own output of run.sh: -r 720 -s 3) and check/*.png (the instructor's), for the four
scenes that exist (in/lines_pointlight has no OBJ).  Compared: (a) the unmodified reference compiled here (oracle/_ref/
raytrace_ref), (b) the C oracle + reference tonemap, and — with --gpu on a GPU box that has the goldens — (c) the CUDA path.
Run in the container that has /root/reference:   python tools/report_vs_author.py [--out profiles/r2_vs_author.md]
Nothing here is a gate: SURVEY finding 1 — check/*.png is not reproducible from the reference's own code."""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("YRT_REFERENCE", "/root/reference")
SCENES = {"simple": "simple_pointlight", "basic": "basic_pointlight", "refl": "refl_pointlight", "instance": "instance10000_pointlight"}


def stats(a, b):
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    d = np.abs(a[..., :3].astype(int) - b[..., :3].astype(int)).max(axis=-1)
    return f"identical {100 * (d == 0).mean():.3f} %, within 1/255 {100 * (d <= 1).mean():.3f} %, max delta {int(d.max())}, mean |delta| {np.abs(a[..., :3].astype(int) - b[..., :3].astype(int)).mean():.4f}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_vs_author.md"))
    ap.add_argument("--gpu", action="store_true", help="also render with the CUDA path (needs a GPU)")
    ap.add_argument("--frames-dir", default=os.path.join(ROOT, "gpurun_out", "frames"), help="PNGs of the CUDA path rendered on the GPU box (tools/gpu_dump_frames.py)")
    ap.add_argument("--only-frames", action="store_true", help="skip the CPU renderers, compare only the PNGs in --frames-dir")
    a = ap.parse_args()
    from oracle import oracle
    from yocto_raytracing_b200.scene import FlatScene
    oracle.build()
    lines = ["# Report-only: images of this repo's paths vs the PNGs the reference repo ships (run.sh: -r 720 -s 3, 1280x720)", "",
             "`out/` = the author's own outputs, `check/` = the instructor's images (not reproducible from the reference's code, SURVEY finding 1).",
             "Rows: which renderer produced the compared image.  Nothing here gates a test.", "",
             "| scene | renderer | vs out/*.png | vs check/*.png |", "|---|---|---|---|"]
    cli = os.path.join(ROOT, "oracle", "_ref", "raytrace_ref")
    for short, name in SCENES.items():
        obj = f"{REF}/in/{name}/{name}.obj"
        out_png = np.array(Image.open(f"{REF}/out/{short}.png").convert("RGBA"))
        chk_png = np.array(Image.open(f"{REF}/check/{short}.png").convert("RGBA"))
        imgs = {}
        fp = os.path.join(a.frames_dir, short + ".png")
        if os.path.exists(fp):
            imgs["CUDA path on B200 (libyrt_b200.so, tools/gpu_dump_frames.py) + reference tonemap"] = np.array(Image.open(fp).convert("RGBA"))
        if a.only_frames:
            for k, v in imgs.items():
                lines.append(f"| {short} | {k} | {stats(v, out_png)} | {stats(v, chk_png)} |")
                print(lines[-1], flush=True)
            continue
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "ref.png")
            subprocess.run([cli, "-r", "720", "-s", "3", "-o", p, os.path.basename(obj)], cwd=os.path.dirname(obj), check=True, stdout=subprocess.DEVNULL)
            imgs["unmodified reference compiled here (oracle/_ref/raytrace_ref)"] = np.array(Image.open(p).convert("RGBA"))
        flat = FlatScene.load(os.path.join(ROOT, "tests", "golden", ("instance10000" if short == "instance" else short) + ".scene.npz"))
        w = flat.image_width(720)
        img, _ = oracle.OracleScene(flat).render(w, 720, 3, 0.1, max_depth=10 ** 6, threads=os.cpu_count() or 4)
        imgs["C oracle (oracle/yrt_oracle.c) + reference tonemap"] = oracle.tonemap(img)
        if a.gpu:
            import yocto_raytracing_b200 as y
            y.init(1)
            with y.Scene(flat) as scn:
                g, _ = scn.render(w, 720, 3, 0.1)
            imgs["CUDA path (libyrt_b200.so) + reference tonemap"] = oracle.tonemap(g)
        for k, v in imgs.items():
            lines.append(f"| {short} | {k} | {stats(v, out_png)} | {stats(v, chk_png)} |")
            print(lines[-1], flush=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
