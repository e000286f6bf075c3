"""Run oracle/_ref/ref_probe (the unmodified reference compiled from /root/reference/src) and parse its dumps."""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBE = os.path.join(ROOT, "oracle", "_ref", "ref_probe")
REC = np.dtype([("inst", "<i4"), ("shape", "<i4"), ("ei", "<i4"), ("dist", "<f4"), ("w1", "<f4"), ("w2", "<f4")])


def available():
    return os.path.exists(PROBE)


def ids(obj_path, resolution, samples, mode="ids"):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "ids.bin")
        subprocess.run([PROBE, mode, os.path.basename(obj_path), str(resolution), str(samples), out], check=True,
                       cwd=os.path.dirname(obj_path) or ".", stdout=subprocess.PIPE)
        raw = np.fromfile(out, dtype=np.uint8)
    hdr = raw[:16].view("<i4")
    rec = raw[16:].view(REC)
    return int(hdr[0]), int(hdr[1]), rec


def image(obj_path, resolution, samples, amb=0.1):
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "img.bin")
        p = subprocess.run([PROBE, "image", os.path.basename(obj_path), str(resolution), str(samples), repr(float(amb)), out],
                           check=True, cwd=os.path.dirname(obj_path) or ".", stdout=subprocess.PIPE)
        info = json.loads(p.stdout.decode().strip().splitlines()[-1])
        raw = np.fromfile(out, dtype=np.uint8)
    w, h = raw[:8].view("<i4")
    img = raw[8:].view("<f4").reshape(int(h), int(w), 4).copy()
    return img, info
