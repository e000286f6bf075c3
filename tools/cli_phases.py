"""Wall-clock phases of the drop-in CLI (bin/raytrace --stats) on the headline config, with the reference's own host code
for load / tonemap / PNG and with the SURVEY 8f replacements (--cache, --device-ldr, --fast-png).  python tools/cli_phases.py"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yocto_raytracing_b200 import synth

exe = os.path.join(ROOT, "bin", "raytrace")
td = tempfile.mkdtemp()
sc = synth.instance_grid_scene(100)
obj = sc.write_obj(td)
name = os.path.basename(obj)
base = ["-r", "1080", "-s", "4", "--stats"]
runs = [("reference host code (OBJ load, host tonemap, stb PNG)", []),
        ("--cache (first run: writes the cache)", ["--cache"]),
        ("--cache --device-ldr --fast-png", ["--cache", "--device-ldr", "--fast-png"]),
        ("--cache --device-ldr --fast-png (again)", ["--cache", "--device-ldr", "--fast-png"])]
for label, extra in runs:
    r = subprocess.run([exe] + base + extra + ["-o", "out.png", name], cwd=td, capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if l.startswith("phases") or l.startswith("rays")]
    print(f"[{label}] rc={r.returncode}")
    for l in lines:
        print("   ", l)
    if r.returncode != 0:
        print(r.stdout[-400:], r.stderr[-400:])
