"""The five configs of BASELINE.json as flattened scenes that travel to the GPU box.

configs 1-3 and 5 exist as the reference's own scenes (tests/golden/<name>.scene.npz: flattened by the reference's loader,
bin/yrt_flatten); config 4 (lines) has no OBJ in the reference and is generated (synth.lines_config4); the headline bench
scene is the synthetic instance10000-shaped one (synth.instance_grid_scene) because BASELINE asks for synthetic scenes of
the named shape at 1/2/4/8 GPUs.  load(name) -> (FlatScene, resolution, samples per axis, description)."""
import os

from . import synth
from .scene import FlatScene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NAMES = ("simple", "basic", "refl", "lines", "instance", "instance_real")


def load(name: str):
    if name in ("simple", "basic", "refl"):
        flat = FlatScene.load(os.path.join(GOLDEN, name + ".scene.npz"))
        return flat, 720, 3, f"in/{name}_pointlight (the reference's scene), run.sh: -r 720 -s 3"
    if name == "lines":
        return synth.lines_config4().flat(), 720, 3, "lines config (SURVEY 8d config 4, synthetic: 2 x 65 536 hairs x 8 segments), -r 720 -s 3"
    if name == "instance_real":
        flat = FlatScene.load(os.path.join(GOLDEN, "instance10000.scene.npz"))
        return flat, 1080, 4, "in/instance10000_pointlight (the reference's scene), -r 1080 -s 4"
    if name == "instance":
        return synth.instance_grid_scene(100).flat(), 1080, 4, "instance10000_pointlight-shaped synthetic scene, -r 1080 -s 4"
    raise ValueError(f"unknown config {name!r}; one of {NAMES}")
