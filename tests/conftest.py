import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
RIGID_GOLDEN_CASES = ["simple", "basic", "refl", "instance10000", "lines_synth", "mixed7", "gltf7"]
# + scenes with scaled / sheared instance frames (OBJ `i` lines; glTF node scales), traced through the reference's own instance tree
GOLDEN_CASES = RIGID_GOLDEN_CASES + ["nonrigid31", "gltf23s"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    from yocto_raytracing_b200.scene import FlatScene
    flat = FlatScene.load(os.path.join(GOLDEN, name + ".scene.npz"))
    with np.load(os.path.join(GOLDEN, name + ".ref.npz")) as z:
        ref = {k: z[k] for k in z.files}
    return flat, ref


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu():
    """The product library bound to cuda:0; GPU tests fail (not skip) if it cannot be loaded."""
    import yocto_raytracing_b200 as y
    y.init(1)
    return y


def ldr_stats(a8, b8):
    """Fraction of pixels whose RGBA8 channels all differ by <= 1 (and == 0), and the max difference."""
    d = np.abs(a8.astype(np.int32) - b8.astype(np.int32)).max(axis=-1)
    return float((d <= 1).mean()), float((d == 0).mean()), int(d.max())


def id_match(ids, ref_ids):
    return float((ids == ref_ids).all(axis=1).mean())
