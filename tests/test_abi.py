"""The C-ABI library loads and exports every symbol include/yrt_b200.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from yocto_raytracing_b200 import _lib
from yocto_raytracing_b200 import synth


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "yrt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yrt_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/yrt_b200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype"
    assert lib.yrt_abi_version() == 3


def test_struct_layouts_match_header():
    # sizes implied by the header on LP64
    assert C.sizeof(_lib.Camera) == 16 * 4
    assert C.sizeof(_lib.Stats) == 4 * 8 + 6 * 4 + 8 * 4 + 8
    assert C.sizeof(_lib.SceneDesc) == 6 * 4 + 26 * 8 + 8


def test_image_width_is_reference_rounding():
    lib = _lib.load()
    flat, _ = load_golden("simple")
    cam = flat.camera_struct()
    for res, w in ((720, 1280), (1080, 1920), (90, 160), (1, 2), (7, 12)):
        assert lib.yrt_image_width(C.byref(cam), res) == w == flat.image_width(res)


def test_rows_owned_partition():
    lib = _lib.load()
    for h, tr, world in ((1080, 16, 8), (720, 16, 3), (37, 8, 4), (5, 16, 8)):
        assert sum(lib.yrt_rows_owned(h, tr, r, world) for r in range(world)) == h
    assert lib.yrt_rows_owned(100, 16, 8, 8) == 0      # bad rank


def test_scene_validation_errors_without_device():
    """Malformed descriptions are rejected by the host-side validation before any device work."""
    lib = _lib.load()
    flat = synth.mixed_scene(3, n_objects=2).flat()
    h = C.c_void_p()
    bad = flat.desc()
    bad.n_instances = -1
    assert lib.yrt_scene_create(C.byref(bad), C.byref(h)) == -1
    assert b"negative" in lib.yrt_last_error()
    flat2 = synth.mixed_scene(3, n_objects=2).flat()
    flat2.arrays["elem_idx"][0] = 10 ** 6
    assert lib.yrt_scene_create(C.byref(flat2.desc()), C.byref(h)) == -1
    assert b"out of range" in lib.yrt_last_error()
    flat3 = synth.mixed_scene(3, n_objects=2).flat()
    flat3.arrays["inst_mat"][0] = 99
    assert lib.yrt_scene_create(C.byref(flat3.desc()), C.byref(h)) == -1
    assert lib.yrt_scene_create(None, C.byref(h)) == -1


def test_no_cpu_fallback():
    """Without a CUDA device the compute entry points must fail loudly, never fall back."""
    lib = _lib.load()
    if lib.yrt_device_count() > 0:
        pytest.skip("a GPU is visible")
    assert lib.yrt_init(1) == -2
    assert b"no CPU path" in lib.yrt_last_error()
    flat = synth.mixed_scene(3, n_objects=2).flat()
    h = C.c_void_p()
    assert lib.yrt_scene_create(C.byref(flat.desc()), C.byref(h)) == -2
    assert not h.value
    out = np.zeros(16, np.uint8)
    assert lib.yrt_tonemap(C.c_void_p(np.zeros(16, np.float32).ctypes.data), 2, 2, C.c_void_p(out.ctypes.data)) == -2
    import yocto_raytracing_b200 as y
    with pytest.raises(y.YrtError):
        y.Scene(flat)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "yocto_raytracing_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "liboracle" not in txt and "yrt_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
