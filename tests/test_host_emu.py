"""The device code (csrc/*.cuh: math, LBVH items, traversal, tie ranks, shading) compiled for the CPU
(tests/host_emu) against the reference's golden outputs — what can be checked here without a GPU."""
import numpy as np
import pytest

import _emu
from conftest import GOLDEN_CASES, id_match, ldr_stats, load_golden

pytestmark = pytest.mark.skipif(not _emu.available(), reason="tests/host_emu/libyrt_hostemu.so not built (make hostemu)")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_device_hit_ids(name):
    flat, ref = load_golden(name)
    w, h = int(ref["ids_width"]), int(ref["ids_height"])
    ids, dist, uv, ctr = _emu.EmuScene(flat).trace_primary(w, h, 1)
    assert id_match(ids, ref["ids"]) >= 0.9999          # north_star: >= 99.99 % of primary rays
    same = (ids == ref["ids"]).all(axis=1)
    assert np.array_equal(dist[same], ref["dist"][same]) and np.array_equal(uv[same], ref["uv"][same])
    assert ctr[3] < 64
    # the fused (FFMA) slab test must accept a superset of the reference's slab test (scene.cpp:371-383)
    assert ctr[4] == 0, f"{ctr[4]} boxes culled that the reference would enter"
    assert ctr[5] <= 1e-3 * ctr[0]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_emulated_device_image(oracle_mod, name):
    flat, ref = load_golden(name)
    h, w = ref["image"].shape[:2]
    img, rays = _emu.EmuScene(flat).render(w, h, int(ref["image_samples"]), float(ref["ambient"]), max_depth=10 ** 6)
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), oracle_mod.tonemap(ref["image"]))
    assert within1 >= 0.999
    # on the host the same libm is used, so apart from exact-tie/grazing rays the floats are identical
    assert (img == ref["image"]).all(axis=2).mean() >= 0.999


@pytest.mark.parametrize("leaf_blas,leaf_tlas", [(1, 1), (8, 4), (2, 8)])
def test_result_independent_of_tree_shape(leaf_blas, leaf_tlas):
    """Different leaf sizes = different trees and visit orders; hits (incl. tie winners) must not change."""
    flat, ref = load_golden("instance10000")
    ids0, d0, _, _ = _emu.EmuScene(flat).trace_primary(96, 54, 1)
    ids1, d1, _, _ = _emu.EmuScene(flat, leaf_blas, leaf_tlas).trace_primary(96, 54, 1)
    assert np.array_equal(ids0, ids1) and np.array_equal(d0, d1)
