"""The device code (csrc/*.cuh: math, LBVH items, traversal, tie ranks, shading) compiled for the CPU
(tests/host_emu) against the reference's golden outputs — what can be checked here without a GPU."""
import numpy as np
import pytest

import _emu
from conftest import GOLDEN_CASES, RIGID_GOLDEN_CASES, id_match, ldr_stats, load_golden

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


def test_axis_parallel_rays_slab_superset():
    """Direction components that are exactly 0 (invd = +-inf): the slab test must still never cull a box the
    reference would enter.  Uses a camera looking straight down the -z axis so that central rays are exact."""
    from yocto_raytracing_b200 import synth
    sc = synth.mixed_scene(7)
    fr = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0.0, 1.0, 9.0], np.float32)      # identity rotation: d = (x, y, -focus)
    sc.camera = np.concatenate([fr, np.array([0.6, 16 / 9, 0, 9.0], np.float32)])
    flat = sc.flat()
    from oracle import oracle
    oracle.build()
    # odd width/height with 1 sample put (u, v) = (0.5, 0.5) on a pixel centre: d = (0, 0, -1) exactly there,
    # and whole rows / columns with one exactly-zero component
    w, h = 97, 55
    ids, dist, uv, ctr = _emu.EmuScene(flat).trace_primary(w, h, 1)
    rids, rdist, _ = oracle.OracleScene(flat).trace_primary(w, h, 1)
    assert ctr[4] == 0
    assert id_match(ids, rids) >= 0.9995


def test_generic_rays_incl_axis_aligned_vs_oracle(oracle_mod):
    """Same rays as the GPU test: random rays plus exactly axis-aligned ones (invd = +-inf paths)."""
    from yocto_raytracing_b200 import synth
    flat = synth.mixed_scene(21).flat()
    rng = np.random.RandomState(0)
    n = 20000
    o = rng.uniform(-6, 6, (n, 3)); o[:, 1] = rng.uniform(0.2, 6, n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.full((n, 1), 1e-4), rng.uniform(0.5, 30, (n, 1))], 1).astype(np.float32)
    rays[:100, 3:6] = [0, -1, 0]
    rays[100:200, 3:6] = [1, 0, 0]
    rays[200:300, 3:6] = [0, 0, -1]
    ids, dist, occ, false_rejects = _emu.EmuScene(flat).intersect(rays)
    oc = oracle_mod.OracleScene(flat)
    rids, rdist, _ = oc.intersect_first(rays)
    rocc = oc.intersect_any(rays)
    assert false_rejects == 0
    assert id_match(ids, rids) >= 0.9999
    assert id_match(ids[:300], rids[:300]) == 1.0
    assert (occ == rocc).mean() >= 0.9999


@pytest.mark.parametrize("name", ["instance10000", "mixed7", "lines_synth"])
@pytest.mark.parametrize("variant", ["bin", "wide"])
def test_other_arity_assignments_give_the_same_hits(name, variant):
    """The default build walks binary node records with closest-hit rays and 4-wide records with any-hit rays; the other
    assignments (build options: both binary = round 1's tree, both 4-wide) must return the same hits as the reference and
    as the default, never cull a box the reference would enter, and the 4-wide closest-hit walk must fetch about half as
    many node records as the binary one."""
    if not (_emu.available() and _emu.available(variant)):
        pytest.skip(f"host emulation ({variant} variant) not built")
    flat, ref = load_golden(name)
    w, h = int(ref["ids_width"]), int(ref["ids_height"])
    ids, dist, uv, ctr = _emu.EmuScene(flat, variant=variant).trace_primary(w, h, 1)
    ids0, dist0, uv0, ctr0 = _emu.EmuScene(flat).trace_primary(w, h, 1)
    assert id_match(ids, ref["ids"]) >= 0.9999
    assert np.array_equal(ids, ids0) and np.array_equal(dist.view(np.uint32), dist0.view(np.uint32))
    assert ctr[4] == 0 and ctr0[4] == 0
    if variant == "wide":
        assert ctr[7] < 0.65 * ctr0[7], (ctr[7], ctr0[7])
    # whole frames (shadow rays walk the other array in the default build): same image
    img, _ = _emu.EmuScene(flat, variant=variant).render(64, 36, 2, 0.1)
    img0, _ = _emu.EmuScene(flat).render(64, 36, 2, 0.1)
    assert np.array_equal(img.view(np.uint32), img0.view(np.uint32))


def test_stack_need_bounds_the_stack_actually_used():
    """stackneed_item (csrc/yrt_lbvh.cuh): the bound the build checks against YRT_STACK_CAP is never exceeded by a ray."""
    from yocto_raytracing_b200 import synth
    for flat in (load_golden("instance10000")[0], synth.hair_scene(2048, seed=3).flat(), synth.mixed_scene(5).flat()):
        for variant in ("", "bin", "wide"):
            if not _emu.available(variant):
                continue
            es = _emu.EmuScene(flat, variant=variant)
            ids, dist, uv, ctr = es.trace_primary(128, 72, 1)
            assert ctr[3] <= es.info()[6] <= 128, (ctr[3], es.info())


@pytest.mark.parametrize("case", ["instance10000_1080p", "lines_config4"])
def test_emulated_device_hit_ids_full_size(case):
    """The device traversal code (host build) on the two full-size goldens of the unmodified reference: the real
    instance10000 scene at 1920x1080 (2 073 600 primary rays) and the 1 048 576-segment lines config at 1280x720."""
    import os
    from conftest import GOLDEN
    from yocto_raytracing_b200 import synth
    flat = load_golden("instance10000")[0] if case == "instance10000_1080p" else synth.lines_config4().flat()
    with np.load(os.path.join(GOLDEN, case + ".ref.npz")) as z:
        inst, ei, dist, w, h = z["inst"], z["ei"], z["dist"], int(z["ids_width"]), int(z["ids_height"])
    es = _emu.EmuScene(flat)
    ids, d, uv, ctr = es.trace_primary(w, h, 1)
    same = (ids[:, 0] == inst) & (ids[:, 2] == ei)
    assert same.mean() >= 0.9999, (same.mean(), int((~same).sum()))
    assert np.array_equal(d[same].view(np.uint32), dist[same].view(np.uint32))
    assert ctr[4] == 0 and ctr[3] <= es.info()[6] <= 128


# ---- apex grids (csrc/yrt_pgrid.cuh): rays of the camera / towards a point light start at the root of their cell ----------
@pytest.mark.parametrize("name", RIGID_GOLDEN_CASES)     # (scenes with non-rigid frames have no grids)
def test_apex_grids_change_nothing(name):
    """Same hits (ids, distances, barycentrics, tie winners) and the same image, bit for bit, with and without the grids —
    i.e. every cell's candidate list holds all instances a ray of that cell can hit — and against the reference's goldens."""
    flat, ref = load_golden(name)
    w, h = int(ref["ids_width"]), int(ref["ids_height"])
    ih, iw = ref["image"].shape[:2]
    e = _emu.EmuScene(flat)
    ids0, d0, uv0, c0 = e.trace_primary(w, h, 1)
    img0, r0 = e.render(iw, ih, int(ref["image_samples"]), float(ref["ambient"]), max_depth=64)
    for light_R, shift in ((16, 3), (64, 1), (128, 5)):
        entries, fallback, n_grids, nodes = _emu.set_grids(e, light_R, shift)
        assert n_grids == min(e.info()[4], 8)                 # every point light of the golden scenes gets a grid
        ids1, d1, uv1, c1 = e.trace_primary(w, h, 1)
        img1, r1 = e.render(iw, ih, int(ref["image_samples"]), float(ref["ambient"]), max_depth=64)
        assert np.array_equal(ids0, ids1) and np.array_equal(d0, d1) and np.array_equal(uv0, uv1)
        assert np.array_equal(img0, img1) and r0[:3] == r1[:3]
        assert c1[4] == 0                                      # the slab audit holds on the chain nodes too
    assert id_match(ids1, ref["ids"]) >= 0.9999


def test_apex_grids_do_less_work():
    """instance10000: the box tests of the instance level shrink to a few per ray (what the grids are for)."""
    flat, ref = load_golden("instance10000")
    e = _emu.EmuScene(flat)
    w, h = 480, 270
    _, _, _, c0 = e.trace_primary(w, h, 1)
    _, r0 = e.render(w // 2, h // 2, 1)
    _emu.set_grids(e, 128, 3)
    _, _, _, c1 = e.trace_primary(w, h, 1)
    _, r1 = e.render(w // 2, h // 2, 1)
    n = w * h
    assert c0[6] / n > 30 and c1[6] / n < 12, (c0[6] / n, c1[6] / n)          # instance-level box tests per camera ray
    assert r0[4] / r0[2] > 18 and r1[4] / r1[2] < 6, (r0[4] / r0[2], r1[4] / r1[2])   # ... per shadow ray
    assert c1[0] < 0.75 * c0[0] and r1[3] < 0.65 * r0[3]                       # all box tests


def test_apex_grids_moving_camera_and_odd_sizes():
    """Camera grids for cameras inside / beside / far from the scene, odd image sizes and every cell size: same hits as the tree walk."""
    from yocto_raytracing_b200 import synth
    rng = np.random.default_rng(5)
    sc = synth.mixed_scene(7)
    for trial in range(6):
        eye = rng.uniform(-6, 6, 3).astype(np.float32)
        if trial == 0:
            eye = np.array([0.0, 1.0, 0.0], np.float32)       # inside the scene: boxes around and behind the pinhole
        at = rng.uniform(-1, 1, 3).astype(np.float32)
        z = (eye - at) / np.linalg.norm(eye - at)
        x = np.cross(np.array([0, 1, 0], np.float32), z); x /= np.linalg.norm(x)
        yv = np.cross(z, x)
        fr = np.concatenate([x, yv, z, eye]).astype(np.float32)
        sc.camera = np.concatenate([fr, np.array([rng.uniform(0.3, 1.2), 16 / 9, 0, rng.uniform(0.5, 9.0)], np.float32)])
        flat = sc.flat()
        e = _emu.EmuScene(flat)
        w, h = int(rng.integers(33, 160)), int(rng.integers(17, 90))
        ids0, d0, _, _ = e.trace_primary(w, h, 2)
        img0, _ = e.render(w, h, 1)
        for shift in (0, 2, 4):
            _emu.set_grids(e, 32, shift)
            ids1, d1, _, _ = e.trace_primary(w, h, 2)
            img1, _ = e.render(w, h, 1)
            assert np.array_equal(ids0, ids1) and np.array_equal(d0, d1), (trial, shift)
            assert np.array_equal(img0, img1), (trial, shift)


def test_apex_grids_skip_rotated_lights():
    """shade() aims at transform_point(frame, pos - p) (raytrace.cpp:129-130): only a light whose frame does not rotate
    converges its shadow rays on one point.  A rotated light must keep the instance tree — and the frame must not change."""
    flat, ref = load_golden("instance10000")
    a = {k: v.copy() for k, v in flat.arrays.items()}
    lights = flat.light_instances()
    fr = a["inst_frame"].reshape(-1, 12)
    c, s = np.float32(np.cos(0.3)), np.float32(np.sin(0.3))
    fr[lights[1], :9] = np.array([c, 0, s, 0, 1, 0, -s, 0, c], np.float32)        # rotation about y
    from yocto_raytracing_b200.scene import FlatScene
    flat2 = FlatScene._normalise(a)
    e = _emu.EmuScene(flat2)
    img0, r0 = e.render(160, 90, 1)
    entries, fallback, n_grids, nodes = _emu.set_grids(e, 64, 3)
    assert n_grids == len(lights) - 1
    img1, r1 = e.render(160, 90, 1)
    assert np.array_equal(img0, img1) and r0[:3] == r1[:3]


def test_apex_grids_without_room(monkeypatch):
    """Cells that find no room for their keys or chain nodes (and cells with too many candidates) start at the tree's root:
    tiny capacities must cost speed only, never a hit."""
    flat, ref = load_golden("instance10000")
    e = _emu.EmuScene(flat)
    ids0, d0, _, _ = e.trace_primary(160, 90, 1)
    img0, r0 = e.render(160, 90, 1)
    for env in ({"YRT_LIGHT_GRID_NODES": "500"}, {"YRT_LIGHT_GRID_KEYS": "3000"}, {"YRT_LIGHT_GRID_NODES": "1", "YRT_LIGHT_GRID_KEYS": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        entries, fallback, n_grids, nodes = _emu.set_grids(e, 64, 6)        # (64-pixel camera cells: many hold more than 16 candidates)
        for k in env:
            monkeypatch.delenv(k)
        assert fallback > 0, env
        ids1, d1, _, _ = e.trace_primary(160, 90, 1)
        img1, r1 = e.render(160, 90, 1)
        assert np.array_equal(ids0, ids1) and np.array_equal(d0, d1) and np.array_equal(img0, img1), env


def test_apex_grids_lights_anywhere():
    """Point lights inside the object layer, inside / at an instance, and very far away: frames with the light grids equal the
    tree walk's bit for bit (boxes around the light go into every cell; margins scale with the extent of the scene)."""
    from yocto_raytracing_b200 import synth
    from yocto_raytracing_b200.scene import FlatScene
    rng = np.random.default_rng(11)
    for trial in range(8):
        flat = synth.instance_grid_scene(12, seed=3 + trial).flat()
        a = {k: v.copy() for k, v in flat.arrays.items()}
        fr = a["inst_frame"].reshape(-1, 12)
        for li in flat.light_instances():
            mode = trial % 4
            if mode == 0:
                fr[li, 9:12] = rng.uniform(-10, 10, 3) * np.array([1, 0.15, 1]) + np.array([0, 1.0, 0])
            elif mode == 1:
                fr[li, 9:12] = fr[int(rng.integers(0, len(fr) - 4)), 9:12] + rng.uniform(-0.3, 0.3, 3)
            elif mode == 2:
                fr[li, 9:12] = rng.uniform(-1, 1, 3) * 1e4
            else:
                fr[li, 9:12] = rng.uniform(-20, 20, 3) + np.array([0, 25, 0])
        e = _emu.EmuScene(FlatScene._normalise(a))
        img0, r0 = e.render(96, 54, 1)
        _emu.set_grids(e, int(rng.choice([8, 64, 128])), int(rng.integers(0, 5)))
        img1, r1 = e.render(96, 54, 1)
        assert np.array_equal(img0, img1) and r0[:3] == r1[:3], trial


@pytest.mark.parametrize("seed,frame_seed,every,mirror_floor", [(31, 5, 3, False), (7, 1, 2, False), (101, 9, 1, False), (31, 5, 3, True)])
def test_nonrigid_frames_walk_the_reference_instance_tree(oracle_mod, monkeypatch, seed, frame_seed, every, mirror_floor):
    """Scaled / sheared instance frames: transform_ray_inverse (src/vmath.h:275-278) does not invert them, so what the reference
    returns depends on which instances its own tree lets a ray test and in which order (src/scene.cpp:446-479).  The device
    path walks a copy of that tree (RefTlas, trace_ray_ref) and must agree with the oracle — which is pinned to the
    unmodified reference on such a scene (test_oracle_vs_live_reference_on_fresh_scenes) — on every ray: ids, distances,
    barycentrics, shadow rays, mirror rays, the image.  The LBVH alone does not (second half)."""
    from yocto_raytracing_b200 import synth
    flat = synth.nonrigid_scene(seed, frame_seed, every, mirror_floor).flat()
    assert flat.nonrigid_instances() > 0
    w, h = 128, 72
    o = oracle_mod.OracleScene(flat)
    e = _emu.EmuScene(flat)
    rids, rdist, ruv = o.trace_primary(w, h, 1)
    ids, dist, uv, ctr = e.trace_primary(w, h, 1)
    assert np.array_equal(ids, rids) and np.array_equal(dist, rdist) and np.array_equal(uv, ruv) and ctr[4] == 0
    rimg, rc = o.render(w, h, 2, 0.1, threads=4)
    img, rays = e.render(w, h, 2, 0.1, max_depth=10 ** 6)
    assert np.array_equal(img.view(np.uint32), rimg.view(np.uint32))
    assert rays[:3] == [rc["primary_rays"], rc["reflection_rays"], rc["shadow_rays"]] and (rc["reflection_rays"] > 0) == mirror_floor
    rng = np.random.RandomState(seed)
    n = 8000
    ro = rng.uniform(-6, 6, (n, 3)); ro[:, 1] = rng.uniform(0.2, 6, n)
    rd = rng.normal(size=(n, 3)); rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    gen = np.concatenate([ro, rd, np.full((n, 1), 1e-4), rng.uniform(0.5, 30, (n, 1))], 1).astype(np.float32)
    gen[:50, 3:6] = [0, -1, 0]                       # exactly axis-parallel rays: the reference's slab formula inside the shapes too
    gids, gdist, gocc, false_rejects = e.intersect(gen)
    oids, odist, _ = o.intersect_first(gen)
    assert np.array_equal(gids, oids) and np.array_equal(gdist, odist) and np.array_equal(gocc, o.intersect_any(gen)) and false_rejects == 0
    # the same scene through the LBVH's own instance tree: different instances tested, different local distances kept
    monkeypatch.setenv("YRT_EMU_NO_REF_TLAS", "1")
    ids2, _, _, _ = _emu.EmuScene(flat).trace_primary(w, h, 1)
    assert not np.array_equal(ids2, rids)


@pytest.mark.parametrize("nonrigid", [False, True])
def test_instances_of_an_empty_shape(oracle_mod, nonrigid):
    """An instance whose shape has no elements: the reference keeps it in its instance tree with the world box of an invalid
    bbox (src/scene.cpp:558-562, bbox_to_world of +-FLT_MAX corners: infinities and NaNs in the partition and in the node
    boxes) and can never hit it.  Same frames with and without it on the device path — through the LBVH (which drops such
    instances) and through the copy of the reference's tree (which keeps them, like the reference)."""
    from yocto_raytracing_b200 import synth
    sc = synth.nonrigid_scene(31, 5, 3) if nonrigid else synth.mixed_scene(31, reflective_floor=False)
    F = np.float32
    sc.shapes.append(synth.Shape("void", 0, np.zeros((3, 3), F), np.tile(np.array([0, 0, 1], F), (3, 1)), np.zeros((0, 3), np.int32), "matte", np.zeros((3, 2), F)))
    sc.instances.append(("void", len(sc.shapes) - 1, synth.translation_frame((0.5, 1.0, 0.5))))
    sc.instances.insert(3, ("void2", len(sc.shapes) - 1, synth.translation_frame((-1.5, 0.3, 2.5))))
    flat = sc.flat()
    assert (flat.nonrigid_instances() > 0) == nonrigid
    w, h = 128, 72
    o, e = oracle_mod.OracleScene(flat), _emu.EmuScene(flat)
    rids, rdist, _ = o.trace_primary(w, h, 1)
    ids, dist, _, _ = e.trace_primary(w, h, 1)
    assert np.array_equal(ids, rids) and np.array_equal(dist, rdist)
    rimg, _ = o.render(w, h, 2, 0.1, threads=4)
    img, _ = e.render(w, h, 2, 0.1, max_depth=10 ** 6)
    assert np.array_equal(img.view(np.uint32), rimg.view(np.uint32))


def test_nonrigid_frames_in_a_large_instance_tree(oracle_mod):
    """The instance10000-shaped scene with every seventh instance scaled and sheared (1 429 of 10 004 frames): the copy of the
    reference's instance tree is 13 levels deep there and leaves hold up to four instances; every primary ray and every pixel
    must still equal the oracle's."""
    from yocto_raytracing_b200 import synth
    sc = synth.instance_grid_scene(100)
    rng = np.random.default_rng(3)
    for k, (iname, si, fr) in enumerate(list(sc.instances)):
        if k % 7 == 0 and k > 0 and not iname.startswith("light") and not iname.startswith("floor"):
            f = np.array(fr, np.float32).reshape(4, 3).copy()
            f[0] *= rng.uniform(0.6, 1.6); f[1] *= rng.uniform(0.6, 1.6); f[2] *= rng.uniform(0.6, 1.6)
            f[0] += 0.2 * f[2]
            sc.instances[k] = (iname, si, f.reshape(-1))
    flat = sc.flat()
    assert flat.nonrigid_instances() > 1000
    w, h = 160, 90
    o, e = oracle_mod.OracleScene(flat), _emu.EmuScene(flat)
    rids, rdist, _ = o.trace_primary(w, h, 1)
    ids, dist, _, ctr = e.trace_primary(w, h, 1)
    assert np.array_equal(ids, rids) and np.array_equal(dist, rdist) and ctr[4] == 0
    rimg, rc = o.render(w // 2, h // 2, 2, 0.1, threads=4)
    img, rays = e.render(w // 2, h // 2, 2, 0.1, max_depth=10 ** 6)
    assert np.array_equal(img.view(np.uint32), rimg.view(np.uint32)) and rays[2] == rc["shadow_rays"]


def test_randomized_scenes_and_cameras_vs_oracle(oracle_mod):
    """A sweep nobody hand-picked: 40 seeded scenes of every kind (mixed, non-rigid, instance grids, hair, untextured), random
    cameras — inside the scene, looking straight down — with and without apex grids.  Per scene: closest-hit ids on >= 99.99 %
    of the rays (bit-identical distances where the ids agree), no box culled that the reference would enter, float image
    identical on >= 99.9 % of the pixels.  (A 150-scene run of the same loop found one differing pixel: seed 1047, an
    exact-distance tie between two instances whose later-ranked candidate the reference's own tree culls before it is tested —
    the residual DESIGN.md section 1 describes; it is part of this sweep.)"""
    from yocto_raytracing_b200 import synth
    total = mismatched = 0
    for seed in list(range(1000, 1039)) + [1047]:
        rng = np.random.RandomState(seed)
        kind = seed % 5
        if kind == 0: sc = synth.mixed_scene(seed, n_objects=int(rng.randint(1, 40)))
        elif kind == 1: sc = synth.nonrigid_scene(seed, seed + 1, 1 + seed % 3, mirror_floor=bool(seed & 4))
        elif kind == 2: sc = synth.instance_grid_scene(int(rng.randint(2, 16)), seed=seed)
        elif kind == 3: sc = synth.hair_scene(int(rng.randint(8, 300)), seed=seed)
        else: sc = synth.mixed_scene(seed, reflective_floor=False, textured=False)
        eye = rng.uniform(-6, 6, 3); eye[1] = rng.uniform(0.05, 8)
        tgt = rng.uniform(-2, 2, 3); tgt[1] = rng.uniform(0, 2)
        if seed % 7 == 0: tgt = eye + np.array([0, -1.0, 0.0]) + 1e-3 * rng.normal(size=3)
        sc.camera = synth.make_camera(tuple(eye), tuple(tgt), float(rng.uniform(0.2, 1.2)))
        flat = sc.flat()
        w, h = 80, 45
        o, e = oracle_mod.OracleScene(flat), _emu.EmuScene(flat)
        if seed % 2: _emu.set_grids(e, 16 << (seed % 3), 1 + seed % 4)
        rids, rdist, _ = o.trace_primary(w, h, 2)
        ids, dist, _, ctr = e.trace_primary(w, h, 2)
        same = (ids == rids).all(axis=1)
        assert same.mean() >= 0.9999 and np.array_equal(dist[same], rdist[same]) and ctr[4] == 0, seed
        rimg, _ = o.render(w, h, 2, 0.1, threads=4, max_depth=40)
        img, _ = e.render(w, h, 2, 0.1, max_depth=40)
        assert (img == rimg).all(axis=2).mean() >= 0.999, seed
        total += same.size; mismatched += int((~same).sum())
    assert total == 40 * 80 * 45 * 4 and mismatched <= 3, mismatched      # (2 of 576 000 rays, both exact ties)

