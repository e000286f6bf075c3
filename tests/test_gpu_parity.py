"""Parity of the CUDA path (through the C ABI) against the reference: golden outputs of the unmodified
reference (tests/golden) and the C oracle on seeded synthetic scenes.  Run on the B200 box: pytest -m gpu.

Bars (BASELINE.json north_star): closest-hit (instance, shape, element) equal on >= 99.99 % of primary rays;
final RGBA8 after the reference tonemap within 1/255 per channel on >= 99.9 % of pixels.  Integer/index work
(sort, partition, gather) is bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN_CASES, id_match, ldr_stats, load_golden
from yocto_raytracing_b200 import FlatScene, _lib, synth

pytestmark = pytest.mark.gpu

ID_BAR = 0.9999
PIXEL_BAR = 0.999


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_hit_ids_vs_reference(gpu, name):
    flat, ref = load_golden(name)
    w, h = int(ref["ids_width"]), int(ref["ids_height"])
    with gpu.Scene(flat) as scn:
        ids, dist, uv = scn.trace_primary(w, h, 1)
    assert id_match(ids, ref["ids"]) >= ID_BAR
    same = (ids == ref["ids"]).all(axis=1)
    # same primitive => bit-identical distance and barycentrics (no FMA contraction on the device)
    assert np.array_equal(dist[same].view(np.uint32), ref["dist"][same].view(np.uint32))
    assert np.array_equal(uv[same].view(np.uint32), ref["uv"][same].view(np.uint32))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_image_vs_reference(gpu, oracle_mod, name):
    flat, ref = load_golden(name)
    h, w = ref["image"].shape[:2]
    with gpu.Scene(flat) as scn:
        img, st = scn.render(w, h, int(ref["image_samples"]), float(ref["ambient"]))
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), oracle_mod.tonemap(ref["image"]))
    assert within1 >= PIXEL_BAR, (within1, ident, mx)
    # float image: only the specular powf (CUDA vs glibc, a few ulp) may differ; mixed7 is a hall of mirrors (the
    # reference recurses 637 levels there, we stop at YRT_MAX_DEPTH = 64: what is cut carries less than 0.7^64 of its light)
    close = np.isclose(img, ref["image"], rtol=2e-5, atol=1e-6).all(axis=2).mean()
    assert close >= PIXEL_BAR, close
    assert (img[..., 3] == 1.0).all()
    assert st.primary_rays == w * h * int(ref["image_samples"]) ** 2


def test_ray_counts_match_oracle(gpu, oracle_mod):
    flat, _ = load_golden("refl")
    with gpu.Scene(flat) as scn:
        img, st = scn.render(96, 54, 2, 0.1)
    _, cnt = oracle_mod.OracleScene(flat).render(96, 54, 2, 0.1, max_depth=64)
    assert st.primary_rays == cnt["primary_rays"]
    assert abs(st.reflection_rays - cnt["reflection_rays"]) <= 2 and abs(st.shadow_rays - cnt["shadow_rays"]) <= 4
    assert st.max_depth == cnt["max_depth"]


@pytest.mark.parametrize("maker,res,smp", [(lambda: synth.instance_grid_scene(24, seed=3), 120, 2), (lambda: synth.hair_scene(512, seed=5), 120, 2),
                                           (lambda: synth.mixed_scene(11), 120, 3), (lambda: synth.mixed_scene(12, reflective_floor=False, textured=False), 64, 1)])
def test_synthetic_scenes_vs_oracle(gpu, oracle_mod, maker, res, smp):
    flat = maker().flat()
    w = flat.image_width(res)
    o = oracle_mod.OracleScene(flat)
    ref_ids, ref_dist, _ = o.trace_primary(w, res, 1)
    ref_img, cnt = o.render(w, res, smp, 0.1, max_depth=64, threads=8)
    with gpu.Scene(flat) as scn:
        ids, dist, _ = scn.trace_primary(w, res, 1)
        img, st = scn.render(w, res, smp, 0.1)
    assert id_match(ids, ref_ids) >= ID_BAR
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), oracle_mod.tonemap(ref_img))
    assert within1 >= PIXEL_BAR, (within1, ident, mx)


def test_generic_ray_queries_vs_oracle(gpu, oracle_mod):
    flat = synth.mixed_scene(21).flat()
    rng = np.random.RandomState(0)
    n = 20000
    o = rng.uniform(-6, 6, (n, 3)); o[:, 1] = rng.uniform(0.2, 6, n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.full((n, 1), 1e-4), rng.uniform(0.5, 30, (n, 1))], 1).astype(np.float32)
    rays[:100, 3:6] = [0, -1, 0]          # axis-aligned directions: invd = +-inf paths of the slab test
    rays[100:200, 3:6] = [1, 0, 0]
    oc = oracle_mod.OracleScene(flat)
    with gpu.Scene(flat) as scn:
        ids, dist, uv = scn.intersect_first(rays)
        occ = scn.intersect_any(rays)
    rids, rdist, ruv = oc.intersect_first(rays)
    rocc = oc.intersect_any(rays)
    assert id_match(ids, rids) >= ID_BAR
    same = (ids == rids).all(axis=1)
    assert np.array_equal(dist[same], rdist[same])
    assert (occ == rocc).mean() >= ID_BAR
    assert np.array_equal(occ != 0, ids[:, 0] >= 0) or ((occ != 0) == (ids[:, 0] >= 0)).mean() >= ID_BAR   # any-hit <=> closest-hit exists


@pytest.mark.parametrize("seed,frame_seed,every,mirror_floor", [(31, 5, 3, False), (101, 9, 1, False), (31, 5, 3, True)])
def test_nonrigid_instance_frames_vs_oracle(gpu, oracle_mod, seed, frame_seed, every, mirror_floor):
    """Scaled / sheared instance frames (OBJ `i` lines and glTF node transforms may carry them): transform_ray_inverse
    (src/vmath.h:275-278) does not invert them, so the reference's result depends on its own instance tree and visit order
    (src/scene.cpp:446-479).  The library traces such scenes through a copy of that tree (RefTlas, k_trace_*_ref) and is held
    to the same bars as every other scene; the oracle is pinned to the unmodified reference on such a scene
    (tests/test_oracle.py::test_oracle_vs_live_reference_on_fresh_scenes, nonrigid31).  The LBVH alone misses 3 - 8 % of the
    rays of these scenes (tests/test_host_emu.py)."""
    flat = synth.nonrigid_scene(seed, frame_seed, every, mirror_floor).flat()
    assert flat.nonrigid_instances() > 0
    w, h = 192, 108
    o = oracle_mod.OracleScene(flat)
    rids, rdist, ruv = o.trace_primary(w, h, 1)
    rimg, rc = o.render(w, h, 2, 0.1, threads=8)
    rng = np.random.RandomState(seed)
    n = 20000
    ro = rng.uniform(-6, 6, (n, 3)); ro[:, 1] = rng.uniform(0.2, 6, n)
    rd = rng.normal(size=(n, 3)); rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    rays = np.concatenate([ro, rd, np.full((n, 1), 1e-4), rng.uniform(0.5, 30, (n, 1))], 1).astype(np.float32)
    rays[:100, 3:6] = [0, -1, 0]
    with gpu.Scene(flat) as scn:
        ids, dist, uv = scn.trace_primary(w, h, 1)
        img, st = scn.render(w, h, 2, 0.1)
        img2, _ = scn.render(w, h, 2, 0.1, want_stats=False)      # two pipelines
        gids, gdist, _ = scn.intersect_first(rays)
        gocc = scn.intersect_any(rays)
    assert id_match(ids, rids) >= ID_BAR
    same = (ids == rids).all(axis=1)
    assert np.array_equal(dist[same], rdist[same]) and np.array_equal(uv[same], ruv[same])
    assert (st.primary_rays, st.reflection_rays, st.shadow_rays) == (rc["primary_rays"], rc["reflection_rays"], rc["shadow_rays"])
    assert (st.reflection_rays > 0) == mirror_floor
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), oracle_mod.tonemap(rimg))
    assert within1 >= PIXEL_BAR, (within1, ident, mx)
    assert np.array_equal(img.view(np.uint32), img2.view(np.uint32))
    oids, odist, _ = o.intersect_first(rays)
    assert id_match(gids, oids) >= ID_BAR
    gsame = (gids == oids).all(axis=1)
    assert np.array_equal(gdist[gsame], odist[gsame])
    assert (gocc == o.intersect_any(rays)).mean() >= ID_BAR


def test_radix_sort_is_stable_and_exact(gpu):
    lib = _lib.load()
    rng = np.random.RandomState(1)
    for n in (1, 2, 33, 1024, 1025, 100003):
        keys = rng.randint(0, 2 ** 62, n, dtype=np.int64).astype(np.uint64)
        keys[::3] = keys[0]               # many duplicates: stability matters
        if n > 10:
            keys[5:10] |= np.uint64(1) << np.uint64(63)
        vals = np.arange(n, dtype=np.int32)
        k2, v2 = keys.copy(), vals.copy()
        assert lib.yrt_debug_sort_pairs(C.c_void_p(k2.ctypes.data), C.c_void_p(v2.ctypes.data), n) == 0
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order].astype(np.int32))


def test_row_tiles_are_bit_identical_to_whole_frame(gpu):
    """Multi-GPU partition emulated on one GPU: every (rank, world) renders its interleaved tiles; assembled
    frames must equal the world=1 frame bit for bit (pixels are independent, per-pixel sum order is fixed)."""
    import torch
    from yocto_raytracing_b200 import distributed as D
    flat, _ = load_golden("instance10000")
    w, h, s = 160, 90, 2
    with gpu.Scene(flat) as scn:
        whole, _ = scn.render(w, h, s, 0.1)
        for world, tr in ((2, 16), (3, 8), (8, 16)):
            full = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
            for rank in range(world):
                own = D.rows_owned(h, tr, rank, world)
                if own == 0:
                    continue
                packed = torch.empty((own, w, 4), dtype=torch.float32, device="cuda")
                scn.render_rows_into(packed.data_ptr(), w, h, s, 0.1, tr, rank, world, torch.cuda.current_stream().cuda_stream, True)
                torch.cuda.synchronize()
                D._unpack(packed, full, w, h, tr, rank, world)
            torch.cuda.synchronize()
            assert np.array_equal(full.cpu().numpy().view(np.uint32), whole.view(np.uint32)), (world, tr)


def test_determinism_and_batching(gpu, monkeypatch):
    flat = synth.mixed_scene(11).flat()
    with gpu.Scene(flat) as scn:
        a, _ = scn.render(128, 72, 2, 0.1)
        b, _ = scn.render(128, 72, 2, 0.1)
        monkeypatch.setenv("YRT_BATCH_SLOTS", "5000")      # many small batches instead of one
        c, _ = scn.render(128, 72, 2, 0.1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))


def test_batch_size_and_grid_do_not_change_the_frame(gpu, monkeypatch):
    """Which warp traces which ray (batch size, resident CTAs per SM) must not change the frame, the hit ids or the ray
    counts; a reflective scene exercises the queue launches whose item counts live in device memory."""
    flat = synth.mixed_scene(11).flat()
    w, h, s = 131, 73, 3
    with gpu.Scene(flat) as scn:
        a, sa = scn.render(w, h, s, 0.1)
        ia, da, _ = scn.trace_primary(w, h, s)
        for env in ({"YRT_BATCH_SLOTS": "7000"}, {"YRT_BLOCKS_PER_SM": "1"}, {"YRT_BATCH_SLOTS": "333", "YRT_BLOCKS_PER_SM": "3"}):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            b, sb = scn.render(w, h, s, 0.1)
            ib, db, _ = scn.trace_primary(w, h, s)
            for k in env:
                monkeypatch.delenv(k)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), env
            assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32)), env
            assert (sa.primary_rays, sa.shadow_rays, sa.reflection_rays, sa.max_depth) == (sb.primary_rays, sb.shadow_rays, sb.reflection_rays, sb.max_depth), env


def test_edge_cases(gpu, oracle_mod):
    # empty scene: no shapes, no instances -> black, alpha 1
    sc = synth.SynthScene(name="empty")
    sc.camera = synth.make_camera((0, 1, 5), (0, 0, 0), 0.5)
    flat = sc.flat()
    with gpu.Scene(flat) as scn:
        img, st = scn.render(32, 18, 2, 0.1)
        ids, _, _ = scn.trace_primary(32, 18, 1)
    assert not img[..., :3].any() and (img[..., 3] == 1).all() and (ids == -1).all() and st.shadow_rays == 0
    # one triangle, no lights: ambient only; single-element shapes have no internal BVH node
    sc = synth.SynthScene(name="one")
    sc.materials.append(synth.Material("m", kd=(0.5, 0.25, 1.0)))
    sc.shapes.append(synth.Shape("t", 0, np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], np.float32), np.tile(np.array([0, 0, 1], np.float32), (3, 1)),
                                 np.array([[0, 1, 2]], np.int32), "m", np.zeros((3, 2), np.float32)))
    sc.instances.append(("t", 0, synth.translation_frame((0, 0, 0))))
    sc.camera = synth.make_camera((0, 0, 5), (0, 0, 0), 0.5)
    flat = sc.flat()
    ref, _ = oracle_mod.OracleScene(flat).render(64, 36, 2, 0.2)
    with gpu.Scene(flat) as scn:
        img, st = scn.render(64, 36, 2, 0.2)
    assert np.array_equal(img.view(np.uint32), ref.view(np.uint32)) and st.shadow_rays == 0
    assert img[18, 32, 0] == np.float32(0.2) * np.float32(0.5)
    # instances of a shape without elements (the reference keeps them in its instance tree with an infinite / NaN box and never
    # hits them), in a rigid scene (LBVH: dropped) and in a non-rigid one (copy of the reference's tree: kept)
    for sc in (synth.mixed_scene(31, reflective_floor=False), synth.nonrigid_scene(31, 5, 3)):
        F = np.float32
        sc.shapes.append(synth.Shape("void", 0, np.zeros((3, 3), F), np.tile(np.array([0, 0, 1], F), (3, 1)), np.zeros((0, 3), np.int32), "matte", np.zeros((3, 2), F)))
        sc.instances.append(("void", len(sc.shapes) - 1, synth.translation_frame((0.5, 1.0, 0.5))))
        sc.instances.insert(3, ("void2", len(sc.shapes) - 1, synth.translation_frame((-1.5, 0.3, 2.5))))
        flat = sc.flat()
        rids, rdist, _ = oracle_mod.OracleScene(flat).trace_primary(128, 72, 1)
        with gpu.Scene(flat) as scn:
            ids, dist, _ = scn.trace_primary(128, 72, 1)
        assert id_match(ids, rids) >= ID_BAR
        same = (ids == rids).all(axis=1)
        assert np.array_equal(dist[same], rdist[same])


def test_truncated_paths_are_reported_without_statistics(gpu, monkeypatch):
    """The recursion cap is this path's one deliberate difference from the reference's unbounded shade() recursion
    (src/raytrace.cpp:190-204): a frame rendered WITHOUT per-call statistics must still be able to say how many mirror
    bounces it dropped (yrt_frame_truncated_paths; the CLI's warning), and the count equals the one in yrt_stats."""
    flat, _ = load_golden("refl")             # a mirror floor under matte objects: exactly one bounce
    with gpu.Scene(flat) as scn:
        assert scn.truncated_paths() == 0                       # nothing rendered yet
        a, sa = scn.render(160, 90, 2, 0.1)
        assert sa.truncated_paths == 0 and scn.truncated_paths() == 0 and sa.max_depth == 2 and sa.reflection_rays > 0
        monkeypatch.setenv("YRT_MAX_DEPTH", "1")                # no bounce at all: every mirror ray of frame a is dropped
        b, sb = scn.render(160, 90, 2, 0.1)
        assert sb.truncated_paths == sa.reflection_rays and sb.reflection_rays == 0 and sb.max_depth == 1
        assert scn.truncated_paths() == sb.truncated_paths
        c, _ = scn.render(160, 90, 2, 0.1, want_stats=False)    # two pipelines, nothing read back by the call itself
        assert scn.truncated_paths() == sb.truncated_paths
        assert np.array_equal(b.view(np.uint32), c.view(np.uint32)) and not np.array_equal(a.view(np.uint32), b.view(np.uint32))
        monkeypatch.delenv("YRT_MAX_DEPTH")
        d, _ = scn.render(160, 90, 2, 0.1, want_stats=False)
        assert scn.truncated_paths() == 0 and np.array_equal(a.view(np.uint32), d.view(np.uint32))
    flat, _ = load_golden("mixed7")           # facing mirrors: the reference recurses hundreds of levels, the default cap is 64
    with gpu.Scene(flat) as scn:
        _, st = scn.render(160, 90, 2, 0.1)
        assert st.truncated_paths > 0 and st.max_depth == 64
        scn.render(160, 90, 2, 0.1, want_stats=False)
        assert scn.truncated_paths() == st.truncated_paths


def test_device_tonemap_vs_reference_tonemap(gpu, oracle_mod):
    _, ref = load_golden("simple")
    a = gpu.tonemap(ref["image"])
    b = oracle_mod.tonemap(ref["image"])
    within1, ident, mx = ldr_stats(a, b)
    assert within1 == 1.0 and ident >= 0.995      # CUDA powf vs glibc powf at a truncation boundary: <= 1 level


def test_in_process_api_matches_reference_signature(gpu):
    flat, ref = load_golden("basic")
    with gpu.Scene(flat) as scn:
        hdr = scn.raytrace(0.1, 90, 2)               # raytrace(scn, {amb,amb,amb}, resolution, samples)
        info = scn.info()
    assert hdr.shape == ref["image"].shape and info["lights"] == 2 and info["prims"] == flat.n_elements
    assert info["blas_depth"] + info["tlas_depth"] + 4 <= 128


def test_full_size_properties(gpu):
    """BASELINE size (1920x1080, 16 spp, 10 004 instances): properties that need no CPU reference —
    alpha = 1, finite, all primary rays counted, shadow rays = hits x lights, and the 1080p frame equals the
    frame assembled from 8 ranks' tiles."""
    import torch
    from yocto_raytracing_b200 import distributed as D
    flat = synth.instance_grid_scene(100).flat()
    w, h, s = 1920, 1080, 4
    with gpu.Scene(flat) as scn:
        img, st = scn.render(w, h, s, 0.1)
        assert st.primary_rays == w * h * 16 and st.shadow_rays % 3 == 0 and st.shadow_rays <= 3 * st.primary_rays
        assert st.shadow_rays >= 0.99 * 3 * st.primary_rays      # the floor fills the view
        assert np.isfinite(img).all() and (img[..., 3] == 1).all() and img[..., :3].min() >= 0
        full = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        for rank in range(8):
            own = D.rows_owned(h, 16, rank, 8)
            packed = torch.empty((own, w, 4), dtype=torch.float32, device="cuda")
            scn.render_rows_into(packed.data_ptr(), w, h, s, 0.1, 16, rank, 8, torch.cuda.current_stream().cuda_stream, True)
            torch.cuda.synchronize()
            D._unpack(packed, full, w, h, 16, rank, 8)
        torch.cuda.synchronize()
        assert np.array_equal(full.cpu().numpy().view(np.uint32), img.view(np.uint32))


def test_drop_in_cli_matches_reference_cli(gpu, tmp_path):
    """bin/raytrace (the reference's main() with build_bvh + raytrace swapped for the C ABI, same loader, same
    PNG writer, same flags) against the unmodified reference binary on the same OBJ: PNG pixels within 1/255 on
    >= 99.9 % of pixels.  Both binaries link the reference's loader, so they exist only where it was built."""
    import os
    import subprocess
    from PIL import Image
    from conftest import ROOT
    ours, ref = os.path.join(ROOT, "bin", "raytrace"), os.path.join(ROOT, "oracle", "_ref", "raytrace_ref")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("bin/raytrace / oracle/_ref/raytrace_ref not built (need the reference sources at build time)")
    # (nonrigid31: scaled + sheared `i` lines — neither binary refuses it, and the PNGs agree like any other scene's)
    for sc, res, smp in ((synth.instance_grid_scene(20, seed=9), 180, 2), (synth.mixed_scene(7), 120, 2), (synth.nonrigid_scene(31, 5, 3), 120, 2), (synth.hair_scene(256), 120, 2)):
        obj = sc.write_obj(str(tmp_path / sc.name))
        cwd = os.path.dirname(obj)
        a, b = os.path.join(cwd, "ours.png"), os.path.join(cwd, "ref.png")
        args = ["-r", str(res), "-s", str(smp), "-a", "0.1"]
        r1 = subprocess.run([ours] + args + ["-o", a, os.path.basename(obj)], cwd=cwd, capture_output=True, text=True)
        assert r1.returncode == 0, r1.stdout + r1.stderr
        # same four progress lines as the reference (src/raytrace.cpp:273-285)
        assert [l.split()[0] for l in r1.stdout.strip().splitlines()[:4]] == ["loading", "creating", "tracing", "saving"]
        subprocess.run([ref] + args + ["-o", b, os.path.basename(obj)], cwd=cwd, check=True, capture_output=True)
        ia, ib = np.array(Image.open(a)), np.array(Image.open(b))
        assert ia.shape == ib.shape
        within1, ident, mx = ldr_stats(ia, ib)
        assert within1 >= PIXEL_BAR, (sc.name, within1, ident, mx)
    # additive flag --gpus: same PNG from 2 GPUs (interleaved rows, peer stores into GPU 0's frame) as from 1
    if gpu.device_count() >= 2:
        c = os.path.join(cwd, "ours2.png")
        r2 = subprocess.run([ours] + args + ["--gpus", "2", "--stats", "-o", c, os.path.basename(obj)], cwd=cwd, capture_output=True, text=True)
        assert r2.returncode == 0, r2.stdout + r2.stderr
        assert "on 2 GPU(s)" in r2.stdout
        assert np.array_equal(np.array(Image.open(c)), ia)
    # unknown option: usage + non-zero exit like yu::cmdline (src/ext/yocto_utils.h:1157-1174)
    assert subprocess.run([ours, "--bogus", "x.obj"], capture_output=True).returncode != 0


def test_cli_scene_cache_and_device_ldr(gpu, tmp_path):
    """SURVEY 8f.1 / 8f.2 through the CLI: --device-ldr (tonemap on the GPU, RGBA8 out) and --cache (flattened scene
    <scene>.yrts written on the first run, read instead of the OBJ on the second).  The PNG of both runs must agree with
    the unmodified reference binary's within 1/255 on >= 99.9 % of pixels, and the cached run must be identical to the
    uncached one."""
    import os
    import subprocess
    from PIL import Image
    from conftest import ROOT
    ours, ref = os.path.join(ROOT, "bin", "raytrace"), os.path.join(ROOT, "oracle", "_ref", "raytrace_ref")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("bin/raytrace / oracle/_ref/raytrace_ref not built (need the reference sources at build time)")
    sc = synth.mixed_scene(7)
    obj = sc.write_obj(str(tmp_path / sc.name))
    cwd, name = os.path.dirname(obj), os.path.basename(obj)
    args = ["-r", "120", "-s", "2", "-a", "0.1"]
    outs = []
    for i in range(2):
        out = os.path.join(cwd, f"ours{i}.png")
        r = subprocess.run([ours] + args + ["--cache", "--device-ldr", "--stats", "-o", out, name], cwd=cwd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert [l.split()[0] for l in r.stdout.strip().splitlines()[:4]] == ["loading", "creating", "tracing", "saving"]
        assert ("(scene cache)" in r.stdout) == (i == 1), r.stdout
        outs.append(np.array(Image.open(out)))
    assert os.path.exists(obj + ".yrts")
    assert np.array_equal(outs[0], outs[1])
    # the cache file is what FlatScene.load reads: same scene through Python
    flat = FlatScene.load(obj + ".yrts")
    assert flat.arrays["inst_shape"].size == sc.flat().arrays["inst_shape"].size
    b = os.path.join(cwd, "ref.png")
    subprocess.run([ref] + args + ["-o", b, name], cwd=cwd, check=True, capture_output=True)
    within1, ident, mx = ldr_stats(outs[0], np.array(Image.open(b)))
    assert within1 >= PIXEL_BAR, (within1, ident, mx)
    # a stale cache (older than the OBJ) is ignored and rewritten
    os.utime(obj + ".yrts", (1, 1))
    r = subprocess.run([ours] + args + ["--cache", "--stats", "-o", os.path.join(cwd, "ours3.png"), name], cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0 and "(scene cache)" not in r.stdout
    assert os.stat(obj + ".yrts").st_mtime > 1
    # ... and so is a cache older than the material library or a texture next to the scene
    import time
    for dep in (obj[:-4] + ".mtl", os.path.join(cwd, "grid.png")):
        r = subprocess.run([ours] + args + ["--cache", "--stats", "-o", os.path.join(cwd, "ours3.png"), name], cwd=cwd, capture_output=True, text=True)
        assert r.returncode == 0 and "(scene cache)" in r.stdout          # fresh now
        future = max(time.time(), os.stat(obj + ".yrts").st_mtime) + 5      # (whole seconds later than the cache, whatever the runs took)
        os.utime(dep, (future, future))
        r = subprocess.run([ours] + args + ["--cache", "--stats", "-o", os.path.join(cwd, "ours3.png"), name], cwd=cwd, capture_output=True, text=True)
        assert r.returncode == 0 and "(scene cache)" not in r.stdout, dep
        os.utime(obj + ".yrts", (future + 1, future + 1))
    # .hdr output keeps the float path even with --device-ldr
    r = subprocess.run([ours] + args + ["--device-ldr", "-o", os.path.join(cwd, "ours.hdr"), name], cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0 and os.path.getsize(os.path.join(cwd, "ours.hdr")) > 0
    # --fast-png (parallel encoder instead of stb_image_write): the decoded pixels are those of the stb-written files,
    # with the host tonemap (== ours3.png, same float frame) and with the device tonemap (== ours0.png)
    for extra, same_as in (([], np.array(Image.open(os.path.join(cwd, "ours3.png")))), (["--device-ldr"], outs[0])):
        out = os.path.join(cwd, "fast.png")
        r = subprocess.run([ours] + args + ["--fast-png", "--stats"] + extra + ["-o", out, name], cwd=cwd, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert np.array_equal(np.array(Image.open(out)), same_as), extra


def test_fused_gather_into_one_frame(gpu):
    """yrt_render_rows_into_frame: every rank's resolve stores its rows at their final position of ONE full frame
    (what the ranks do over NVLink peer memory); emulated here by rendering all ranks of several partitions into
    one buffer.  Must equal the whole-frame render bit for bit."""
    import ctypes as C
    import torch
    lib = _lib.load()
    flat, _ = load_golden("instance10000")
    w, h, s = 160, 90, 2
    with gpu.Scene(flat) as scn:
        whole, _ = scn.render(w, h, s, 0.1)
        for world, tr in ((2, 1), (8, 1), (3, 7)):
            ptr = C.c_void_p()
            assert lib.yrt_frame_alloc(w, h, C.byref(ptr)) == 0
            for rank in range(world):
                scn.render_rows_into_frame(ptr.value, w, h, s, 0.1, tr, rank, world, 0, rank == world - 1)
            torch.cuda.synchronize()
            iface = {"shape": (h, w, 4), "typestr": "<f4", "data": (ptr.value, False), "version": 3, "strides": None}
            t = torch.as_tensor(type("F", (), {"__cuda_array_interface__": iface})(), device="cuda")
            got = t.cpu().numpy()
            assert lib.yrt_frame_free(ptr) == 0
            assert np.array_equal(got.view(np.uint32), whole.view(np.uint32)), (world, tr)


def test_large_scene_and_deep_trees(gpu, oracle_mod):
    """90 000 instances (300 x 300 grid): the LBVH depth check, the radix sort at > 64 K keys and the tie ranks at scale."""
    flat = synth.instance_grid_scene(300, seed=5).flat()
    w, h = 160, 90
    with gpu.Scene(flat) as scn:
        info = scn.info()
        ids, dist, _ = scn.trace_primary(w, h, 1)
    assert info["tlas_nodes"] == flat.n_instances - 1 and info["blas_depth"] + info["tlas_depth"] + 4 <= 128
    rids, rdist, _ = oracle_mod.OracleScene(flat).trace_primary(w, h, 1)
    assert id_match(ids, rids) >= ID_BAR
    same = (ids == rids).all(axis=1)
    assert np.array_equal(dist[same], rdist[same])


def test_in_process_multi_gpu_matches_single(gpu):
    """yrt_init(n) + yrt_render / yrt_render_ldr: rows interleaved over the GPUs of this process, every GPU copies its own
    rows into the caller's host frame; bit-identical to one GPU, for the float frame, the RGBA8 frame, a moving camera and a
    frame height that the GPU count does not divide."""
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs in this process")
    flat, _ = load_golden("instance10000")
    cams = [flat.arrays["camera"].copy() for _ in range(3)]
    cams[1][9] += 3.0
    cams[2][10] -= 5.0
    sizes = [(320, 180), (322, 181), (64, 3)]
    one = []
    with gpu.Scene(flat) as scn:
        for cam in cams:
            scn.set_camera(cam)
            for (w, h) in sizes:
                one.append((scn.render(w, h, 2, 0.1)[0].copy(), scn.render_ldr(w, h, 2, 0.1)[0].copy()))
    try:
        gpu.init(min(n, 4))
        with gpu.Scene(flat) as scn:
            k = 0
            for cam in cams:
                scn.set_camera(cam)
                for (w, h) in sizes:
                    many, st = scn.render(w, h, 2, 0.1)
                    ldr, _ = scn.render_ldr(w, h, 2, 0.1)
                    assert st.n_gpus == min(n, 4)
                    assert np.array_equal(one[k][0].view(np.uint32), many.view(np.uint32)), (k, w, h)
                    assert np.array_equal(one[k][1], ldr), (k, w, h)
                    k += 1
    finally:
        gpu.init(1)


def test_render_ldr_device_tonemap(gpu, oracle_mod):
    """yrt_render_ldr: float frame + device tonemap vs the reference float image through the host tonemap."""
    flat, ref = load_golden("simple")
    h, w = ref["image"].shape[:2]
    with gpu.Scene(flat) as scn:
        ldr, _ = scn.render_ldr(w, h, int(ref["image_samples"]), float(ref["ambient"]))
    within1, ident, mx = ldr_stats(ldr, oracle_mod.tonemap(ref["image"]))
    assert within1 >= PIXEL_BAR and ident >= 0.99, (within1, ident, mx)
    assert (ldr[..., 3] == 255).all()


def test_many_samples_per_pixel(gpu, oracle_mod):
    """-s 8 (64 spp): slots per pixel exceed a warp; ordered per-pixel sum must still match the oracle bit for bit
    on a scene without specular powf (basic has Ks 0.8 -> compare tonemapped)."""
    flat, _ = load_golden("basic")
    w, h, s = 48, 27, 8
    ref, _ = oracle_mod.OracleScene(flat).render(w, h, s, 0.1, threads=8)
    with gpu.Scene(flat) as scn:
        img, st = scn.render(w, h, s, 0.1)
    assert st.primary_rays == w * h * 64
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), oracle_mod.tonemap(ref))
    assert within1 >= PIXEL_BAR


@pytest.mark.parametrize("name", ["simple", "basic", "refl", "instance10000"])
def test_run_sh_configs_full_size(gpu, oracle_mod, name):
    """The reference's own configs (run.sh: -r 720 -s 3 -> 1280x720, 9 spp) on the reference's scenes, at full size:
    GPU image vs the C oracle (pinned bit-exactly to the reference) — pixels within 1/255 on >= 99.9 %, ray counts equal."""
    import os
    flat, _ = load_golden(name)
    w, h, s = flat.image_width(720), 720, 3
    assert w == 1280
    ref, cnt = oracle_mod.OracleScene(flat).render(w, h, s, 0.1, max_depth=64, threads=os.cpu_count() or 4)
    with gpu.Scene(flat) as scn:
        img, st = scn.render(w, h, s, 0.1)
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), oracle_mod.tonemap(ref))
    assert within1 >= PIXEL_BAR, (name, within1, ident, mx)
    assert st.primary_rays == cnt["primary_rays"] == w * h * 9
    assert abs(st.shadow_rays - cnt["shadow_rays"]) <= 1e-5 * cnt["shadow_rays"] + 4
    assert abs(st.reflection_rays - cnt["reflection_rays"]) <= 1e-5 * cnt["reflection_rays"] + 4


def _full_golden(name):
    import os
    from conftest import GOLDEN
    with np.load(os.path.join(GOLDEN, name + ".ref.npz")) as z:
        return {k: z[k] for k in z.files}


def test_target_config_on_the_real_scene_1080p_16spp(gpu, oracle_mod):
    """BASELINE target: the reference's own in/instance10000_pointlight at 1920x1080.  (a) closest-hit (instance, element) of
    all 2 073 600 primary rays at 1 spp against ref_probe's dump, distances bit-identical on equal ids; (b) the 16 spp frame,
    through the reference's tonemap, against the PNG the unmodified reference CLI wrote (`-r 1080 -s 4`; SURVEY 8c digest
    22bc1ac0e6f98ba9): >= 99.9 % of pixels within 1/255.  Goldens: tools/make_golden_full.py."""
    import hashlib
    import os
    from PIL import Image
    from conftest import GOLDEN
    flat, _ = load_golden("instance10000")
    ref = _full_golden("instance10000_1080p")
    w, h = int(ref["ids_width"]), int(ref["ids_height"])
    assert (w, h) == (1920, 1080) == (flat.image_width(1080), 1080)
    png = np.array(Image.open(os.path.join(GOLDEN, "instance10000_1080p_s4.png")))
    assert hashlib.sha256(png.tobytes()).hexdigest()[:16] == "22bc1ac0e6f98ba9"
    with gpu.Scene(flat) as scn:
        ids, dist, _ = scn.trace_primary(w, h, 1)
        img, st = scn.render(w, h, 4, 0.1)
        ldr_dev, _ = scn.render_ldr(w, h, 4, 0.1)
    same = (ids[:, 0] == ref["inst"]) & (ids[:, 2] == ref["ei"])
    assert same.mean() >= ID_BAR, (same.mean(), int((~same).sum()))
    assert np.array_equal(dist[same].view(np.uint32), ref["dist"][same].view(np.uint32))
    assert st.primary_rays == w * h * 16 and st.reflection_rays == 0 and st.truncated_paths == 0
    assert st.total_rays == 132710394          # 33 177 600 primary + 99 532 794 shadow rays per frame
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), png)
    assert within1 >= PIXEL_BAR, (within1, ident, mx)
    assert ident >= 0.999, (within1, ident, mx)
    within1d, identd, mxd = ldr_stats(ldr_dev, png)      # device tonemap (yrt_render_ldr) against the same file
    assert within1d >= PIXEL_BAR, (within1d, identd, mxd)


def test_lines_config4_full_size(gpu, oracle_mod):
    """BASELINE configs[3] at the size SURVEY 8d specifies: 2 x 65 536 hairs x 8 segments (1 048 576 line elements in two
    bottom-level trees), 1280x720: ids of every primary ray at 1 spp and the 9 spp frame against the unmodified reference
    (ref_probe dump / CLI PNG of the same generated scene, tools/make_golden_full.py)."""
    import os
    from PIL import Image
    from conftest import GOLDEN
    flat = synth.lines_config4().flat()
    ref = _full_golden("lines_config4")
    w, h = int(ref["ids_width"]), int(ref["ids_height"])
    assert (w, h) == (1280, 720) and flat.n_elements > 1048576
    png = np.array(Image.open(os.path.join(GOLDEN, "lines_config4_720p_s3.png")))
    with gpu.Scene(flat) as scn:
        info = scn.info()
        ids, dist, _ = scn.trace_primary(w, h, 1)
        img, st = scn.render(w, h, 3, 0.1)
    same = (ids[:, 0] == ref["inst"]) & (ids[:, 2] == ref["ei"])
    assert same.mean() >= ID_BAR, (same.mean(), int((~same).sum()), info)
    assert np.array_equal(dist[same].view(np.uint32), ref["dist"][same].view(np.uint32))
    within1, ident, mx = ldr_stats(oracle_mod.tonemap(img), png)
    assert within1 >= PIXEL_BAR, (within1, ident, mx)
    assert st.primary_rays == w * h * 9 and st.shadow_rays % 2 == 0


def test_torchrun_moving_camera_frames_match_single_gpu(gpu):
    """One process per GPU (torchrun, NCCL), a different camera every frame, both multi-GPU frame paths (peer stores into
    rank 0's device frame; per-rank copies into a shared host frame): every frame bit-identical to the single-GPU render.
    A missing barrier (frame k+1 overtaking the consumer of frame k) shows up as a torn frame here."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(ROOT, "tests", "_mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "MGPU_OK" in r.stdout, r.stdout[-2000:]


def test_rows_to_host_assembles_the_frame(gpu):
    """yrt_render_rows_to_host: every rank's pitched device->host copy lands its rows at their final positions of ONE host
    frame (what the ranks do into shared memory); all ranks of several partitions, ragged last tiles included, emulated on
    one GPU.  Must equal the whole-frame render bit for bit."""
    import torch
    flat, _ = load_golden("instance10000")
    w, h, s = 160, 91, 2
    with gpu.Scene(flat) as scn:
        whole, _ = scn.render(w, h, s, 0.1)
        for world, tr in ((2, 1), (8, 1), (3, 7), (4, 16), (2, 91), (5, 100)):
            host = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
            for rank in range(world):
                scn.render_rows_to_host(host.data_ptr(), w, h, s, 0.1, tr, rank, world, 0, rank == world - 1)
            torch.cuda.synchronize()
            assert np.array_equal(host.numpy().view(np.uint32), whole.view(np.uint32)), (world, tr)


@pytest.mark.parametrize("maker", [lambda: load_golden("instance10000")[0], lambda: synth.mixed_scene(7).flat(), lambda: synth.hair_scene(4096, seed=2).flat(),
                                   lambda: synth.instance_grid_scene(60, seed=4).flat()])
def test_device_lbvh_equals_serial_execution_bit_for_bit(gpu, maker):
    """Race detector of the GPU build (compute-sanitizer is not available on this pool): Morton keys, the multi-CTA stable
    radix sort, Karras topology and the lock-free bottom-up passes (refit, tree rotations, re-layout, stack need — arrival
    counters + fences) are deterministic, so the node arrays the device builds must equal, bit for bit, those of a SERIAL
    execution of the same per-item functions on the host (tests/host_emu) — for both node arities, on every one of 10
    builds of the same scene (a stale read of a child's box shows up as a different box)."""
    import _emu
    if not _emu.available():
        pytest.skip("host emulation not built")
    flat = maker()
    es = _emu.EmuScene(flat)
    want = {a: es.nodes(a) for a in (2, 4)}
    for rep in range(10):
        with gpu.Scene(flat) as scn:
            info = scn.info()
            for a in (2, 4):
                got = scn.debug_nodes(a)
                assert got.shape == want[a].shape, (a, got.shape, want[a].shape, info)
                same = (got.view(np.uint32) == want[a].view(np.uint32)).all(axis=1)
                assert same.all(), (rep, a, int((~same).sum()), np.flatnonzero(~same)[:8])


@pytest.mark.parametrize("maker,w,h,s", [(lambda: load_golden("instance10000")[0], 640, 360, 2), (lambda: synth.mixed_scene(11).flat(), 131, 73, 3),
                                         (lambda: load_golden("refl")[0], 320, 180, 2), (lambda: synth.hair_scene(512, seed=5).flat(), 160, 90, 2)])
def test_apex_grids_do_not_change_the_frame(gpu, monkeypatch, maker, w, h, s):
    """The apex grids (csrc/yrt_pgrid.cuh: camera rays / shadow rays start at the root of their cell's short chain of
    instance-level nodes instead of the instance tree's root) must not change a single bit: frame, hit ids, distances and
    ray counts with the grids (default), without them, and with other cell sizes."""
    flat = maker()
    monkeypatch.setenv("YRT_PGRID_MIN_INSTANCES", "1")         # grids on small scenes too
    with gpu.Scene(flat) as scn:
        a, sa = scn.render(w, h, s, 0.1)
        ia, da, _ = scn.trace_primary(w, h, s)
    for env in ({"YRT_PGRID": "0"}, {"YRT_CAM_CELL_SHIFT": "0", "YRT_LIGHT_GRID_R": "16"}, {"YRT_CAM_CELL_SHIFT": "5", "YRT_LIGHT_GRID_R": "256"},
                {"YRT_PGRID_MIN_INSTANCES": "1000000"}, {"YRT_LIGHT_GRID_R": "8", "YRT_CAM_CELL_SHIFT": "8"},
                {"YRT_LIGHT_GRID_NODES": "300", "YRT_LIGHT_GRID_KEYS": "2000"}, {"YRT_GRID_STREAM": "0", "YRT_STREAMS": "3"}):   # no room for most cells' chains; grid on the frame's stream
        monkeypatch.setenv("YRT_PGRID_MIN_INSTANCES", "1")
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with gpu.Scene(flat) as scn:
            b, sb = scn.render(w, h, s, 0.1)
            ib, db, _ = scn.trace_primary(w, h, s)
        for k in env:
            monkeypatch.delenv(k)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), env
        assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32)), env
        assert (sa.primary_rays, sa.shadow_rays, sa.reflection_rays, sa.max_depth) == (sb.primary_rays, sb.shadow_rays, sb.reflection_rays, sb.max_depth), env
