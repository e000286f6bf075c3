"""Host-side logic: scene container, synthetic generators, row-tile partition, and the world_size-2 gather
(gloo, CPU) of the one-process-per-GPU path."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from conftest import ROOT, load_golden
from yocto_raytracing_b200 import distributed as D
from yocto_raytracing_b200 import synth
from yocto_raytracing_b200.scene import FlatScene


def test_flat_scene_npz_roundtrip(tmp_path):
    flat = synth.mixed_scene(5).flat()
    p = str(tmp_path / "s.npz")
    flat.save_npz(p)
    back = FlatScene.load(p)
    for k, v in flat.arrays.items():
        assert np.array_equal(v, back.arrays[k]), k
    assert back.n_shapes == 9 and back.n_instances == flat.n_instances


def test_lights_are_all_positive_ke_instances():
    flat = synth.mixed_scene(5).flat()
    lights = flat.light_instances()
    # 'half_emitter' has ke = (5,0,5): AND of the three channels (raytrace.cpp:126) -> not a light
    assert len(lights) == 2
    flat2, _ = load_golden("instance10000")
    assert len(flat2.light_instances()) == 3 and flat2.n_instances == 10004 and flat2.n_elements == 41987


def test_instance_grid_has_named_shape():
    sc = synth.instance_grid_scene(100)
    flat = sc.flat()
    assert flat.n_instances == 10004 and flat.n_shapes == 14
    tri = flat.arrays["shape_elem_cnt"][:11]
    assert tri[0] == 8192 and all(t in (3072, 4096) for t in tri[1:])
    assert flat.image_width(1080) == 1920


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "bin", "yrt_flatten")), reason="needs the reference loader (built only where /root/reference exists)")
def test_direct_flat_equals_reference_loader_on_written_obj(tmp_path):
    """The OBJ our generator writes, loaded by the reference's own loader, gives the arrays we build directly."""
    for sc in (synth.instance_grid_scene(6), synth.hair_scene(64), synth.mixed_scene(7)):
        obj = sc.write_obj(str(tmp_path / sc.name))
        y = obj[:-4] + ".yrts"
        subprocess.run([os.path.join(ROOT, "bin", "yrt_flatten"), os.path.basename(obj), y], check=True, cwd=os.path.dirname(obj),
                       stdout=subprocess.DEVNULL)
        a, b = sc.flat(), FlatScene.load(y)
        for k in a.arrays:
            if k == "uv":   # the loader computes 1-(1-v): one rounding
                assert np.allclose(a.arrays[k], b.arrays[k], atol=1e-6)
            else:
                assert np.array_equal(a.arrays[k], b.arrays[k]), (sc.name, k)


@pytest.mark.parametrize("h,tr,world", [(1080, 16, 8), (720, 16, 3), (37, 8, 4), (5, 16, 8), (90, 16, 1)])
def test_row_tile_partition_is_a_partition(h, tr, world):
    seen = np.concatenate([D.global_rows(h, tr, r, world) for r in range(world)])
    assert sorted(seen.tolist()) == list(range(h))
    for r in range(world):
        assert D.rows_owned(h, tr, r, world) == len(D.global_rows(h, tr, r, world))
    assert max(D.rows_owned(h, tr, r, world) for r in range(world)) - min(D.rows_owned(h, tr, r, world) for r in range(world)) <= tr


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import load_golden
from oracle import oracle
from yocto_raytracing_b200 import distributed as D
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
flat, _ = load_golden("basic")
W, H, S, TR = 48, 27, 1, 4
o = oracle.OracleScene(flat)
full_ref, _ = o.render(W, H, S)                      # every rank can compute the expected frame
rows = D.global_rows(H, TR, rank, world)
packed = torch.from_numpy(np.ascontiguousarray(full_ref[rows]))   # "this rank's render": its own rows only
out = D.gather_rows(packed, W, H, TR, rank, world)
if rank == 0:
    assert out is not None and np.array_equal(out.numpy(), full_ref), "gathered frame differs"
    print("GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
'''


def test_world2_gather_gloo():
    """N>1 host path on CPU: 2 ranks (gloo), each holds only its interleaved row tiles; rank 0 must end up with
    the full frame."""
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "w.py")
        open(script, "w").write(_WORKER)
        procs = []
        for r in range(2):
            env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", OMP_NUM_THREADS="2")
            procs.append(subprocess.Popen([sys.executable, script, ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
        outs = [p.communicate(timeout=300)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        assert "GATHER_OK" in outs[0]


def test_nonrigid_instance_frames_are_counted():
    """Scaled / sheared frames are counted on the host (include/yrt_b200.h, yrt_desc_nonrigid_instances): a scene that has any
    is traced through a copy of the reference's own instance tree instead of the LBVH."""
    sc = synth.mixed_scene(31)
    assert sc.flat().nonrigid_instances() == 0
    inst = list(sc.instances)
    name, si, fr = inst[3]
    f = np.array(fr, np.float32).reshape(4, 3).copy()
    f[0] *= 1.5
    inst[3] = (name, si, f.reshape(-1))
    name, si, fr = inst[5]
    f = np.array(fr, np.float32).reshape(4, 3).copy()
    f[0] += 0.3 * f[1]
    inst[5] = (name, si, f.reshape(-1))
    sc.instances = inst
    assert sc.flat().nonrigid_instances() == 2


@pytest.mark.parametrize("w,h,threads", [(1, 1, 0), (3, 2, 1), (257, 129, 3), (1920, 1080, 0), (1920, 1080, 1), (640, 2000, 8)])
def test_parallel_png_decodes_to_the_same_pixels(tmp_path, w, h, threads):
    """yrt_write_png (replaces stbi_write_png behind save_image, src/image.cpp:41-44): whatever the band
    split, the file is a valid PNG that decodes to exactly the RGBA8 pixels handed in."""
    from PIL import Image
    import yocto_raytracing_b200 as y
    rng = np.random.default_rng(w * 7 + h)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([xx * 255 // max(w - 1, 1), yy * 255 // max(h - 1, 1), (xx + yy) % 256, np.full_like(xx, 255)], -1).astype(np.uint8)
    img[h // 4: h // 2, w // 4: w // 2] = rng.integers(0, 256, (h // 2 - h // 4, w // 2 - w // 4, 4))   # noise block, alpha included
    p = str(tmp_path / "a.png")
    y.write_png(p, img, threads=threads)
    with Image.open(p) as im:
        assert im.mode == "RGBA" and im.size == (w, h)
        assert np.array_equal(np.array(im), img)


def test_parallel_png_matches_reference_png_pixels(tmp_path):
    """Re-encode a reference output (out/*.png fixtures are not shipped; use a golden float image through the oracle's
    tonemap): decoded pixels equal the tonemapped frame."""
    from PIL import Image
    import yocto_raytracing_b200 as y
    from oracle import oracle
    _, ref = load_golden("simple")
    ldr = oracle.tonemap(ref["image"])
    p = str(tmp_path / "simple.png")
    y.write_png(p, ldr)
    with Image.open(p) as im:
        assert np.array_equal(np.array(im), ldr)
    with pytest.raises(y.YrtError):
        y.write_png(str(tmp_path / "no_such_dir" / "x.png"), ldr)
    with pytest.raises(ValueError):
        y.write_png(p, ldr[..., :3])


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the reference's own CPU implementation, oracle/_ref, or the C port where it is absent) on a
    tiny bounded sample: stdout carries the one JSON line of the contract and nothing else; no GPU involved."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-baseline-resolution", "24", "--config", "basic"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["higher_is_better"] is True and "workload" in d["config"]
    # without a GPU the product arm refuses to run: there is no CPU fallback
    import torch
    if not torch.cuda.is_available():
        r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300)
        assert r2.returncode != 0 and r2.stdout.strip() == ""


def test_scene_cache_loader_rejects_corrupt_counts(tmp_path):
    """The CLI's scene cache (<scene>.yrts, host/yrt_flatten.cpp): a count that promises more bytes than the file holds is
    rejected before anything is allocated; an intact file reads back with the counts it was written with."""
    import struct
    tool = os.path.join(ROOT, "bin", "yrt_flatten")
    if not os.path.exists(tool):
        pytest.skip("bin/yrt_flatten not built (needs the reference sources at build time)")
    from yocto_raytracing_b200 import synth
    obj = synth.mixed_scene(7).write_obj(str(tmp_path))
    good = str(tmp_path / "s.yrts")
    assert subprocess.run([tool, os.path.basename(obj), good], cwd=str(tmp_path), capture_output=True).returncode == 0
    r = subprocess.run([tool, "--check", good], capture_output=True, text=True)
    assert r.returncode == 0 and "instances" in r.stdout
    b = bytearray(open(good, "rb").read())
    struct.pack_into("<q", b, 16 + 24 + 4, 1 << 40)          # count of the first array: 2^40 elements
    bad = str(tmp_path / "bad.yrts")
    open(bad, "wb").write(b)
    r = subprocess.run([tool, "--check", bad], capture_output=True, text=True, timeout=30)
    assert r.returncode != 0 and "malformed" in r.stderr


_BARRIER_WORKER = r"""
import ctypes as C, sys
from multiprocessing import shared_memory
sys.path.insert(0, sys.argv[1])
from yocto_raytracing_b200 import _lib
import numpy as np
rank, world, name = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
shm = shared_memory.SharedMemory(name=name)
try:                                       # (Python < 3.13 registers attached segments for unlinking at exit: the owner unlinks)
    from multiprocessing import resource_tracker
    resource_tracker.unregister(shm._name, "shared_memory")
except Exception:
    pass
ctr = np.ndarray((8,), np.int64, buffer=shm.buf)
data = np.ndarray((world,), np.int64, buffer=shm.buf, offset=64)
lib = _lib.load()
bad = 0
for gen in range(1, 301):
    data[rank] = gen                       # "my rows of frame gen have landed"
    assert lib.yrt_host_barrier(ctr.ctypes.data, world, gen) == 0
    if (data[:world] < gen).any():         # after the barrier every rank's write of this generation is visible
        bad += 1
    assert lib.yrt_host_barrier(ctr[1:].ctypes.data, world, gen) == 0     # second counter: nobody runs ahead into gen + 1 while we check
print("BARRIER_OK" if bad == 0 else f"BARRIER_BAD {bad}")
del ctr, data
shm.close()
"""


def test_host_barrier_orders_processes(tmp_path):
    """yrt_host_barrier (the per-frame barrier of distributed.SharedHostFrame): three processes, 300 generations, a counter in
    POSIX shared memory — after the barrier of generation g every process sees every other process's writes of g."""
    from multiprocessing import shared_memory
    shm = shared_memory.SharedMemory(create=True, size=256)
    try:
        shm.buf[:256] = bytes(256)
        script = tmp_path / "w.py"
        script.write_text(_BARRIER_WORKER)
        procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "3", shm.name], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(3)]
        outs = [p.communicate(timeout=120)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), outs
        assert all("BARRIER_OK" in o for o in outs), outs
    finally:
        shm.close()
        shm.unlink()
