#!/usr/bin/env python
"""bench.py — Mrays/s & ms/frame of the hot path on BASELINE.json's headline config:
an instance10000_pointlight-shaped scene (10 004 instances of 14 shapes, 3 point lights) at 1920x1080,
16 spp (-r 1080 -s 4), on 1/2/4/8 B200 of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's own CPU implementation on this box's host cores

A step is one frame.  Rays are counted by reference semantics (intersect_first + intersect_any calls the
reference would make: primary + reflection + shadow).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

# per-ray algorithmic work of the reference on this config (measured with counters in the
# reference's BVH): 2.84 KB of node/primitive/instance bytes and 2.45 kflop per ray
ALG_BYTES_PER_RAY = 2840.0
ALG_FLOPS_PER_RAY = 2450.0
# dram bytes of one k_trace_any_lights launch (whole 1080p/16spp frame) from the ncu --set full capture in profiles/
NCU_DRAM_BYTES_PER_ANY_LAUNCH = 1.1833e9   # profiles/r1j_final_ncu_summary.txt: 1.0755 GB read + 0.1078 GB write (algorithmic: 33.2 M hits x 35 B)
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--resolution", type=int, default=1080)
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--n-side", type=int, default=100, help="instances per grid side (100 -> 10 004 instances)")
    ap.add_argument("--tile-rows", type=int, default=1, help="rows per interleaved tile (1: rows r, r+N, r+2N, ...)")
    ap.add_argument("--gather", default="ipc", choices=["ipc", "nccl"],
                    help="multi-GPU framebuffer exchange: ipc = every rank's resolve kernel stores its rows into rank 0's frame over "
                         "NVLink peer memory + 1-element all-reduce; nccl = packed rows, NCCL gather, unpack on rank 0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-resolution", type=int, default=0, help="0 = sized for ~15 s of CPU work")
    return ap.parse_args()


def workload_name(args, flat):
    w = flat.image_width(args.resolution)
    return (f"instance10000_pointlight-shaped synthetic scene ({flat.n_instances} instances, {flat.n_shapes} shapes, "
            f"{flat.n_elements} elements, {len(flat.light_instances())} point lights), {w}x{args.resolution}, {args.samples ** 2} spp")


# ---- clocks during the timed region -----------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(pw))}


# ---- the reference's CPU implementation, bounded sample ------------------------------------------------
def cpu_reference_sample(synth_scene, flat, resolution, samples, ray_counter):
    """Time the reference's raytrace() on this box's host cores for one bounded frame.  Prefers the UNMODIFIED
    reference (oracle/_ref/ref_probe, compiled from /root/reference/src; single-threaded like the reference),
    else the C port (oracle/yrt_oracle.c) on one thread.  ray_counter(resolution, samples) -> rays by
    reference semantics."""
    import ref_probe
    w = flat.image_width(resolution)
    cores_total = os.cpu_count() or 1
    if ref_probe.available():
        with tempfile.TemporaryDirectory() as td:
            obj = synth_scene.write_obj(td)
            _, info = ref_probe.image(obj, resolution, samples, 0.1)
        secs, kind = float(info["raytrace_s"]), "reference"
        rays = ray_counter(resolution, samples)
    else:
        from oracle import oracle
        o = oracle.OracleScene(flat)
        t0 = time.perf_counter()
        _, cnt = o.render(w, resolution, samples, 0.1, threads=1)
        secs, kind = time.perf_counter() - t0, "port"
        rays = cnt["primary_rays"] + cnt["reflection_rays"] + cnt["shadow_rays"]
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": 1, "kind": kind, "host_cores_total": cores_total,
            "sample": f"same scene, one frame at {w}x{resolution}, {samples * samples} spp = {rays} rays in {secs:.2f} s "
                      f"(raytrace() phase only, single thread: the reference is single-threaded, src/raytrace.cpp:228-251)",
            "seconds": secs, "rays": rays}


def oracle_ray_count(flat, resolution, samples):
    """Ray count by reference semantics without a GPU and without rendering: for this scene family every hit
    casts n_lights shadow rays and nothing reflects; hits come from the C oracle's closest-hit pass."""
    from oracle import oracle
    o = oracle.OracleScene(flat)
    w = flat.image_width(resolution)
    ids, _, _ = o.trace_primary(w, resolution, samples)
    hits = int((ids[:, 0] >= 0).sum())
    return ids.shape[0] + hits * len(flat.light_instances())


def sample_resolution_for(seconds, mrays_per_s=0.16, n_lights=3):
    rays = seconds * mrays_per_s * 1e6
    r = int((rays / ((1 + n_lights) * 16.0 / 9.0)) ** 0.5)
    return max(36, min(1080, r - r % 2))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0      # one CPU arm per job: rank 0 alone runs it
    from yocto_raytracing_b200 import synth
    sc = synth.instance_grid_scene(args.n_side)
    flat = sc.flat()
    total = max(1, args.steps + args.warmup)
    res = args.cpu_baseline_resolution or sample_resolution_for(min(20.0, 150.0 / total))
    rays = oracle_ray_count(flat, res, 1)
    times, last = [], None
    for i in range(total):
        last = cpu_reference_sample(sc, flat, res, 1, lambda r, s: rays)
        if i >= args.warmup:
            times.append(last["seconds"])
    secs = float(np.mean(times))
    value = rays / secs / 1e6
    w = flat.image_width(args.resolution)
    full_rays = None
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, flat), "step": f"bounded sample per step: one frame of the same scene at "
                   f"{flat.image_width(res)}x{res}, 1 spp ({rays} rays); Mrays/s is resolution-independent for this path"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": 1, "kind": last["kind"], "host_cores_total": last["host_cores_total"],
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---- the B200 arm -----------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import yocto_raytracing_b200 as y
    from yocto_raytracing_b200 import distributed as D
    from yocto_raytracing_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU fallback")
    torch.cuda.set_device(local)
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    y.init_device(local)

    sc = synth.instance_grid_scene(args.n_side)
    flat = sc.flat()
    W, H, S = flat.image_width(args.resolution), args.resolution, args.samples
    t_up = time.perf_counter()
    scene = y.Scene(flat)           # validate + upload + GPU LBVH build (reported separately, not in the timed region)
    info = scene.info()
    info["scene_create_ms_wall"] = round((time.perf_counter() - t_up) * 1e3, 3)
    dev = torch.device("cuda", local)
    tr = args.tile_rows if multi else H

    shared = None
    if args.gather == "ipc":
        # CUDA IPC can be unavailable (container policy); every rank must take the same path
        ok = torch.ones(1, device=dev)
        try:
            shared = D.SharedFrame(W, H)
        except Exception as e:      # noqa: BLE001
            ok.zero_()
            print(f"[bench] rank {rank}: shared frame over CUDA IPC unavailable ({e}); falling back to the NCCL gather", file=sys.stderr)
        if multi:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            shared = None
            args.gather = "nccl"

    def frame(want_stats=True):
        if shared is not None:
            st = shared.render(scene, S, 0.1, tr, want_stats)
            return (shared.tensor() if rank == 0 else None), st
        return D.render_sharded(scene, W, H, S, 0.1, tr, None, want_stats)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ----
    # the clock sampler starts BEFORE the warm-up so that no idle gap (clock ramp-down) precedes the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # deferred statistics: per-launch CUDA events and ray counters are recorded in-stream, nothing synchronises
    # with the host inside the timed region; the totals over the K frames are read after it
    scene.stats_begin()
    e0.record()
    for _ in range(args.steps):
        full, _ = frame(False)
    e1.record()
    barrier()
    tot_stats = scene.stats_end()
    assert tot_stats.frames == args.steps, (tot_stats.frames, args.steps)
    per_frame = {k: (v / args.steps if isinstance(v, (int, float)) else v) for k, v in tot_stats.as_dict().items()}
    kernel_timing = "CUDA events around every launch inside the timed region"
    if multi and os.environ.get("YRT_STREAMS", "2") != "1":
        # a rank's share runs as two pipelines on two streams, so kernels of the two half shares overlap and an event
        # span includes queueing behind the other pipeline: per-kernel durations come from a single-pipeline pass of
        # 3 frames right after the timed region (same buffers, same kernels)
        os.environ["YRT_STREAMS"] = "1"
        scene.stats_begin()
        for _ in range(3):
            frame(False)
        barrier()
        k3 = scene.stats_end()
        del os.environ["YRT_STREAMS"]
        for key in ("ms_trace_closest", "ms_trace_any", "ms_shade", "ms_other", "n_closest", "n_any", "n_shade", "n_other"):
            per_frame[key] = getattr(k3, key) / 3.0
        kernel_timing = "single-pipeline pass of 3 frames right after the timed region (in the timed region two pipelines overlap kernels)"
    stats_all = [per_frame]
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    rays_local = float(tot_stats.primary_rays + tot_stats.reflection_rays + tot_stats.shadow_rays)
    tot = torch.tensor([rays_local, float(tot_stats.launches)], device=dev, dtype=torch.float64)
    if multi:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    # per-rank device time of the frames (render only, from the in-stream frame events) — shows rank imbalance
    mine = torch.tensor([tot_stats.ms_total / args.steps, (tot_stats.ms_trace_closest + tot_stats.ms_trace_any + tot_stats.ms_shade + tot_stats.ms_other) / args.steps],
                        device=dev, dtype=torch.float64)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if multi:
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    per_rank_ms = [[round(float(x[0]), 4), round(float(x[1]), 4)] for x in per_rank]
    clocks = sampler.stop() if rank == 0 else None
    ms_total = float(ms.item())
    rays_total, launches = float(tot[0].item()), int(tot[1].item())
    if multi:
        launches += args.steps * (world if shared is not None else world + 1)      # + the exchange: all-reduce per rank / NCCL gather + unpack kernels
    value = rays_total / (ms_total * 1e-3) / 1e6

    # ---- end to end through the public API with HOST buffers (`e2e`) ----
    pinned = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None
    host_np = pinned.numpy() if rank == 0 else None

    def frame_e2e():
        if not multi:
            scene.render(W, H, S, 0.1, out=host_np, want_stats=False)      # yrt_render: camera in, host framebuffer out
        else:
            full, _ = frame(False)
            if rank == 0:
                pinned.copy_(full, non_blocking=True)
            torch.cuda.synchronize()
    for _ in range(2):
        frame_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        frame_e2e()
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if multi:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = rays_total / float(dt.item()) / 1e6

    if rank != 0:
        if multi:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_trace_any_lights: 75 % of the rays), rank 0 ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        hbm_peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    any_ms = float(np.mean([s["ms_trace_any"] for s in stats_all]))
    any_n = max(1, int(round(np.mean([s["n_any"] for s in stats_all]))))
    closest_ms = float(np.mean([s["ms_trace_closest"] for s in stats_all]))
    shade_ms = float(np.mean([s["ms_shade"] for s in stats_all]))
    other_ms = float(np.mean([s["ms_other"] for s in stats_all]))
    shadow_per_frame = float(np.mean([s["shadow_rays"] for s in stats_all]))
    primary_per_frame = float(np.mean([s["primary_rays"] for s in stats_all]))
    any_launch_ms = any_ms / any_n
    rays_per_launch = shadow_per_frame / any_n
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # this path is bound by FP32 issue and by L1/L2 node fetch, NOT by HBM or tensor cores, so the
    # roofline is stated against the FP32 pipe: algorithmic flops = 2.45 kflop per ray (the reference's own
    # box/triangle/transform counts) over 148 SMs x 128 lanes x 2 flop x the SM clock seen during the run.
    # Exact (unfused) arithmetic in the primitive tests and min/max/select-heavy slab tests cap what is reachable.
    fp32_peak_tflops = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    achieved_tflops = rays_per_launch * ALG_FLOPS_PER_RAY / (any_launch_ms * 1e-3) / 1e12
    cache_gbs = rays_per_launch * ALG_BYTES_PER_RAY / (any_launch_ms * 1e-3) / 1e9
    roofline = {
        "kernel": "k_trace_any_lights", "bound": "fp32", "bound_note": "FP32 instruction issue — neither hbm nor tensor (SURVEY 8d): the 3 MB scene is cache resident and nothing is a contraction", "achieved": achieved_tflops,
        "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak_tflops,
        "peak_source": f"148 SMs x 128 FP32 lanes x 2 flop x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
        "traffic": NCU_DRAM_BYTES_PER_ANY_LAUNCH, "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/",
        "launch_ms": any_launch_ms, "launch_timing": kernel_timing, "launches_per_frame": any_n, "rays_per_launch": rays_per_launch,
        "algorithmic_flops_per_ray": ALG_FLOPS_PER_RAY, "algorithmic_bytes_per_ray": ALG_BYTES_PER_RAY,
        "hbm": {"note": "literal bytes roofline: 2.84 KB/ray is CACHE-level (L1/L2) node+primitive traffic of the reference's traversal; "
                        "compulsory HBM traffic is ~11 B/ray (hit record + position + visibility), so frac > 1 here only says the "
                        "scene is cache-resident, not that HBM binds",
                "achieved": cache_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": cache_gbs / hbm_peak, "peak_source": peak_src},
        "kernel_share_of_step": any_ms / (ms_total / args.steps),
    }

    cpu_baseline = None
    if not multi and not args.no_cpu_baseline:
        res = args.cpu_baseline_resolution or sample_resolution_for(15.0)

        def count(r, s):
            _, st = scene.render(flat.image_width(r), r, s, 0.1)
            return st.total_rays
        cb = cpu_reference_sample(sc, flat, res, 1, count)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "host_cores_total", "sample")}

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, flat), "rays_per_frame": rays_total / args.steps, "parallelism": f"row-tiles x{world}", "exchange": args.gather if multi else "none",
                   "tile_rows": tr, "l2": "per-frame ray/hit queues (~1.7 GB streamed per frame) exceed the 126 MB L2; the 3 MB scene is "
                   "cache-resident by design", "lbvh": info},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 64 + 12, "d2h_bytes_per_step": W * H * 16,
                "ms_per_step": float(dt.item()) * 1e3 / args.steps,
                "api": "Scene.render -> yrt_render (host framebuffer out)" if not multi else
                (f"distributed.SharedFrame.render (peer stores into rank 0's frame + all-reduce) + D2H on rank 0" if shared is not None
                 else "distributed.render_sharded + NCCL gather + D2H on rank 0")},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "per_rank_render_ms_and_kernel_sum": per_rank_ms,
        "breakdown_ms_per_frame_rank0": {"trace_closest": closest_ms, "trace_any": any_ms, "shade": shade_ms, "resolve": other_ms},
        "mrays_s_by_kernel_rank0": {"closest": primary_per_frame / (closest_ms * 1e-3) / 1e6 if closest_ms else None,
                                    "any": shadow_per_frame / (any_ms * 1e-3) / 1e6 if any_ms else None},
    }
    emit(line)
    scene.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Libraries underneath (NCCL with NCCL_DEBUG=VERSION/INFO, the CUDA runtime) write to file descriptor 1 behind Python's
# back; the contract is one JSON line on stdout, so everything else that lands on fd 1 is sent to stderr.
_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
