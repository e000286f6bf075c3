"""Per-frame device time of the multi-GPU frame loop (SharedFrame, peer-memory barrier) under torchrun: prints every rank's
per-frame CUDA-event times — shows whether a slow step is one stalled frame or a uniform slowdown.  --barrier peer|nccl"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import configs, distributed as D
ap = argparse.ArgumentParser(); ap.add_argument("--barrier", default="peer"); ap.add_argument("--frames", type=int, default=40)
ap.add_argument("--stats", action="store_true", help="deferred statistics on (per-launch events), like bench.py's timed region")
ap.add_argument("--tensor", action="store_true", help="rank 0 takes a tensor view of the frame after every render call, like bench.py")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
y.init_device(local)
flat, res, smp, name = configs.load("instance")
w = flat.image_width(res)
scene = y.Scene(flat)
fr = D.SharedFrame(w, res, barrier=a.barrier)
for _ in range(5):
    fr.render(scene, smp, 0.1, 1)
torch.cuda.synchronize(); dist.barrier()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.frames + 1)]
if a.stats:
    scene.stats_begin()
ev[0].record()
for k in range(a.frames):
    fr.render(scene, smp, 0.1, 1)
    if a.tensor and rank == 0:
        t = fr.tensor()
    ev[k + 1].record()
torch.cuda.synchronize(); dist.barrier()
if a.stats:
    scene.stats_end()
ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(a.frames)]
print(f"rank {rank} [{a.barrier} stats={a.stats} tensor={a.tensor}] mean {np.mean(ms):.3f} median {np.median(ms):.3f} max {np.max(ms):.3f} | " + " ".join(f"{m:.2f}" for m in ms), flush=True)
fr.close(); scene.close(); dist.barrier(); dist.destroy_process_group()
