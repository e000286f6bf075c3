"""Worker of test_torchrun_moving_camera_frames_match_single_gpu (one process per GPU under torchrun, NCCL): several frames
with a DIFFERENT camera each, through both multi-GPU frame paths — rank 0's device frame fed by peer stores
(distributed.SharedFrame) and the host frame shared by the ranks (distributed.SharedHostFrame).  Rank 0 compares every
frame bitwise with its own single-GPU render of the same camera, consuming each frame only AFTER the next render call
has been issued by the other ranks would be a race — the barriers of the two classes are what is under test."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yocto_raytracing_b200 as y  # noqa: E402
from yocto_raytracing_b200 import distributed as D, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    y.init_device(local)
    sc = synth.instance_grid_scene(24, seed=3)
    flat = sc.flat()
    w, h, s = 384, 216, 2
    scene = y.Scene(flat)
    dev_frame = D.SharedFrame(w, h)                         # two frames, one peer-memory barrier kernel per frame
    nccl_frame = D.SharedFrame(w, h, barrier="nccl")        # round 1's protocol: one frame between two all-reduces
    host_frame = D.SharedHostFrame(w, h)
    cams = [synth.make_camera((12.0 * np.cos(a), 14.0 + 3 * k, 12.0 * np.sin(a)), (0, 1, 0), 0.6) for k, a in enumerate(np.linspace(0.3, 2.5, 5))]
    bad = 0
    got_dev, got_host, got_nccl = [], [], []
    for cam in cams:
        scene.set_camera(cam)
        dev_frame.render(scene, s, 0.1, 1)
        if rank == 0:
            got_dev.append(dev_frame.tensor().cpu().numpy().copy())     # read the frame; the next render call follows at once
        nccl_frame.render(scene, s, 0.1, 1)
        if rank == 0:
            got_nccl.append(nccl_frame.tensor().cpu().numpy().copy())
        host_frame.render(scene, s, 0.1, 1)
        if rank == 0:
            got_host.append(host_frame.array.copy())
    dist.barrier()
    if rank == 0:
        for k, cam in enumerate(cams):
            scene.set_camera(cam)
            ref, _ = scene.render(w, h, s, 0.1, want_stats=False)
            if not np.array_equal(ref.view(np.uint32), got_dev[k].view(np.uint32)):
                bad += 1
                print(f"frame {k}: device frame differs from the single-GPU frame", flush=True)
            if not np.array_equal(ref.view(np.uint32), got_nccl[k].view(np.uint32)):
                bad += 1
                print(f"frame {k}: device frame (NCCL barriers) differs from the single-GPU frame", flush=True)
            if not np.array_equal(ref.view(np.uint32), got_host[k].view(np.uint32)):
                bad += 1
                print(f"frame {k}: host frame differs from the single-GPU frame", flush=True)
        assert len({a.tobytes() for a in got_dev}) == len(cams)            # the cameras really differ
        print("MGPU_OK" if bad == 0 else f"MGPU_BAD {bad}", flush=True)
    host_frame.close()
    nccl_frame.close()
    dev_frame.close()
    scene.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
