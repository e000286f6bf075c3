"""Frame time of one rank's share (rank 0 of `world`) with one vs two pipelines (YRT_STREAMS), one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import synth
y.init(1)
flat = synth.instance_grid_scene(100).flat()
W, H, S = 1920, 1080, 4
buf = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
ref = None
with y.Scene(flat) as scn:
    for world in (8, 4, 1):
        for streams in (1, 2, 3, 4):
            os.environ["YRT_STREAMS"] = str(streams)
            for _ in range(3):
                scn.render_rows_into_frame(buf.data_ptr(), W, H, S, 0.1, 1, 0, world, 0, False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                scn.render_rows_into_frame(buf.data_ptr(), W, H, S, 0.1, 1, 0, world, 0, False)
            e1.record(); torch.cuda.synchronize()
            print(f"world {world} streams {streams}: {e0.elapsed_time(e1) / 20:.4f} ms / frame", flush=True)
        if world == 1:
            a = buf.cpu().numpy().copy()
            os.environ["YRT_STREAMS"] = "1"
            scn.render_rows_into_frame(buf.data_ptr(), W, H, S, 0.1, 1, 0, 1, 0, False); torch.cuda.synchronize()
            print("two-stream frame identical to one-stream:", np.array_equal(a.view(np.uint32), buf.cpu().numpy().view(np.uint32)))
