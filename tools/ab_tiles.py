"""A/B sweep of the work-distribution schemes (csrc/yrt_work.cuh) on one GPU, one process: whole frames (N = 1) and
rank 0's share of an 8-GPU frame, per-kernel CUDA-event times, and a bit-identity check of every frame against the
linear scheme.  python tools/ab_tiles.py [--frames 6]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=6)
ap.add_argument("--share-frames", type=int, default=20)
a = ap.parse_args()
y.init(1)
flat = synth.instance_grid_scene(100).flat()
W, H, S = 1920, 1080, 4
KEYS = ("YRT_TILE", "YRT_TILE_W", "YRT_TILE_H", "YRT_TILES_PER_SM", "YRT_BLOCKS_PER_SM", "YRT_CHUNK_ITEMS")
configs = [("linear", {"YRT_TILE": "0"})]
for tw, th in ((2, 1), (4, 2), (4, 4), (8, 4), (8, 8), (16, 8), (16, 16), (32, 8), (32, 16), (64, 16), (64, 32)):
    configs.append((f"tile {tw}x{th}", {"YRT_TILE": "1", "YRT_TILE_W": str(tw), "YRT_TILE_H": str(th), "YRT_TILES_PER_SM": "0"}))
configs.append(("tile 16x8 adaptive", {"YRT_TILE": "1"}))
configs.append(("tile 16x8, 7 CTAs/SM", {"YRT_TILE": "1", "YRT_TILE_W": "16", "YRT_TILE_H": "8", "YRT_TILES_PER_SM": "0", "YRT_BLOCKS_PER_SM": "7"}))
configs.append(("tile 8x8, 7 CTAs/SM", {"YRT_TILE": "1", "YRT_TILE_W": "8", "YRT_TILE_H": "8", "YRT_TILES_PER_SM": "0", "YRT_BLOCKS_PER_SM": "7"}))
configs.append(("linear again", {"YRT_TILE": "0"}))
buf = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
ref = None
with y.Scene(flat) as scn:
    print(scn.info(), flush=True)
    for name, env in configs:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        ts = []
        for f in range(a.frames):
            img, st = scn.render(W, H, S, 0.1)
            ts.append((st.ms_total, st.ms_trace_closest, st.ms_trace_any, st.ms_shade, st.ms_other))
        t = np.median(np.array(ts[2:]), axis=0)
        if ref is None:
            ref = img.copy()
        same = np.array_equal(ref.view(np.uint32), img.view(np.uint32))
        # rank 0's share of an 8-GPU frame (two pipelines, rows r = 0 mod 8), stored into a full frame
        for _ in range(3):
            scn.render_rows_into_frame(buf.data_ptr(), W, H, S, 0.1, 1, 0, 8, 0, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.share_frames):
            scn.render_rows_into_frame(buf.data_ptr(), W, H, S, 0.1, 1, 0, 8, 0, False)
        e1.record(); torch.cuda.synchronize()
        share = e0.elapsed_time(e1) / a.share_frames
        rows = buf[0::8].cpu().numpy()
        same_share = np.array_equal(rows.view(np.uint32), ref[0::8].view(np.uint32))
        print(f"{name:24s} frame {t[0]:7.3f} ms | closest {t[1]:6.3f} any {t[2]:6.3f} shade {t[3]:6.3f} other {t[4]:6.3f} | 1/8 share {share:6.3f} ms | "
              f"identical {same} {same_share}", flush=True)
