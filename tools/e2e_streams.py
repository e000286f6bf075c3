"""End-to-end time of yrt_render (host frame out, page-locked buffer) for 1..4 pipelines per frame (YRT_STREAMS): the host copy
of one batch runs under the kernels of the next.  GPU box: python tools/e2e_streams.py"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import yocto_raytracing_b200 as y
from yocto_raytracing_b200 import configs

y.init(1)
flat, res, smp, name = configs.load("instance")
w = flat.image_width(res)
buf = torch.empty((res, w, 4), dtype=torch.float32).pin_memory().numpy()
with y.Scene(flat) as scn:
    for streams, slots in (("1", 0), ("2", 0), ("3", 0), ("4", 0), ("2", 0), ("2", 8400000), ("2", 5600000), ("2", 4200000), ("3", 5600000)):
        os.environ["YRT_STREAMS"] = streams
        if slots:
            os.environ["YRT_BATCH_SLOTS"] = str(slots)       # more, smaller batches over the same pipelines: a smaller last copy is left exposed
        else:
            os.environ.pop("YRT_BATCH_SLOTS", None)
        for _ in range(3):
            scn.render(w, res, smp, 0.1, out=buf, want_stats=False)
        t0 = time.perf_counter()
        for _ in range(20):
            scn.render(w, res, smp, 0.1, out=buf, want_stats=False)
        print(f"YRT_STREAMS={streams} YRT_BATCH_SLOTS={slots or 'default'}: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per frame end to end")
