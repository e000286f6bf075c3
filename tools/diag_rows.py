"""Per-image-row render cost (one GPU): finds rows with pathological rays."""
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
with y.Scene(flat) as scn:
    t = np.zeros((H, 3))
    for r in range(H):
        for _ in range(2):
            st = scn.render_rows_into(buf.data_ptr(), W, H, S, 0.1, 1, r, H, 0, True)
        t[r] = (st.ms_total, st.ms_trace_closest, st.ms_trace_any)
    order = np.argsort(-t[:, 0])[:12]
    print("median row ms", np.median(t, axis=0))
    for r in order:
        print("row", r, "mod8", r % 8, "ms total/closest/any", t[r].round(4))
    print("sum by mod 8:", [round(float(t[r::8, 0].sum()), 3) for r in range(8)])
