"""Diagnose multi-GPU imbalance: (a) the same full frame on every visible GPU, (b) every rank's row set on GPU 0."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1:      # child: one device
    dev = int(sys.argv[1])
    import torch
    import yocto_raytracing_b200 as y
    from yocto_raytracing_b200 import synth
    torch.cuda.set_device(dev)
    y.init_device(dev)
    flat = synth.instance_grid_scene(100).flat()
    W, H, S = 1920, 1080, 4
    with y.Scene(flat) as scn:
        for _ in range(3):
            img, st = scn.render(W, H, S, 0.1)
        print(f"gpu {dev}: full frame {st.ms_total:.3f} ms", flush=True)
        if dev == 0:
            buf = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
            for r in range(8):
                for _ in range(3):
                    st = scn.render_rows_into(buf.data_ptr(), W, H, S, 0.1, 1, r, 8, 0, True)
                print(f"  rows of rank {r}/8 on gpu 0: {st.ms_total:.3f} ms  (closest {st.ms_trace_closest:.3f} any {st.ms_trace_any:.3f})", flush=True)
else:
    import torch
    n = torch.cuda.device_count()
    for d in range(n):
        subprocess.run([sys.executable, __file__, str(d)], check=False)
