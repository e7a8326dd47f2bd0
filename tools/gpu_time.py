#!/usr/bin/env python
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1] if len(sys.argv) > 1 else "p3_t09.cli"
cols, rows, spp = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
accel = int(sys.argv[6]) if len(sys.argv) > 6 else 0
ctx = drt.Context(device=0, cols=cols, rows=rows, counters=False)
t0 = time.time(); s = drt.Scene.from_cli(ctx, name, spp=spp, accel=accel); print("load+upload s", round(time.time() - t0, 2), s.info(), flush=True)
for i in range(reps):
    t0 = time.time(); argb, st = s.draw(); dt = time.time() - t0
    print(json.dumps(dict(ms_total=round(st.ms_total, 1), wall_ms=round(dt * 1e3, 1), trace=round(st.ms_trace, 1), shade=round(st.ms_shade, 1), light=round(st.ms_light, 1), other=round(st.ms_other, 1),
                          rays=st.rays_total, mrays_s=round(st.rays_total / st.ms_total / 1e3, 1), launches=st.kernel_launches)), flush=True)
s.save("gpurun_out/time_%s.png" % name.replace(".cli", ""), argb)
