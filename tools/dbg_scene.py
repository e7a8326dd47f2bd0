#!/usr/bin/env python
"""Debugging aid: render one scene through the C ABI and print stats.  python tools/dbg_scene.py scene.cli accel cols rows [spp] [photons]
With DRT_SYNC_DEBUG=1 the library syncs after every launch and names a faulting kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distraytracer_old_b200 as drt
name, accel, cols, rows = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
spp = int(sys.argv[5]) if len(sys.argv) > 5 else 0
photons = int(sys.argv[6]) if len(sys.argv) > 6 else -1
ctx = drt.Context(device=0, cols=cols, rows=rows)
s = drt.Scene.from_cli(ctx, name, spp=spp, photons=photons, accel=accel)
for k in range(2):
    argb, st = s.draw()
    print(name, "accel", accel, "frame", k, {k2: v for k2, v in st.as_dict().items() if v}, flush=True)
ctx.close()
