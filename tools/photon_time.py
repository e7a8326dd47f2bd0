#!/usr/bin/env python
"""Development aid: timing of the photon pass (emission, grid build) and of a render with the gather."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import distraytracer_old_b200 as drt
name = sys.argv[1]; n = int(sys.argv[2]); cols, rows, spp = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
ctx = drt.Context(device=0, cols=cols, rows=rows)
s = drt.Scene.from_cli(ctx, name, spp=spp, photons=n)
for i in range(3):
    t0 = time.perf_counter(); st = s.emit_photons_range(0, n); t1 = time.perf_counter()
    argb, rs = s.draw(); t2 = time.perf_counter()
    print("emit: wall %.1f ms, events %.1f ms, stored %d, segments %d, launches %d | render wall %.1f ms: total %.1f trace %.1f shade+gather %.1f light %.1f other %.1f" %
          ((t1 - t0) * 1e3, st.ms_total, st.photons_stored, st.rays_photon, st.kernel_launches, (t2 - t1) * 1e3, rs.ms_total, rs.ms_trace, rs.ms_shade, rs.ms_light, rs.ms_other), flush=True)
