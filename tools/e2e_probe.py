#!/usr/bin/env python
"""Development aid: where does the end-to-end (host buffers) step spend its wall time?"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import distraytracer_old_b200 as drt
ctx = drt.Context(device=0, cols=3840, rows=2160)
s = drt.Scene.from_cli(ctx, "p3_t09.cli", spp=16)
host = torch.empty(3840 * 2160, dtype=torch.int32).pin_memory()
for i in range(4):
    t0 = time.perf_counter(); s.reupload(); t1 = time.perf_counter(); st = s.draw_into(host.data_ptr()); t2 = time.perf_counter()
    print("reupload %.1f ms  draw_into %.1f ms (gpu %.1f)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, st.ms_total), flush=True)
