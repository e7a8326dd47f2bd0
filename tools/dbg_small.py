#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1]; accel = int(sys.argv[2]); n = int(sys.argv[3])
rng = np.random.default_rng(3)
org = np.zeros((n, 3)); tgt = rng.uniform(-4, 4, size=(n, 3)) + np.array([0, 1.5, -13.5])
ctx = drt.Context(device=0, counters=False); s = drt.Scene.from_cli(ctx, name, accel=accel); i, t = s.trace_rays(org, tgt - org); ctx.close()
print("hits", (i[:, 0] >= 0).sum(), "checksum", int(i.sum()), float(t.sum()))
