#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1]
rng = np.random.default_rng(3); n = 200000
org = np.zeros((n, 3)); tgt = rng.uniform(-4, 4, size=(n, 3)) + np.array([0, 1.5, -13.5])
d = tgt - org
res = {}
for accel in (0, 1, 2):
    for cnt in (True, False):
        ctx = drt.Context(device=0, counters=cnt); s = drt.Scene.from_cli(ctx, name, accel=accel); res[(accel, cnt)] = s.trace_rays(org, d); ctx.close()
ref = res[(0, True)]
for k, v in res.items():
    print(k, "differences vs (0,True):", int(((v[0] != ref[0]).any(axis=1) | (v[1] != ref[1])).sum()))
