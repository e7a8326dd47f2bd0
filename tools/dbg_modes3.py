#!/usr/bin/env python
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1]; accel = int(sys.argv[2])
rng = np.random.default_rng(3); n = 200000
org = np.zeros((n, 3)); tgt = rng.uniform(-4, 4, size=(n, 3)) + np.array([0, 1.5, -13.5])
d = tgt - org
res = []
for cnt in (True, False):
    ctx = drt.Context(device=0, counters=cnt); s = drt.Scene.from_cli(ctx, name, accel=accel); res.append(s.trace_rays(org, d)); ctx.close()
(i0, t0), (i1, t1) = res
bad = np.nonzero((i0 != i1).any(axis=1) | (t0 != t1))[0]
print("differences", len(bad))
for k in bad[:14]:
    print(k, "counters", i0[k], repr(float(t0[k])), "nocounters", i1[k], repr(float(t1[k])))
