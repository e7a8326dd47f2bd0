#!/usr/bin/env python
"""Development aid: compare closest-hit results of two traversal modes on random rays and print the first differences."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1]; m0, m1 = int(sys.argv[2]), int(sys.argv[3])
rng = np.random.default_rng(3); n = 400000
org = np.zeros((n, 3)); tgt = rng.uniform(-4, 4, size=(n, 3)) + np.array([0, 1.5, -13.5])
d = tgt - org
res = []
for accel in (m0, m1):
    ctx = drt.Context(device=0, counters=bool(int(os.environ.get("DBG_COUNTERS", "0")))); s = drt.Scene.from_cli(ctx, name, accel=accel); res.append(s.trace_rays(org, d)); ctx.close()
(i0, t0), (i1, t1) = res
bad = np.nonzero((i0 != i1).any(axis=1) | (t0 != t1))[0]
print("hits", (i0[:, 0] >= 0).sum(), "differences", len(bad))
for k in bad[:12]:
    print(k, "dir", d[k] / np.linalg.norm(d[k]), "mode%d" % m0, i0[k], repr(t0[k]), "mode%d" % m1, i1[k], repr(t1[k]))
