#!/usr/bin/env python
"""Development aid: which gather tier do queries on the box floor / walls take, and how long does the probe take per tier mix?"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1]; n = int(sys.argv[2])
ctx = drt.Context(device=0, cols=64, rows=64)
s = drt.Scene.from_cli(ctx, name, photons=n)
rng = np.random.default_rng(1)
m = 2000000
pts = np.stack([rng.uniform(-1.2, 1.2, m), np.full(m, -1.0), rng.uniform(-5.7, 0.7, m)], axis=1)
pts = pts[np.argsort((pts[:, 2] * 200).astype(int) * 1000 + (pts[:, 0] * 200).astype(int))]     # coherent order, like pixels
s.photon_probe(pts[:1000])
t0 = time.perf_counter(); out = s.photon_probe(pts); dt = time.perf_counter() - t0
print(name, n, "tiers", dict(zip(*[x.tolist() for x in np.unique(out[:, 4], return_counts=True)])), "probe wall %.1f ms for %d queries" % (dt * 1e3, m))
