import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
from oracle import orc
rng = np.random.default_rng(5)
name = "p3_t08"
n = 20000
org = rng.uniform(-6, 6, size=(n, 3)); org[:, 2] = rng.uniform(-2, 6, size=n)
tgt = rng.uniform(-1.5, 1.5, size=(n, 3)) + np.array([0, 0, -4.0]); d = tgt - org
o = orc.OracleScene(name + ".cli"); oi, ot = o.trace_rays(org, d)
for cnt in (True, False):
    ctx = drt.Context(device=0, counters=cnt); s = drt.Scene.from_cli(ctx, name + ".cli")
    gi, gt = s.trace_rays(org, d)
    bad = np.nonzero(gt != ot)[0]
    print("counters", cnt, "id mismatches", (gi != oi).any(axis=1).sum(), "t mismatches", len(bad))
    for b in bad[:8]:
        print("  ray", b, "ids", gi[b], oi[b], "t gpu %.17g orc %.17g rel %.3g" % (gt[b], ot[b], abs(gt[b]-ot[b])/max(abs(ot[b]),1e-300)))
    ctx.close()
