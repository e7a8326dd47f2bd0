import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
name = sys.argv[1] if len(sys.argv) > 1 else "p3_t08"
dt = np.dtype([("t", "f8"), ("prim", "i4"), ("arg0", "i4"), ("arg1", "i4"), ("state", "i4"), ("hitXform", "i4"), ("shaderOverride", "i4"), ("inst", "i4"), ("pad", "i4"), ("loc", "f8", 3), ("rawDir", "f8", 3), ("pad1", "f8")])
print("itemsize", dt.itemsize)
res = {}
for mode in (0, 512):
    os.environ["DRT_DUMP_HITS"] = "/tmp/hits_%d.bin" % mode
    ctx = drt.Context(device=0); s = drt.Scene.from_cli(ctx, name + ".cli", accel=mode); s.draw(aov=True); ctx.close()
    res[mode] = np.fromfile("/tmp/hits_%d.bin" % mode, dtype=dt)
a, b = res[0], res[512]
hit = a["prim"] >= 0
for f in ("t", "prim", "arg0", "arg1", "state", "hitXform", "shaderOverride", "inst", "loc", "rawDir"):
    x, y = a[f][hit], b[f][hit]
    ne = (x != y).reshape(len(x), -1).any(axis=1)
    print(f, "mismatch", ne.sum())
    if ne.sum():
        i = np.nonzero(ne)[0][:4]
        print("   false-variant:", x[i].tolist()); print("   true-variant :", y[i].tolist())
