import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from PIL import Image
import distraytracer_old_b200 as drt
from oracle import orc
name = sys.argv[1] if len(sys.argv) > 1 else "p3_t08"
o = orc.OracleScene(name + ".cli"); r = o.render(threads=8)
oa = orc.argb_to_rgb8(r["argb"]).astype(int)
for mode in (0, 256, 512, 768):
    ctx = drt.Context(device=0); s = drt.Scene.from_cli(ctx, name + ".cli", accel=mode)
    g = s.draw(aov=True)
    d = np.abs(orc.argb_to_rgb8(g["argb"]).astype(int) - oa).max(axis=-1)
    print("mode", mode, "px_ne", (d > 0).sum(), "hit mismatch", (g["hit_prim"] != r["hit_prim"]).sum(), "t mismatch", (g["t"] != r["t"]).sum(), "shadow rays", g["stats"].rays_shadow, r["stats"]["shadow"])
    Image.fromarray(np.minimum(d * 40, 255).astype(np.uint8)).save("gpurun_out/dbg_%s_%d.png" % (name, mode))
    ctx.close()
