import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import distraytracer_old_b200 as drt
ctx = drt.Context(device=0, cols=100, rows=100); s = drt.Scene.from_cli(ctx, "p3_t08.cli")
g = s.draw(aov=True); print("done", g["stats"].rays_total)
ctx.close()
