#!/usr/bin/env python
"""Development aid: render scenes with libdrt (GPU) and with the oracle (CPU) and print parity + timing."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import distraytracer_old_b200 as drt
from oracle import orc

def compare(name, cols=300, rows=300, spp=0, photons=-1, accel=0, save=None):
    ctx = drt.Context(device=0, cols=cols, rows=rows, counters=os.environ.get("DRT_COUNTERS", "1") == "1")
    t0 = time.time(); s = drt.Scene.from_cli(ctx, name, spp=spp, photons=photons, accel=accel); tl = time.time() - t0
    g = s.draw(aov=True); g2 = s.draw(aov=True)
    o = orc.OracleScene(name, cols=cols, rows=rows, spp=spp if spp > 0 else -1, photons=photons)
    r = o.render(threads=os.cpu_count())
    hp = (g["hit_prim"] != r["hit_prim"]).sum(); hi = (g["hit_inst"] != r["hit_inst"]).sum()
    ga, oa = orc.argb_to_rgb8(g["argb"]).astype(int), orc.argb_to_rgb8(r["argb"]).astype(int)
    d = np.abs(ga - oa).max(axis=-1)
    st = g2["stats"]
    res = dict(scene=name, res=[cols, rows], spp=s.info()["spp"], load_s=round(tl, 3), hit_prim_mismatch=int(hp), hit_inst_mismatch=int(hi),
               px_gt2=int((d > 2).sum()), px_ne=int((d > 0).sum()), max_diff=int(d.max()), rgb_maxabs=float(np.nanmax(np.abs(g["rgb"] - r["rgb"]))),
               gpu_ms=round(st.ms_total, 3), ms_trace=round(st.ms_trace, 3), ms_shade=round(st.ms_shade, 3), ms_light=round(st.ms_light, 3),
               rays=st.rays_total, oracle_rays=sum(r["stats"][k] for k in ("primary", "shadow", "reflect", "refract")), cpu_s=round(r["seconds"], 3),
               gpu_box=st.box_tests, orc_box=r["stats"]["box_tests"], gpu_prim=st.prim_tests, orc_prim=r["stats"]["prim_tests"])
    print(json.dumps(res), flush=True)
    if save:
        from PIL import Image
        os.makedirs(save, exist_ok=True)
        Image.fromarray(ga.astype(np.uint8)).save(os.path.join(save, name.replace(".cli", "") + "_gpu.png"))
        Image.fromarray((np.minimum(d * 40, 255)).astype(np.uint8)).save(os.path.join(save, name.replace(".cli", "") + "_diff.png"))
    ctx.close()
    return res

if __name__ == "__main__":
    scenes = sys.argv[1:] or ["t01.cli", "t03.cli", "p3_t05.cli", "p3_t06.cli", "p3_t08.cli", "p3_t12.cli", "p3_t02_sierp.cli", "p3_t09.cli"]
    out = []
    for sc in scenes:
        try:
            out.append(compare(sc, save="gpurun_out/img"))
        except Exception as e:
            print("FAIL", sc, repr(e), flush=True)
    json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
