"""Generates ref_sky_mask.npz (see ref_masks.py): the pixels of the reference's t11_sierp.png that equal a sky-only render of the same
camera + skydome, computed with the ORACLE (test infrastructure).  Run from the repo root: python tests/golden/make_ref_masks.py"""
import os, sys, tempfile
import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    ref = np.asarray(Image.open(os.path.join(here, "ref_t11_sierp.png")).convert("RGB")).astype(int)
    d = tempfile.mkdtemp()
    open(os.path.join(d, "sky_only.cli"), "w").write("fov 30\nbackground texture nightSky.png 2000 0 -1 -1000\nwrite x.png\n")   # camera + skydome lines of p3_t11_sierp.cli
    S = orc.argb_to_rgb8(orc.OracleScene("sky_only.cli", data_dir=d).render(threads=os.cpu_count())["argb"]).astype(int)
    diff = np.abs(ref - S).max(axis=-1)
    sky = diff <= 2
    print("sky pixels: %.4f of the frame, exact matches among them: %.6f" % (sky.mean(), (diff[sky] == 0).mean()))
    np.savez_compressed(os.path.join(here, "ref_sky_mask.npz"), sky=np.packbits(sky.reshape(-1)))
