"""Masks that pin oracle and CUDA path to the reference's own deterministic render t11_sierp.png
(<- data/project3/p3_t11_sierp.cli, Sierpinski depth 6 of bunny instances over the nightSky skydome).

The reference's mesh (bun69k.cli) is absent from its checkout, so mesh pixels can only be compared loosely; but
  * every pixel the reference shows as SKY is independent of the mesh: it pins the camera model (myScene.java:1367-1380), the skydome
    lookup (myScene.java:1104-1149) and the texel decode bit for bit;
  * the SILHOUETTE of the 1365 instances pins the Sierpinski transform chain (myScene.java:339-392) and the instance-level BVH.
`ref_sky_mask.npz` holds the sky mask: pixels where the reference PNG equals a sky-only render of the same camera (made by
make_ref_masks.py with the oracle; 62.2 % of the frame, every one of them an EXACT match).  `core` = that mask eroded by 2 pixels, which
removes the band where the stand-in mesh's outline may differ from the real bunny's."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def erode(mask, n):
    m = mask.copy()
    for _ in range(n):
        p = np.pad(m, 1, constant_values=True)
        m = p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
    return m


def load_ref():
    from PIL import Image
    ref = np.asarray(Image.open(os.path.join(HERE, "ref_t11_sierp.png")).convert("RGB")).astype(int)
    sky = np.unpackbits(np.load(os.path.join(HERE, "ref_sky_mask.npz"))["sky"])[:ref.shape[0] * ref.shape[1]].reshape(ref.shape[:2]).astype(bool)
    return {"ref": ref, "sky": sky, "core": erode(sky, 2)}


def compare(ref, rgb8, hit_mask):
    """rgb8: HxWx3 render of p3_t11_sierp_d6.cli at 300x300; hit_mask: primary ray hit geometry."""
    d = np.abs(rgb8.astype(int) - ref["ref"]).max(axis=-1)
    geom_ref = ~ref["sky"]
    return {"sky_frac": float(ref["sky"].mean()),
            "sky_bad_frac": float((d[ref["core"]] > 2).mean()),        # away from the outline: must be 0 up to the 1e-3 budget
            "sky_exact_frac": float((d[ref["core"]] == 0).mean()),
            "sky_raw_bad_frac": float((d[ref["sky"]] > 2).mean()),     # including the outline band (stand-in mesh): informational
            "iou": float((geom_ref & hit_mask).sum() / max(1, (geom_ref | hit_mask).sum()))}
