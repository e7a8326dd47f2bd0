#!/usr/bin/env python
"""Decode the scene textures (PNG/JPEG) into raw ARGB int32 files '<name>.argb'.

Layout: b'ARGB', int32 width, int32 height, width*height int32 (0xAARRGGBB, little endian) -- the
layout of Processing's PImage.pixels, which is what the reference hands to its texture lookups
(myRTFileReader.java:133,265).  Both the oracle and the product's C++ harness read these files so
that they see identical texels (JPEG decoders differ by +-1 LSB).  The Python host
decodes with PIL directly and passes the same ints through the C ABI.
"""
import struct, sys, os
import numpy as np
from PIL import Image

def decode(path):
    im = Image.open(path).convert("RGBA")
    a = np.asarray(im, dtype=np.uint32)
    argb = (np.uint32(0xFF) << 24) | (a[..., 0] << 16) | (a[..., 1] << 8) | a[..., 2]   # PImage RGB images carry alpha 0xFF
    return im.width, im.height, argb.astype("<u4")

def main(src_dir, dst_dir):
    os.makedirs(dst_dir, exist_ok=True)
    for fn in sorted(os.listdir(src_dir)):
        if not fn.lower().endswith((".png", ".jpg", ".jpeg")):
            continue
        w, h, px = decode(os.path.join(src_dir, fn))
        with open(os.path.join(dst_dir, fn + ".argb"), "wb") as f:
            f.write(b"ARGB" + struct.pack("<ii", w, h) + px.tobytes())
        print(fn, w, h)

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "scenes", "txtrs")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "..", "scenes", "txtrs_argb")
    main(src, dst)
