#!/usr/bin/env python
"""Stand-in for the reference's missing data/bun69k.cli (listed in /root/reference/.MISSING_LARGE_BLOBS).

bun500.cli (966 triangles, same bunny, same extents) is subdivided 1->4 by edge midpoints three times
-> 61 824 triangles, written in the same bare `begin / vertex x3 / end` block format.  If a real
bun69k.cli is placed in scenes/ it is used as is (this script never overwrites an existing file unless --force).
"""
import os, sys

def read_tris(path):
    tris, cur = [], []
    for line in open(path):
        t = line.split()
        if not t: continue
        if t[0] == "vertex": cur.append(tuple(float(x) for x in t[1:4]))
        elif t[0] == "end":
            if len(cur) == 3: tris.append(tuple(cur))
            cur = []
    return tris

def mid(a, b): return ((a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5, (a[2] + b[2]) * 0.5)

def subdivide(tris):
    out = []
    for a, b, c in tris:
        ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
        out += [(a, ab, ca), (ab, b, bc), (ca, bc, c), (ab, bc, ca)]   # same winding as the parent
    return out

def main():
    here = os.path.dirname(os.path.abspath(__file__))
    scenes = os.path.join(here, "..", "scenes")
    dst = os.path.join(scenes, "bun69k.cli")
    if os.path.exists(dst) and "--force" not in sys.argv:
        print("bun69k.cli exists, kept"); return
    tris = read_tris(os.path.join(scenes, "bun500.cli"))
    for _ in range(3): tris = subdivide(tris)
    with open(dst, "w") as f:
        f.write("# STAND-IN for the reference's bun69k.cli: bun500.cli subdivided 3x (%d triangles); see tools/make_bun69k.py\n" % len(tris))
        for tri in tris:
            f.write("begin\n")
            for v in tri: f.write("vertex %r %r %r\n" % v)
            f.write("end\n\n")
    print("wrote", dst, len(tris), "triangles")

if __name__ == "__main__":
    main()
