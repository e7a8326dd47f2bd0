#!/usr/bin/env python
"""Synthetic scaling scenes of SURVEY 8(d), written as ordinary .cli files into scenes/gen/ (generated, git-ignored):
  soup N        N uniformly random triangles (centre ~U[-1,1]^3, edge vectors ~U[-s,s]^3, s = N^(-1/3), seed 1234) in one BVH
                under the config-2 camera / lights / floor (data/p3_t09.cli without the texture)
  grid K        K x K x K instances of the bun69k stand-in inside one instance BVH (begin_list ... end_accel)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "scenes", "gen")
HEAD = """fov 60
background 0.2 0.2 1
point_light  3 4  0  .8 .8 .8
point_light -3 4  0  .2 .2 .2
diffuse  .8 .8 .8  .2 .2 .2
begin
vertex -100 -1 -100
vertex  100 -1 -100
vertex  100 -1  100
end
begin
vertex  100 -1  100
vertex -100 -1  100
vertex -100 -1 -100
end
"""


def soup(n):
    os.makedirs(GEN, exist_ok=True)
    rng = np.random.default_rng(1234)
    s = n ** (-1.0 / 3.0)
    c = rng.uniform(-1, 1, size=(n, 3))
    e1 = rng.uniform(-s, s, size=(n, 3))
    e2 = rng.uniform(-s, s, size=(n, 3))
    v = np.stack([c, c + e1, c + e2], axis=1)                      # [n, 3, 3]
    mesh = os.path.join(GEN, "soup_%d_mesh.cli" % n)
    with open(mesh, "w") as f:
        for t in v:
            f.write("begin\nvertex %.9g %.9g %.9g\nvertex %.9g %.9g %.9g\nvertex %.9g %.9g %.9g\nend\n" % tuple(t.reshape(-1)))
    path = os.path.join(GEN, "soup_%d.cli" % n)
    with open(path, "w") as f:
        f.write(HEAD + "translate 0 0 -3\ndiffuse 0.9 0.9 0.9 0 0 0\nbegin_list\nread gen/soup_%d_mesh.cli\nend_accel\nwrite soup.png\n" % n)
    return path


def grid(k):
    os.makedirs(GEN, exist_ok=True)
    path = os.path.join(GEN, "grid_%d.cli" % k)
    sc = 0.9 / k
    with open(path, "w") as f:
        f.write(HEAD + "diffuse 0.9 0.9 0.9 0 0 0\nbegin_list\nread bun69k.cli\nend_accel\nnamed_object bun\npush\ntranslate 0 0 -3.5\nbegin_list\n")
        for i in range(k):
            for j in range(k):
                for l in range(k):
                    x, y, z = ((a + 0.5) / k * 2 - 1 for a in (i, j, l))
                    f.write("push\ntranslate %.9g %.9g %.9g\nscale %.9g %.9g %.9g\ninstance bun\npop\n" % (x, y, z, sc, sc, sc))
        f.write("end_accel\npop\nwrite grid.png\n")
    return path


if __name__ == "__main__":
    kind, n = sys.argv[1], int(sys.argv[2])
    print({"soup": soup, "grid": grid}[kind](n))
