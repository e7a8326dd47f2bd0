#!/usr/bin/env python
"""Regenerates tests/golden/golden.npz from the CPU oracle.

PROVENANCE: the reference ships no tests, golden vectors or known-answer files, and no JVM exists in the
authoring container, so these vectors are ORACLE outputs frozen at authoring time (they pin the oracle and the
CUDA path against regressions; they are not reference-generated).  The only reference-produced artefact that
maps to a scene in the checkout is the 300x300 render t11c.png (data/t11.cli:84), kept here verbatim as
ref_t11c.png and used as a loose PSNR pin of the oracle (tests/test_oracle.py).
Independent known answers also checked by the tests: java.util.Random(0).nextDouble() = 0.730967787376657
(published JDK behaviour) and the Philox4x32-10 reference vectors of Salmon et al. (Random123 kat_vectors).
"""
import os, sys, zlib
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import orc

TEX_PROBES = [("p4_st0%d" % i, 2) for i in range(1, 10)] + [("p3_t08", 2), ("p4_t03", 0), ("p4_t04", 0), ("p4_t02", 0)]


def main():
    g = {}
    for name, n in (("t01", 64), ("t03", 64), ("p3_t08", 64), ("p3_t02_sierp", 64), ("p3_t12", 64), ("p3_t06", 48), ("c5Fish", 48), ("planets3Ortho", 40), ("p2_t06", 32), ("t06", 48)):
        s = orc.OracleScene(name + ".cli", cols=n, rows=n)
        r = s.render()
        g[name + "_argb"] = r["argb"]; g[name + "_hit_prim"] = r["hit_prim"]; g[name + "_hit_inst"] = r["hit_inst"]
        g[name + "_stats"] = np.array([r["stats"][k] for k in ("primary", "shadow", "reflect", "refract")], dtype=np.int64)
    s = orc.OracleScene("p3_t08.cli")
    d, box = s.dump_bvh(2); g["p3_t08_bvh_crc"] = np.array([zlib.crc32(d.tobytes())], dtype=np.int64); g["p3_t08_bvh_box"] = box
    s = orc.OracleScene("p3_t02_sierp.cli")
    d, box = s.dump_bvh(0); g["sierp_bvh_crc"] = np.array([zlib.crc32(d.tobytes())], dtype=np.int64); g["sierp_bvh_box"] = box; g["sierp_bvh_head"] = d[:64]
    rng = np.random.default_rng(7)
    pts = rng.uniform(-3, 3, size=(64, 3))
    g["probe_pts"] = pts
    g["perlin"] = np.array([orc.perlin(*[np.float32(v) for v in p]) for p in pts], dtype=np.float32)
    for sc, serial in TEX_PROBES:
        o = orc.OracleScene(sc + ".cli")
        g["tex_" + sc] = o.eval_texture(serial, pts)
    g["philox"] = np.array([orc.u01(0x5EED, 0, a, b, 1, d) for a in (0, 1, 77777) for b in (0, 3) for d in (0, 1, 20)])
    g["javarand0"] = orc.java_random(0, 4); g["javarand42"] = orc.java_random(42, 4)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **g)
    print("wrote golden.npz with", len(g), "arrays")

if __name__ == "__main__":
    main()
