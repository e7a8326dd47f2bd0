#!/usr/bin/env python
"""Builds tuning variants of libdrt.so into build/variants/ (never the in-tree library):  python tools/build_variants.py name:-DA=1,-DB=2 ...
Every variant must pass tests/test_gpu_accel_modes.py before its numbers mean anything (DRT_LIB=<path> selects it)."""
import os, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distraytracer_old_b200 import build as B

def one(spec):
    name, _, defs = spec.partition(":")
    out = os.path.join(ROOT, "build", "variants", "libdrt_%s.so" % name)
    B.build(force=True, out=out, defines=[d[2:] for d in defs.split(",") if d.startswith("-D")])
    return out

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        for p in ex.map(one, sys.argv[1:]):
            print(p)
