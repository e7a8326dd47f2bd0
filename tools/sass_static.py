#!/usr/bin/env python
"""Static SASS instruction counts per source region of the lean kernels (no ncu capture needed): quick check of a kernel edit before GPU time."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_model as sm
ranges = sm.marker_ranges()
for k, mangled in sm.KERNELS.items():
    with tempfile.TemporaryDirectory() as wd:
        insts = sm.disassemble(os.path.join(sm.ROOT, "distraytracer_old_b200", "libdrt.so"), mangled, wd)
    c = {"node": 0, "leaf": 0, "fixed": 0}
    for i in insts:
        c[sm.region_of(i, ranges)] += 1
    print(k, len(insts), c)
