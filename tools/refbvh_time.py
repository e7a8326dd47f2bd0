#!/usr/bin/env python
"""Times the object ordering of the reference-topology median-split BVH: device builder (csrc/refbvh.cuh) vs host recursion, random keys.
  python tools/refbvh_time.py [n ...]   ->  one JSON line per n (wall clock through the C ABI: H2D of the keys and D2H of the order included)"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import distraytracer_old_b200 as d

ctx = d.Context(device=0, cols=16, rows=16)
for n in [int(a) for a in sys.argv[1:]] or [65536, 262144, 1 << 20, 1 << 22, 1 << 24]:
    keys = np.random.default_rng(n).normal(size=(n, 3))
    ctx.bvh_order(keys[:4096], True)                       # warm-up (allocations are grow-only, the second call of a size reuses them)
    ctx.bvh_order(keys, True)
    t = time.perf_counter(); od = ctx.bvh_order(keys, True); td = time.perf_counter() - t
    th = None
    if n <= 1 << 22:
        t = time.perf_counter(); oh = ctx.bvh_order(keys, False); th = time.perf_counter() - t
        assert (od == oh).all()
    print(json.dumps({"objects": n, "device_ms": round(1e3 * td, 2), "host_ms": None if th is None else round(1e3 * th, 1), "identical": th is not None}), flush=True)
