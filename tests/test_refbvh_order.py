"""Object order of the reference's median-split BVH (myBVH.addObjList / buildSortedObjAras, myGeomBase.java:338-386;
DistRayTracer.getIDXofMaxBVHSpan, DistRayTracer.java:409-418).

CPU: the library's host recursion (one in-place stable sort per node) against a LITERAL restatement of the reference's three-lists-per-node
algorithm written here in Python.  GPU: the device builder (csrc/refbvh.cuh, segmented radix sorts) against the host recursion, bit for bit,
and whole scenes: BVH dumps of a device-ordered context == host-only context."""
import os
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def java_key(v):
    """Double.compareTo order: -0.0 < 0.0 (no NaNs in these tests)"""
    b = struct.unpack("<q", struct.pack("<d", float(v)))[0]
    return (float(v), b)


def literal_order(keys):
    """The reference, literally: every node holds three lists, list i = the parent's split-axis sublist re-sorted on coordinate i through a
    TreeMap<Double, List> (equal keys keep their arrival order); split axis = widest (last - first) of the three lists, strict '<' from -1;
    leaves keep list 0; objListSize at the root is size - 1 (Q2)."""
    n = len(keys)

    def resort(lst, skip):
        out = [None] * 3
        for i in range(3):
            if i == skip:
                out[i] = list(lst)
                continue
            groups = {}
            for o in lst:
                groups.setdefault(java_key(keys[o][i]), []).append(o)
            out[i] = [o for k in sorted(groups) for o in groups[k]]
        return out

    res = []

    def add(lists, count):
        if count <= 5:
            res.extend(lists[0])
            return
        split = int(.5 * count)
        widest, axis = -1.0, -1
        for i in range(3):
            d = keys[lists[i][-1]][i] - keys[lists[i][0]][i]
            if widest < d:
                widest, axis = d, i
        add(resort(lists[axis][:split], axis), split)
        add(resort(lists[axis][split:count], axis), count - split)
        return lists[axis][count:]

    dropped = add(resort(list(range(n)), -1), n - 1) or []
    return np.array(res + list(dropped), dtype=np.int32)


def key_sets():
    rng = np.random.default_rng(7)
    out = []
    for n in (1, 2, 5, 6, 7, 11, 12, 13, 23, 24, 25, 47, 100, 257, 1000):
        out.append(("uniform_%d" % n, rng.uniform(-3, 3, size=(n, 3))))
        out.append(("ties_%d" % n, rng.integers(-2, 3, size=(n, 3)).astype(np.float64)))          # many equal keys: arrival order decides
    z = rng.integers(-1, 2, size=(300, 3)).astype(np.float64) * 0.0
    z[::3] = -0.0                                                                                   # -0.0 sorts before +0.0 in a TreeMap<Double>
    out.append(("signed_zeros", z))
    g = np.stack(np.meshgrid(np.arange(9.0), np.arange(7.0), np.arange(5.0), indexing="ij"), -1).reshape(-1, 3)
    out.append(("lattice", g[rng.permutation(len(g))]))
    flat = rng.uniform(-1, 1, size=(500, 3)); flat[:, 1] = 0.25                                     # one axis of zero span
    out.append(("flat_y", flat))
    return out


@pytest.mark.parametrize("name,keys", key_sets(), ids=[k[0] for k in key_sets()])
def test_host_order_equals_literal_three_list_algorithm(drt, name, keys):
    ctx = drt.Context(device=-1, cols=16, rows=16)
    got = ctx.bvh_order(keys, on_device=False)
    ref = literal_order(keys)
    assert got.shape == ref.shape and (got == ref).all(), name
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,keys", key_sets()[2:], ids=[k[0] for k in key_sets()[2:]])
def test_device_order_equals_host_order(drt, gpu_ctx_factory, name, keys):
    ctx = gpu_ctx_factory(cols=16, rows=16)
    assert (ctx.bvh_order(keys, on_device=True) == ctx.bvh_order(keys, on_device=False)).all(), name
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,ties", [(65536, False), (65536, True), (1 << 20, False), (1 << 20, True), (3000001, False)])
def test_device_order_large(drt, gpu_ctx_factory, n, ties):
    rng = np.random.default_rng(n + ties)
    keys = rng.integers(-40, 41, size=(n, 3)).astype(np.float64) / 8 if ties else rng.normal(size=(n, 3))
    ctx = gpu_ctx_factory(cols=16, rows=16)
    d, h = ctx.bvh_order(keys, on_device=True), ctx.bvh_order(keys, on_device=False)
    assert (d == h).all()
    assert (np.sort(d) == np.arange(n)).all()              # a permutation
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["p3_t08.cli", "p3_t02_sierp.cli", "p3_t11_sierp.cli", "gen/soup_65536.cli"])
def test_device_built_scene_equals_host_built_scene(drt, gpu_ctx_factory, scene, monkeypatch):
    if scene.startswith("gen/") and not os.path.exists(os.path.join(ROOT, "scenes", scene)):
        import subprocess, sys
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_synth.py"), "soup", "65536"])
    monkeypatch.setenv("DRT_REFBVH_MIN", "8")              # every list of >= 8 objects is ordered on the device
    dev = gpu_ctx_factory(cols=32, rows=32)
    sd = drt.Scene.from_cli(dev, scene)
    host = drt.Context(device=-1, cols=32, rows=32)
    sh = drt.Scene.from_cli(host, scene)
    assert sd.build_info()["bvh_device_builds"] >= 1 and sh.build_info()["bvh_device_builds"] == 0
    n_top = sd.info()["top"]
    found = 0
    for t in range(n_top):
        a, b = sd.dump_bvh(t), sh.dump_bvh(t)
        assert (a[0] is None) == (b[0] is None)
        if a[0] is not None:
            found += 1
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all()          # node / leaf / object sequence and the root box
    assert found >= 1
    dev.close(); host.close()
