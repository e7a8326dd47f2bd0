"""CPU restatement of the performance-mode acceleration structure (BASELINE north star, subsystem 2: "GPU LBVH from 30-bit Morton codes with
radix sort ... Morton/BVH ordering must be bit-exact").  TEST INFRASTRUCTURE: imported by tests/ only.

The reference has no LBVH -- its tree is the median split restated in oracle/orc_core.hpp (myGeomBase.java:338-386) -- so this file restates the
published algorithms the product's csrc/lbvh.cuh implements, independently, in numpy / plain Python:
  * 30-bit Morton code of the triangle's box centre inside the BVH's root box (10 bits per axis, x most significant), FP64 arithmetic
  * stable sort by code (equal codes keep the input order = the reference tree's DFS leaf order)
  * leaves = 4 consecutive sorted triangles, leaf code = code of its first triangle
  * Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees": radix tree over the keys (code, index)
  * leaf box = min/max over its vertices, padded outward by 1e-12 * max(1, |coordinate|); node boxes = fmin / fmax of the children
"""
import numpy as np


def _expand10(v):
    v = v.astype(np.uint64) & 1023
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v & 0xFFFFFFFF


def morton30(verts, box):
    """verts [n, 9] (three vertices), box = (min xyz, max xyz) of the BVH root"""
    v = verts.reshape(-1, 3, 3)
    lo, hi = v.min(axis=1), v.max(axis=1)
    c = 0.5 * (lo + hi)
    mn, mx = np.asarray(box[:3]), np.asarray(box[3:])
    ext = mx - mn
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.where(ext > 0, (c - mn) / ext, 0.0)
    q = np.minimum(np.maximum(u * 1024.0, 0.0), 1023.0).astype(np.uint32)          # C cast: truncation of a non-negative value
    return ((_expand10(q[:, 0]) << 2) | (_expand10(q[:, 1]) << 1) | _expand10(q[:, 2])).astype(np.uint32)


def build(verts, box):
    """-> order (stable Morton order of the input triangles), links [nLeaves-1, 2], boxes [nLeaves-1, 12] (lmin lmax rmin rmax)"""
    n = len(verts)
    code = morton30(verts, box)
    order = np.argsort(code, kind="stable")
    sv, sc = verts[order].reshape(n, 3, 3), code[order]
    nl = (n + 3) // 4
    leaf_code = [int(sc[4 * j]) for j in range(nl)]
    lmin = np.array([sv[4 * j:4 * j + 4].reshape(-1, 3).min(axis=0) for j in range(nl)])
    lmax = np.array([sv[4 * j:4 * j + 4].reshape(-1, 3).max(axis=0) for j in range(nl)])
    lmin = lmin - 1e-12 * np.maximum(1.0, np.abs(lmin))
    lmax = lmax + 1e-12 * np.maximum(1.0, np.abs(lmax))
    key = [(leaf_code[j] << 32) | j for j in range(nl)]

    def delta(i, j):
        if j < 0 or j >= nl:
            return -1
        x = key[i] ^ key[j]
        return 64 - x.bit_length()

    links = np.zeros((max(nl - 1, 0), 2), dtype=np.int32)
    for i in range(nl - 1):
        d = 1 if delta(i, i + 1) - delta(i, i - 1) >= 0 else -1
        dmin = delta(i, i - d)
        lmax_ = 2
        while delta(i, i + lmax_ * d) > dmin:
            lmax_ <<= 1
        l, t = 0, lmax_ >> 1
        while t >= 1:
            if delta(i, i + (l + t) * d) > dmin:
                l += t
            t >>= 1
        j = i + l * d
        dnode = delta(i, j)
        s, t = 0, l
        while True:
            t = (t + 1) >> 1
            if delta(i, i + (s + t) * d) > dnode:
                s += t
            if t <= 1:
                break
        gamma = i + s * d + min(d, 0)
        lo, hi = min(i, j), max(i, j)
        links[i, 0] = -(1 + gamma) if lo == gamma else gamma
        links[i, 1] = -(1 + gamma + 1) if hi == gamma + 1 else gamma + 1
    boxes = np.zeros((max(nl - 1, 0), 12))
    done = {}

    def box_of(ref):                      # iterative post-order (trees over sorted meshes are deep)
        stack = [(ref, False)]
        while stack:
            r, seen = stack.pop()
            if r < 0 or r in done:
                continue
            if not seen:
                stack.append((r, True)); stack.append((int(links[r, 0]), False)); stack.append((int(links[r, 1]), False))
            else:
                bb = []
                for c in (int(links[r, 0]), int(links[r, 1])):
                    bb.append((lmin[-1 - c], lmax[-1 - c]) if c < 0 else done[c])
                boxes[r, 0:3], boxes[r, 3:6], boxes[r, 6:9], boxes[r, 9:12] = bb[0][0], bb[0][1], bb[1][0], bb[1][1]
                done[r] = (np.minimum(bb[0][0], bb[1][0]), np.maximum(bb[0][1], bb[1][1]))
    if nl > 1:
        box_of(0)
    return order, links, boxes
