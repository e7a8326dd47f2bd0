"""ctypes wrapper over oracle/liborc.so -- ORACLE, test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (distraytracer_old_b200) never does.
PARITY UNPINNED: the reference has no tests and cannot run here (no JDK); see DESIGN.md.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "liborc.so")
SCENES = os.path.join(ROOT, "scenes")
TEX = os.path.join(SCENES, "txtrs_argb")


class _Opts(C.Structure):
    _fields_ = [("cols", C.c_int), ("rows", C.c_int), ("spp", C.c_int), ("literal_renorm", C.c_int),
                ("seed", C.c_uint64), ("photons", C.c_longlong), ("data_dir", C.c_char_p), ("tex_dir", C.c_char_p)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_load.restype = C.c_void_p
        L.orc_load.argtypes = [C.c_char_p, C.POINTER(_Opts), C.c_char_p, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.orc_emit_photons.restype = C.c_longlong
        L.orc_emit_photons.argtypes = [C.c_void_p]
        L.orc_get_photons.restype = C.c_longlong
        L.orc_get_photons.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong]
        L.orc_photon_probe.restype = C.c_longlong
        L.orc_photon_probe.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]
        L.orc_render.restype = C.c_double
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
        L.orc_render_chunks.restype = C.c_double
        L.orc_render_chunks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_trace_rays.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_dump_bvh.restype = C.c_longlong
        L.orc_dump_bvh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        L.orc_eval_texture.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_java_random.argtypes = [C.c_longlong, C.c_int, C.c_void_p]
        L.orc_perlin.restype = C.c_float
        L.orc_perlin.argtypes = [C.c_float, C.c_float, C.c_float]
        L.orc_obj_ctm.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_warnings.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


STAT_NAMES = ["primary", "shadow", "reflect", "refract", "photon_seg", "box_tests", "prim_tests",
              "box_tests_primary", "prim_tests_primary", "photons_stored"]


class OracleScene:
    """One parsed scene in the CPU oracle (mirrors myRTFileReader.readRTFile + myScene.draw)."""

    def __init__(self, scene_file, cols=300, rows=300, spp=-1, seed=0x5EED, literal_renorm=False,
                 photons=-1, data_dir=SCENES, tex_dir=TEX):
        L = lib()
        o = _Opts(cols, rows, spp, int(literal_renorm), seed, photons, data_dir.encode(), tex_dir.encode())
        err = C.create_string_buffer(512)
        self.h = L.orc_load(scene_file.encode(), C.byref(o), err, 512)
        if not self.h:
            raise RuntimeError("oracle: " + err.value.decode())
        info = (C.c_int * 8)()
        L.orc_info(self.h, info)
        (self.cols, self.rows, self.spp, self.n_objs, self.n_lights, self.n_prims, self.n_insts, self.photon_kind) = list(info)

    def emit_photons(self):
        return lib().orc_emit_photons(self.h)

    def photons(self):
        n = self.emit_photons()
        out = np.zeros((max(n, 1), 6), dtype=np.float64)
        m = lib().orc_get_photons(self.h, out.ctypes.data, n)
        return out[:m]

    def photon_probe(self, pts):
        """find_near at world points: rows of {sum r, sum g, sum b, d2 of the farthest of the k}."""
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.zeros((pts.shape[0], 4), dtype=np.float64)
        lib().orc_photon_probe(self.h, pts.shape[0], pts.ctypes.data, out.ctypes.data)
        return out

    def render(self, rect=None, threads=1, want=("argb", "hit_prim", "hit_inst", "rgb", "t")):
        x0, y0, x1, y1 = rect if rect else (0, 0, self.cols, self.rows)
        n = (x1 - x0) * (y1 - y0)
        res = {}
        bufs = {}
        for name, dt, k in (("argb", np.int32, 1), ("hit_prim", np.int32, 1), ("hit_inst", np.int32, 1), ("rgb", np.float64, 3), ("t", np.float64, 1)):
            bufs[name] = np.zeros((y1 - y0, x1 - x0) + ((k,) if k > 1 else ()), dtype=dt) if name in want else None
        st = np.zeros(10, dtype=np.uint64)
        p = lambda a: a.ctypes.data if a is not None else None
        secs = lib().orc_render(self.h, x0, y0, x1, y1, threads, p(bufs["argb"]), p(bufs["hit_prim"]), p(bufs["hit_inst"]), p(bufs["rgb"]), p(bufs["t"]), st.ctypes.data)
        for k, v in bufs.items():
            if v is not None:
                res[k] = v
        res["seconds"] = secs
        res["stats"] = dict(zip(STAT_NAMES, [int(x) for x in st]))
        res["pixels"] = n
        return res

    def render_chunks(self, chunk_rows, stride, phase=0, threads=1, want_argb=True):
        """Interleaved row chunks {c : c % stride == phase}: the pixel set of GPU rank `phase` of `stride` (compact, chunks back to back)."""
        n_chunks = len(range(phase, (self.rows + chunk_rows - 1) // chunk_rows, stride))
        argb = np.zeros((n_chunks * chunk_rows, self.cols), dtype=np.int32) if want_argb else None
        st = np.zeros(10, dtype=np.uint64)
        secs = lib().orc_render_chunks(self.h, chunk_rows, stride, phase, threads, argb.ctypes.data if argb is not None else None, st.ctypes.data)
        return {"argb": argb, "seconds": secs, "stats": dict(zip(STAT_NAMES, [int(x) for x in st]))}

    def trace_rays(self, org, dirs):
        org = np.ascontiguousarray(org, dtype=np.float64)
        dirs = np.ascontiguousarray(dirs, dtype=np.float64)
        n = org.shape[0]
        ids = np.zeros((n, 2), dtype=np.int32)
        t = np.zeros(n, dtype=np.float64)
        lib().orc_trace_rays(self.h, n, org.ctypes.data, dirs.ctypes.data, ids.ctypes.data, t.ctypes.data)
        return ids, t

    def dump_bvh(self, obj_idx):
        box = np.zeros(6)
        n = lib().orc_dump_bvh(self.h, obj_idx, None, 0, box.ctypes.data)
        if n < 0:
            return None, None
        out = np.zeros(n, dtype=np.int32)
        lib().orc_dump_bvh(self.h, obj_idx, out.ctypes.data, n, box.ctypes.data)
        return out, box

    def eval_texture(self, shader_serial, hit_loc, fwd_loc=None):
        hit_loc = np.ascontiguousarray(hit_loc, dtype=np.float64)
        fwd_loc = hit_loc if fwd_loc is None else np.ascontiguousarray(fwd_loc, dtype=np.float64)
        out = np.zeros_like(hit_loc)
        r = lib().orc_eval_texture(self.h, shader_serial, hit_loc.shape[0], hit_loc.ctypes.data, fwd_loc.ctypes.data, out.ctypes.data)
        if r != 0:
            raise IndexError("no such shader" if r == -1 else "image textures cannot be probed")
        return out

    def obj_ctm(self, idx):
        m = np.zeros(16)
        if lib().orc_obj_ctm(self.h, idx, m.ctypes.data) != 0:
            raise IndexError(idx)
        return m.reshape(4, 4)

    def warnings(self):
        b = C.create_string_buffer(8192)
        lib().orc_warnings(self.h, b, 8192)
        return [w for w in b.value.decode().split("\n") if w]


def argb_to_rgb8(argb):
    a = argb.astype(np.uint32)
    return np.stack([(a >> 16) & 255, (a >> 8) & 255, a & 255], axis=-1).astype(np.uint8)


def u01(seed, stream, a, b, c, d):
    return lib().orc_u01(seed, stream, a, b, c, d)


def java_random(seed, n):
    out = np.zeros(n)
    lib().orc_java_random(seed, n, out.ctypes.data)
    return out


def perlin(x, y, z):
    return lib().orc_perlin(x, y, z)
