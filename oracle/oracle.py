"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/liboracle.so (the plain-C restatement of the reference's
render path, oracle/yrt_oracle.c).  Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs;
the product package yocto_raytracing_b200 never imports it."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    """Compile the C restatement (gcc, baseline x86-64, no FMA contraction, no fast-math)."""
    src = os.path.join(_HERE, "yrt_oracle.c")
    if os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(_HERE, "yrt_oracle.h"))):
        return
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp", "-o", LIB_PATH, src, "-lm"],
                   check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
    return _lib


class OracleScene:
    def __init__(self, flat):
        self.flat = flat
        self._desc = flat.desc()
        self.h = C.c_void_p()
        st = lib().oracle_scene_create(C.byref(self._desc), C.byref(self.h))
        if st != 0:
            raise RuntimeError(f"oracle_scene_create failed: {st}")

    def __del__(self):
        try:
            if self.h:
                lib().oracle_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def info(self):
        out = (C.c_int64 * 4)()
        lib().oracle_scene_info(self.h, out)
        return {"scene_nodes": out[0], "shape_nodes": out[1], "lights": out[2]}

    def render(self, width, height, samples, amb=0.1, max_depth=0, threads=1, rows=None):
        """raytrace() (src/raytrace.cpp:213).  Returns (float32 image H x W x 4, counts dict)."""
        img = np.zeros((height, width, 4), np.float32)
        cam = self.flat.camera_struct()
        a = (C.c_float * 3)(amb, amb, amb)
        cnt = (C.c_int64 * 4)()
        r0, r1 = rows if rows is not None else (0, height)
        st = lib().oracle_render_rows(self.h, C.byref(cam), a, width, height, samples, max_depth, threads, r0, r1,
                                      C.c_void_p(img.ctypes.data), cnt)
        if st != 0:
            raise RuntimeError(f"oracle_render failed: {st}")
        return img, {"primary_rays": cnt[0], "reflection_rays": cnt[1], "shadow_rays": cnt[2], "max_depth": cnt[3]}

    def trace_primary(self, width, height, samples, brute_force=False):
        n = width * height * samples * samples
        ids = np.empty((n, 3), np.int32)
        dist = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32)
        cam = self.flat.camera_struct()
        st = lib().oracle_trace_primary(self.h, C.byref(cam), width, height, samples, int(brute_force), C.c_void_p(ids.ctypes.data),
                                        C.c_void_p(dist.ctypes.data), C.c_void_p(uv.ctypes.data))
        assert st == 0
        return ids, dist, uv

    def intersect_first(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        ids = np.empty((n, 3), np.int32)
        dist = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32)
        st = lib().oracle_intersect_first(self.h, C.c_void_p(rays.ctypes.data), C.c_int64(n), C.c_void_p(ids.ctypes.data),
                                          C.c_void_p(dist.ctypes.data), C.c_void_p(uv.ctypes.data))
        assert st == 0
        return ids, dist, uv

    def intersect_any(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        occ = np.empty(n, np.uint8)
        st = lib().oracle_intersect_any(self.h, C.c_void_p(rays.ctypes.data), C.c_int64(n), C.c_void_p(occ.ctypes.data))
        assert st == 0
        return occ


def tonemap(img):
    """tonemap(hdr, 0, false) of src/image.cpp:55-78 with the host libm."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape[:2]
    out = np.empty((h, w, 4), np.uint8)
    lib().oracle_tonemap(C.c_void_p(img.ctypes.data), w, h, C.c_void_p(out.ctypes.data))
    return out
