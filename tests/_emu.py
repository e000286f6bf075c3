"""ctypes binding of tests/host_emu/libyrt_hostemu.so (device code compiled for the CPU; tests only)."""
import ctypes as C
import os

import numpy as np

from yocto_raytracing_b200 import _lib

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu")
_PATH = os.path.join(_DIR, "libyrt_hostemu.so")


def _path(variant=""):
    return os.path.join(_DIR, f"libyrt_hostemu_{variant}.so") if variant else _PATH


def available(variant=""):
    return os.path.exists(_path(variant))


_libs = {}


def lib(variant=""):
    """The emulation library; variants: "bin" = both ray kinds on binary node records, "wide" = both on 4-wide records
    (the default build: closest-hit rays binary, any-hit rays 4-wide)."""
    if variant not in _libs:
        l = C.CDLL(_path(variant))
        l.emu_last_error.restype = C.c_char_p
        _libs[variant] = l
    return _libs[variant]


class EmuScene:
    def __init__(self, flat, leaf_blas=0, leaf_tlas=0, variant=""):
        self.flat = flat
        self._lib = lib(variant)
        self._desc = flat.desc()
        self.h = C.c_void_p()
        st = self._lib.emu_scene_create(C.byref(self._desc), leaf_blas, leaf_tlas, C.byref(self.h))
        if st != 0:
            raise RuntimeError(self._lib.emu_last_error().decode())

    def info(self):
        out = (C.c_int64 * 8)()
        self._lib.emu_scene_info(self.h, out)
        return list(out)

    def nodes(self, arity):
        """Node records (arity 2 or 4) of the emulated build as a float32 array (n, 4)."""
        self._lib.emu_read_nodes.restype = C.c_int64
        n = self._lib.emu_read_nodes(self.h, arity, None)
        out = np.zeros((max(n, 0), 4), np.float32)
        if n > 0:
            self._lib.emu_read_nodes(self.h, arity, C.c_void_p(out.ctypes.data))
        return out

    def trace_primary(self, width, height, samples):
        n = width * height * samples * samples
        ids = np.empty((n, 3), np.int32)
        dist = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32)
        ctr = (C.c_int64 * 8)()
        cam = self.flat.camera_struct()
        st = self._lib.emu_trace_primary(self.h, C.byref(cam), width, height, samples, C.c_void_p(ids.ctypes.data),
                                     C.c_void_p(dist.ctypes.data), C.c_void_p(uv.ctypes.data), ctr)
        assert st == 0
        return ids, dist, uv, list(ctr)

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        ids = np.empty((n, 3), np.int32)
        dist = np.empty(n, np.float32)
        occ = np.empty(n, np.uint8)
        fr = C.c_int64(0)
        st = self._lib.emu_intersect(self.h, C.c_void_p(rays.ctypes.data), C.c_int64(n), C.c_void_p(ids.ctypes.data), C.c_void_p(dist.ctypes.data),
                                 C.c_void_p(occ.ctypes.data), C.byref(fr))
        assert st == 0
        return ids, dist, occ, fr.value

    def render(self, width, height, samples, amb=0.1, max_depth=16):
        img = np.empty((height, width, 4), np.float32)
        cam = self.flat.camera_struct()
        a = (C.c_float * 3)(amb, amb, amb)
        rc = (C.c_int64 * 9)()
        st = self._lib.emu_render(self.h, C.byref(cam), a, width, height, samples, max_depth, C.c_void_p(img.ctypes.data), rc)
        assert st == 0
        return img, list(rc)

    def __del__(self):
        try:
            if self.h:
                self._lib.emu_scene_destroy(self.h)
        except Exception:
            pass


def set_grids(scene, light_R=0, cam_shift=-1):
    """Apex grids of the emulation on / off (csrc/yrt_pgrid.cuh): light_R > 0 = cube grids of light_R x light_R cells per face
    around the point lights, cam_shift >= 0 = a camera grid of 2^cam_shift-pixel cells per trace_primary / render call.
    Returns (light-grid list entries, cells not served, lights with a grid, chain nodes)."""
    st = (C.c_int64 * 4)()
    assert scene._lib.emu_set_grids(scene.h, light_R, cam_shift, st) == 0
    return st[0], st[1], st[2], st[3]
