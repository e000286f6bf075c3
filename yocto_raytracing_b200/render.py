"""Host-side mirror of the reference's render interface on top of the C ABI.

Reference                                    here
  build_bvh(scn, false)   (raytrace.cpp:278)   Scene(flat)                      -> yrt_scene_create
  raytrace(scn, amb, resolution, samples)      Scene.raytrace(amb, resolution, samples)   -> yrt_render
                          (raytrace.cpp:213)     (same argument meaning: amb scalar -> {amb,amb,amb},
                                                  resolution = image height, samples = per-axis count)
  intersect_first / intersect_any (scene.h:236) Scene.intersect_first / intersect_any
  tonemap(hdr, 0, false)  (image.cpp:55)       tonemap(img)

Everything computes on the GPU through libyrt_b200.so; there is no CPU path (load() / init() raise).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import Stats, YrtError, check
from .scene import FlatScene


def device_count() -> int:
    return _lib.load().yrt_device_count()


def init(n_gpus: int = 1) -> None:
    """Use GPUs 0..n_gpus-1 of this process for Scene.raytrace (interleaved row tiles, gathered on GPU 0)."""
    check(_lib.load().yrt_init(int(n_gpus)))


def init_device(device: int) -> None:
    """One process per GPU (torchrun): bind this process to `device`."""
    check(_lib.load().yrt_init_device(int(device)))


class Scene:
    """Device-resident scene + two-level LBVH (replaces build_bvh, src/scene.cpp:554)."""

    def __init__(self, flat: FlatScene):
        self.flat = flat
        self._desc = flat.desc()
        self._h = C.c_void_p()
        check(_lib.load().yrt_scene_create(C.byref(self._desc), C.byref(self._h)))
        self._cam = flat.camera_struct()

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.load().yrt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def info(self) -> dict:
        out = (C.c_int64 * 8)()
        check(_lib.load().yrt_scene_info(self._h, out))
        keys = ("blas_nodes", "tlas_nodes", "blas_depth", "tlas_depth", "lights", "prims", "build_us", "reflective_materials")
        return dict(zip(keys, list(out)))

    def image_size(self, resolution: int) -> Tuple[int, int]:
        """(width, height) the reference allocates: (int)round(aspect*resolution) x resolution (raytrace.cpp:216)."""
        return _lib.load().yrt_image_width(C.byref(self._cam), int(resolution)), int(resolution)

    # ---- raytrace() -------------------------------------------------------------------------
    def render(self, width: int, height: int, samples: int, amb=0.1, out: Optional[np.ndarray] = None,
               want_stats: bool = True):
        """Render into a HOST float32 (height, width, 4) array through yrt_render; returns (image, Stats|None)."""
        if out is None:
            out = np.empty((height, width, 4), np.float32)
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and out.size == width * height * 4
        a = np.broadcast_to(np.asarray(amb, np.float32), (3,))
        amb3 = (C.c_float * 3)(*[float(x) for x in a])
        st = Stats() if want_stats else None
        check(_lib.load().yrt_render(self._h, C.byref(self._cam), amb3, int(width), int(height), int(samples),
                                     C.c_void_p(out.ctypes.data), C.byref(st) if st is not None else None))
        return out, st

    def render_ldr(self, width: int, height: int, samples: int, amb=0.1, want_stats: bool = False):
        """raytrace() + tonemap(hdr, 0, false) on the device -> uint8 (height, width, 4), what the reference saves as PNG."""
        out = np.empty((height, width, 4), np.uint8)
        a = np.broadcast_to(np.asarray(amb, np.float32), (3,))
        amb3 = (C.c_float * 3)(*[float(x) for x in a])
        st = Stats() if want_stats else None
        check(_lib.load().yrt_render_ldr(self._h, C.byref(self._cam), amb3, int(width), int(height), int(samples),
                                         C.c_void_p(out.ctypes.data), None, C.byref(st) if st is not None else None))
        return out, st

    def truncated_paths(self) -> int:
        """yrt_frame_truncated_paths: mirror bounces the last frame dropped at the recursion cap (0 = the frame is what the
        reference's unbounded recursion gives, src/raytrace.cpp:190-204)."""
        n = C.c_int64(0)
        check(_lib.load().yrt_frame_truncated_paths(self._h, C.byref(n)))
        return int(n.value)

    def raytrace(self, amb: float, resolution: int, samples: int) -> np.ndarray:
        """image4f raytrace(scn, {amb,amb,amb}, resolution, samples) — src/raytrace.cpp:213."""
        w, h = self.image_size(resolution)
        img, _ = self.render(w, h, samples, amb, want_stats=False)
        return img

    def render_rows_into(self, d_ptr: int, width: int, height: int, samples: int, amb=0.1, tile_rows: int = 16, rank: int = 0,
                         world: int = 1, stream: int = 0, want_stats: bool = False):
        """yrt_render_rows: this rank's interleaved row tiles into DEVICE memory at d_ptr (packed rows)."""
        a = np.broadcast_to(np.asarray(amb, np.float32), (3,))
        amb3 = (C.c_float * 3)(*[float(x) for x in a])
        st = Stats() if want_stats else None
        check(_lib.load().yrt_render_rows(self._h, C.byref(self._cam), amb3, int(width), int(height), int(samples), int(tile_rows),
                                          int(rank), int(world), C.c_void_p(d_ptr), C.c_void_p(stream),
                                          C.byref(st) if st is not None else None))
        return st

    def render_rows_into_frame(self, d_full: int, width: int, height: int, samples: int, amb=0.1, tile_rows: int = 1, rank: int = 0,
                               world: int = 1, stream: int = 0, want_stats: bool = False):
        """yrt_render_rows_into_frame: this rank's rows straight into the FULL frame at d_full (own or peer memory)."""
        a = np.broadcast_to(np.asarray(amb, np.float32), (3,))
        amb3 = (C.c_float * 3)(*[float(x) for x in a])
        st = Stats() if want_stats else None
        check(_lib.load().yrt_render_rows_into_frame(self._h, C.byref(self._cam), amb3, int(width), int(height), int(samples),
                                                     int(tile_rows), int(rank), int(world), C.c_void_p(d_full), C.c_void_p(stream),
                                                     C.byref(st) if st is not None else None))
        return st

    def render_rows_to_host(self, h_ptr: int, width: int, height: int, samples: int, amb=0.1, tile_rows: int = 1, rank: int = 0,
                            world: int = 1, stream: int = 0, want_stats: bool = False):
        """yrt_render_rows_to_host: this rank's rows to their final positions of a HOST frame at h_ptr (whole image, float4)."""
        a = np.broadcast_to(np.asarray(amb, np.float32), (3,))
        amb3 = (C.c_float * 3)(*[float(x) for x in a])
        st = Stats() if want_stats else None
        check(_lib.load().yrt_render_rows_to_host(self._h, C.byref(self._cam), amb3, int(width), int(height), int(samples), int(tile_rows),
                                                  int(rank), int(world), C.c_void_p(h_ptr), C.c_void_p(stream),
                                                  C.byref(st) if st is not None else None))
        return st

    def set_camera(self, cam16) -> None:
        """Replace the camera (16 floats: frame x, y, z, o, then fovy, aspect, aperture, focus) used by the render calls."""
        c = np.asarray(cam16, np.float32).reshape(16)
        for i in range(12):
            self._cam.frame[i] = float(c[i])
        self._cam.fovy, self._cam.aspect, self._cam.aperture, self._cam.focus = (float(x) for x in c[12:16])

    def prepare(self, width: int, height: int, samples: int) -> None:
        """yrt_scene_prepare: allocate the render workspace of that frame size now."""
        check(_lib.load().yrt_scene_prepare(self._h, int(width), int(height), int(samples)))

    def debug_nodes(self, arity: int) -> np.ndarray:
        """yrt_debug_read_nodes: the node records (arity 2 or 4) of device 0 as a float32 array (n, 4)."""
        lib = _lib.load()
        n = lib.yrt_debug_read_nodes(self._h, int(arity), None, 0)
        if n < 0:
            check(int(n))
        out = np.zeros((n, 4), np.float32)
        if n:
            r = lib.yrt_debug_read_nodes(self._h, int(arity), C.c_void_p(out.ctypes.data), n)
            if r < 0:
                check(int(r))
        return out

    def stats_begin(self) -> None:
        """Open deferred statistics: following render_rows_into frames record events/counters without host syncs."""
        check(_lib.load().yrt_stats_begin(self._h))

    def stats_end(self) -> Stats:
        """Wait for the device and return the totals over the frames since stats_begin (Stats.frames of them)."""
        st = Stats()
        check(_lib.load().yrt_stats_end(self._h, C.byref(st)))
        return st

    # ---- queries ------------------------------------------------------------------------------
    def trace_primary(self, width: int, height: int, samples: int):
        """(ids[n,3] = (instance, shape, element) or -1, dist[n], uv[n,2]) per primary ray, ray order
        ((j*width+i)*samples+jj)*samples+ii."""
        n = width * height * samples * samples
        ids = np.empty((n, 3), np.int32)
        dist = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32)
        check(_lib.load().yrt_trace_primary(self._h, C.byref(self._cam), int(width), int(height), int(samples),
                                            C.c_void_p(ids.ctypes.data), C.c_void_p(dist.ctypes.data), C.c_void_p(uv.ctypes.data)))
        return ids, dist, uv

    def intersect_first(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        ids = np.empty((n, 3), np.int32)
        dist = np.empty(n, np.float32)
        uv = np.empty((n, 2), np.float32)
        check(_lib.load().yrt_intersect_first(self._h, C.c_void_p(rays.ctypes.data), C.c_int64(n), C.c_void_p(ids.ctypes.data),
                                              C.c_void_p(dist.ctypes.data), C.c_void_p(uv.ctypes.data)))
        return ids, dist, uv

    def intersect_any(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        occ = np.empty(n, np.uint8)
        check(_lib.load().yrt_intersect_any(self._h, C.c_void_p(rays.ctypes.data), C.c_int64(n), C.c_void_p(occ.ctypes.data)))
        return occ


def tonemap(img: np.ndarray) -> np.ndarray:
    """tonemap(hdr, 0, false) (src/image.cpp:55-78) on the device: float32 (h,w,4) -> uint8 (h,w,4)."""
    img = np.ascontiguousarray(img, np.float32)
    h, w = img.shape[:2]
    out = np.empty((h, w, 4), np.uint8)
    check(_lib.load().yrt_tonemap(C.c_void_p(img.ctypes.data), w, h, C.c_void_p(out.ctypes.data)))
    return out


def write_png(path: str, rgba8: np.ndarray, threads: int = 0, level: int = 1) -> None:
    """yrt_write_png: the RGBA8 frame (h,w,4) as a PNG encoded on `threads` host threads (0 = all); decodes to the same
    pixels as the reference's save_image (src/image.cpp:41-44).  Host only."""
    a = np.ascontiguousarray(rgba8, np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("write_png expects an (h, w, 4) uint8 array")
    check(_lib.load().yrt_write_png(str(path).encode(), C.c_void_p(a.ctypes.data), int(a.shape[1]), int(a.shape[0]), int(threads), int(level)))
