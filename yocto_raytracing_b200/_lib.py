"""ctypes binding of libyrt_b200.so (the C ABI of include/yrt_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, the
compute entry points raise.  PyTorch is not needed to use this module.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("YRT_B200_LIB") or os.path.join(_HERE, "libyrt_b200.so")   # override: A/B runs of build variants

YRT_OK = 0


class YrtError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"yrt error {status}: {message}")
        self.status = status
        self.message = message


class SceneDesc(C.Structure):
    """yrt_scene_desc (include/yrt_b200.h)."""
    _fields_ = [
        ("n_shapes", C.c_int32), ("n_instances", C.c_int32), ("n_materials", C.c_int32), ("n_textures", C.c_int32),
        ("n_verts", C.c_int32), ("n_elem_idx", C.c_int32),
        ("shape_kind", C.c_void_p), ("shape_elem_off", C.c_void_p), ("shape_elem_cnt", C.c_void_p),
        ("shape_vert_off", C.c_void_p), ("shape_vert_cnt", C.c_void_p), ("shape_has_uv", C.c_void_p),
        ("shape_has_radius", C.c_void_p),
        ("elem_idx", C.c_void_p), ("pos", C.c_void_p), ("norm", C.c_void_p), ("uv", C.c_void_p), ("radius", C.c_void_p),
        ("inst_frame", C.c_void_p), ("inst_shape", C.c_void_p), ("inst_mat", C.c_void_p),
        ("mat_ke", C.c_void_p), ("mat_kd", C.c_void_p), ("mat_ks", C.c_void_p), ("mat_kr", C.c_void_p), ("mat_rs", C.c_void_p),
        ("mat_kd_tex", C.c_void_p), ("mat_ks_tex", C.c_void_p),
        ("tex_w", C.c_void_p), ("tex_h", C.c_void_p), ("tex_off", C.c_void_p), ("tex_rgba8", C.c_void_p),
        ("tex_bytes", C.c_int64),
    ]


class Camera(C.Structure):
    """yrt_camera."""
    _fields_ = [("frame", C.c_float * 12), ("fovy", C.c_float), ("aspect", C.c_float), ("aperture", C.c_float),
                ("focus", C.c_float)]


class Stats(C.Structure):
    """yrt_stats."""
    _fields_ = [
        ("primary_rays", C.c_int64), ("reflection_rays", C.c_int64), ("shadow_rays", C.c_int64), ("launches", C.c_int64),
        ("ms_total", C.c_float), ("ms_trace_closest", C.c_float), ("ms_trace_any", C.c_float), ("ms_shade", C.c_float),
        ("ms_other", C.c_float), ("ms_gather", C.c_float), ("max_depth", C.c_int32), ("n_gpus", C.c_int32),
        ("n_closest", C.c_int32), ("n_any", C.c_int32), ("n_shade", C.c_int32), ("n_other", C.c_int32),
        ("frames", C.c_int32), ("reserved", C.c_int32), ("truncated_paths", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}

    @property
    def total_rays(self) -> int:
        return self.primary_rays + self.reflection_rays + self.shadow_rays


# every symbol include/yrt_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "yrt_abi_version": (C.c_int, []),
    "yrt_last_error": (C.c_char_p, []),
    "yrt_device_count": (C.c_int, []),
    "yrt_init": (C.c_int, [C.c_int]),
    "yrt_init_device": (C.c_int, [C.c_int]),
    "yrt_scene_create": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(C.c_void_p)]),
    "yrt_scene_destroy": (None, [C.c_void_p]),
    "yrt_scene_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "yrt_image_width": (C.c_int, [C.POINTER(Camera), C.c_int]),
    "yrt_desc_nonrigid_instances": (C.c_int, [C.POINTER(SceneDesc)]),
    "yrt_write_png": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "yrt_render": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_void_p,
                             C.POINTER(Stats)]),
    "yrt_render_ldr": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.POINTER(Stats)]),
    "yrt_render_rows": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "yrt_rows_owned": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "yrt_render_rows_to_host": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "yrt_render_rows_into_frame": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Stats)]),
    "yrt_host_barrier": (C.c_int, [C.c_void_p, C.c_int, C.c_int64]),
    "yrt_frame_alloc": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "yrt_frame_free": (C.c_int, [C.c_void_p]),
    "yrt_frame_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "yrt_frame_import": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "yrt_frame_release": (C.c_int, [C.c_void_p]),
    "yrt_frame_barrier": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    "yrt_stats_begin": (C.c_int, [C.c_void_p]),
    "yrt_stats_end": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "yrt_unpack_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "yrt_trace_primary": (C.c_int, [C.c_void_p, C.POINTER(Camera), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "yrt_intersect_first": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "yrt_intersect_any": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "yrt_tonemap": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "yrt_scene_prepare": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "yrt_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "yrt_frame_truncated_paths": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "yrt_counters_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "yrt_debug_sort_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "yrt_debug_read_nodes": (C.c_int64, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64]),
}

_lib = None


def load() -> C.CDLL:
    """Load libyrt_b200.so (built by `make lib` / __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not found: build it with `make lib` (or __graft_entry__.build()). "
            "There is no CPU fallback for the render path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.yrt_abi_version() != 3:
        raise RuntimeError("libyrt_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != YRT_OK:
        msg = load().yrt_last_error()
        raise YrtError(status, msg.decode("utf-8", "replace") if msg else "")
