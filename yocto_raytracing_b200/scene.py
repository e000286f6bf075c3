"""Flattened scene container — the Python mirror of yrt_scene_desc (include/yrt_b200.h).

A FlatScene is produced either by the reference's own loader through bin/yrt_flatten (".yrts" file,
written by yocto_raytracing_b200/host/yrt_flatten.cpp from the reference's `scene`, src/scene.h:136),
by the synthetic generators in synth.py, or from a committed .npz fixture.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass, field
from typing import Dict

import numpy as np

from . import _lib

TRIANGLES, LINES, POINTS = 0, 1, 2

_I32 = ("shape_kind", "shape_elem_off", "shape_elem_cnt", "shape_vert_off", "shape_vert_cnt", "shape_has_uv",
        "shape_has_radius", "elem_idx", "inst_shape", "inst_mat", "mat_kd_tex", "mat_ks_tex", "tex_w", "tex_h")
_F32 = ("pos", "norm", "uv", "radius", "inst_frame", "mat_ke", "mat_kd", "mat_ks", "mat_kr", "mat_rs", "camera")
_DTYPES = {0: np.int32, 1: np.float32, 2: np.uint8, 3: np.int64}


@dataclass
class FlatScene:
    arrays: Dict[str, np.ndarray] = field(default_factory=dict)

    # ---- construction -------------------------------------------------------------------
    @staticmethod
    def _normalise(arrays: Dict[str, np.ndarray]) -> "FlatScene":
        a = {}
        for k in _I32:
            a[k] = np.ascontiguousarray(arrays.get(k, np.zeros(0)), dtype=np.int32).reshape(-1)
        for k in _F32:
            a[k] = np.ascontiguousarray(arrays.get(k, np.zeros(0)), dtype=np.float32).reshape(-1)
        a["tex_off"] = np.ascontiguousarray(arrays.get("tex_off", np.zeros(0)), dtype=np.int64).reshape(-1)
        a["tex_rgba8"] = np.ascontiguousarray(arrays.get("tex_rgba8", np.zeros(0)), dtype=np.uint8).reshape(-1)
        if a["camera"].size != 16:
            raise ValueError("scene needs a camera: 12 frame floats + fovy, aspect, aperture, focus")
        return FlatScene(a)

    @staticmethod
    def load(path: str) -> "FlatScene":
        """Read a .yrts container (yrt_flat_save) or a .npz fixture."""
        if path.endswith(".npz"):
            with np.load(path) as z:
                return FlatScene._normalise({k: z[k] for k in z.files})
        with open(path, "rb") as f:
            buf = f.read()
        if buf[:8] != b"YRTSCN01":
            raise ValueError(f"{path}: not a YRTSCN01 container")
        n_arrays = struct.unpack_from("<i", buf, 8)[0]
        off = 16
        arrays = {}
        for _ in range(n_arrays):
            name = buf[off:off + 24].split(b"\0", 1)[0].decode()
            dt, count = struct.unpack_from("<iq", buf, off + 24)
            off += 36
            dtype = np.dtype(_DTYPES[dt])
            nbytes = count * dtype.itemsize
            arrays[name] = np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()
            off += nbytes + ((8 - nbytes % 8) % 8)
        return FlatScene._normalise(arrays)

    def save_npz(self, path: str) -> None:
        np.savez_compressed(path, **self.arrays)

    # ---- accessors ----------------------------------------------------------------------
    def __getattr__(self, name):
        try:
            return self.__dict__["arrays"][name]
        except KeyError:
            raise AttributeError(name)

    @property
    def n_shapes(self): return int(self.arrays["shape_kind"].size)
    @property
    def n_instances(self): return int(self.arrays["inst_shape"].size)
    @property
    def n_materials(self): return int(self.arrays["mat_rs"].size)
    @property
    def n_textures(self): return int(self.arrays["tex_w"].size)
    @property
    def n_verts(self): return int(self.arrays["pos"].size // 3)
    @property
    def n_elements(self): return int(self.arrays["shape_elem_cnt"].sum())

    def light_instances(self) -> np.ndarray:
        """Instances whose material has ke.x>0 && ke.y>0 && ke.z>0 (src/raytrace.cpp:126), in instance order."""
        ke = self.arrays["mat_ke"].reshape(-1, 3)
        lit = (ke > 0).all(axis=1)
        return np.nonzero(lit[self.arrays["inst_mat"]])[0]

    def camera_struct(self) -> _lib.Camera:
        c = self.arrays["camera"]
        cam = _lib.Camera()
        for i in range(12):
            cam.frame[i] = float(c[i])
        cam.fovy, cam.aspect, cam.aperture, cam.focus = float(c[12]), float(c[13]), float(c[14]), float(c[15])
        return cam

    def image_width(self, resolution: int) -> int:
        """(int)std::round(cam->aspect * resolution), src/raytrace.cpp:216 (float32 product, round half away)."""
        v = np.float32(self.arrays["camera"][13]) * np.float32(resolution)
        return int(np.floor(np.float32(v) + np.float32(0.5))) if v >= 0 else int(np.ceil(v - np.float32(0.5)))

    def nonrigid_instances(self) -> int:
        """yrt_desc_nonrigid_instances: instances whose frame is not orthonormal (the reference's result for those depends
        on its own BVH visit order, include/yrt_b200.h); host-only, needs no GPU."""
        d = self.desc()
        return int(_lib.load().yrt_desc_nonrigid_instances(C.byref(d)))

    def desc(self) -> _lib.SceneDesc:
        """ctypes view (borrows the numpy buffers: keep this FlatScene alive while it is used)."""
        a = self.arrays
        d = _lib.SceneDesc()
        d.n_shapes, d.n_instances, d.n_materials, d.n_textures = self.n_shapes, self.n_instances, self.n_materials, self.n_textures
        d.n_verts, d.n_elem_idx = self.n_verts, int(a["elem_idx"].size)
        for name, _ in _lib.SceneDesc._fields_:
            if name in a:
                setattr(d, name, a[name].ctypes.data if a[name].size else None)
        if a["uv"].size == 0:
            d.uv = None
        if a["radius"].size == 0:
            d.radius = None
        d.tex_bytes = int(a["tex_rgba8"].size)
        return d
