"""Synthetic scenes "of the named shape" (BASELINE.json configs).

Each generator returns a SynthScene that can be turned into
  * a FlatScene directly (what the C ABI consumes), and
  * an OBJ + MTL (+ PNG) in the dialect the reference's loader reads (src/ext/yocto_obj.cpp:401-497:
    `c` cameras, `i` instances, `vr` radii, `p`/`l`/`f` with pos/tex/norm/color/radius slots), so the
    unmodified reference binary can render the very same scene on the host.
There is no network and the reference's `in/` directory does not travel to the GPU box, so benchmarks
and most parity tests run on these.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .scene import FlatScene, LINES, POINTS, TRIANGLES

F = np.float32


@dataclass
class Shape:
    name: str
    kind: int
    pos: np.ndarray                 # (n,3) float32
    norm: np.ndarray                # (n,3) float32 (tangent for lines)
    elems: np.ndarray               # (m,3|2|1) int32, local vertex ids
    material: str
    uv: Optional[np.ndarray] = None     # (n,2)
    radius: Optional[np.ndarray] = None # (n,)


@dataclass
class Material:
    name: str
    kd: tuple = (0.5, 0.5, 0.5)
    ks: tuple = (0.0, 0.0, 0.0)
    ke: tuple = (0.0, 0.0, 0.0)
    kr: tuple = (0.0, 0.0, 0.0)
    ns: float = 1.0
    map_kd: Optional[str] = None
    map_ks: Optional[str] = None


@dataclass
class SynthScene:
    shapes: List[Shape] = field(default_factory=list)
    materials: List[Material] = field(default_factory=list)
    textures: Dict[str, np.ndarray] = field(default_factory=dict)   # name -> (h,w,4) uint8
    instances: List[tuple] = field(default_factory=list)            # (name, shape index, frame(12,))
    camera: np.ndarray = None                                       # 16 floats: frame, fovy, aspect, aperture, focus
    name: str = "synth"

    # ---- flat --------------------------------------------------------------------------------
    def flat(self) -> FlatScene:
        mat_index = {m.name: i for i, m in enumerate(self.materials)}
        tex_names = list(self.textures.keys())
        tex_index = {n: i for i, n in enumerate(tex_names)}
        a = {k: [] for k in ("shape_kind", "shape_elem_off", "shape_elem_cnt", "shape_vert_off", "shape_vert_cnt", "shape_has_uv",
                             "shape_has_radius")}
        elem_idx, pos, norm, uv, radius = [], [], [], [], []
        eo = vo = 0
        for s in self.shapes:
            n = s.pos.shape[0]
            a["shape_kind"].append(s.kind)
            a["shape_elem_off"].append(eo)
            a["shape_elem_cnt"].append(s.elems.shape[0])
            a["shape_vert_off"].append(vo)
            a["shape_vert_cnt"].append(n)
            a["shape_has_uv"].append(int(s.uv is not None))
            a["shape_has_radius"].append(int(s.radius is not None))
            elem_idx.append(s.elems.astype(np.int32).reshape(-1))
            pos.append(s.pos.astype(F).reshape(-1))
            norm.append(s.norm.astype(F).reshape(-1))
            uv.append((s.uv if s.uv is not None else np.zeros((n, 2))).astype(F).reshape(-1))
            radius.append((s.radius if s.radius is not None else np.zeros(n)).astype(F).reshape(-1))
            eo += s.elems.size
            vo += n
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
        a["elem_idx"], a["pos"], a["norm"], a["uv"], a["radius"] = cat(elem_idx, np.int32), cat(pos, F), cat(norm, F), cat(uv, F), cat(radius, F)
        a["inst_frame"] = cat([np.asarray(f, F).reshape(-1) for _, _, f in self.instances], F)
        a["inst_shape"] = np.array([si for _, si, _ in self.instances], np.int32)
        a["inst_mat"] = np.array([mat_index[self.shapes[si].material] for _, si, _ in self.instances], np.int32)
        for key in ("ke", "kd", "ks", "kr"):
            a["mat_" + key] = np.array([getattr(m, key) for m in self.materials], F).reshape(-1)
        # roughness as the loader derives it from Ns (src/ext/yocto_scn.cpp:253): rs = (2/(Ns+2))^(1/4)
        a["mat_rs"] = np.array([np.power(F(2) / (F(m.ns) + F(2)), F(0.25)) for m in self.materials], F)
        a["mat_kd_tex"] = np.array([tex_index.get(m.map_kd, -1) if m.map_kd else -1 for m in self.materials], np.int32)
        a["mat_ks_tex"] = np.array([tex_index.get(m.map_ks, -1) if m.map_ks else -1 for m in self.materials], np.int32)
        tw, th, toff, tdata, off = [], [], [], [], 0
        for n_ in tex_names:
            t = np.ascontiguousarray(self.textures[n_], np.uint8)
            th.append(t.shape[0]); tw.append(t.shape[1]); toff.append(off)
            tdata.append(t.reshape(-1)); off += t.size
        a["tex_w"], a["tex_h"] = np.array(tw, np.int32), np.array(th, np.int32)
        a["tex_off"] = np.array(toff, np.int64)
        a["tex_rgba8"] = cat(tdata, np.uint8)
        a["camera"] = np.asarray(self.camera, F)
        return FlatScene._normalise(a)

    # ---- OBJ dialect of the reference's loader ---------------------------------------------------
    def write_obj(self, directory: str, name: Optional[str] = None) -> str:
        """Write <name>.obj/.mtl (+ textures as PNG); returns the .obj path."""
        name = name or self.name
        os.makedirs(directory, exist_ok=True)
        g = lambda v: "%.9g" % float(v)
        with open(os.path.join(directory, name + ".mtl"), "w") as f:
            for m in self.materials:
                f.write(f"newmtl {m.name}\n  illum 2\n")
                f.write("  Kd %s %s %s\n" % tuple(map(g, m.kd)))
                if any(m.ks): f.write("  Ks %s %s %s\n" % tuple(map(g, m.ks)))
                if any(m.ke): f.write("  Ke %s %s %s\n" % tuple(map(g, m.ke)))
                if any(m.kr): f.write("  Kr %s %s %s\n" % tuple(map(g, m.kr)))
                f.write("  Ns %s\n" % g(m.ns))
                if m.map_kd: f.write(f"  map_Kd {m.map_kd}\n")
                if m.map_ks: f.write(f"  map_Ks {m.map_ks}\n")
                f.write("\n")
        if self.textures:
            from PIL import Image
            for tn, t in self.textures.items():
                Image.fromarray(np.ascontiguousarray(t, np.uint8), "RGBA").save(os.path.join(directory, tn))
        path = os.path.join(directory, name + ".obj")
        with open(path, "w") as f:
            f.write(f"mtllib {name}.mtl\n")
            c = self.camera
            f.write("c cam 0 %s %s %s %s %s\n" % (g(c[12]), g(c[13]), g(c[14]), g(c[15]), " ".join(g(x) for x in c[:12])))
            for iname, si, fr in self.instances:
                f.write("i %s %s %s\n" % (iname, self.shapes[si].name, " ".join(g(x) for x in np.asarray(fr).reshape(-1))))
            # all vertex data first (global 1-based indices), then the objects
            vbase, rbase, tbase = [], [], []
            nv = nr = nt = 0
            for s in self.shapes:
                vbase.append(nv); rbase.append(nr); tbase.append(nt)
                for p in s.pos: f.write("v %s %s %s\n" % (g(p[0]), g(p[1]), g(p[2])))
                nv += s.pos.shape[0]
            for s in self.shapes:
                for p in s.norm: f.write("vn %s %s %s\n" % (g(p[0]), g(p[1]), g(p[2])))
            for k, s in enumerate(self.shapes):
                tbase[k] = nt
                if s.uv is not None:
                    # the loader flips v (obj_flip_texcoord, src/ext/yocto_scn.h:424): pre-flip so it lands on s.uv
                    for p in s.uv: f.write("vt %s %s\n" % (g(p[0]), g(F(1) - F(p[1]))))
                    nt += s.uv.shape[0]
            for k, s in enumerate(self.shapes):
                rbase[k] = nr
                if s.radius is not None:
                    for r in s.radius: f.write("vr %s\n" % g(r))
                    nr += s.radius.shape[0]
            for k, s in enumerate(self.shapes):
                f.write(f"o {s.name}\nusemtl {s.material}\n")
                tag = {TRIANGLES: "f", LINES: "l", POINTS: "p"}[s.kind]

                def ref(v):
                    p = vbase[k] + v + 1
                    t = str(tbase[k] + v + 1) if s.uv is not None else ""
                    r = str(rbase[k] + v + 1) if s.radius is not None else ""
                    out = f"{p}/{t}/{p}"
                    if r: out += f"//{r}"
                    return out
                for e in s.elems:
                    f.write(tag + " " + " ".join(ref(int(v)) for v in e) + "\n")
        return path


# ---- helpers -------------------------------------------------------------------------------------
def lookat_frame(eye, target, up=(0, 1, 0)) -> np.ndarray:
    eye, target, up = (np.asarray(v, np.float64) for v in (eye, target, up))
    z = eye - target; z /= np.linalg.norm(z)
    x = np.cross(up, z); x /= np.linalg.norm(x)
    y = np.cross(z, x); y /= np.linalg.norm(y)
    return np.concatenate([x, y, z, eye]).astype(F)


def make_camera(eye, target, fovy, aspect=16.0 / 9.0, aperture=0.0) -> np.ndarray:
    fr = lookat_frame(eye, target)
    focus = float(np.linalg.norm(np.asarray(eye, np.float64) - np.asarray(target, np.float64)))
    return np.concatenate([fr, np.array([fovy, aspect, aperture, focus], F)]).astype(F)


def translation_frame(o) -> np.ndarray:
    return np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, o[0], o[1], o[2]], F)


def rotation_frame(axis, angle, o) -> np.ndarray:
    axis = np.asarray(axis, np.float64); axis /= np.linalg.norm(axis)
    c, s = math.cos(angle), math.sin(angle)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + s * K + (1 - c) * (K @ K)
    return np.concatenate([R[:, 0], R[:, 1], R[:, 2], np.asarray(o, np.float64)]).astype(F)


def _first_use_order(shape: Shape) -> Shape:
    """Renumber vertices in order of first use by the elements — the order in which the reference's loader
    meets (and numbers) them (src/ext/yocto_scn.cpp:311-318), so loader output == our arrays."""
    flat = shape.elems.reshape(-1)
    _, first = np.unique(flat, return_index=True)
    order = flat[np.sort(first)]                      # old ids in first-use order (unused vertices dropped)
    remap = -np.ones(shape.pos.shape[0], np.int64)
    remap[order] = np.arange(order.size)
    return Shape(shape.name, shape.kind, shape.pos[order], shape.norm[order], remap[shape.elems].astype(np.int32), shape.material,
                 None if shape.uv is None else shape.uv[order], None if shape.radius is None else shape.radius[order])


def grid_floor(name, material, half, n, uv_scale) -> Shape:
    """n x n quads over [-half, half]^2 at y = 0, two triangles per quad (fan order of the loader)."""
    t = np.linspace(-half, half, n + 1, dtype=np.float64)
    x, z = np.meshgrid(t, t, indexing="xy")
    pos = np.stack([x, np.zeros_like(x), z], -1).reshape(-1, 3).astype(F)
    norm = np.tile(np.array([0, 1, 0], F), (pos.shape[0], 1))
    uv = (np.stack([x, z], -1).reshape(-1, 2) + half) * (uv_scale / (2 * half))
    idx = lambda i, j: j * (n + 1) + i
    tris = []
    for j in range(n):
        for i in range(n):
            a, b, c, d = idx(i, j), idx(i, j + 1), idx(i + 1, j + 1), idx(i + 1, j)
            tris += [(a, b, c), (a, c, d)]
    return _first_use_order(Shape(name, TRIANGLES, pos, norm, np.array(tris, np.int32), material, uv.astype(F)))


def cube_sphere(name, material, n, roundness) -> Shape:
    """Subdivided cube in [-1,1]^3 (6 n^2 quads) blended towards the unit sphere; roundness 0 = cube
    (coplanar faces across instances, like the reference's shapes), 1 = sphere."""
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    pos, norm, uv, tris = [], [], [], []
    t = np.linspace(-1, 1, n + 1)
    for fi, (nrm, du, dv) in enumerate(faces):
        nrm, du, dv = (np.asarray(v, np.float64) for v in (nrm, du, dv))
        base = len(pos)
        for j in range(n + 1):
            for i in range(n + 1):
                c = nrm + du * t[i] + dv * t[j]
                s = c / np.linalg.norm(c)
                p = (1 - roundness) * c + roundness * s
                nn = (1 - roundness) * nrm + roundness * s
                pos.append(p); norm.append(nn / np.linalg.norm(nn)); uv.append(((t[i] + 1) / 2, (t[j] + 1) / 2))
        idx = lambda i, j: base + j * (n + 1) + i
        for j in range(n):
            for i in range(n):
                a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
                tris += [(a, b, c), (a, c, d)]
    return _first_use_order(Shape(name, TRIANGLES, np.array(pos, F), np.array(norm, F), np.array(tris, np.int32), material, np.array(uv, F)))


def uv_sphere(name, material, nu, nv, scale=(1, 1, 1)) -> Shape:
    pos, norm, uv, tris = [], [], [], []
    sc = np.asarray(scale, np.float64)
    for j in range(nv + 1):
        th = math.pi * j / nv
        for i in range(nu + 1):
            ph = 2 * math.pi * i / nu
            d = np.array([math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)])
            pos.append(d * sc)
            nn = d / sc
            norm.append(nn / np.linalg.norm(nn))
            uv.append((i / nu, j / nv))
    idx = lambda i, j: j * (nu + 1) + i
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            tris += [(a, b, c), (a, c, d)]
    return _first_use_order(Shape(name, TRIANGLES, np.array(pos, F), np.array(norm, F), np.array(tris, np.int32), material, np.array(uv, F)))


def point_light(name, material, with_uv=False) -> Shape:
    """A light as the reference scenes model it: one point of radius 0.001 (it IS in the BVH and hittable).
    The reference's lights carry no texcoords, which is undefined behaviour there if one is ever hit
    (src/raytrace.cpp:113 -> scene.h:195); with_uv avoids that for scenes rendered by the reference binary."""
    return Shape(name, POINTS, np.zeros((1, 3), F), np.array([[0, 0, 1]], F), np.array([[0]], np.int32), material,
                 np.zeros((1, 2), F) if with_uv else None, np.array([0.001], F))


def checker_texture(size=64, cells=8, a=(230, 230, 230, 255), b=(60, 60, 60, 255), line=(200, 30, 30, 255)) -> np.ndarray:
    t = np.zeros((size, size, 4), np.uint8)
    c = size // cells
    yy, xx = np.mgrid[0:size, 0:size]
    t[:] = np.where((((xx // c) + (yy // c)) % 2 == 0)[..., None], np.array(a, np.uint8), np.array(b, np.uint8))
    t[(xx % c == 0) | (yy % c == 0)] = np.array(line, np.uint8)
    return t


# ---- config 5: N instances of 10 meshes on a jittered grid (instance10000_pointlight's shape) ---------
def instance_grid_scene(n_side: int = 100, seed: int = 10000, mesh_n: int = 16) -> SynthScene:
    """n_side^2 instances (spacing 2, jitter +-0.5, y = 1) of 10 meshes (3072-4096 triangles each at the
    default mesh_n = 16) over a 64x64-quad floor, 3 point lights at height 50, camera like cam01 — the
    structure of in/instance10000_pointlight (10 004 instances, 14 shapes, 41 984 triangles + 3 points)."""
    rng = np.random.RandomState(seed)
    sc = SynthScene(name=f"instance{n_side * n_side}_synth")
    half = 1.2 * n_side
    sc.materials.append(Material("floor_txt", kd=(0.2, 0.2, 0.2), ns=1))
    sc.shapes.append(grid_floor("floor", "floor_txt", half, 64, 120.0))
    for k in range(10):
        kd = tuple(float(x) for x in rng.uniform(0.5, 1.0, 3))
        glossy = k % 5 in (2, 3, 4)
        sc.materials.append(Material(f"mat{k:03d}", kd=kd, ks=(0.04, 0.04, 0.04) if glossy else (0, 0, 0),
                                     ns=float(rng.uniform(2000, 5000)) if glossy else 1.0))
        if k % 3 == 1:
            sc.shapes.append(uv_sphere(f"shp{k:03d}", f"mat{k:03d}", 4 * mesh_n, 2 * mesh_n, scale=(1.0, 1.0 - 0.05 * k, 1.0)))
        else:
            sc.shapes.append(cube_sphere(f"shp{k:03d}", f"mat{k:03d}", mesh_n, roundness=[0.0, 0.0, 0.35, 0.0, 0.7, 1.0, 0.0, 0.5, 0.15, 0.0][k]))
    for k, (ke, o) in enumerate([(2000, (0, 50, 50)), (1000, (50, 50, 0)), (1000, (0, 50, -50))]):
        sc.materials.append(Material(f"pointlight{k + 1:02d}", kd=(0, 0, 0), ke=(ke, ke, ke), ns=1))
        sc.shapes.append(point_light(f"pointlight{k + 1:02d}", f"pointlight{k + 1:02d}"))
    sc.instances.append(("floor", 0, translation_frame((0, 0, 0))))
    for j in range(n_side):
        for i in range(n_side):
            x = -n_side + 1 + 2 * i + rng.uniform(-0.5, 0.5)
            z = -n_side + 1 + 2 * j + rng.uniform(-0.5, 0.5)
            sc.instances.append((f"ist{j * n_side + i:06d}", 1 + int(rng.randint(0, 10)), translation_frame((F(x), 1, F(z)))))
    s = n_side / 100.0
    for k, o in enumerate([(0, 50, 50), (50, 50, 0), (0, 50, -50)]):
        sc.instances.append((f"pointlight{k + 1:02d}", 11 + k, translation_frame((o[0] * s, 50 * s if s > 1 else 50, o[2] * s))))
    sc.camera = make_camera((0, 75 * s, 75 * s), (0, 1, 0), 0.471239)
    return sc


# ---- config 4: hair (lines) --------------------------------------------------------------------------
def hair_scene(n_hairs: int = 4096, segments: int = 8, seed: int = 1234) -> SynthScene:
    """Two unit spheres (x = +-1.25, y = 1) each with n_hairs polylines of `segments` segments (length 0.3,
    radius 0.001 root -> 0.0005 tip) over a textured 64x64 floor, two point lights; stands in for the
    missing in/lines_pointlight/lines_pointlight.obj (SURVEY.md finding 2).  Hair vertices carry texcoords:
    the reference evaluates eval_texcoord unconditionally (src/raytrace.cpp:113)."""
    rng = np.random.RandomState(seed)
    sc = SynthScene(name="lines_synth")
    sc.textures["grid.png"] = checker_texture(64, 8)
    sc.materials += [Material("floor_txt", kd=(1, 1, 1), ns=1, map_kd="grid.png"), Material("lines", kd=(0.2, 0.2, 0.2), ns=1),
                     Material("interior", kd=(0.2, 0.2, 0.2), ns=1), Material("pointlight", kd=(0, 0, 0), ke=(100, 100, 100), ns=1)]
    sc.shapes.append(grid_floor("floor", "floor_txt", 20.0, 64, 40.0))
    sc.shapes.append(uv_sphere("interior", "interior", 64, 32))
    # hairs in the sphere's object space
    k = np.arange(n_hairs) + 0.5
    phi = np.arccos(1 - 2 * k / n_hairs)
    theta = math.pi * (1 + 5 ** 0.5) * k
    roots = np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], -1)
    dirs = roots + 0.35 * rng.normal(size=roots.shape)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    pos, tan, uv, rad, lines = [], [], [], [], []
    for h in range(n_hairs):
        d = dirs[h].copy()
        p = roots[h].copy()
        base = len(pos)
        for s in range(segments + 1):
            pos.append(p.copy()); tan.append(d / np.linalg.norm(d)); uv.append((h / n_hairs, s / segments))
            rad.append(0.001 + (0.0005 - 0.001) * s / segments)
            d = d + 0.25 * np.array([0, -1, 0]) * (0.3 / segments) * 4 + 0.05 * rng.normal(size=3)
            d /= np.linalg.norm(d)
            p = p + d * (0.3 / segments)
        lines += [(base + s, base + s + 1) for s in range(segments)]
    sc.shapes.append(Shape("hair", LINES, np.array(pos, F), np.array(tan, F), np.array(lines, np.int32), "lines", np.array(uv, F), np.array(rad, F)))
    sc.shapes.append(point_light("pointlight01", "pointlight"))
    sc.shapes.append(point_light("pointlight02", "pointlight"))
    sc.instances += [("floor", 0, translation_frame((0, 0, 0))),
                     ("sphere_l", 1, translation_frame((-1.25, 1, 0))), ("sphere_r", 1, translation_frame((1.25, 1, 0))),
                     ("hair_l", 2, translation_frame((-1.25, 1, 0))), ("hair_r", 2, rotation_frame((0, 1, 0), 0.7, (1.25, 1, 0))),
                     ("pointlight01", 3, translation_frame((-2.0, 7.0, 4.0))), ("pointlight02", 4, translation_frame((3.0, 6.0, 5.0)))]
    sc.camera = make_camera((0, 4, 10), (0, 1, 0), 0.261799)
    return sc


def _hair_shape(name: str, n_hairs: int, segments: int, rng: np.random.RandomState) -> Shape:
    """n_hairs polylines of `segments` segments (length 0.3) rooted on a Fibonacci lattice of the unit sphere; direction =
    normal + seeded noise, bending down along the hair; radius 0.001 root -> 0.0005 tip; every vertex carries a texcoord."""
    k = np.arange(n_hairs) + 0.5
    phi = np.arccos(1 - 2 * k / n_hairs)
    theta = math.pi * (1 + 5 ** 0.5) * k
    roots = np.stack([np.cos(theta) * np.sin(phi), np.cos(phi), np.sin(theta) * np.sin(phi)], -1)
    d = roots + 0.35 * rng.normal(size=roots.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = roots.copy()
    pos = np.empty((n_hairs, segments + 1, 3))
    tan = np.empty((n_hairs, segments + 1, 3))
    step = 0.3 / segments
    for s_ in range(segments + 1):
        pos[:, s_] = p
        tan[:, s_] = d
        d = d + np.array([0, -step, 0]) + 0.05 * rng.normal(size=d.shape)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = p + d * step
    hs, ss = np.meshgrid(np.arange(n_hairs), np.arange(segments + 1), indexing="ij")
    uv = np.stack([hs / n_hairs, ss / segments], -1)
    rad = 0.001 + (0.0005 - 0.001) * ss / segments
    base = (hs * (segments + 1) + ss)[:, :-1].reshape(-1)
    lines = np.stack([base, base + 1], 1)
    return Shape(name, LINES, pos.reshape(-1, 3).astype(F), tan.reshape(-1, 3).astype(F), lines.astype(np.int32), "lines",
                 uv.reshape(-1, 2).astype(F), rad.reshape(-1).astype(F))


def lines_config4(n_hairs: int = 65536, segments: int = 8, seed: int = 1234, texture: Optional[np.ndarray] = None) -> SynthScene:
    """SURVEY 8d config 4 (BASELINE.json configs[3]) at its specified size — the reference's in/lines_pointlight has its
    materials and texture but no OBJ, so the scene is synthetic: camera and lights of simple_pointlight (cam: fovy 0.261799,
    focus 10.4403 at (0, 4, 10); two point lights Ke 100 at (+-1.4, 8, 6)), a 64x64-quad floor over [-20, 20]^2 with a grid
    texture, two unit spheres at x = +-1.25, y = 1 (material `interior`), each with its OWN n_hairs hairs of 8 segments
    (MT19937 seeded 1234): 2 x 65 536 x 8 = 1 048 576 line elements in two BLAS of 524 288 — the one config whose
    bottom-level trees, not the instance tree, are large."""
    rng = np.random.RandomState(seed)
    sc = SynthScene(name="lines_config4")
    sc.textures["grid.png"] = texture if texture is not None else checker_texture(512, 16)
    sc.materials += [Material("floor_txt", kd=(1, 1, 1), ns=1, map_kd="grid.png"), Material("lines", kd=(0.2, 0.2, 0.2), ns=1),
                     Material("interior", kd=(0.2, 0.2, 0.2), ns=1), Material("pointlight", kd=(0, 0, 0), ke=(100, 100, 100), ns=1)]
    sc.shapes.append(grid_floor("floor", "floor_txt", 20.0, 64, 40.0))
    sc.shapes.append(uv_sphere("interior", "interior", 64, 32))
    sc.shapes.append(_hair_shape("hair_l", n_hairs, segments, rng))
    sc.shapes.append(_hair_shape("hair_r", n_hairs, segments, rng))
    sc.shapes.append(point_light("pointlight01", "pointlight", with_uv=True))
    sc.shapes.append(point_light("pointlight02", "pointlight", with_uv=True))
    sc.instances += [("floor", 0, translation_frame((0, 0, 0))),
                     ("sphere_l", 1, translation_frame((-1.25, 1, 0))), ("sphere_r", 1, translation_frame((1.25, 1, 0))),
                     ("hair_l", 2, translation_frame((-1.25, 1, 0))), ("hair_r", 3, translation_frame((1.25, 1, 0))),
                     ("pointlight01", 4, translation_frame((1.4, 8.0, 6.0))), ("pointlight02", 5, translation_frame((-1.4, 8.0, 6.0)))]
    sc.camera = np.array([1, 0, 0, 0, 0.957826, -0.287348, 0, 0.287348, 0.957826, 0, 4, 10, 0.261799, 1.77778, 0, 10.4403], F)
    return sc


# ---- small mixed scene for edge cases ------------------------------------------------------------------
def mixed_scene(seed: int = 7, n_objects: int = 24, reflective_floor: bool = True, textured: bool = True) -> SynthScene:
    """Small scene touching every code path: textured + mirror floor (recursion), rotated instances of
    cubes/spheres (non-identity frames), a hair tuft (lines), free-standing points, several lights, glossy
    and textured-specular materials."""
    rng = np.random.RandomState(seed)
    sc = SynthScene(name=f"mixed{seed}")
    if textured:
        sc.textures["grid.png"] = checker_texture(64, 8)
        sc.textures["spec.png"] = checker_texture(32, 4, a=(255, 255, 255, 255), b=(20, 20, 20, 255), line=(128, 128, 128, 255))
    sc.materials += [
        Material("floor", kd=(1, 1, 1), ns=1, kr=(0.4, 0.4, 0.4) if reflective_floor else (0, 0, 0), map_kd="grid.png" if textured else None),
        Material("glossy", kd=(0.9, 0.5, 0.3), ks=(0.8, 0.8, 0.8), ns=198, map_ks="spec.png" if textured else None),
        Material("matte", kd=(0.3, 0.6, 0.9), ns=1),
        Material("mirror", kd=(0.1, 0.1, 0.1), ks=(0.2, 0.2, 0.2), ns=19998, kr=(0.7, 0.7, 0.7)),
        Material("hairmat", kd=(0.4, 0.3, 0.2), ks=(0.3, 0.3, 0.3), ns=50),
        Material("dots", kd=(0.9, 0.9, 0.1), ns=1),
        Material("light_a", kd=(0, 0, 0), ke=(60, 60, 60), ns=1), Material("light_b", kd=(0, 0, 0), ke=(30, 40, 50), ns=1),
        Material("half_emitter", kd=(0.5, 0.5, 0.5), ke=(5, 0, 5), ns=1),   # ke.y == 0: NOT a light (raytrace.cpp:126)
    ]
    sc.shapes.append(grid_floor("floor", "floor", 8.0, 16, 8.0))
    sc.shapes.append(cube_sphere("cube", "glossy", 4, 0.0))
    sc.shapes.append(uv_sphere("ball", "matte", 16, 8))
    sc.shapes.append(cube_sphere("blob", "mirror", 6, 0.6))
    # hair tuft
    pos, tan, uv, rad, lines = [], [], [], [], []
    for h in range(96):
        p = np.array([rng.uniform(-0.4, 0.4), 0.0, rng.uniform(-0.4, 0.4)])
        d = np.array([rng.normal() * 0.3, 1.0, rng.normal() * 0.3]); d /= np.linalg.norm(d)
        base = len(pos)
        for s in range(5):
            pos.append(p.copy()); tan.append(d.copy()); uv.append((h / 96, s / 4)); rad.append(0.02 - 0.003 * s)
            p = p + d * 0.25
            d = d + rng.normal(size=3) * 0.15; d /= np.linalg.norm(d)
        lines += [(base + s, base + s + 1) for s in range(4)]
    sc.shapes.append(Shape("tuft", LINES, np.array(pos, F), np.array(tan, F), np.array(lines, np.int32), "hairmat", np.array(uv, F), np.array(rad, F)))
    npts = 40
    ppos = rng.uniform(-0.5, 0.5, (npts, 3)); ppos[:, 1] += 0.5
    sc.shapes.append(Shape("dots", POINTS, ppos.astype(F), np.tile(np.array([0, 1, 0], F), (npts, 1)), np.arange(npts, dtype=np.int32).reshape(-1, 1),
                           "dots", rng.uniform(0, 1, (npts, 2)).astype(F), rng.uniform(0.02, 0.06, npts).astype(F)))
    sc.shapes.append(point_light("light_a", "light_a", with_uv=True))
    sc.shapes.append(point_light("light_b", "light_b", with_uv=True))
    sc.shapes.append(Shape("half", POINTS, np.zeros((1, 3), F), np.array([[0, 1, 0]], F), np.array([[0]], np.int32), "half_emitter",
                           np.zeros((1, 2), F), np.array([0.05], F)))
    sc.instances.append(("floor", 0, translation_frame((0, 0, 0))))
    for k in range(n_objects):
        si = 1 + int(rng.randint(0, 3))
        o = (rng.uniform(-5, 5), 1.0 + rng.uniform(0, 0.6), rng.uniform(-5, 3))
        sc.instances.append((f"obj{k:03d}", si, rotation_frame(rng.normal(size=3), rng.uniform(0, 6.28), o) if k % 2 else translation_frame(o)))
    sc.instances += [("tuft_a", 4, translation_frame((1.5, 0, 3.5))), ("tuft_b", 4, rotation_frame((0, 1, 0), 1.0, (-1.5, 0, 3.0))),
                     ("dots_a", 5, translation_frame((0, 0.2, 4.0))),
                     ("light_a", 6, translation_frame((-3, 6, 5))), ("light_b", 7, rotation_frame((1, 0, 0), 0.0, (4, 5, 3))),
                     ("half", 8, translation_frame((0, 3, 0)))]
    sc.camera = make_camera((0, 4, 10), (0, 1, 0), 0.6)
    return sc


def nonrigid_scene(seed: int = 31, frame_seed: int = 5, every: int = 3, mirror_floor: bool = False) -> SynthScene:
    """mixed_scene with every `every`-th object instance scaled per axis and sheared: frames transform_ray_inverse
    (src/vmath.h:275-278) does not invert.  The reference's result for such a scene depends on its own instance tree and visit
    order (src/scene.cpp:446-479) — the case the RefTlas path of the library exists for.  Object mirrors are switched off (a sheared
    mirror traps rays and the reference recurses until its stack overflows); mirror_floor keeps the rigid floor reflective, so
    that mirror rays meet the non-rigid instances too."""
    sc = mixed_scene(seed, reflective_floor=mirror_floor)
    sc.name = f"nonrigid{seed}"
    rng = np.random.default_rng(frame_seed)
    for m in sc.materials:
        if m.name != "floor":
            m.kr = (0.0, 0.0, 0.0)
    for k, (iname, si, fr) in enumerate(list(sc.instances)):
        if iname.startswith("obj") and k % every == 0:
            f = np.array(fr, np.float32).reshape(4, 3).copy()
            f[0] *= rng.uniform(0.5, 1.8); f[1] *= rng.uniform(0.5, 1.8); f[2] *= rng.uniform(0.6, 1.5)
            f[0] += 0.3 * f[1]
            sc.instances[k] = (iname, si, f.reshape(-1))
    return sc

