"""Write a small glTF 2.0 scene (scene.gltf + scene.bin) that the reference's loader accepts (src/ext/yocto_scn.cpp:723-1083):
meshes with POSITION / NORMAL / TEXCOORD_0 and uint32 indices, KHR_materials_pbrSpecularGlossiness materials (the loader
maps diffuseFactor -> kd, specularFactor -> ks, glossinessFactor -> rs, emissiveFactor -> ke), point lights as POINTS
primitives with an emissive material, nodes with translation + rotation quaternions (rigid; optionally node scales), one perspective camera node.
Used by the tests and tools/make_golden.py to cover the SURVEY 8f.4 input path."""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yocto_raytracing_b200 import synth  # noqa: E402


def quat_from_axis_angle(axis, angle):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    s = np.sin(angle / 2)
    return [float(a[0] * s), float(a[1] * s), float(a[2] * s), float(np.cos(angle / 2))]


def quat_from_frame(x, y, z):
    m = np.stack([x, y, z], 1).astype(np.float64)      # columns = axes
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
    else:
        i = int(np.argmax(np.diag(m)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
        q = [0, 0, 0, 0]
        q[i] = 0.25 * s
        q[j] = (m[j, i] + m[i, j]) / s
        q[k] = (m[k, i] + m[i, k]) / s
        q[3] = (m[k, j] - m[j, k]) / s
    return [float(v) for v in q]


class Builder:
    def __init__(self):
        self.bin = bytearray()
        self.views, self.accessors, self.meshes, self.materials, self.nodes, self.cameras = [], [], [], [], [], []

    def _accessor(self, arr, comp, typ, target=None, minmax=False):
        arr = np.ascontiguousarray(arr)
        while len(self.bin) % 4:
            self.bin.append(0)
        off = len(self.bin)
        self.bin += arr.tobytes()
        v = {"buffer": 0, "byteOffset": off, "byteLength": arr.nbytes}
        if target:
            v["target"] = target
        self.views.append(v)
        a = {"bufferView": len(self.views) - 1, "componentType": comp, "count": int(arr.shape[0]), "type": typ}
        if minmax:
            a["min"] = [float(x) for x in arr.min(0)]
            a["max"] = [float(x) for x in arr.max(0)]
        self.accessors.append(a)
        return len(self.accessors) - 1

    def material(self, name, kd=(0.5, 0.5, 0.5), ks=(0, 0, 0), rs=1.0, ke=(0, 0, 0)):
        self.materials.append({"name": name, "emissiveFactor": [float(x) for x in ke],
                               "extensions": {"KHR_materials_pbrSpecularGlossiness": {"diffuseFactor": [float(kd[0]), float(kd[1]), float(kd[2]), 1.0],
                                                                                      "specularFactor": [float(x) for x in ks], "glossinessFactor": float(rs)}}})
        return len(self.materials) - 1

    def mesh(self, name, shape, material):
        attrs = {"POSITION": self._accessor(shape.pos.astype(np.float32), 5126, "VEC3", 34962, True),
                 "NORMAL": self._accessor(shape.norm.astype(np.float32), 5126, "VEC3", 34962)}
        uv = shape.uv if shape.uv is not None else np.zeros((shape.pos.shape[0], 2), np.float32)
        attrs["TEXCOORD_0"] = self._accessor(uv.astype(np.float32), 5126, "VEC2", 34962)
        prim = {"attributes": attrs, "material": material}
        if shape.kind == synth.TRIANGLES:
            prim["mode"] = 4
            prim["indices"] = self._accessor(shape.elems.astype(np.uint32).reshape(-1), 5125, "SCALAR", 34963)
        elif shape.kind == synth.POINTS:
            prim["mode"] = 0
        else:
            prim["mode"] = 1
            prim["indices"] = self._accessor(shape.elems.astype(np.uint32).reshape(-1), 5125, "SCALAR", 34963)
        self.meshes.append({"name": name, "primitives": [prim]})
        return len(self.meshes) - 1

    def node(self, name, mesh=None, camera=None, translation=(0, 0, 0), rotation=(0, 0, 0, 1), children=None, scale=None):
        n = {"name": name, "translation": [float(x) for x in translation], "rotation": [float(x) for x in rotation]}
        if scale is not None:     # a node scale makes the instance frame non-rigid (the reference keeps it in the frame's axes)
            n["scale"] = [float(x) for x in scale]
        if mesh is not None:
            n["mesh"] = mesh
        if camera is not None:
            n["camera"] = camera
        if children:
            n["children"] = children
        self.nodes.append(n)
        return len(self.nodes) - 1

    def write(self, directory, name, roots):
        os.makedirs(directory, exist_ok=True)
        doc = {"asset": {"version": "2.0"}, "extensionsUsed": ["KHR_materials_pbrSpecularGlossiness"], "scene": 0, "scenes": [{"nodes": roots}],
               "nodes": self.nodes, "meshes": self.meshes, "materials": self.materials, "cameras": self.cameras, "accessors": self.accessors,
               "bufferViews": self.views, "buffers": [{"uri": name + ".bin", "byteLength": len(self.bin)}]}
        with open(os.path.join(directory, name + ".bin"), "wb") as f:
            f.write(bytes(self.bin))
        p = os.path.join(directory, name + ".gltf")
        with open(p, "w") as f:
            json.dump(doc, f)
        return p


def gltf_scene(directory, seed=7, n_objects=10, name="gltf7", scales=False):
    """Floor + rotated round cubes / spheres (a child-node hierarchy included) + two point lights + a look-at camera.
    scales: every third object node also carries a non-uniform scale (children inherit it) — the usual case in glTF files,
    and one the reference's ray transform does not invert (non-rigid instance frames)."""
    rng = np.random.default_rng(seed)
    b = Builder()
    m_floor = b.material("floor", kd=(0.6, 0.6, 0.55))
    m_a = b.material("glossy", kd=(0.7, 0.25, 0.2), ks=(0.3, 0.3, 0.3), rs=0.35)
    m_b = b.material("matte", kd=(0.2, 0.45, 0.75))
    m_l = b.material("light", kd=(0, 0, 0), ke=(60, 60, 60))
    m_l2 = b.material("light2", kd=(0, 0, 0), ke=(30, 45, 60))
    floor = b.mesh("floor", synth.grid_floor("floor", "floor", 8.0, 8, 4.0), m_floor)
    cube = b.mesh("cube", synth.cube_sphere("cube", "glossy", 6, 0.4), m_a)
    ball = b.mesh("ball", synth.uv_sphere("ball", "matte", 16, 8, scale=(1.0, 0.7, 1.0)), m_b)
    light = b.mesh("light", synth.point_light("light", "light", with_uv=True), m_l)
    light2 = b.mesh("light2", synth.point_light("light2", "light2", with_uv=True), m_l2)
    roots = [b.node("floor", mesh=floor)]
    for k in range(n_objects):
        o = (rng.uniform(-5, 5), rng.uniform(0.6, 1.6), rng.uniform(-5, 5))
        q = quat_from_axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        child = None
        if k % 4 == 0:   # a satellite in the parent's frame: exercises the node hierarchy (xform = parent * local)
            child = [b.node(f"sat{k}", mesh=ball if k % 8 else cube, translation=(1.6, 0.4, 0.0), rotation=quat_from_axis_angle((0, 0, 1), 0.5))]
        sc = (rng.uniform(0.6, 1.7), rng.uniform(0.6, 1.7), rng.uniform(0.6, 1.7)) if scales and k % 3 == 0 else None
        roots.append(b.node(f"obj{k}", mesh=cube if k % 2 else ball, translation=o, rotation=q, children=child, scale=sc))
    roots.append(b.node("light_a", mesh=light, translation=(-3, 6, 5)))
    roots.append(b.node("light_b", mesh=light2, translation=(4, 5, 3)))
    b.cameras.append({"type": "perspective", "perspective": {"yfov": 0.6, "aspectRatio": 16.0 / 9.0, "znear": 0.1, "zfar": 100.0}})
    fr = synth.lookat_frame((0.0, 6.0, 13.0), (0.0, 1.0, 0.0)).reshape(4, 3)
    roots.append(b.node("camera", camera=0, translation=fr[3], rotation=quat_from_frame(fr[0], fr[1], fr[2])))
    return b.write(directory, name, roots)


if __name__ == "__main__":
    print(gltf_scene(sys.argv[1] if len(sys.argv) > 1 else "."))
