"""B200-native render path of yocto_raytracing.

    from yocto_raytracing_b200 import FlatScene, Scene
    with Scene(FlatScene.load("scene.yrts")) as scn:          # upload + GPU LBVH build
        hdr = scn.raytrace(amb=0.1, resolution=720, samples=3)   # == reference raytrace()
"""
from .scene import FlatScene, TRIANGLES, LINES, POINTS   # noqa: F401
from .render import Scene, device_count, init, init_device, tonemap, write_png   # noqa: F401
from ._lib import Stats, YrtError   # noqa: F401

__all__ = ["FlatScene", "Scene", "Stats", "YrtError", "device_count", "init", "init_device", "tonemap", "write_png", "TRIANGLES", "LINES", "POINTS"]
