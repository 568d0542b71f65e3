"""One render of one config for ncu (run on the GPU box): python scripts/prof_one.py C5 8"""
import os, sys, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, build
CFG = {"C1": ("rtow_final", 1200, 675, 20), "C2": ("rtow_lights", 1920, 1080, 20), "C3": ("cornell", 1024, 1024, 1000),
       "C4": ("xmas", 1920, 1080, 50), "C4b": ("raining", 1920, 1080, 50), "C5": ("millions_lights", 3840, 2160, 20)}
tag = sys.argv[1]; spp = int(sys.argv[2]); variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
name, w, h, depth = CFG[tag]
p = os.path.join(tempfile.mkdtemp(), name + ".scene")
subprocess.run([build.build_host(), name, "dump", p], check=True, capture_output=True)
s = scene_io.load_scene(p)
cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth)
with rt.DeviceSceneHandle(s) as d:
    d.render(rt.camera_with(cam, spp=1), variant=variant)
    _, st = d.render(cam, variant=variant)
    print(tag, st["kernel_ms"], st["paths"] / st["kernel_ms"] / 1e3)
