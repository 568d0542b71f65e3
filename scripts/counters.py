"""Traversal work per ray (B200RT_FLAG_COUNTERS) for the bench scenes: python scripts/counters.py [spp]"""
import os, sys, subprocess, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, build, capi
CFG = {"C1": ("rtow_final", 1200, 675, 20), "C2": ("rtow_lights", 1920, 1080, 20), "C3": ("cornell", 1024, 1024, 1000),
       "C4": ("xmas", 1920, 1080, 50), "C4b": ("raining", 1920, 1080, 50), "C5": ("millions_lights", 3840, 2160, 20)}
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for tag in os.environ.get("ONLY", "C1,C2,C3,C4,C5").split(","):
    name, w, h, depth = CFG[tag]
    p = os.path.join(tempfile.mkdtemp(), name + ".scene")
    subprocess.run([build.build_host(), name, "dump", p], check=True, capture_output=True)
    s = scene_io.load_scene(p)
    cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth)
    with rt.DeviceSceneHandle(s) as d:
        _, st = d.render(cam, flags=capi.FLAG_COUNTERS)
        print(json.dumps({"cfg": tag, "rays_per_path": st["rays"] / st["paths"], "nodes_per_ray": st["node_visits"] / st["rays"],
                          "prim_tests_per_ray": st["prim_tests"] / st["rays"], "depth": d.info()["tree_depth"]}))
