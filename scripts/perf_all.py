"""Perf survey over all BASELINE configs at reduced spp (run on the GPU box).
   python scripts/perf_all.py [spp] [variant]"""
import os, sys, json, subprocess, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, capi, build
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
leaf = int(sys.argv[3]) if len(sys.argv) > 3 else 0
FLAGS = int(os.environ.get("FLAGS", "0"))     # e.g. FLAGS=16: B200RT_FLAG_THREAD_PIXELS (A/B against the tile work pool)
CFG = [("C1", "rtow_final", 1200, 675, 20), ("C2", "rtow_lights", 1920, 1080, 20), ("C3", "cornell", 1024, 1024, 1000),
       ("C4", "xmas", 1920, 1080, 50), ("C4b", "raining", 1920, 1080, 50), ("C5", "millions_lights", 3840, 2160, 20)]
only = os.environ.get("ONLY")
binp = build.build_host()
tmp = tempfile.mkdtemp()
for tag, name, w, h, depth in CFG:
    if only and tag not in only.split(","):
        continue
    p = os.path.join(tmp, name + ".scene")
    subprocess.run([binp, name, "dump", p], check=True, capture_output=True)
    s = scene_io.load_scene(p)
    cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth)
    if os.environ.get("BG"):                         # experiment: override the background (e.g. BG=0 or BG=0.7,0.8,1)
        bg = [float(x) for x in os.environ["BG"].split(",")]
        cam = rt.camera_with(cam, background=bg * 3 if len(bg) == 1 else bg)
    t0 = time.time()
    with rt.DeviceSceneHandle(s, max_leaf_prims=leaf) as d:
        info = d.info()
        d.render(rt.camera_with(cam, spp=max(1, spp // 8)), variant=variant, flags=FLAGS)
        _, st = d.render(cam, variant=variant, flags=FLAGS)
        _, sc = d.render(rt.camera_with(cam, spp=max(1, spp // 8)), flags=capi.FLAG_COUNTERS | FLAGS, variant=variant)
    print(json.dumps({"cfg": tag, "scene": name, "res": f"{w}x{h}", "spp": spp, "prims": info["n_prims"], "nodes": info["n_nodes"],
                      "depth": info["tree_depth"], "build_ms": round(info["build_ms"]), "MB": round(info["device_bytes"] / 1e6, 1),
                      "kernel_ms": round(st["kernel_ms"], 2), "Mpaths/s": round(st["paths"] / st["kernel_ms"] / 1e3, 1),
                      "Mrays/s": round(st["rays"] / st["kernel_ms"] / 1e3, 1), "rays/path": round(st["rays"] / st["paths"], 3),
                      "nodes/ray": round(sc["node_visits"] / sc["rays"], 2), "prims/ray": round(sc["prim_tests"] / sc["rays"], 2)}), flush=True)
