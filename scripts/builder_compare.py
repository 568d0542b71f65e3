"""Host SAH vs GPU LBVH: build time, tree size, traversal work and render rate (run on the GPU box)."""
import os, sys, json, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, capi, build
CFG = [("C2", "rtow_lights", 1920, 1080, 20, 64), ("C4", "xmas", 1920, 1080, 50, 64), ("C4b", "raining", 1920, 1080, 50, 64),
       ("C5", "millions_lights", 3840, 2160, 20, 32)]
binp = build.build_host()
tmp = tempfile.mkdtemp()
for tag, name, w, h, depth, spp in CFG:
    p = os.path.join(tmp, name + ".scene")
    subprocess.run([binp, name, "dump", p], check=True, capture_output=True)
    s = scene_io.load_scene(p)
    cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth)
    for bname, b in (("host_sah", capi.BUILDER_HOST_SAH), ("gpu_lbvh", capi.BUILDER_GPU_LBVH)):
        for rep in range(2):
            with rt.DeviceSceneHandle(s, builder=b) as d:
                info = d.info()
                if rep == 0:
                    continue
                d.render(rt.camera_with(cam, spp=4))
                _, st = d.render(cam)
                _, sc = d.render(rt.camera_with(cam, spp=4), flags=capi.FLAG_COUNTERS)
        print(json.dumps({"cfg": tag, "builder": bname, "build_ms": round(info["build_ms"], 2), "upload_ms": round(info["upload_ms"], 2),
                          "nodes": info["n_nodes"], "depth": info["tree_depth"], "stack": info["stack_entries"],
                          "Mpaths/s": round(st["paths"] / st["kernel_ms"] / 1e3, 1), "nodes/ray": round(sc["node_visits"] / sc["rays"], 2),
                          "prims/ray": round(sc["prim_tests"] / sc["rays"], 2)}), flush=True)
