"""Tiny end-to-end run for compute-sanitizer (raycast + both render variants + tonemap)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, capi
G = os.path.join(ROOT, "tests", "golden")
for name in ("xmas", "cornell", "pathological"):
    s = scene_io.load_scene(f"{G}/{name}.scene.gz")
    rays, tmin, tmax = scene_io.load_rays(f"{G}/{name}.rays.gz")
    with rt.DeviceSceneHandle(s) as d:
        p, t = d.raycast(rays[:2000], tmin, tmax)
        if name != "pathological":
            cam = rt.camera_with(s.camera, image_w=40, image_h=24, spp=4, max_depth=12)
            a, _ = d.render(cam, variant=capi.VARIANT_MEGAKERNEL)
            b, _ = d.render(cam, variant=capi.VARIANT_WAVEFRONT)
            c, _ = d.render(cam, variant=capi.VARIANT_MEGAKERNEL_VOTED)
            rt.tonemap(a)
            print(name, float(a.mean()), float(b.mean()), float(c.mean()))
print("sanitize run done")
