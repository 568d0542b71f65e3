import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io
for name in ("rtow_lights", "xmas"):
    s = scene_io.load_scene(os.path.join(ROOT, "tests", "golden", f"{name}.scene.gz"))
    cam = rt.camera_with(s.camera, image_w=1920, image_h=1080, spp=64, max_depth=50 if name == "xmas" else 20)
    out = np.empty((1080, 1920, 3), np.float32)
    for i in range(3):
        t = time.perf_counter()
        _, st, info = rt.render_scene(s, cam, out=out)
        dt = (time.perf_counter() - t) * 1e3
        print(name, i, "wall %.1f ms" % dt, {k: round(v, 2) if isinstance(v, float) else v for k, v in st.items() if k in ("kernel_ms", "d2h_ms", "total_ms", "h2d_ms")},
              "build %.2f upload %.2f" % (info["build_ms"], info["upload_ms"]))
    with rt.DeviceSceneHandle(s) as d:
        for i in range(2):
            t = time.perf_counter(); _, st = d.render(cam); dt = (time.perf_counter() - t) * 1e3
            print(name, "resident", "wall %.1f ms" % dt, round(st["kernel_ms"], 2), round(st["d2h_ms"], 2))
