"""End to end through the ONE-CALL C-ABI entry on N GPUs of this process (b200rt_render_scene_multi: build once on
devices[0] + copy to peers + sample-split render + exchange + read back), host buffers in, host buffer out:
    python scripts/multi_e2e.py C5 [n_devices] [spp] [iters]"""
import os, sys, json, subprocess, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, build, capi
CFG = {"C1": ("rtow_final", 1200, 675, 500, 20), "C2": ("rtow_lights", 1920, 1080, 1024, 20), "C3": ("cornell", 1024, 1024, 4096, 1000),
       "C4": ("xmas", 1920, 1080, 1024, 50), "C4b": ("raining", 1920, 1080, 1024, 50), "C5": ("millions_lights", 3840, 2160, 1024, 20)}
tag = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else capi.device_count()
name, w, h, spp, depth = CFG[tag]
if len(sys.argv) > 3 and int(sys.argv[3]) > 0:
    spp = int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
p = os.path.join(tempfile.mkdtemp(), name + ".scene")
subprocess.run([build.build_host(), name, "dump", p], check=True, capture_output=True)
s = scene_io.load_scene(p)
cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth)
out = np.empty((h, w, 3), np.float32)
for i in range(iters + 1):
    t0 = time.perf_counter()
    _, st, info = rt.render_scene(s, cam, out=out, devices=list(range(n)))
    ms = (time.perf_counter() - t0) * 1e3
    if i == 0:
        continue   # first call pays pool growth + pinned staging allocation
    print(json.dumps({"cfg": tag, "n_devices": n, "spp": spp, "wall_ms": round(ms, 2), "total_ms": round(st["total_ms"], 2),
                      "build_ms": round(st["build_ms"], 2), "replicate_ms": round(st["replicate_ms"], 2),
                      "kernel_ms_max": round(st["kernel_ms"], 2), "exchange_ms": round(st["exchange_ms"], 3),
                      "d2h_ms": round(st["d2h_ms"], 2), "peer_exchange": st["peer_exchange"],
                      "Mpaths/s_e2e": round(st["paths"] / ms / 1e3, 1), "Mpaths/s_kernel": round(st["paths"] / st["kernel_ms"] / 1e3, 1),
                      "mean": float(out.mean())}), flush=True)
