"""Image convergence at a config's own resolution and (near) its own spp: GPU frame vs two
independent renders of the unmodified reference.   python scripts/converge_full.py C2 1024"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, build
CFG = {"C1": ("rtow_final", 1200, 675, 20), "C2": ("rtow_lights", 1920, 1080, 20), "C3": ("cornell", 1024, 1024, 1000),
       "C4": ("xmas", 1920, 1080, 50), "C4b": ("raining", 1920, 1080, 50), "C5": ("millions_lights", 3840, 2160, 20)}
tag, spp = sys.argv[1], int(sys.argv[2])
name, w, h, depth = CFG[tag]
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "ref_bridge")
tmp = tempfile.mkdtemp()
p = os.path.join(tmp, name + ".scene")
subprocess.run([build.build_host(), name, "dump", p], check=True, capture_output=True)
s = scene_io.load_scene(p)
a, b = os.path.join(tmp, "a.hdr"), os.path.join(tmp, "b.hdr")
t0 = time.time()
res = subprocess.run([BRIDGE, name, "--w", str(w), "--h", str(h), "--spp", str(spp), "--depth", str(depth), "--render-seed", "5",
                      "render", a, "render", b], capture_output=True, text=True)
cpu = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
with rt.DeviceSceneHandle(s) as d:
    G, st = d.render(rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth), seed=2026)
def tone(img):
    img = np.asarray(img, np.float64); lum = 0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2]
    return np.sqrt(np.clip(img / (1 + lum[..., None]), 0, None))
A, B = scene_io.load_hdr(a), scene_io.load_hdr(b)
tA, tB, tG = tone(A), tone(B), tone(G)
rm = lambda x, y: float(np.sqrt(np.mean((x - y) ** 2)))
blk = lambda x, k: x[: h // k * k, : w // k * k].reshape(h // k, k, w // k, k, 3).mean(axis=(1, 3))
lum = lambda x: float((0.2126 * x[..., 0] + 0.7152 * x[..., 1] + 0.0722 * x[..., 2]).mean())
out = {"config": tag, "scene": name, "res": f"{w}x{h}", "spp": spp, "depth": depth,
       "gpu_kernel_ms": round(st["kernel_ms"], 1), "gpu_mpaths_s": round(st["paths"] / st["kernel_ms"] / 1e3, 1),
       "cpu_seconds_per_render": [round(c["seconds"], 1) for c in cpu], "cpu_mpaths_s": [round(c["mpaths_per_s"], 1) for c in cpu],
       "rmse_gpu_vs_ref": round(max(rm(tG, tA), rm(tG, tB)), 6), "rmse_ref_vs_ref": round(rm(tA, tB), 6),
       "ratio": round(max(rm(tG, tA), rm(tG, tB)) / rm(tA, tB), 4),
       "ratio_8x8_blocks": round(max(rm(blk(tG, 8), blk(tA, 8)), rm(blk(tG, 8), blk(tB, 8))) / rm(blk(tA, 8), blk(tB, 8)), 4),
       "ratio_32x32_blocks": round(max(rm(blk(tG, 32), blk(tA, 32)), rm(blk(tG, 32), blk(tB, 32))) / rm(blk(tA, 32), blk(tB, 32)), 4),
       "mean_linear_luminance": {"gpu": round(lum(G), 6), "refA": round(lum(A), 6), "refB": round(lum(B), 6)},
       "rays_per_path_gpu": round(st["rays"] / st["paths"], 4)}
print(json.dumps(out))
