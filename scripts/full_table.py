"""All BASELINE configs: GPU (this repo) vs the unmodified reference on this box's host cores.
   python scripts/full_table.py [gpu_spp] > profiles/table.md   (run on the GPU box)"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, build
gpu_spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
CFG = [("C1", "rtow_final", 1200, 675, 500, 20, 16), ("C2", "rtow_lights", 1920, 1080, 1024, 20, 16),
       ("C3", "cornell", 1024, 1024, 4096, 1000, 16), ("C4", "xmas", 1920, 1080, 1024, 50, 8),
       ("C4b", "raining", 1920, 1080, 1024, 50, 8), ("C5", "millions_lights", 3840, 2160, 1024, 20, 2)]
BRIDGE = os.path.join(ROOT, "oracle", "_ref", "ref_bridge")
cores = len(os.sched_getaffinity(0))
binp = build.build_host()
tmp = tempfile.mkdtemp()
print(f"| config | scene (prims) | resolution, spp, depth | GPU Mpaths/s | GPU Mrays/s | BVH build ms (ours) | reference CPU Mpaths/s ({cores} threads) | reference BVH build ms | speed-up |")
print("|---|---|---|---|---|---|---|---|---|")
for tag, name, w, h, spp, depth, cpu_spp in CFG:
    p = os.path.join(tmp, name + ".scene")
    subprocess.run([binp, name, "dump", p], check=True, capture_output=True)
    s = scene_io.load_scene(p)
    cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=gpu_spp, max_depth=depth)
    with rt.DeviceSceneHandle(s) as d:
        info = d.info()
        d.render(rt.camera_with(cam, spp=8))
        _, st = d.render(cam)
    res = subprocess.run([BRIDGE, name, "--w", str(w), "--h", str(h), "--spp", str(cpu_spp), "--depth", str(depth),
                          "--threads", str(cores), "render", "-"], capture_output=True, text=True)
    r = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    g = st["paths"] / st["kernel_ms"] / 1e3
    print(f"| {tag} | {name} ({info['n_prims']}) | {w}x{h}, {spp} spp (GPU timed at {gpu_spp}, CPU at {cpu_spp}), depth {depth} | {g:.0f} | "
          f"{st['rays'] / st['kernel_ms'] / 1e3:.0f} | {info['build_ms']:.0f} | {r['mpaths_per_s']:.1f} | {r['bvh_build_ms']:.0f} | {g / r['mpaths_per_s']:.0f}x |", flush=True)
