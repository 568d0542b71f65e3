"""Where the 32 lanes of a warp are at every executed step of the path kernel's schedule
(b200rt_debug_lane_accounting), all BASELINE configs at full resolution, reduced spp:
    python scripts/lane_accounting.py [spp] > gpurun_out/lanes.jsonl"""
import os, sys, json, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import scene_io, build
CFG = {"C1": ("rtow_final", 1200, 675, 20), "C2": ("rtow_lights", 1920, 1080, 20), "C3": ("cornell", 1024, 1024, 1000),
       "C4": ("xmas", 1920, 1080, 50), "C4b": ("raining", 1920, 1080, 50), "C5": ("millions_lights", 3840, 2160, 20)}
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
FLAGS = int(os.environ.get("FLAGS", "0"))     # 16 = B200RT_FLAG_THREAD_PIXELS (round-1 work distribution)
# instruction weights per executed step (SASS counts, profiles/r1_final_megakernel_C2_1024spp_ncu.md)
W_NODE, W_PRIM_SPHERE, W_PRIM_QUAD, W_SHADE = 160.0, 135.0, 150.0, 350.0
for tag in os.environ.get("ONLY", "C1,C2,C3,C4,C4b,C5").split(","):
    name, w, h, depth = CFG[tag]
    p = os.path.join(tempfile.mkdtemp(), name + ".scene")
    subprocess.run([build.build_host(), name, "dump", p], check=True, capture_output=True)
    s = scene_io.load_scene(p)
    cam = rt.camera_with(s.camera, image_w=w, image_h=h, spp=spp, max_depth=depth)
    with rt.DeviceSceneHandle(s) as d:
        a = [int(x) for x in d.lane_accounting(cam, flags=FLAGS)]
    wp = W_PRIM_QUAD if len(s.quads) > len(s.spheres) else W_PRIM_SPHERE
    slots = 32.0 * (a[0] * W_SHADE + a[2] * W_NODE + a[12] * wp)          # lane-slots issued, instruction weighted
    used = a[1] * W_SHADE + a[3] * W_NODE + a[13] * wp
    out = {"cfg": tag, "spp": spp, "work_distribution": "thread owns pixel" if FLAGS & 16 else "tile work pool", "raw": a,
           "lanes_per_shade": a[1] / max(1, a[0]), "lanes_per_node_step": a[3] / max(1, a[2]),
           "lanes_per_leaf_step": a[8] / max(1, a[7]), "lanes_per_prim_test": a[13] / max(1, a[12]),
           "node_execs_per_ray": 32.0 * a[2] / max(1, a[15]), "prim_execs_per_ray": 32.0 * a[12] / max(1, a[15]),
           "node_step_idle": {"at_leaf": a[4] / (32.0 * a[2]), "done_waiting": a[5] / (32.0 * a[2]), "finished_pixel": a[6] / (32.0 * a[2])},
           "leaf_step_idle": {"at_node": a[9] / (32.0 * max(1, a[7])), "done_waiting": a[10] / (32.0 * max(1, a[7])),
                              "finished_pixel": a[11] / (32.0 * max(1, a[7]))},
           "weighted_lane_utilisation": used / slots,
           "weighted_share": {"shade": a[0] * W_SHADE * 32 / slots, "node": a[2] * W_NODE * 32 / slots, "prim": a[12] * wp * 32 / slots}}
    print(json.dumps(out), flush=True)
