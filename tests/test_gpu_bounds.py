"""Memory-safety net in place of compute-sanitizer (which the GPU pool does not allow; SURVEY section 5 planned it).

cpp_raytracer_b200/libb200rt_dbg.so is the SAME source built with -DB200RT_DEBUG_BOUNDS: every traversal-stack push,
node index, primitive / material index and frame store is range-checked on the device, violations are counted and the
access skipped.  The suite renders and ray-casts every fixture scene plus a GPU-built (LBVH) scene through that build in a
child process (B200RT_LIB selects the library) and requires zero violations and the same image as the production build;
a second child lowers the enforced stack capacity (B200RT_DEBUG_STACK_CAP=2) to prove the check is live."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, RENDER_SCENES

pytestmark = pytest.mark.gpu

CHILD = r'''
import json, os, sys, hashlib
import numpy as np
sys.path.insert(0, os.environ["B200RT_ROOT"])
import cpp_raytracer_b200 as rt
from cpp_raytracer_b200 import capi, scene_io
g = os.path.join(os.environ["B200RT_ROOT"], "tests", "golden")
out = {}
for name in os.environ["SCENES"].split(","):
    scene = scene_io.load_scene(os.path.join(g, name + ".scene.gz"))
    rays, tmin, tmax = scene_io.load_rays(os.path.join(g, name + ".rays.gz"))
    with rt.DeviceSceneHandle(scene) as dev:
        prim, t = dev.raycast(rays, tmin, tmax)
        cam = rt.camera_with(scene.camera, image_w=64, image_h=48, spp=8, max_depth=12)
        img, st = dev.render(cam, seed=3)
        wf, _ = dev.render(cam, seed=3, variant=capi.VARIANT_WAVEFRONT)
        try:
            viol = [int(x) for x in dev.bounds_violations()]
        except capi.B200rtError:
            viol = None
    with rt.DeviceSceneHandle(scene, devices=[0, 0]) as dev2:      # the multi-device path (device listed twice), checked too
        img2, st2 = dev2.render(cam, seed=3)
        try:
            viol2 = [int(x) for x in dev2.bounds_violations()]
        except capi.B200rtError:
            viol2 = None
    out[name] = {"viol": viol, "viol_multi": viol2, "img": hashlib.sha256(img.tobytes()).hexdigest(),
                 "img_multi": hashlib.sha256(img2.tobytes()).hexdigest(),
                 "hits": hashlib.sha256(prim.tobytes() + t.tobytes()).hexdigest(), "rays": st["rays"], "rays_multi": st2["rays"]}
# a scene big enough for the GPU LBVH builder (AUTO switches at 65536 primitives)
rng = np.random.default_rng(5)
n = 70000
sph = np.zeros(n, dtype=capi.SPHERE_DTYPE)
sph["c"] = rng.uniform(-30, 30, size=(n, 3)); sph["r"] = rng.uniform(0.05, 0.4, size=n); sph["mat"] = rng.integers(0, 4, size=n); sph["prim"] = np.arange(n)
mats = np.zeros(4, dtype=capi.MATERIAL_DTYPE)
mats["kind"] = [0, 1, 2, 3]; mats["rgb"] = 0.7; mats["param"] = [0, 0.2, 1.5, 4.0]
base = scene_io.load_scene(os.path.join(g, "rtow_final.scene.gz"))
big = capi.HostScene(mats, sph, np.zeros(0, dtype=capi.QUAD_DTYPE), base.camera, "lbvh")
with rt.DeviceSceneHandle(big) as dev:
    img, st = dev.render(rt.camera_with(base.camera, image_w=96, image_h=64, spp=4), seed=1)
    try:
        viol = [int(x) for x in dev.bounds_violations()]
    except capi.B200rtError:
        viol = None
    out["lbvh"] = {"viol": viol, "img": hashlib.sha256(img.tobytes()).hexdigest(), "rays": st["rays"], "depth": dev.info()["tree_depth"]}
print(json.dumps(out))
'''


def run_child(lib, scenes, extra_env=None):
    env = dict(os.environ, B200RT_ROOT=ROOT, SCENES=",".join(scenes))
    if lib:
        env["B200RT_LIB"] = lib
    env.update(extra_env or {})
    res = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env)
    assert res.returncode == 0, res.stderr[-3000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.fixture(scope="module")
def dbg_lib():
    from cpp_raytracer_b200 import build
    return build.build_debug()


def test_bounds_checked_build_sees_no_violation_and_the_same_results(dbg_lib):
    scenes = RENDER_SCENES + ["pathological"]
    dbg = run_child(dbg_lib, scenes)
    prod = run_child(None, scenes)
    for name in scenes + ["lbvh"]:
        assert dbg[name]["viol"] == [0, 0, 0, 0], (name, dbg[name]["viol"])
        assert prod[name]["viol"] is None                       # the production build has no checks compiled in
        assert dbg[name]["img"] == prod[name]["img"], name      # same arithmetic, bit for bit
        assert dbg[name]["rays"] == prod[name]["rays"], name
        if "viol_multi" in dbg[name]:
            assert dbg[name]["viol_multi"] == [0, 0, 0, 0], (name, dbg[name]["viol_multi"])
            assert dbg[name]["img_multi"] == prod[name]["img_multi"] and dbg[name]["rays_multi"] == prod[name]["rays"], name
        if "hits" in dbg[name]:
            assert dbg[name]["hits"] == prod[name]["hits"], name


def test_bounds_check_is_live(dbg_lib):
    """With the enforced stack capacity lowered to 2 entries the same renders MUST report stack violations (and
    still terminate: the offending pushes are dropped)."""
    got = run_child(dbg_lib, ["rtow_final"], {"B200RT_DEBUG_STACK_CAP": "2"})
    assert got["rtow_final"]["viol"][0] > 0
    assert got["rtow_final"]["viol"][1:] == [0, 0, 0]
