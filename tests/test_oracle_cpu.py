"""Pins the oracle (oracle/pt_oracle.c, the plain-C restatement of the reference's path) against
the REFERENCE ITSELF: tests/golden/* was produced by oracle/_ref/ref_bridge, i.e. the unmodified
reference headers compiled in this container (tests/golden/make_golden.py).  The reference has
no tests or golden vectors of its own, so these are the pins.
CPU only; the oracle is never used by the product."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, SMALL_SCENES

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pt_oracle  # noqa: E402


@pytest.mark.parametrize("name", SMALL_SCENES)
def test_oracle_closest_hit_bit_exact_vs_reference(golden, name):
    scene = golden.scene(name)
    rays, tmin, tmax = golden.rays(name)
    want_p, want_t = golden.hits(name, brute=True)     # reference Scene::hit_by (scene.h:59-75)
    if name == "xmas":                                 # 4202 prims x 5745 rays brute force: keep it quick
        rays, want_p, want_t = rays[::3], want_p[::3], want_t[::3]
    p, t = pt_oracle.raycast_brute(scene, rays, tmin, tmax)
    assert np.array_equal(p, want_p)
    assert np.array_equal(t, want_t)                   # bit-exact hit times


@pytest.mark.parametrize("name", ["rtow_final", "rtow_lights", "quads", "cornell", "xmas"])
def test_oracle_render_bit_exact_vs_reference_single_thread(name):
    """Same LCG state in, the same double-precision pixels out as Camera::render<BVH> on one
    thread: pins ray generation (incl. defocus disk), all four materials, recursion depth
    handling, the RNG draw ORDER (camera.h:197-198 is unsequenced; g++ draws the delta_y factor
    first) and the accumulation."""
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import scene_io
    g = np.load(os.path.join(GOLDEN, f"{name}.render1t.npz"))
    scene = scene_io.load_scene(os.path.join(GOLDEN, f"{name}.scene.gz"))
    cam = rt.camera_with(scene.camera, image_w=int(g["w"]), image_h=int(g["h"]), spp=int(g["spp"]), max_depth=int(g["max_depth"]))
    img, rays, _ = pt_oracle.render(scene, cam, int(g["lcg_state"]))
    assert rays > 0
    assert np.array_equal(img, g["image"]), f"{name}: {(img != g['image']).any(axis=2).sum()} pixels differ"


@pytest.mark.parametrize("name", SMALL_SCENES)
def test_oracle_camera_init_bit_exact(golden, name):
    scene = golden.scene(name)
    cam = scene.camera.copy()
    want = {k: cam[k].copy() for k in ("pixel00", "delta_x", "delta_y", "disk_x", "disk_y")}
    for k in want:
        cam[k] = 0
    got = pt_oracle.camera_init(cam)
    for k, v in want.items():
        assert np.array_equal(got[k], v)


def test_oracle_lcg_known_answers(golden):
    """rand_util.h:51-117: thread seed = SeedSeqGenerator::next_seed() after set_seed(12345), then
    the 32-bit LCG; also the closed forms SURVEY.md section 4 quotes."""
    kat = golden.kat()
    seed = kat["rand_double_seed"]
    first = pt_oracle.seed_sequence_next(seed)
    assert first == (2483477 * seed + 2987434823) % 2**32
    vals, _ = pt_oracle.rand_doubles(first, 16)
    assert vals == kat["rand_double_next16"]
    s, want = first, []
    for _ in range(16):
        s = (1664525 * s + 1013904223) % 2**32
        want.append(s / (2**32 - 2))
    assert np.allclose(vals, want, rtol=0, atol=1e-16)


def test_oracle_reflect_refract_schlick_tonemap_known_answers(golden):
    kat = golden.kat()
    d, n = kat["reflect"]["d"], kat["reflect"]["n"]
    assert pt_oracle.reflected(d, n).tolist() == kat["reflect"]["out"]
    for case in kat["refract"]:
        out = pt_oracle.refracted(d, n, case["eta"])
        assert (out is not None) == case["ok"]
        if case["ok"]:
            assert out.tolist() == case["out"]
    for c, eta, want in kat["reflectance"]:
        assert pt_oracle.reflectance(c, eta) == want
    rgb = np.array([k["rgb"] for k in kat["tonemap"]])
    want = np.array([[int(x) for x in k["out"].split()] for k in kat["tonemap"]])
    assert np.array_equal(pt_oracle.tonemap(rgb), want)


def test_oracle_ties_go_to_lowest_index(golden):
    """Two coincident spheres: Scene::hit_by keeps the first (later ones must be strictly closer)."""
    from cpp_raytracer_b200 import capi
    sph = np.zeros(2, capi.SPHERE_DTYPE)
    sph["c"] = [[0, 0, -5], [0, 0, -5]]
    sph["r"] = 1.0
    sph["prim"] = [0, 1]
    scene = capi.HostScene(np.zeros(1, capi.MATERIAL_DTYPE), sph, np.zeros(0, capi.QUAD_DTYPE), np.zeros(1, capi.CAMERA_DTYPE))
    p, t = pt_oracle.raycast_brute(scene, np.array([[0.0, 0, 0, 0, 0, -1]]))
    assert p[0] == 0 and t[0] == 4.0
