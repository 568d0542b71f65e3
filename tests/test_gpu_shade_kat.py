"""GPU parity, deterministic per-ray check of the surface interaction.

The path kernels draw their random numbers from Philox, the reference from its LCG, so images can
only be compared statistically (test_gpu_render.py).  b200rt_debug_shade runs the SAME device
function the kernels call (closest hit + shade_hit) with the caller's random words, which makes every
branch checkable ray by ray against the oracle's pinned restatement of the reference:

    hit point      ray(t) = origin + t * dir                       ray3d.h:16
    normal         (P - C) / r  |  unit(s1 x s2)                   sphere.h:94, parallelogram.h:234,273
    face rule      dot(dir, n) > 0  =>  inside, normal flipped     hittable.h:56-70
    Lambertian     n + unit-sphere sample, near-zero guard          material.h:64-86
    Metal          reflected(unit(dir), n) + fuzz * sample, absorbed if n . s < 0   material.h:116-139
    Dielectric     TIR / Schlick choice, reflected | refracted      material.h:175-218, vec3d.h:144-200
    DiffuseLight   emits intensity * colour on both faces, ends     material.h:248-263

Tolerances: hit index, t and hit point bit-exact; reflected / refracted directions 1e-12 relative
(FP64, same formulas, operation order of a sum may differ); directions that add a sampled unit vector
1e-6 absolute (the sample is FP32: sincospif vs numpy); colours 1e-6 relative (FP32).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


def _u01(w):
    return np.float32(int(w) >> 8) * np.float32(1.0 / 16777216.0)


def _sample_unit_sphere(w0, w1):
    """shade.cuh sample_unit_sphere in FP32: z = 1 - 2 u0, azimuth 2 pi u1."""
    u1, u2 = _u01(w0), _u01(w1)
    cz = np.float32(1.0) - np.float32(2.0) * u1
    r = np.sqrt(np.maximum(np.float32(0.0), np.float32(1.0) - cz * cz), dtype=np.float32)
    ang = np.float64(2.0) * np.float64(u2) * np.pi
    return np.array([np.float64(r) * np.cos(ang), np.float64(r) * np.sin(ang), np.float64(cz)])


def _unit(v):
    return v * (1.0 / np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))       # operator/= is *= 1/d (vec3d.h:31)


@pytest.mark.parametrize("name", ["rtow_lights", "rtow_final", "cornell", "xmas", "quads"])
def test_one_surface_interaction_per_ray_matches_oracle(golden, name):
    import cpp_raytracer_b200 as rt
    from cpp_raytracer_b200 import capi
    import pt_oracle

    if not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", f"{name}.scene.gz")):
        pytest.skip(f"no golden scene {name}")
    scene = golden.scene(name)
    rays, tmin, tmax = golden.rays(name)
    n = min(len(rays), 6000)
    rays = rays[:n]
    rng = np.random.default_rng(20231225)
    rnd = rng.integers(0, 2 ** 32, size=(n, 4), dtype=np.uint32)
    rnd[0::3, 2] = 0                      # u = 0: Dielectric reflects whenever the Schlick term is positive
    rnd[1::3, 2] = 0xFFFFFFFF             # u ~ 1: refracts unless total internal reflection
    with rt.DeviceSceneHandle(scene) as dev:
        rec = dev.debug_shade(rays, rnd, tmin, tmax)
        prim_rc, t_rc = dev.raycast(rays, tmin, tmax)
    prim_ref, t_ref = golden.hits(name, brute=True)
    assert np.array_equal(rec["prim"], prim_rc) and np.array_equal(rec["t"][prim_rc >= 0], t_rc[prim_rc >= 0])
    assert np.array_equal(rec["prim"], prim_ref[:n]) and np.array_equal(rec["t"][prim_ref[:n] >= 0], t_ref[:n][prim_ref[:n] >= 0])

    where = {}
    for i, s in enumerate(scene.spheres):
        where[int(s["prim"])] = ("s", i)
    for i, q in enumerate(scene.quads):
        where[int(q["prim"])] = ("q", i)
    seen = {"lambert": 0, "metal": 0, "metal_absorbed": 0, "glass_reflect": 0, "glass_refract": 0, "glass_tir": 0, "light": 0,
            "inside": 0, "miss": 0}
    for i in range(n):
        r = rec[i]
        if r["prim"] < 0:
            seen["miss"] += 1
            assert r["flags"] == 0 and not r["scattered"].any() and not r["emit"].any()
            continue
        o, d, t = rays[i, :3], rays[i, 3:], r["t"]
        P = o + t * d
        kind, idx = where[int(r["prim"])]
        if kind == "s":
            s = scene.spheres[idx]
            nrm = (P - s["c"]) * (1.0 / s["r"])
            mat = scene.materials[int(s["mat"])]
        else:
            q = scene.quads[idx]
            nrm = _unit(np.cross(q["s1"], q["s2"]))
            mat = scene.materials[int(q["mat"])]
        inside = float(d @ nrm) > 0
        if inside:
            nrm = -nrm
            seen["inside"] += 1
        k = int(mat["kind"])
        rgb = mat["rgb"].astype(np.float64)
        if k == capi.MAT_LIGHT:
            seen["light"] += 1
            assert r["flags"] == 0
            assert np.allclose(r["emit"], rgb * mat["param"], rtol=1e-6)
            continue
        assert not r["emit"].any()
        if k == capi.MAT_LAMBERTIAN:
            seen["lambert"] += 1
            want = nrm + _sample_unit_sphere(rnd[i, 0], rnd[i, 1])
            if np.all(np.abs(want) < 1e-8):
                want = nrm
            assert r["flags"] == 1 and np.array_equal(r["scattered"][:3], P)
            assert np.allclose(r["scattered"][3:], want, rtol=0, atol=1e-6)
            assert np.allclose(r["atten"], rgb, rtol=1e-6)
            continue
        v = _unit(d)
        refl = pt_oracle.reflected(v, nrm)
        if k == capi.MAT_METAL:
            fuzz = min(float(mat["param"]), 1.0)
            want = refl + fuzz * _sample_unit_sphere(rnd[i, 0], rnd[i, 1])
            margin = float(nrm @ want)
            if abs(margin) < 1e-5:
                continue                     # on the absorption boundary within the FP32 sample's accuracy
            if margin < 0:
                seen["metal_absorbed"] += 1
                assert r["flags"] == 0
            else:
                seen["metal"] += 1
                assert r["flags"] == 1 and np.array_equal(r["scattered"][:3], P)
                assert np.allclose(r["scattered"][3:], want, rtol=0, atol=1e-6 if fuzz > 0 else 1e-12)
                assert np.allclose(r["atten"], rgb, rtol=1e-6)
            continue
        assert k == capi.MAT_DIELECTRIC
        ior = float(mat["param"])
        eta = ior if inside else 1.0 / ior
        cos_theta = min(float(-(v @ nrm)), 1.0)
        sin_theta = np.sqrt(1.0 - cos_theta * cos_theta)
        u = float(_u01(rnd[i, 2]))
        if eta * sin_theta > 1.0:
            seen["glass_tir"] += 1
            want = refl
        else:
            schlick = pt_oracle.reflectance(cos_theta, eta)
            if abs(u - schlick) < 1e-9:
                continue
            if u < schlick:
                seen["glass_reflect"] += 1
                want = refl
            else:
                seen["glass_refract"] += 1
                want = pt_oracle.refracted(v, nrm, eta)
        assert r["flags"] == 1 and np.array_equal(r["scattered"][:3], P)
        assert np.allclose(r["scattered"][3:], want, rtol=1e-12, atol=1e-14), (i, r["scattered"][3:], want)
        assert np.array_equal(r["atten"], np.ones(3, dtype=np.float32))
    print(name, seen)
    if name == "rtow_lights":
        for key in ("lambert", "metal", "glass_reflect", "glass_refract", "light", "inside"):
            assert seen[key] > 0, (key, seen)
    if name == "cornell":
        assert seen["lambert"] > 0 and seen["light"] > 0


@pytest.mark.parametrize("name", ["rtow_lights", "rtow_final", "cornell", "xmas"])
def test_primary_rays_match_oracle_draw_for_draw(golden, name):
    """camera_ray on the device vs the oracle's random_ray_through_pixel (camera.h:184-200; pinned bit-exactly
    against the reference by the single-thread render test) fed the same draws.  Pinhole cameras: bit-exact.
    Defocus: the disk point is an FP32 sqrt / sincospif product on the device, so the origin agrees to 1e-6 of
    the disk radius and the direction follows (target point bit-exact)."""
    from cpp_raytracer_b200 import capi
    import pt_oracle
    scene = golden.scene(name)
    cam = scene.camera.copy()
    w, h = int(cam["image_w"][0]), int(cam["image_h"][0])
    rng = np.random.default_rng(99)
    n = 4000
    pix = np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], axis=1).astype(np.uint32)
    pix[:4] = [[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]]
    rnd = rng.integers(0, 2 ** 32, size=(n, 4), dtype=np.uint32)
    rnd[0] = 0
    rnd[1] = 0xFFFFFFFF
    got = capi.debug_camera_rays(cam, pix, rnd)
    defocus = float(cam["defocus_angle"][0]) > 0
    disk = max(np.linalg.norm(cam["disk_x"][0]), np.linalg.norm(cam["disk_y"][0]))
    for i in range(n):
        u = [np.float32(int(x) >> 8) * np.float32(1.0 / 16777216.0) for x in rnd[i]]
        r1, r2 = float(u[0]) - 0.5, float(u[1]) - 0.5
        rad = np.sqrt(u[2], dtype=np.float32)
        ang = 2.0 * float(u[3]) * np.pi
        vx, vy = float(rad) * np.cos(ang), float(rad) * np.sin(ang)
        want = pt_oracle.camera_ray(cam, int(pix[i, 1]), int(pix[i, 0]), vx, vy, r1, r2)
        if not defocus:
            assert np.array_equal(got[i], want), (i, got[i], want)
        else:
            assert np.allclose(got[i, :3], want[:3], rtol=0, atol=2e-6 * disk + 1e-15)
            assert np.allclose(got[i, :3] + got[i, 3:], want[:3] + want[3:], rtol=1e-15, atol=1e-15)   # the target point
