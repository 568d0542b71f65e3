"""GPU tests of the API-compatible C++ host layer: a reference-style program
using BVH reuse (`BVH bvh(world); cam.render(bvh)`, camera.h:264-297 / bvh.h:754-776), closest-hit
queries through `BVH::hit_by` (bvh.h:585-715 -> b200rt_raycast) and progressive rendering, compiled
with g++ against host/include + libb200rt.so and run on the device."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include <cstdio>
#include "util/rand_util.h"
#include "base/scene.h"
#include "base/material.h"
#include "base/camera.h"
#include "acceleration/bvh.h"
#include "shapes/shapes.h"

static bool same(const Image &a, const Image &b) {
    for (size_t r = 0; r < a.height(); ++r)
        for (size_t c = 0; c < a.width(); ++c)
            if (a[r][c].r != b[r][c].r || a[r][c].g != b[r][c].g || a[r][c].b != b[r][c].b) return false;
    return true;
}
static double maxrel(const Image &a, const Image &b) {
    double m = 0;
    for (size_t r = 0; r < a.height(); ++r)
        for (size_t c = 0; c < a.width(); ++c) {
            const double x[3] = {a[r][c].r, a[r][c].g, a[r][c].b}, y[3] = {b[r][c].r, b[r][c].g, b[r][c].b};
            for (int k = 0; k < 3; ++k) m = std::max(m, std::fabs(x[k] - y[k]) / (std::fabs(y[k]) + 1e-6));
        }
    return m;
}

int main() {
    SeedSeqGenerator::get_instance().set_seed(7);
    Scene world;
    auto ground = std::make_shared<Lambertian>(RGB::from_mag(0.5, 0.5, 0.5));
    auto glass = std::make_shared<Dielectric>(1.5);
    auto metal = std::make_shared<Metal>(RGB::from_mag(0.7, 0.6, 0.5), 0.1);
    auto light = std::make_shared<DiffuseLight>(RGB::from_mag(1), 4);
    world.add(std::make_shared<Sphere>(Point3D(0, -1000, 0), 1000, ground));          // primitive 0
    world.add(std::make_shared<Sphere>(Point3D(0, 1, 0), 1.0, glass));                // primitive 1
    world.add(std::make_shared<Parallelogram>(Point3D(-2, 0, -3), Vec3D(4, 0, 0), Vec3D(0, 4, 0), metal));   // primitive 2
    world.add(std::make_shared<Box>(Point3D(2, 0, 2), Point3D(3, 1, 3), light));      // primitives 3..8
    Camera cam;
    cam.set_image_by_width_and_aspect_ratio(96, 16. / 9.).set_vertical_fov(30).set_camera_center(Point3D{10, 3, 6})
       .set_camera_lookat(Point3D{0, 1, 0}).set_samples_per_pixel(24).set_max_depth(8).set_background(RGB::from_mag(0.7, 0.8, 1));

    BVH bvh(world);
    const Image direct = cam.render(world);            // build + upload + render + free in one call
    const Image resident = cam.render(bvh);            // resident scene
    const Image again = cam.render(bvh);
    BVH copy = bvh;                                     // copies share the device scene
    const Image from_copy = cam.render(copy);
    size_t passes = 0, last_done = 0;
    const Image prog = cam.render_progressive(bvh, 5, [&](const Image &, size_t done) { ++passes; last_done = done; });

    // closest-hit queries (hit_info semantics, hittable.h:46-71)
    const auto down = bvh.hit_by(Ray3D{Point3D{0, 5, 0}, Vec3D{0, -1, 0}}, Interval::with_min(0.00001));       // top of the glass sphere
    const auto inside = bvh.hit_by(Ray3D{Point3D{0, 1, 0}, Vec3D{0, 0, 2}}, Interval::with_min(0.00001));      // from its centre outwards
    const auto quad = bvh.hit_by(Ray3D{Point3D{0.5, 1, 5}, Vec3D{0, 0, -1}}, Interval(6.0, 100.0));            // sphere excluded by the interval
    const auto miss = bvh.hit_by(Ray3D{Point3D{0, 5, 0}, Vec3D{0, 1, 0}}, Interval::with_min(0.00001));
    std::vector<Ray3D> many;
    for (int i = 0; i < 1000; ++i) many.push_back(Ray3D{Point3D{10, 3, 6}, Vec3D{-10 + 0.004 * i, -2.5, -6 + 0.003 * i}});
    const auto batch = bvh.hit_by(std::span<const Ray3D>(many), Interval::with_min(0.00001));
    size_t hits = 0, agree = 0;
    for (size_t i = 0; i < many.size(); ++i) {
        if (batch[i]) ++hits;
        const auto one = bvh.hit_by(many[i], Interval::with_min(0.00001));
        if (one.has_value() == batch[i].has_value() && (!one || (one->hit_time == batch[i]->hit_time && one->material == batch[i]->material))) ++agree;
    }

    // a single primitive rendered through render<T> (camera.h:264-266)
    Sphere lone(Point3D(0, 1, 0), 1.0, ground);
    Camera cam2 = cam;
    cam2.set_samples_per_pixel(8);
    Scene lone_scene;
    lone_scene.add(std::make_shared<Sphere>(Point3D(0, 1, 0), 1.0, ground));
    const bool lone_same = same(cam2.render(lone), cam2.render(lone_scene));

    std::printf("{\"direct_eq_resident\": %d, \"resident_repeatable\": %d, \"copy_eq\": %d, \"passes\": %zu, \"last_done\": %zu, "
                "\"prog_maxrel\": %.3g, \"down_t\": %.17g, \"down_ny\": %.17g, \"down_outside\": %d, \"down_is_glass\": %d, "
                "\"inside_t\": %.17g, \"inside_nz\": %.17g, \"inside_outside\": %d, \"quad_t\": %.17g, \"quad_nz\": %.17g, "
                "\"quad_is_metal\": %d, \"miss\": %d, \"batch_hits\": %zu, \"batch_agree\": %zu, \"lone_same\": %d, \"nodes\": %llu}\n",
                (int)same(direct, resident), (int)same(resident, again), (int)same(resident, from_copy), passes, last_done,
                maxrel(prog, resident), down->hit_time, down->unit_surface_normal.y, (int)down->hit_from_outside,
                (int)(down->material == glass.get()), inside->hit_time, inside->unit_surface_normal.z, (int)inside->hit_from_outside,
                quad->hit_time, quad->unit_surface_normal.z, (int)(quad->material == metal.get()), (int)!miss.has_value(), hits, agree,
                (int)lone_same, (unsigned long long)bvh.info().n_nodes);
    return 0;
}
'''


def test_reference_style_bvh_reuse_hit_by_and_progressive(tmp_path):
    src = tmp_path / "prog.cpp"
    src.write_text(PROGRAM)
    exe = str(tmp_path / "prog")
    env = dict(os.environ)
    env.pop("CXX", None); env.pop("CC", None)
    inc = os.path.join(ROOT, "cpp_raytracer_b200", "host", "include")
    lib = os.path.join(ROOT, "cpp_raytracer_b200")
    subprocess.run(["g++", "-std=c++20", "-O1", f"-I{inc}", "-o", exe, str(src), f"-L{lib}", "-lb200rt", f"-Wl,-rpath,{lib}"],
                   check=True, env=env)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["direct_eq_resident"] and res["resident_repeatable"] and res["copy_eq"]
    assert res["passes"] == 5 and res["last_done"] == 24          # 5+5+5+5+4
    assert res["prog_maxrel"] < 1e-5                               # same samples, FP32 summation order only
    # ray from (0,5,0) straight down: enters the unit sphere centred (0,1,0) at y = 2, t = 3, from outside
    assert res["down_t"] == 3.0 and res["down_ny"] == 1.0 and res["down_outside"] and res["down_is_glass"]
    # from the centre along +z with |d| = 2: leaves at t = 0.5; the reported normal is flipped inwards
    assert res["inside_t"] == 0.5 and res["inside_nz"] == -1.0 and not res["inside_outside"]
    # both sphere roots (t = 5 -/+ sqrt(0.75)) lie below the interval (6, 100): the mirror behind it at z = -3 is hit
    assert res["quad_is_metal"] and res["quad_t"] == 8.0 and res["quad_nz"] == 1.0
    assert res["miss"] and res["batch_agree"] == 1000 and 0 < res["batch_hits"] <= 1000 and res["lone_same"]
