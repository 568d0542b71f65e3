// scenes.hpp -- the reference's scene set (reference src/main.cpp:13-650), restated against the
// API-compatible host headers.  Only public API is used (Scene::add, Sphere, Parallelogram, Box,
// the four materials, RGB, rand_double / rand_int, the Camera setters), so every function here
// would also compile against the reference's own headers.
//
// The reference repeats the "RTOW grid" loop four times with different extents and material
// thresholds; here it is one parameterised helper.  Where the reference leaves the order of
// random draws to the compiler (function-argument evaluation order), the order g++ produces is
// written out explicitly, so the scenes are identical under any compiler; tests compare the
// flattened result byte for byte with dumps of reference-built scenes.
#pragma once
#include <string>

#include "base/camera.h"
#include "base/material.h"
#include "base/scene.h"
#include "shapes/shapes.h"
#include "util/rand_util.h"

namespace b200rt_scenes {

template <typename T, typename... Args>
std::shared_ptr<T> ms(Args &&...args) { return std::make_shared<T>(std::forward<Args>(args)...); }

struct Built {
    Scene world;
    Camera camera;
};

struct GridParams {
    int a_lo, a_hi, b_lo, b_hi;     // small spheres at (a + 0.9u, 0.2, b + 0.9u) for a in [a_lo,a_hi), b in [b_lo,b_hi)
    double ground_radius;
    bool lights;                    // 3.5 % emissive spheres, metal band 0.8-0.9 instead of 0.8-0.95
    double light_lo, light_hi;      // intensity range of the small lights
};

// main.cpp:16-54 / 81-130 / 157-197 / 221-262
inline void rtow_grid(Scene &world, const GridParams &g) {
    world.add(ms<Sphere>(Point3D{0, -g.ground_radius, 0}, g.ground_radius, ms<Lambertian>(RGB::from_mag(0.5, 0.5, 0.5))));
    for (int a = g.a_lo; a < g.a_hi; a++) {
        for (int b = g.b_lo; b < g.b_hi; b++) {
            const double choose_mat = rand_double();
            const double cx = a + 0.9 * rand_double();
            const double cz = b + 0.9 * rand_double();
            const Point3D center{cx, 0.2, cz};
            if (!((center - Point3D{4, 0.2, 0}).mag() > 0.9)) continue;
            std::shared_ptr<Material> m;
            const double metal_hi = g.lights ? 0.9 : 0.95;
            if (g.lights && choose_mat < 0.035) {
                const RGB albedo = RGB::random();
                m = ms<DiffuseLight>(albedo, rand_double(g.light_lo, g.light_hi));
            } else if (choose_mat < 0.8) {
                const RGB first = RGB::random(), second = RGB::random();
                m = ms<Lambertian>(first * second);
            } else if (choose_mat < metal_hi) {
                const RGB albedo = RGB::random(0.5, 1);
                const double fuzz = rand_double(0, 0.5);
                m = ms<Metal>(albedo, fuzz);
            } else {
                m = ms<Dielectric>(1.5);
            }
            world.add(ms<Sphere>(center, 0.2, m));
        }
    }
    world.add(ms<Sphere>(Point3D{0, 1, 0}, 1.0, ms<Dielectric>(1.5)));
    world.add(ms<Sphere>(Point3D{-4, 1, 0}, 1.0, ms<Lambertian>(RGB::from_mag(0.4, 0.2, 0.1))));
    world.add(ms<Sphere>(Point3D{4, 1, 0}, 1.0, ms<Metal>(RGB::from_mag(0.7, 0.6, 0.5), 0.0)));
}

inline Built rtow_final_image() {                                   // main.cpp:13-75 (unseeded in the reference)
    Built s;
    rtow_grid(s.world, {-11, 11, -11, 11, 1000, false, 0, 0});
    s.camera.set_image_by_width_and_aspect_ratio(1200, 16. / 9.).set_vertical_fov(20).set_camera_center(Point3D{13, 2, 3})
        .set_camera_lookat(Point3D{0, 0, 0}).set_camera_up_direction(Vec3D{0, 1, 0}).set_defocus_angle(0.6)
        .set_focus_distance(10).set_samples_per_pixel(500).set_max_depth(20).set_background(RGB::from_mag(0.7, 0.8, 1));
    return s;
}

inline Built rtow_final_lights_with_tone_mapping() {                // main.cpp:77-152
    SeedSeqGenerator::get_instance().set_seed(2286021279);
    Built s;
    rtow_grid(s.world, {-11, 11, -11, 11, 1000000, true, 30, 100});
    s.world.add(ms<Sphere>(Point3D{0, 2.5, 2.5}, 0.2, ms<DiffuseLight>(RGB::from_mag(0.380205, 0.680817, 0.385431), 150)));
    s.camera.set_image_by_width_and_aspect_ratio(1080, 16. / 9.).set_vertical_fov(25).set_camera_center(Point3D{13, 2, 3})
        .set_camera_lookat(Point3D{0, 0, 0}).set_camera_up_direction(Vec3D{0, 1, 0}).set_defocus_angle(0.48)
        .set_focus_distance(10).set_samples_per_pixel(2000).set_max_depth(20).set_background(RGB::zero());
    return s;
}

inline Built millions_of_spheres() {                                // main.cpp:154-216 (unseeded in the reference)
    Built s;
    rtow_grid(s.world, {-1001, 1001, -1001, 51, 1000000, false, 0, 0});
    s.camera.set_image_by_width_and_aspect_ratio(2160, 16. / 9.).set_vertical_fov(40).set_camera_center(Point3D{0, 10, 50})
        .set_camera_lookat(Point3D{0, 0, 0}).set_camera_up_direction(Vec3D{0, 1, 0}).set_defocus_angle(0.1)
        .set_focus_distance(51).set_samples_per_pixel(500).set_max_depth(50);
    return s;
}

inline Built millions_of_spheres_with_lights() {                    // main.cpp:218-290
    SeedSeqGenerator::get_instance().set_seed(473654968);
    Built s;
    rtow_grid(s.world, {-1001, 1001, -1501, 51, 1000000, true, 5, 15});
    s.world.add(ms<Sphere>(Point3D{0, 12, 0}, 3, ms<DiffuseLight>(RGB::from_mag(0.380205, 0.680817, 0.385431), 150)));
    s.camera.set_image_by_width_and_aspect_ratio(1080, 16. / 9.).set_vertical_fov(40).set_camera_center(Point3D{0, 12.5, 50})
        .set_camera_lookat(Point3D{0, 0, 0}).set_camera_up_direction(Vec3D{0, 1, 0}).set_defocus_angle(0.1)
        .set_focus_distance(51).set_samples_per_pixel(1000).set_max_depth(20).set_background(RGB::zero());
    return s;
}

inline Built parallelogram_test() {                                 // main.cpp:294-322
    Built s;
    const struct { Point3D v; Vec3D s1, s2; RGB c; } quads[] = {
        {{-3, -2, 5}, {0, 0, -4}, {0, 4, 0}, RGB::from_mag(1.0, 0.2, 0.2)},
        {{-2, -2, 0}, {4, 0, 0}, {0, 4, 0}, RGB::from_mag(0.2, 1.0, 0.2)},
        {{3, -2, 1}, {0, 0, 4}, {0, 4, 0}, RGB::from_mag(0.2, 0.2, 1.0)},
        {{-2, 3, 1}, {4, 0, 0}, {0, 0, 4}, RGB::from_mag(1.0, 0.5, 0.0)},
        {{-2, -3, 5}, {4, 0, 0}, {0, 0, -4}, RGB::from_mag(0.2, 0.8, 0.8)},
    };
    // the reference creates all five materials first, then the quads
    std::shared_ptr<Material> mats[5];
    for (int i = 0; i < 5; ++i) mats[i] = ms<Lambertian>(quads[i].c);
    for (int i = 0; i < 5; ++i) s.world.add(ms<Parallelogram>(quads[i].v, quads[i].s1, quads[i].s2, mats[i]));
    s.camera.set_image_by_width_and_aspect_ratio(1000, 1.).set_samples_per_pixel(100).set_max_depth(50).set_vertical_fov(80)
        .set_camera_center(Point3D{0, 0, 9}).set_camera_direction_towards(Point3D{0, 0, 0}).set_camera_up_direction(Point3D{0, 1, 0})
        .turn_blur_off().set_background(RGB::from_mag(0.7, 0.8, 1));
    return s;
}

inline Built cornell_box_test(bool empty) {                         // main.cpp:326-360
    Built s;
    auto red = ms<Lambertian>(RGB::from_mag(.65, .05, .05));
    auto white = ms<Lambertian>(RGB::from_mag(.73, .73, .73));
    auto green = ms<Lambertian>(RGB::from_mag(.12, .45, .15));
    auto light = ms<DiffuseLight>(RGB::from_mag(1, 1, 1), 15);
    s.world.add(ms<Parallelogram>(Point3D{555, 0, 0}, Vec3D{0, 555, 0}, Vec3D{0, 0, 555}, green));
    s.world.add(ms<Parallelogram>(Point3D{0, 0, 0}, Vec3D{0, 555, 0}, Vec3D{0, 0, 555}, red));
    s.world.add(ms<Parallelogram>(Point3D{343, 554, 332}, Vec3D{-130, 0, 0}, Vec3D{0, 0, -105}, light));
    s.world.add(ms<Parallelogram>(Point3D{0, 0, 0}, Vec3D{555, 0, 0}, Vec3D{0, 0, 555}, white));
    s.world.add(ms<Parallelogram>(Point3D{555, 555, 555}, Vec3D{-555, 0, 0}, Vec3D{0, 0, -555}, white));
    s.world.add(ms<Parallelogram>(Point3D{0, 0, 555}, Vec3D{555, 0, 0}, Vec3D{0, 555, 0}, white));
    if (!empty) {
        s.world.add(ms<Box>(Point3D{130, 0, 65}, Point3D{295, 165, 230}, white));
        s.world.add(ms<Box>(Point3D{265, 0, 295}, Point3D{430, 330, 460}, white));
    }
    s.camera.set_image_by_width_and_aspect_ratio(1000, 1.).set_samples_per_pixel(10).set_max_depth(1000).set_vertical_fov(40)
        .set_camera_center(Point3D{278, 278, -800}).set_camera_direction_towards(Point3D{278, 278, 0})
        .set_camera_up_direction(Point3D{0, 1, 0}).turn_blur_off().set_background(RGB::from_mag(0));
    return s;
}

inline Built raining_on_the_dance_floor() {                         // main.cpp:365-412
    SeedSeqGenerator::get_instance().set_seed(5987634);
    Built s;
    for (int x = -1000; x <= 1000; ++x)
        for (int z = -1000; z <= 100; ++z) {
            // make_shared<DiffuseLight>(RGB::random(), rand_double(0.5, 2)): g++ draws the intensity first
            const double intensity = rand_double(0.5, 2);
            const RGB colour = RGB::random();
            s.world.add(ms<Parallelogram>(Point3D{x + 0.1, 0, z + 0.1}, Point3D{0.8, 0, 0}, Point3D{0, 0, 0.8},
                                          ms<DiffuseLight>(colour, intensity)));
        }
    for (size_t i = 0; i < 25000; ++i) {
        const double choose_material = rand_double();
        std::shared_ptr<Material> material = ms<Dielectric>(rand_double(1.25, 2.5));
        if (choose_material < 0.05) material = ms<Metal>(RGB::random(), 0);
        // make_shared<Sphere>(Point3D{...}, rand_double(0.25, 0.8), material): radius first, then x, y, z
        const double radius = rand_double(0.25, 0.8);
        const double px = rand_double(-1000, 1000), py = rand_double(2, 40), pz = rand_double(-1000, 50);
        s.world.add(ms<Sphere>(Point3D{px, py, pz}, radius, material));
    }
    for (size_t i = 0; i < 50; ++i) {
        const double radius = rand_double(0.25, 0.5);
        const double px = rand_double(-20, 20), py = rand_double(1, 8), pz = rand_double(-50, 50);
        s.world.add(ms<Sphere>(Point3D{px, py, pz}, radius, ms<Dielectric>(1.5)));
    }
    s.camera.set_image_by_width_and_aspect_ratio(2160, 16. / 9.).set_samples_per_pixel(50).set_max_depth(50).set_vertical_fov(40)
        .set_camera_center(Point3D{0, 10, 50}).set_camera_direction_towards(Point3D{0, 0, 0}).set_camera_up_direction(Point3D{0, 1, 0})
        .turn_blur_off().set_background(RGB::from_mag(0));
    return s;
}

inline Built christmas_tree_made_of_spheres() {                     // main.cpp:414-583
    SeedSeqGenerator::get_instance().set_seed(20231225);
    Built s;
    Scene &world = s.world;
    world.add(ms<Parallelogram>(Point3D{-1000000, 0, -1000000}, Vec3D{2000000, 0, 0}, Vec3D{0, 0, 2000000},
                                ms<Lambertian>(RGB::from_mag(0.25))));
    world.add(ms<Sphere>(Point3D{20, 25, -25}, 2.5, ms<DiffuseLight>(RGB::from_mag(0.8), 500)));
    const int apex_y = 20;
    const double radius_to_height = 1. / 3.;
    const std::array colors{RGB::from_rgb(156, 10, 72), RGB::from_rgb(66, 106, 33), RGB::from_rgb(41, 119, 133),
                            RGB::from_mag(0.5), RGB::from_mag(0.5), RGB::from_mag(0.5)};
    auto too_close = [&](const Point3D &c, double r) {
        return std::any_of(world.begin(), world.end(), [&](const std::shared_ptr<Hittable> &obj) {
            auto sp = std::dynamic_pointer_cast<Sphere>(obj);
            return sp != nullptr && (c - sp->center).mag() <= r + sp->radius + 0.1;
        });
    };
    for (int i = 0; i < 200; ++i) {                                  // ornaments on the cone's surface
        while (true) {
            double y = rand_double(0, apex_y);
            if (y > 17) y = rand_double(0, apex_y);
            if (i == 0) y = apex_y;
            const double ring = (20 - y) * radius_to_height;
            const double angle = rand_double(0, 2 * std::numbers::pi);
            const Point3D c{ring * std::sin(angle), y, ring * std::cos(angle)};
            const double r = rand_double(0.25, 0.45);
            if (too_close(c, r)) continue;
            // make_shared<Metal>(colors[rand_int(...)], rand_double(0, 0.1)): g++ draws the fuzz first
            const double fuzz = rand_double(0, 0.1);
            const RGB colour = colors[rand_int(0, static_cast<int>(colors.size() - 1))];
            std::shared_ptr<Material> m = ms<Metal>(colour, fuzz);
            if (i == 0) m = ms<DiffuseLight>(RGB::from_mag(1), 10);
            world.add(ms<Sphere>(c, r, m));
            break;
        }
    }
    std::vector<std::shared_ptr<Sphere>> snow;
    auto snow_material = ms<Lambertian>(RGB::from_mag(1));
    for (int i = 0; i < 4000; ++i) {
        while (true) {
            const double sx = rand_double(-30, 30), sy = rand_double(0, 30), sz = rand_double(-50, 50);
            const Point3D c{sx, sy, sz};
            const double r = (c.z > 35 ? 0.015 : (c.z > 20 ? 0.03 : 0.05));
            if (too_close(c, r)) continue;
            snow.push_back(ms<Sphere>(c, r, snow_material));
            break;
        }
    }
    for (const auto &f : snow) world.add(f);
    s.camera.set_image_by_width_and_aspect_ratio(1080, 16. / 9.).set_background(RGB::zero()).set_camera_center(Point3D{0, 17.5, 50})
        .set_camera_direction_towards(Point3D{0, 10, 0}).set_camera_up_direction(Point3D{0, 1, 0}).set_vertical_fov(35)
        .set_samples_per_pixel(10000).set_max_depth(50);
    return s;
}

inline Built bvh_pathological_test() {                              // main.cpp:585-650 (spheres only; the reference just builds a BVH)
    Built s;
    for (int i = 0; i < 135; ++i)
        s.world.add(ms<Sphere>(Point3D{std::pow(10.7, i), 0, 0}, std::pow(17.3, i), ms<Lambertian>(RGB::zero())));
    s.camera.set_image_dimensions(64, 64).set_vertical_fov(60).set_camera_center(Point3D{-50, 3, 40})
        .set_camera_direction_towards(Point3D{0, 0, 0});
    return s;
}

// Names match oracle/ref_bridge's scene names.  One scene per process when a scene sets a seed
// (a thread's LCG stream starts at its first draw, rand_util.h:106).
inline bool build(const std::string &name, Built &out) {
    if (name == "rtow_final") { SeedSeqGenerator::get_instance().set_seed(1); out = rtow_final_image(); }
    else if (name == "rtow_lights") out = rtow_final_lights_with_tone_mapping();
    else if (name == "millions") { SeedSeqGenerator::get_instance().set_seed(1); out = millions_of_spheres(); }
    else if (name == "millions_lights") out = millions_of_spheres_with_lights();
    else if (name == "quads") out = parallelogram_test();
    else if (name == "cornell_empty") out = cornell_box_test(true);
    else if (name == "cornell") out = cornell_box_test(false);
    else if (name == "raining") out = raining_on_the_dance_floor();
    else if (name == "xmas") out = christmas_tree_made_of_spheres();
    else if (name == "pathological") out = bvh_pathological_test();
    else return false;
    return true;
}

}  // namespace b200rt_scenes
