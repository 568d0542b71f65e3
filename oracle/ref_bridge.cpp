// oracle/ref_bridge.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference (DeltaPavonis/cpp_raytracer, header-only C++20) as the
// parity oracle and as the same-host CPU baseline.  The reference sources are compiled where
// they lie under $(REF)/include and $(REF)/src (see oracle/Makefile); nothing is copied into
// this repository.  The output binary goes to oracle/_ref/ (git-ignored, but it travels to
// the GPU box with the gpurun snapshot).
//
// How the reference is driven:
//  * All reference headers are included first, in the order src/main.cpp:1-5 uses.
//  * src/main.cpp itself is then included with `Camera` renamed to `CapturingCamera`, a
//    recorder defined below that forwards every fluent setter to a real reference `Camera`
//    and, at `.render(world)`, stashes the Scene + Camera instead of rendering.  This way the
//    nine scene functions (src/main.cpp:13-650) run exactly as the reference wrote them
//    (same RNG draws, same argument-evaluation order, same object order).
//  * -fno-access-control lets the bridge read private members (Parallelogram::vertex,
//    Lambertian::intrinsic_color, BVH::linear_bvh_nodes, Camera::pixel00_loc ...): the
//    reference has no getters.  It does not change code generation of the render path.
//
// Commands (chained on one command line, because the reference's thread_local LCG cannot be
// re-seeded, so a process builds exactly ONE scene -- rand_util.h:106):
//   ref_bridge <scene> [--w W --h H --spp S --depth D --threads T --render-seed R] <cmd> [args] ...
//     dump <out.scene>                         flat scene + camera (format: see write_scene)
//     raycast <rays.bin> <out.bin> [brute]     closest hit (prim id, t) through BVH::hit_by
//                                              semantics (bvh.h:585-715) or Scene::hit_by
//     primhit <rays_ids.bin> <out.bin>         t of ONE given primitive per ray (tie forensics)
//     record <n> <stride> <out_rays.bin>       rays the reference's own paths issue (all bounces)
//     render <out.hdr> [seed]                  Camera::render<BVH> unmodified (camera.h:264-297)
//     stats                                    rays/path, node visits/ray, prim tests/ray
//     kat <out.json>                           unit known-answers (reflect/refract/tonemap/LCG)
//     lcg_state                                main thread's LCG state (consumes one draw)
// Every command prints one JSON line on stdout; the reference's own chatter is swallowed.

#include "util/rand_util.h"
#include "base/scene.h"
#include "base/material.h"
#include "base/camera.h"
#include "shapes/shapes.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>
#include <unordered_map>
#include <omp.h>

// ------------------------------------------------------------------------------------------
// Camera recorder: stands in for `Camera` inside src/main.cpp only.
// ------------------------------------------------------------------------------------------
struct CapturedRender {
    bool have = false;
    Scene world;
    Camera cam;
};
static CapturedRender g_capture;

struct DummyImage {
    void send_as_ppm(const std::string &) const {}
};

class CapturingCamera {
    Camera cam;
public:
#define FWD(name) \
    template <typename... A> CapturingCamera &name(A &&...a) { cam.name(std::forward<A>(a)...); return *this; }
    FWD(set_camera_center) FWD(set_camera_direction) FWD(set_camera_direction_towards)
    FWD(set_camera_lookat) FWD(set_focus_distance) FWD(set_defocus_angle) FWD(turn_blur_off)
    FWD(set_camera_up_direction) FWD(set_image_width) FWD(set_image_height)
    FWD(set_image_dimensions) FWD(set_image_by_width_and_aspect_ratio)
    FWD(set_image_by_height_and_aspect_ratio) FWD(set_samples_per_pixel) FWD(set_max_depth)
    FWD(set_vertical_fov) FWD(set_horizontal_fov) FWD(set_background)
#undef FWD
    DummyImage render(const Scene &world) {
        g_capture.have = true;
        g_capture.world = world;  // copies the shared_ptr vector; objects stay alive
        g_capture.cam = cam;
        return {};
    }
};

// The reference's BVH constructor is also exercised directly by bvh_pathological_test();
// we let that one run as is (it only builds a BVH and prints).
#define Camera CapturingCamera
#define main reference_main_unused
#include "../src/main.cpp"   // resolved through -I$(REF)/include -> $(REF)/include/../src/main.cpp
#undef main
#undef Camera

// ------------------------------------------------------------------------------------------
// Flat scene structs (mirrored by include/b200rt.h and cpp_raytracer_b200/scene_io.py).
// ------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct FlatMaterial { uint32_t kind, pad; double rgb[3]; double param; };            // 40 B
struct FlatSphere   { double c[3]; double r; uint32_t mat, prim; };                  // 40 B
struct FlatQuad     { double v[3], s1[3], s2[3]; uint32_t mat, prim; };              // 80 B
struct FlatCamera {
    uint64_t image_w, image_h, spp, max_depth;
    double center[3], dir[3], up[3];
    double focus_dist, defocus_angle, vfov, hfov;   // radians; -1 when not given
    double background[3];
    double pixel00[3], delta_x[3], delta_y[3], disk_x[3], disk_y[3];   // derived by Camera::init
};
#pragma pack(pop)
enum { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_LIGHT = 3 };

static void put3(double *d, const Vec3D &v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
static void put3(double *d, const RGB &v) { d[0] = v.r; d[1] = v.g; d[2] = v.b; }

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };
static NullBuf g_nullbuf;
static std::streambuf *g_cout_buf = nullptr;
static void hush() { if (!g_cout_buf) g_cout_buf = std::cout.rdbuf(&g_nullbuf); }

struct Bridge {
    Scene world;
    Camera cam;
    std::vector<std::shared_ptr<Hittable>> prims;          // canonical order (scene.h:85-105)
    std::unordered_map<const Hittable *, uint32_t> prim_index;
    std::unique_ptr<BVH> bvh;
    double bvh_build_ms = 0;

    void finish_setup() {
        prims = world.get_primitive_components();
        for (uint32_t i = 0; i < prims.size(); ++i) prim_index[prims[i].get()] = i;
        cam.init();
    }
    void need_bvh() {
        if (bvh) return;
        auto t0 = std::chrono::steady_clock::now();
        bvh = std::make_unique<BVH>(world);
        bvh_build_ms = std::chrono::duration<double, std::milli>(
                           std::chrono::steady_clock::now() - t0).count();
    }

    // ---- BVH::hit_by semantics, restated over the reference's own node array so that the
    // primitive INDEX is observable (hit_info carries no primitive pointer).  Every AABB test
    // and primitive test is the reference's own function; verify_against_unmodified() checks
    // the loop against the untouched BVH::hit_by.
    struct WalkStats { uint64_t nodes = 0, prim_tests = 0; };
    int64_t walk(const Ray3D &ray, Interval ray_times, double &t_out, WalkStats *st) const {
        const auto &nodes = bvh->linear_bvh_nodes;
        const auto &bp = bvh->primitives;
        size_t stack[128];
        size_t sp = 0, cur = 0;
        Vec3D inv{1 / ray.dir.x, 1 / ray.dir.y, 1 / ray.dir.z};
        std::array<bool, 3> neg{ray.dir.x < 0, ray.dir.y < 0, ray.dir.z < 0};
        const Hittable *best = nullptr;
        while (true) {
            const auto &n = nodes[cur];
            if (st) st->nodes++;
            bool descend = false;
            if (n.aabb.is_hit_by_optimized(ray, ray_times, inv, neg)) {
                if (n.is_leaf_node()) {
                    for (size_t i = n.first_primitive_index;
                         i < n.first_primitive_index + n.num_primitives; ++i) {
                        if (st) st->prim_tests++;
                        if (auto h = bp[i]->hit_by(ray, ray_times); h) {
                            best = bp[i].get();
                            ray_times.max = h->hit_time;
                        }
                    }
                } else {
                    if (neg[n.split_axis]) { stack[sp++] = cur + 1; cur = n.second_child_index; }
                    else { stack[sp++] = n.second_child_index; cur = cur + 1; }
                    descend = true;
                }
            }
            if (!descend) {
                if (sp == 0) break;
                cur = stack[--sp];
            }
        }
        if (!best) return -1;
        t_out = ray_times.max;
        return prim_index.at(best);
    }
    int64_t brute(const Ray3D &ray, Interval ray_times, double &t_out) const {
        // Scene::hit_by semantics (scene.h:59-75) over the canonical primitive list.
        int64_t best = -1;
        double tmax = ray_times.max;
        for (size_t i = 0; i < prims.size(); ++i) {
            if (auto h = prims[i]->hit_by(ray, Interval(ray_times.min, tmax)); h) {
                best = (int64_t)i;
                tmax = h->hit_time;
            }
        }
        t_out = tmax;
        return best;
    }
};

// A Hittable wrapper the reference's render<T> accepts; it forwards to the BVH and lets the
// bridge observe every ray the reference's own path loop issues.
struct Observer : Hittable {
    const Bridge *b;
    std::function<void(const Ray3D &)> on_ray;
    mutable std::atomic<uint64_t> rays{0}, nodes{0}, prim_tests{0};
    bool count_walk = false;
    std::optional<hit_info> hit_by(const Ray3D &ray, const Interval &rt) const override {
        rays.fetch_add(1, std::memory_order_relaxed);
        if (on_ray) on_ray(ray);
        if (count_walk) {
            Bridge::WalkStats st; double t;
            b->walk(ray, rt, t, &st);
            nodes.fetch_add(st.nodes, std::memory_order_relaxed);
            prim_tests.fetch_add(st.prim_tests, std::memory_order_relaxed);
        }
        return b->bvh->hit_by(ray, rt);
    }
    AABB get_aabb() const override { return b->bvh->get_aabb(); }
    void print_to(std::ostream &) const override {}
};

// ------------------------------------------------------------------------------------------
static bool build_named_scene(const std::string &name, Bridge &B) {
    auto &seeds = SeedSeqGenerator::get_instance();
    if (name == "rtow_final")           { seeds.set_seed(1); rtow_final_image(); }   // unseeded in ref
    else if (name == "rtow_lights")     rtow_final_lights_with_tone_mapping();
    else if (name == "millions")        { seeds.set_seed(1); millions_of_spheres(); }
    else if (name == "millions_lights") millions_of_spheres_with_lights();
    else if (name == "quads")           parallelogram_test();
    else if (name == "cornell_empty")   cornell_box_test(true);
    else if (name == "cornell")         cornell_box_test(false);
    else if (name == "raining")         raining_on_the_dance_floor();
    else if (name == "xmas")            christmas_tree_made_of_spheres();
    else if (name == "pathological") {
        // bvh_pathological_test() (main.cpp:585-650) only builds a BVH; same 135 spheres here,
        // given a camera so it can also be ray-cast.
        Scene w;
        for (int i = 0; i < 135; ++i)
            w.add(std::make_shared<Sphere>(Point3D{std::pow(10.7, i), 0, 0}, std::pow(17.3, i),
                                           std::make_shared<Lambertian>(RGB::zero())));
        g_capture.world = w;
        g_capture.cam = Camera();
        g_capture.cam.set_image_dimensions(64, 64).set_vertical_fov(60)
            .set_camera_center(Point3D{-50, 3, 40}).set_camera_direction_towards(Point3D{0, 0, 0});
        g_capture.have = true;
    } else return false;
    if (!g_capture.have) return false;
    B.world = g_capture.world;
    B.cam = g_capture.cam;
    return true;
}

static int material_of(const Material *m, FlatMaterial &out) {
    std::memset(&out, 0, sizeof out);
    if (auto p = dynamic_cast<const Lambertian *>(m)) { out.kind = MAT_LAMBERTIAN; put3(out.rgb, p->intrinsic_color); return 0; }
    if (auto p = dynamic_cast<const Metal *>(m)) { out.kind = MAT_METAL; put3(out.rgb, p->intrinsic_color); out.param = p->fuzz_factor; return 0; }
    if (auto p = dynamic_cast<const Dielectric *>(m)) { out.kind = MAT_DIELECTRIC; out.rgb[0] = out.rgb[1] = out.rgb[2] = 1; out.param = p->refr_index; return 0; }
    if (auto p = dynamic_cast<const DiffuseLight *>(m)) { out.kind = MAT_LIGHT; put3(out.rgb, p->intrinsic_color); out.param = p->intensity; return 0; }
    return -1;
}

static FlatCamera flat_camera(const Camera &c) {
    FlatCamera f{};
    f.image_w = c.image_w; f.image_h = c.image_h; f.spp = c.samples_per_pixel; f.max_depth = c.max_depth;
    put3(f.center, c.camera.origin); put3(f.dir, c.camera.dir); put3(f.up, c.view_up_dir);
    f.focus_dist = c.focus_dist.value_or(-1); f.defocus_angle = c.defocus_angle;
    f.vfov = c.vertical_fov.value_or(-1); f.hfov = c.horizontal_fov.value_or(-1);
    put3(f.background, c.background);
    put3(f.pixel00, c.pixel00_loc); put3(f.delta_x, c.pixel_delta_x); put3(f.delta_y, c.pixel_delta_y);
    put3(f.disk_x, c.defocus_disk_x); put3(f.disk_y, c.defocus_disk_y);
    return f;
}

static bool write_scene(const Bridge &B, const std::string &path) {
    std::vector<FlatMaterial> mats;
    std::unordered_map<const Material *, uint32_t> mat_id;
    std::vector<FlatSphere> sph;
    std::vector<FlatQuad> quads;
    auto mat_index = [&](const std::shared_ptr<Material> &m) -> uint32_t {
        auto it = mat_id.find(m.get());
        if (it != mat_id.end()) return it->second;
        FlatMaterial fm;
        if (material_of(m.get(), fm)) { std::fprintf(stderr, "unknown material\n"); std::exit(2); }
        mats.push_back(fm);
        return mat_id[m.get()] = (uint32_t)mats.size() - 1;
    };
    for (uint32_t i = 0; i < B.prims.size(); ++i) {
        const Hittable *h = B.prims[i].get();
        if (auto s = dynamic_cast<const Sphere *>(h)) {
            FlatSphere f{}; put3(f.c, s->center); f.r = s->radius; f.mat = mat_index(s->material); f.prim = i;
            sph.push_back(f);
        } else if (auto q = dynamic_cast<const Parallelogram *>(h)) {
            FlatQuad f{}; put3(f.v, q->vertex); put3(f.s1, q->side1); put3(f.s2, q->side2);
            f.mat = mat_index(q->material); f.prim = i;
            quads.push_back(f);
        } else { std::fprintf(stderr, "unknown primitive\n"); return false; }
    }
    std::ofstream out(path, std::ios::binary);
    if (!out) return false;
    uint64_t hdr[3] = {mats.size(), sph.size(), quads.size()};
    FlatCamera fc = flat_camera(B.cam);
    out.write("B2RTSCN1", 8);
    out.write((const char *)hdr, sizeof hdr);
    out.write((const char *)&fc, sizeof fc);
    out.write((const char *)mats.data(), mats.size() * sizeof(FlatMaterial));
    out.write((const char *)sph.data(), sph.size() * sizeof(FlatSphere));
    out.write((const char *)quads.data(), quads.size() * sizeof(FlatQuad));
    std::printf("{\"cmd\":\"dump\",\"materials\":%zu,\"spheres\":%zu,\"quads\":%zu,\"prims\":%zu,\"w\":%zu,\"h\":%zu}\n",
                mats.size(), sph.size(), quads.size(), B.prims.size(), (size_t)fc.image_w, (size_t)fc.image_h);
    return true;
}

struct RayFile { uint64_t n; double tmin, tmax; std::vector<double> rays; };
static bool read_rays(const std::string &path, RayFile &rf, size_t per_ray) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return false;
    in.read((char *)&rf.n, 8); in.read((char *)&rf.tmin, 8); in.read((char *)&rf.tmax, 8);
    rf.rays.resize(rf.n * per_ray);
    in.read((char *)rf.rays.data(), rf.rays.size() * 8);
    return (bool)in;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: ref_bridge <scene> [opts] <cmd>...\n"); return 1; }
    hush();
    Bridge B;
    std::string scene = argv[1];
    if (!build_named_scene(scene, B)) { std::fprintf(stderr, "unknown scene %s\n", argv[1]); return 1; }
    int threads = omp_get_max_threads();
    int i = 2;
    auto next = [&]() -> std::string { if (i >= argc) { std::fprintf(stderr, "missing arg\n"); std::exit(1); } return argv[i++]; };
    // options
    while (i < argc && std::strncmp(argv[i], "--", 2) == 0) {
        std::string o = next();
        if (o == "--w") B.cam.image_w = std::stoull(next());
        else if (o == "--h") B.cam.image_h = std::stoull(next());
        else if (o == "--spp") B.cam.samples_per_pixel = std::stoull(next());
        else if (o == "--depth") B.cam.max_depth = std::stoull(next());
        else if (o == "--threads") threads = std::stoi(next());
        else if (o == "--render-seed") {
            // Re-seeds the per-thread LCGs of threads that have not drawn yet (rand_util.h:106: a
            // thread's stream is fixed at its first draw) and advances the main thread's stream, so
            // that two renders of the same scene are statistically independent.
            uint32_t rs = (uint32_t)std::stoul(next());
            SeedSeqGenerator::get_instance().set_seed(rs);
            for (uint32_t k = 0; k < 1000 + rs % 1000; ++k) rand_double();
        }
        else { std::fprintf(stderr, "unknown option %s\n", o.c_str()); return 1; }
    }
    omp_set_num_threads(threads);
    B.finish_setup();

    while (i < argc) {
        std::string cmd = next();
        if (cmd == "dump") {
            if (!write_scene(B, next())) return 3;
        } else if (cmd == "raycast") {
            std::string in = next(), out = next();
            bool use_brute = (i < argc && std::string(argv[i]) == "brute") ? (++i, true) : false;
            RayFile rf;
            if (!read_rays(in, rf, 6)) { std::fprintf(stderr, "cannot read %s\n", in.c_str()); return 3; }
            if (!use_brute) B.need_bvh();
            std::vector<int32_t> ids(rf.n);
            std::vector<double> ts(rf.n);
            uint64_t mismatch_unmodified = 0;
            double t0 = now_s();
            #pragma omp parallel for schedule(dynamic, 256) reduction(+ : mismatch_unmodified)
            for (uint64_t k = 0; k < rf.n; ++k) {
                const double *r = &rf.rays[6 * k];
                Ray3D ray{Point3D{r[0], r[1], r[2]}, Vec3D{r[3], r[4], r[5]}};
                Interval rt(rf.tmin, rf.tmax);
                double t = 0;
                int64_t id = use_brute ? B.brute(ray, rt, t) : B.walk(ray, rt, t, nullptr);
                if (!use_brute) {   // the loop above must agree with the untouched BVH::hit_by
                    auto h = B.bvh->hit_by(ray, rt);
                    if ((bool)h != (id >= 0) || (h && h->hit_time != t)) mismatch_unmodified++;
                }
                ids[k] = (int32_t)id;
                ts[k] = id >= 0 ? t : 0.0;
            }
            double dt = now_s() - t0;
            std::ofstream o(out, std::ios::binary);
            o.write((const char *)ids.data(), rf.n * 4);
            o.write((const char *)ts.data(), rf.n * 8);
            std::printf("{\"cmd\":\"raycast\",\"mode\":\"%s\",\"n\":%llu,\"seconds\":%.6f,\"mismatch_vs_unmodified_hit_by\":%llu}\n",
                        use_brute ? "brute" : "bvh", (unsigned long long)rf.n, dt, (unsigned long long)mismatch_unmodified);
            if (mismatch_unmodified) return 4;
        } else if (cmd == "primhit") {
            // input: header + n x 7 doubles (ray, prim id as double); output n doubles (t or -1)
            std::string in = next(), out = next();
            RayFile rf;
            if (!read_rays(in, rf, 7)) return 3;
            std::vector<double> ts(rf.n);
            for (uint64_t k = 0; k < rf.n; ++k) {
                const double *r = &rf.rays[7 * k];
                Ray3D ray{Point3D{r[0], r[1], r[2]}, Vec3D{r[3], r[4], r[5]}};
                auto h = B.prims[(size_t)r[6]]->hit_by(ray, Interval(rf.tmin, rf.tmax));
                ts[k] = h ? h->hit_time : -1.0;
            }
            std::ofstream o(out, std::ios::binary);
            o.write((const char *)ts.data(), rf.n * 8);
            std::printf("{\"cmd\":\"primhit\",\"n\":%llu}\n", (unsigned long long)rf.n);
        } else if (cmd == "record") {
            uint64_t want = std::stoull(next()), stride = std::stoull(next());
            std::string out = next();
            B.need_bvh();
            Observer obs; obs.b = &B;
            std::vector<double> rec; rec.reserve(want * 6);
            uint64_t seen = 0;
            obs.on_ray = [&](const Ray3D &r) {
                if (seen++ % stride == 0 && rec.size() < want * 6) {
                    rec.insert(rec.end(), {r.origin.x, r.origin.y, r.origin.z, r.dir.x, r.dir.y, r.dir.z});
                }
            };
            omp_set_num_threads(1);   // deterministic + the lambda is not thread safe
            B.cam.render(obs);
            omp_set_num_threads(threads);
            uint64_t n = rec.size() / 6;
            double tmin = 0.00001, tmax = std::numeric_limits<double>::infinity();   // camera.h:217
            std::ofstream o(out, std::ios::binary);
            o.write((const char *)&n, 8); o.write((const char *)&tmin, 8); o.write((const char *)&tmax, 8);
            o.write((const char *)rec.data(), rec.size() * 8);
            std::printf("{\"cmd\":\"record\",\"n\":%llu,\"rays_seen\":%llu}\n", (unsigned long long)n, (unsigned long long)seen);
        } else if (cmd == "render") {
            std::string out = next();
            B.need_bvh();
            double t0 = now_s();
            auto img = B.cam.render(*B.bvh);     // camera.h:264-297, untouched
            double dt = now_s() - t0;
            size_t w = img.width(), h = img.height();
            if (out != "-") {
                std::vector<float> px(w * h * 3);
                for (size_t r = 0; r < h; ++r)
                    for (size_t c = 0; c < w; ++c) {
                        px[(r * w + c) * 3 + 0] = (float)img[r][c].r;
                        px[(r * w + c) * 3 + 1] = (float)img[r][c].g;
                        px[(r * w + c) * 3 + 2] = (float)img[r][c].b;
                    }
                std::ofstream o(out, std::ios::binary);
                uint64_t hdr[2] = {w, h};
                o.write((const char *)hdr, 16);
                o.write((const char *)px.data(), px.size() * 4);
            }
            double paths = (double)w * h * B.cam.samples_per_pixel;
            std::printf("{\"cmd\":\"render\",\"scene\":\"%s\",\"w\":%zu,\"h\":%zu,\"spp\":%zu,\"max_depth\":%zu,\"threads\":%d,"
                        "\"seconds\":%.6f,\"mpaths_per_s\":%.6f,\"bvh_build_ms\":%.3f,\"prims\":%zu}\n",
                        scene.c_str(), w, h, (size_t)B.cam.samples_per_pixel, (size_t)B.cam.max_depth, threads, dt,
                        paths / dt / 1e6, B.bvh_build_ms, B.prims.size());
        } else if (cmd == "render_f64") {
            // exact double pixels, for bit-level pinning of the C restatement (1 thread)
            std::string out = next();
            B.need_bvh();
            auto img = B.cam.render(*B.bvh);
            size_t w = img.width(), h = img.height();
            std::vector<double> px(w * h * 3);
            for (size_t r = 0; r < h; ++r)
                for (size_t c = 0; c < w; ++c) {
                    px[(r * w + c) * 3 + 0] = img[r][c].r;
                    px[(r * w + c) * 3 + 1] = img[r][c].g;
                    px[(r * w + c) * 3 + 2] = img[r][c].b;
                }
            std::ofstream o(out, std::ios::binary);
            uint64_t hdr[2] = {w, h};
            o.write((const char *)hdr, 16);
            o.write((const char *)px.data(), px.size() * 8);
            std::printf("{\"cmd\":\"render_f64\",\"w\":%zu,\"h\":%zu}\n", w, h);
        } else if (cmd == "lcg_state") {
            // The calling thread's LCG state (rand_util.h:106) is a function-local thread_local and
            // cannot be read; draw once and invert rand_double(): state = v * (2^32 - 2), exact.
            const double v = rand_double();
            const uint32_t st = (uint32_t)std::llround(v * 4294967294.0);
            std::printf("{\"cmd\":\"lcg_state\",\"state_after_this_draw\":%u}\n", st);
        } else if (cmd == "ppm") {
            // the reference's own writer (image.h:38-56) on the reference's own render
            std::string out = next();
            B.need_bvh();
            B.cam.render(*B.bvh).send_as_ppm(out);
            std::printf("{\"cmd\":\"ppm\"}\n");
        } else if (cmd == "stats") {
            B.need_bvh();
            Observer obs; obs.b = &B; obs.count_walk = true;
            B.cam.render(obs);
            double paths = (double)B.cam.image_w * B.cam.image_h * B.cam.samples_per_pixel;
            double rays = (double)obs.rays.load();
            std::printf("{\"cmd\":\"stats\",\"scene\":\"%s\",\"paths\":%.0f,\"rays\":%.0f,\"rays_per_path\":%.4f,"
                        "\"nodes_per_ray\":%.4f,\"prim_tests_per_ray\":%.4f,\"bvh_nodes\":%zu,\"prims\":%zu}\n",
                        scene.c_str(), paths, rays, rays / paths, obs.nodes.load() / rays,
                        obs.prim_tests.load() / rays, (size_t)B.bvh->linear_bvh_nodes.size(), B.prims.size());
        } else if (cmd == "kat") {
            std::string out = next();
            std::ofstream o(out);
            o.precision(17);
            o << "{\n";
            // LCG stream (rand_util.h:51-117).  Only meaningful for scenes that have not drawn a random
            // number yet (quads, cornell*): seed the sequence with 12345, then the first draw of this
            // thread takes its seed from SeedSeqGenerator::next_seed().
            SeedSeqGenerator::get_instance().set_seed(12345);
            o << "\"rand_double_seed\": 12345,\n\"rand_double_next16\": [";
            for (int k = 0; k < 16; ++k) o << (k ? "," : "") << rand_double();
            o << "],\n";
            // reflected / refracted / reflectance on fixed inputs
            Vec3D d = Vec3D{0.3, -0.8, 0.52}.unit_vector(), n = Vec3D{0.1, 1.0, -0.2}.unit_vector();
            auto rfl = reflected(d, n);
            o << "\"reflect\": {\"d\": [" << d.x << "," << d.y << "," << d.z << "], \"n\": [" << n.x << "," << n.y << "," << n.z
              << "], \"out\": [" << rfl.x << "," << rfl.y << "," << rfl.z << "]},\n";
            o << "\"refract\": [";
            bool first = true;
            for (double eta : {1.0 / 1.5, 1.5, 2.5, 0.4}) {
                auto r = refracted(d, n, eta);
                o << (first ? "" : ",") << "{\"eta\": " << eta << ", \"ok\": " << (r ? "true" : "false");
                if (r) o << ", \"out\": [" << r->x << "," << r->y << "," << r->z << "]";
                o << "}";
                first = false;
            }
            o << "],\n\"reflectance\": [";
            first = true;
            for (double c : {0.0, 0.1, 0.5, 0.9, 1.0})
                for (double eta : {1.0 / 1.5, 1.5}) {
                    o << (first ? "" : ",") << "[" << c << "," << eta << "," << Dielectric::reflectance(c, eta) << "]";
                    first = false;
                }
            o << "],\n\"tonemap\": [";
            first = true;
            for (auto c : {RGB::from_mag(0, 0, 0), RGB::from_mag(0.5, 0.25, 0.125), RGB::from_mag(1, 1, 1),
                           RGB::from_mag(10, 0.2, 0.01), RGB::from_mag(500, 500, 500), RGB::from_mag(0.001, 0.9, 3.5)}) {
                o << (first ? "" : ",") << "{\"rgb\": [" << c.r << "," << c.g << "," << c.b << "], \"out\": \"" << c.as_string() << "\"}";
                first = false;
            }
            o << "]\n}\n";
            std::printf("{\"cmd\":\"kat\"}\n");
        } else {
            std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
            return 1;
        }
        std::fflush(stdout);
    }
    return 0;
}
