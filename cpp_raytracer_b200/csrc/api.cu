// api.cu -- implementation of the C ABI declared in include/b200rt.h.
//
// Host side of the drop-in boundary: validates and flattens the scene description, computes
// primitive bounds exactly as the reference's constructors do (sphere.h:112-123,
// parallelogram.h:269-295), builds the 4-wide BVH (bvh_builder.cpp), lays everything out in
// leaf order as SoA device arrays and launches the kernels.  There is no CPU compute path:
// every entry point that produces hits or pixels fails with B200RT_ENODEVICE without a GPU.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "bvh_builder.h"
#include "kernels.h"
#include "thread_pool.h"

using namespace b200rt;

namespace b200rt {
cudaError_t build_lbvh_device(const B200rtSphere *d_sph, uint32_t n_sph, const B200rtQuad *d_quads, uint32_t n_quad,
                              double2 *out_sph, uint2 *out_sph_meta, double2 *out_quads, uint2 *out_quad_meta,
                              float4 **nodes_out, uint32_t *n_nodes_out, uint32_t *depth_out);
}

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            cudaGetLastError();                                                                \
            return fail(e_ == cudaErrorMemoryAllocation ? B200RT_ENOMEM : B200RT_ECUDA,         \
                        std::string(#expr) + ": " + cudaGetErrorString(e_));                   \
        }                                                                                      \
    } while (0)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        ok = cudaSetDevice(dev) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Device memory comes from the device's default stream-ordered pool with its release threshold
// raised, so that the create / render / destroy cycle of the drop-in call (one per frame) reuses
// memory instead of paying cudaMalloc / cudaFree (and their device-wide synchronisation) each time.
cudaError_t dev_alloc(void **p, size_t bytes) {
    static bool pool_ready[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
        }
        pool_ready[dev] = true;
    }
    e = cudaMallocAsync(p, bytes ? bytes : 1, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    return e;
}
template <typename T>
cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc(reinterpret_cast<void **>(p), bytes); }
void dev_free(void *p) {
    if (p) cudaFreeAsync(p, 0);
}

struct SceneImpl {
    uint32_t magic = 0xB200577Eu;
    int device = 0;
    DeviceScene d{};                 // device pointers
    void *allocs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    B200rtSceneInfo info{};
    int stack = 32;
    unsigned long long *d_counters = nullptr;
    float *d_frame = nullptr;        // scratch frame for the host-buffer render entry
    size_t frame_floats = 0;
    double *d_rays = nullptr; int32_t *d_prim = nullptr; double *d_t = nullptr;   // raycast scratch
    size_t ray_capacity = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void *d_pool = nullptr;          // wavefront path-slot pool
    uint32_t pool_slots = 0;
    int sm_count = 148;
};

SceneImpl *as_scene(void *h) {
    SceneImpl *s = static_cast<SceneImpl *>(h);
    return (s && s->magic == 0xB200577Eu) ? s : nullptr;
}

int device_count_quiet() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- reference-exact host geometry helpers ------------------------------------------------
struct V3 { double x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mul(V3 a, double d) { return {a.x * d, a.y * d, a.z * d}; }
inline V3 divv(V3 a, double d) { return mul(a, 1 / d); }   // operator/= multiplies by 1/d (vec3d.h:31)
inline V3 crossv(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double mag2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double mag(V3 a) { return std::sqrt(mag2(a)); }
inline V3 unitv(V3 a) { return divv(a, mag(a)); }
inline V3 v3(const double *p) { return {p[0], p[1], p[2]}; }
inline void put(double *p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

Box3 sphere_box(const B200rtSphere &s) {   // sphere.h:112-123
    Box3 b;
    for (int a = 0; a < 3; ++a) {
        const double p = s.c[a] - s.r, q = s.c[a] + s.r;
        b.lo[a] = std::fmin(p, q);
        b.hi[a] = std::fmax(p, q);
    }
    return b;
}
Box3 quad_box(const B200rtQuad &q) {       // parallelogram.h:281-295 + aabb.h ensure_min_axis_length
    Box3 b;
    const V3 v = v3(q.v), s1 = v3(q.s1), s2 = v3(q.s2);
    const V3 pts[4] = {v, add(v, s1), add(v, s2), add(add(v, s1), s2)};
    for (int a = 0; a < 3; ++a) { b.lo[a] = std::numeric_limits<double>::infinity(); b.hi[a] = -b.lo[a]; }
    for (const V3 &p : pts) {
        const double c[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; ++a) { b.lo[a] = std::fmin(b.lo[a], c[a]); b.hi[a] = std::fmax(b.hi[a], c[a]); }
    }
    for (int a = 0; a < 3; ++a) {
        const double size = b.hi[a] - b.lo[a];
        if (size < 1e-4) { const double pad = (1e-4 - size) / 2; b.lo[a] -= pad; b.hi[a] += pad; }
    }
    return b;
}

int validate_desc(const B200rtSceneDesc *d) {
    if (!d) return fail(B200RT_EINVAL, "scene description is NULL");
    if ((d->n_materials && !d->materials) || (d->n_spheres && !d->spheres) || (d->n_quads && !d->quads))
        return fail(B200RT_EINVAL, "scene description has a count without an array");
    const uint64_t n_prims = d->n_spheres + d->n_quads;
    if (n_prims > 0x7FFFFFFFull) return fail(B200RT_EINVAL, "too many primitives");
    for (uint64_t i = 0; i < d->n_materials; ++i)
        if (d->materials[i].kind > B200RT_MAT_LIGHT)
            return fail(B200RT_EINVAL, "unknown material kind " + std::to_string(d->materials[i].kind) +
                                           " (closed set: Lambertian, Metal, Dielectric, DiffuseLight)");
    for (uint64_t i = 0; i < d->n_spheres; ++i)
        if (d->spheres[i].mat >= d->n_materials || d->spheres[i].prim >= n_prims)
            return fail(B200RT_EINVAL, "sphere " + std::to_string(i) + ": material or primitive index out of range");
    for (uint64_t i = 0; i < d->n_quads; ++i)
        if (d->quads[i].mat >= d->n_materials || d->quads[i].prim >= n_prims)
            return fail(B200RT_EINVAL, "quad " + std::to_string(i) + ": material or primitive index out of range");
    return B200RT_OK;
}

// Runs body(lo, hi) over [0, n) split into chunks on a temporary pool (serial when small).
template <typename F>
void parallel_ranges(uint64_t n, int threads, F body) {
    if (n < (1u << 16) || threads <= 1) { body((uint64_t)0, n); return; }
    ThreadPool pool(threads);
    const int chunks = threads * 4;
    pool.parallel_for(chunks, [&](int c) { body(n * c / chunks, n * (c + 1) / chunks); });
}
int host_threads(const B200rtBuildOpts *o) {
    return (o && o->build_threads > 0) ? o->build_threads : ThreadPool::hardware_threads();
}

void compute_boxes(const B200rtSceneDesc *d, std::vector<Box3> &boxes, int threads) {
    boxes.resize(d->n_spheres + d->n_quads);
    parallel_ranges(d->n_spheres, threads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) boxes[i] = sphere_box(d->spheres[i]);
    });
    parallel_ranges(d->n_quads, threads, [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) boxes[d->n_spheres + i] = quad_box(d->quads[i]);
    });
}

BuildParams build_params(const B200rtBuildOpts *o) {
    BuildParams p;
    if (o) {
        if (o->max_leaf_prims > 0) p.max_leaf_prims = o->max_leaf_prims;
        if (o->sah_bins > 0) p.sah_bins = o->sah_bins;
        if (o->build_threads > 0) p.threads = o->build_threads;
    }
    return p;
}

template <typename T>
int upload(const std::vector<T> &host, const T **dev, void **slot, uint64_t &bytes) {
    *dev = nullptr;
    if (host.empty()) return B200RT_OK;
    void *p = nullptr;
    CUDA_TRY(dev_alloc(&p, host.size() * sizeof(T)));
    *slot = p;
    CUDA_TRY(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T *>(p);
    bytes += host.size() * sizeof(T);
    return B200RT_OK;
}

void free_scene(SceneImpl *s) {
    if (!s) return;
    DeviceGuard g(s->device);
    cudaDeviceSynchronize();   // kernels of any stream may still read the scene
    for (void *&p : s->allocs) { dev_free(p); p = nullptr; }
    dev_free(s->d_counters);
    dev_free(s->d_frame);
    dev_free(s->d_rays);
    dev_free(s->d_prim);
    dev_free(s->d_t);
    dev_free(s->d_pool);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    s->magic = 0;
    delete s;
}

int fill_camera(const B200rtCamera *cam, CameraParams &C) {
    if (!cam) return fail(B200RT_EINVAL, "camera is NULL");
    if (cam->image_w == 0 || cam->image_h == 0 || cam->image_w > 65536 || cam->image_h > 65536 ||
        cam->image_w * cam->image_h > 0x7FFFFFFFull)   // pixel indices are 32-bit
        return fail(B200RT_EINVAL, "image dimensions out of range");
    if (cam->max_depth > 0xFFFFFFFFull) return fail(B200RT_EINVAL, "max_depth out of range");
    for (int a = 0; a < 3; ++a) {
        C.center[a] = cam->center[a]; C.pixel00[a] = cam->pixel00[a];
        C.delta_x[a] = cam->delta_x[a]; C.delta_y[a] = cam->delta_y[a];
        C.disk_x[a] = cam->disk_x[a]; C.disk_y[a] = cam->disk_y[a];
        C.background[a] = (float)cam->background[a];
    }
    C.defocus = cam->defocus_angle > 0 ? 1u : 0u;   // camera.h:186 (defocus_angle <= 0 -> pinhole)
    C.w = (uint32_t)cam->image_w; C.h = (uint32_t)cam->image_h; C.max_depth = (uint32_t)cam->max_depth;
    return B200RT_OK;
}

int render_on_device(SceneImpl *s, const B200rtCamera *cam, const B200rtRenderOpts *opts, float *d_out,
                     cudaStream_t st, B200rtStats *stats, bool sync_for_stats) {
    RenderParams P{};
    if (int rc = fill_camera(cam, P.cam)) return rc;
    B200rtRenderOpts o{};
    if (opts) o = *opts;
    const uint64_t count = o.sample_count ? o.sample_count : cam->spp;
    if (count > 0xFFFFFFFFull || o.sample_offset + count > 0xFFFFFFFFull) return fail(B200RT_EINVAL, "sample range out of range");
    if (o.variant != B200RT_VARIANT_MEGAKERNEL && o.variant != B200RT_VARIANT_MEGAKERNEL_VOTED &&
        o.variant != B200RT_VARIANT_WAVEFRONT)
        return fail(B200RT_EINVAL, "unknown kernel variant");
    P.scene = s->d;
    P.seed = o.seed;
    P.sample_begin = (uint32_t)o.sample_offset;
    P.sample_count = (uint32_t)count;
    P.out = d_out;
    P.flags = o.flags;
    P.scale = (o.flags & B200RT_FLAG_SUM) || count == 0 ? 1.0f : (float)(1.0 / (double)count);
    P.counters = s->d_counters;
    CUDA_TRY(cudaMemsetAsync(s->d_counters, 0, 3 * sizeof(unsigned long long), st));
    unsigned long long launches = 1;
    CUDA_TRY(cudaEventRecord(s->ev0, st));
    if (o.variant == B200RT_VARIANT_WAVEFRONT) {
        // pool of path slots: 2^21 by default (B200RT_WF_SLOTS overrides, for experiments)
        uint32_t want = 1u << 21;
        if (const char *e = std::getenv("B200RT_WF_SLOTS")) want = (uint32_t)std::max(1024ll, std::atoll(e));
        const unsigned long long items = (unsigned long long)cam->image_w * cam->image_h * count;
        if (items < want) want = (uint32_t)std::max(1024ull, items);
        if (want != s->pool_slots) {
            if (s->d_pool) { cudaStreamSynchronize(st); dev_free(s->d_pool); s->d_pool = nullptr; s->pool_slots = 0; }
            CUDA_TRY(dev_alloc(&s->d_pool, wavefront_pool_alloc_bytes(want)));
            s->pool_slots = want;
        }
        WavefrontPool W{};
        wavefront_pool_layout(s->d_pool, s->pool_slots, W);
        launches = 0;
        CUDA_TRY(run_wavefront(s->stack, P, W, (o.flags & B200RT_FLAG_COUNTERS) != 0, st, s->sm_count, &launches));
    } else {
        CUDA_TRY(launch_path_megakernel(s->stack, P, (o.flags & B200RT_FLAG_COUNTERS) != 0,
                                        o.variant == B200RT_VARIANT_MEGAKERNEL_VOTED, s->info.tree_depth <= 3, st));
    }
    CUDA_TRY(cudaEventRecord(s->ev1, st));
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->paths = (uint64_t)cam->image_w * cam->image_h * count;
        stats->kernel_launches = launches;
        if (sync_for_stats) {
            unsigned long long c[3];
            CUDA_TRY(cudaMemcpyAsync(c, s->d_counters, sizeof c, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
            stats->kernel_ms = ms;
            stats->rays = c[0]; stats->node_visits = c[1]; stats->prim_tests = c[2];
        }
    }
    return B200RT_OK;
}

std::vector<DeviceMaterial> device_materials(const B200rtSceneDesc *desc) {
    std::vector<DeviceMaterial> mats(desc->n_materials);
    for (uint64_t i = 0; i < desc->n_materials; ++i) {
        const B200rtMaterial &m = desc->materials[i];
        DeviceMaterial dm{};
        const double k = m.kind == B200RT_MAT_LIGHT ? m.param : 1.0;   // emit() = intensity * colour (material.h:261-263)
        dm.r = (float)(k * m.rgb[0]); dm.g = (float)(k * m.rgb[1]); dm.b = (float)(k * m.rgb[2]);
        dm.kind = m.kind;
        dm.param = m.kind == B200RT_MAT_METAL ? std::fmin(m.param, 1.0) : m.param;   // Metal ctor clamps fuzz (material.h:150-151)
        mats[i] = dm;
    }
    return mats;
}

// Host construction: parallel binned-SAH build + 4-wide collapse (bvh_builder.cpp), leaf-ordered SoA
// arrays assembled on the host, then uploaded.
int scene_build_host(const B200rtSceneDesc *desc, const B200rtBuildOpts *opts, SceneImpl *s) {
    const double t0 = now_ms();
    std::vector<Box3> boxes;
    compute_boxes(desc, boxes, host_threads(opts));
    BuiltBVH bvh;
    const char *err = nullptr;
    if (!build_bvh4(boxes, desc->n_spheres, desc->n_quads, build_params(opts), bvh, &err))
        return fail(B200RT_EINVAL, std::string("BVH build failed: ") + (err ? err : "?"));
    const int need_stack = (int)(3 * bvh.depth);
    if (need_stack > 128) return fail(B200RT_EINTERNAL, "BVH deeper than the largest traversal stack");

    // ---- leaf-ordered SoA arrays ----
    std::vector<double2> sph(desc->n_spheres * 2);
    std::vector<uint2> sph_meta(desc->n_spheres);
    parallel_ranges(desc->n_spheres, host_threads(opts), [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const B200rtSphere &s = desc->spheres[bvh.sphere_order[i]];
            sph[2 * i] = make_double2(s.c[0], s.c[1]);
            sph[2 * i + 1] = make_double2(s.c[2], s.r);
            sph_meta[i] = make_uint2(s.prim, s.mat);
        }
    });
    std::vector<double2> quads(desc->n_quads * 8);
    std::vector<uint2> quad_meta(desc->n_quads);
    parallel_ranges(desc->n_quads, host_threads(opts), [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        const B200rtQuad &q = desc->quads[bvh.quad_order[i]];
        const V3 s1 = v3(q.s1), s2 = v3(q.s2);
        const V3 n = crossv(s1, s2);                 // parallelogram.h:275-279
        const V3 un = unitv(n);
        const V3 w = divv(n, mag2(n));
        const double f[16] = {un.x, un.y, un.z, q.v[0], q.v[1], q.v[2], w.x, w.y, w.z,
                              s1.x, s1.y, s1.z, s2.x, s2.y, s2.z, 0.0};
        for (int k = 0; k < 8; ++k) quads[8 * i + k] = make_double2(f[2 * k], f[2 * k + 1]);
        quad_meta[i] = make_uint2(q.prim, q.mat);
      }
    });
    const std::vector<DeviceMaterial> mats = device_materials(desc);
    const double t1 = now_ms();

    // ---- upload ----
    uint64_t bytes = 0;
    int rc = B200RT_OK;
    const float4 *d_nodes = nullptr;
    {
        std::vector<float4> flat(bvh.nodes.size() * 8);
        std::memcpy(flat.data(), bvh.nodes.data(), bvh.nodes.size() * sizeof(Node4));
        rc = upload(flat, &d_nodes, &s->allocs[0], bytes);
    }
    if (!rc) rc = upload(sph, &s->d.spheres, &s->allocs[1], bytes);
    if (!rc) rc = upload(sph_meta, &s->d.sphere_meta, &s->allocs[2], bytes);
    if (!rc) rc = upload(quads, &s->d.quads, &s->allocs[3], bytes);
    if (!rc) rc = upload(quad_meta, &s->d.quad_meta, &s->allocs[4], bytes);
    if (!rc) rc = upload(mats, &s->d.materials, &s->allocs[5], bytes);
    s->d.nodes = d_nodes;
    if (rc) return rc;
    if (reinterpret_cast<uintptr_t>(d_nodes) & 127)   // trav_node_step forms plane addresses with OR / XOR on the low bits
        return fail(B200RT_ECUDA, "node array is not 128-byte aligned");
    const double t2 = now_ms();
    s->stack = need_stack <= 32 ? 32 : (need_stack <= 64 ? 64 : 128);
    s->info.n_nodes = bvh.nodes.size();
    s->info.device_bytes = bytes;
    s->info.tree_depth = bvh.depth;
    s->info.build_ms = t1 - t0;
    s->info.upload_ms = t2 - t1;
    return B200RT_OK;
}

// GPU construction (lbvh.cu): the caller's flat arrays go to the device as they are; bounds, Morton
// order, tree, 4-wide collapse and the leaf-ordered SoA arrays are all produced there.
// Returns B200RT_OK, or an error; `too_deep` is set when the tree needs a stack beyond 128 entries
// (the caller then falls back to the host SAH builder).
int scene_build_gpu(const B200rtSceneDesc *desc, SceneImpl *s, bool *too_deep) {
    *too_deep = false;
    const double t0 = now_ms();
    const uint32_t n_sph = (uint32_t)desc->n_spheres, n_quad = (uint32_t)desc->n_quads;
    uint64_t bytes = 0;
    B200rtSphere *raw_sph = nullptr;
    B200rtQuad *raw_quad = nullptr;
    auto drop_raw = [&]() { dev_free(raw_sph); dev_free(raw_quad); };
    CUDA_TRY(dev_alloc(&raw_sph, (size_t)n_sph * sizeof(B200rtSphere)));
    CUDA_TRY(dev_alloc(&raw_quad, (size_t)n_quad * sizeof(B200rtQuad)));
    cudaError_t e = cudaSuccess;
    if (n_sph) e = cudaMemcpy(raw_sph, desc->spheres, (size_t)n_sph * sizeof(B200rtSphere), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_quad) e = cudaMemcpy(raw_quad, desc->quads, (size_t)n_quad * sizeof(B200rtQuad), cudaMemcpyHostToDevice);
    double2 *d_sph = nullptr, *d_quads = nullptr;
    uint2 *d_sph_meta = nullptr, *d_quad_meta = nullptr;
    if (e == cudaSuccess) e = dev_alloc(&d_sph, (size_t)n_sph * 2 * sizeof(double2));
    s->allocs[1] = d_sph;
    if (e == cudaSuccess) e = dev_alloc(&d_sph_meta, (size_t)n_sph * sizeof(uint2));
    s->allocs[2] = d_sph_meta;
    if (e == cudaSuccess) e = dev_alloc(&d_quads, (size_t)n_quad * 8 * sizeof(double2));
    s->allocs[3] = d_quads;
    if (e == cudaSuccess) e = dev_alloc(&d_quad_meta, (size_t)n_quad * sizeof(uint2));
    s->allocs[4] = d_quad_meta;
    float4 *d_nodes = nullptr;
    uint32_t n_nodes = 0, depth = 0;
    if (e == cudaSuccess)
        e = build_lbvh_device(raw_sph, n_sph, raw_quad, n_quad, d_sph, d_sph_meta, d_quads, d_quad_meta, &d_nodes, &n_nodes, &depth);
    s->allocs[0] = d_nodes;
    drop_raw();
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200RT_ECUDA, std::string("GPU BVH build: ") + cudaGetErrorString(e)); }
    if (3 * depth > 128) { *too_deep = true; return B200RT_OK; }
    const std::vector<DeviceMaterial> mats = device_materials(desc);
    if (int rc = upload(mats, &s->d.materials, &s->allocs[5], bytes)) return rc;
    if (reinterpret_cast<uintptr_t>(d_nodes) & 127)   // trav_node_step forms plane addresses with OR / XOR on the low bits
        return fail(B200RT_ECUDA, "node array is not 128-byte aligned");
    s->d.nodes = d_nodes;
    s->d.spheres = d_sph; s->d.sphere_meta = d_sph_meta;
    s->d.quads = d_quads; s->d.quad_meta = d_quad_meta;
    CUDA_TRY(cudaDeviceSynchronize());
    bytes += (uint64_t)n_nodes * sizeof(Node4) + (uint64_t)n_sph * (2 * sizeof(double2) + sizeof(uint2)) +
             (uint64_t)n_quad * (8 * sizeof(double2) + sizeof(uint2));
    const int need_stack = (int)(3 * depth);
    s->stack = need_stack <= 32 ? 32 : (need_stack <= 64 ? 64 : 128);
    s->info.n_nodes = n_nodes;
    s->info.device_bytes = bytes;
    s->info.tree_depth = depth;
    s->info.build_ms = now_ms() - t0;
    s->info.upload_ms = 0;
    return B200RT_OK;
}

}  // namespace

// ==========================================================================================
extern "C" {

int b200rt_version(void) { return B200RT_VERSION; }
const char *b200rt_last_error(void) { return g_last_error.c_str(); }
int b200rt_device_count(void) { return device_count_quiet(); }

int b200rt_camera_init(B200rtCamera *c) {
    if (!c) return fail(B200RT_EINVAL, "camera is NULL");
    if (c->image_w == 0 || c->image_h == 0) return fail(B200RT_EINVAL, "image dimensions must be positive");
    if ((c->vfov >= 0) == (c->hfov >= 0)) return fail(B200RT_EINVAL, "exactly one of vfov / hfov must be given");
    // camera.h:87-157, same operations in the same order, in double
    const double aspect = (double)c->image_w / (double)c->image_h;
    const V3 dir = v3(c->dir), center = v3(c->center);
    if (c->focus_dist < 0) c->focus_dist = mag(dir);
    const double focal = c->focus_dist;
    double vw, vh;
    if (c->vfov >= 0) { vh = 2 * focal * std::tan(c->vfov / 2); vw = vh * aspect; }
    else { vw = 2 * focal * std::tan(c->hfov / 2); vh = vw / aspect; }
    const V3 u = unitv(dir);
    const V3 bz = {-u.x, -u.y, -u.z};
    const V3 bx = unitv(crossv(v3(c->up), bz));
    const V3 by = crossv(bz, bx);
    const V3 x_vec = mul(bx, vw), y_vec = mul(by, -vh);
    const V3 dx = divv(x_vec, (double)c->image_w), dy = divv(y_vec, (double)c->image_h);
    const V3 ulc = sub(sub(sub(center, mul(bz, focal)), divv(x_vec, 2)), divv(y_vec, 2));
    const V3 p00 = add(add(ulc, divv(dx, 2)), divv(dy, 2));
    const double rad = focal * std::tan(c->defocus_angle / 2);
    put(c->delta_x, dx); put(c->delta_y, dy); put(c->pixel00, p00);
    put(c->disk_x, mul(bx, rad)); put(c->disk_y, mul(by, rad));
    return B200RT_OK;
}

int b200rt_scene_create(const B200rtSceneDesc *desc, const B200rtBuildOpts *opts, void **scene_out) {
    if (!scene_out) return fail(B200RT_EINVAL, "scene_out is NULL");
    *scene_out = nullptr;
    if (int rc = validate_desc(desc)) return rc;
    if (device_count_quiet() == 0) return fail(B200RT_ENODEVICE, "no CUDA device: libb200rt has no CPU path");
    int dev = opts ? opts->device : -1;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; } }
    if (dev >= device_count_quiet()) return fail(B200RT_EINVAL, "device ordinal out of range");

    SceneImpl *s = new SceneImpl();
    s->device = dev;
    DeviceGuard g(dev);
    if (!g.ok) { delete s; return fail(B200RT_ECUDA, "cudaSetDevice failed"); }
    const uint64_t n_prims = desc->n_spheres + desc->n_quads;
    int builder = opts ? opts->builder : B200RT_BUILDER_AUTO;
    // AUTO: SAH on the host for small scenes (sub-millisecond, slightly better trees: +5 % on the
    // 4 k-sphere scene), Morton LBVH on the GPU from 64 k primitives up (2.2 M / 3.1 M primitives:
    // 57 / 90 ms incl. the upload vs 0.8 / 1.0 s, with equal or better render rates)
    if (builder == B200RT_BUILDER_AUTO) builder = n_prims >= 65536 ? B200RT_BUILDER_GPU_LBVH : B200RT_BUILDER_HOST_SAH;
    bool built = false;
    if (builder == B200RT_BUILDER_GPU_LBVH && n_prims >= 2) {
        bool too_deep = false;
        if (int rc = scene_build_gpu(desc, s, &too_deep)) { free_scene(s); return rc; }
        if (too_deep) {   // pathological depth: start over with the depth-capped host builder
            for (void *&p : s->allocs) { dev_free(p); p = nullptr; }
        } else {
            built = true;
        }
    }
    if (!built) {
        if (int rc = scene_build_host(desc, opts, s)) { free_scene(s); return rc; }
    }
    {
        cudaError_t e = dev_alloc(&s->d_counters, 3 * sizeof(unsigned long long));
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaEventCreate(&s->ev0);
        if (e == cudaSuccess) e = cudaEventCreate(&s->ev1);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { free_scene(s); return fail(B200RT_ECUDA, std::string("scene setup: ") + cudaGetErrorString(e)); }
    }
    s->info.n_prims = desc->n_spheres + desc->n_quads;
    s->info.n_spheres = desc->n_spheres; s->info.n_quads = desc->n_quads; s->info.n_materials = desc->n_materials;
    s->info.stack_entries = (uint32_t)s->stack;
    *scene_out = s;
    return B200RT_OK;
}

int b200rt_scene_info(void *scene, B200rtSceneInfo *info) {
    SceneImpl *s = as_scene(scene);
    if (!s || !info) return fail(B200RT_EINVAL, "bad scene handle");
    *info = s->info;
    return B200RT_OK;
}

void b200rt_scene_destroy(void *scene) { free_scene(as_scene(scene)); }

int b200rt_raycast(void *scene, const double *rays, int64_t n, double tmin, double tmax, int32_t *prim_out, double *t_out) {
    SceneImpl *s = as_scene(scene);
    if (!s) return fail(B200RT_EINVAL, "bad scene handle");
    if (n < 0 || (n && (!rays || !prim_out || !t_out))) return fail(B200RT_EINVAL, "bad ray buffers");
    if (n == 0) return B200RT_OK;
    DeviceGuard g(s->device);
    if ((size_t)n > s->ray_capacity) {
        if (s->d_rays) { dev_free(s->d_rays); dev_free(s->d_prim); dev_free(s->d_t); s->d_rays = nullptr; s->d_prim = nullptr; s->d_t = nullptr; }
        s->ray_capacity = 0;
        CUDA_TRY(dev_alloc(&s->d_rays, (size_t)n * 6 * sizeof(double)));
        CUDA_TRY(dev_alloc(&s->d_prim, (size_t)n * sizeof(int32_t)));
        CUDA_TRY(dev_alloc(&s->d_t, (size_t)n * sizeof(double)));
        s->ray_capacity = (size_t)n;
    }
    CUDA_TRY(cudaMemcpy(s->d_rays, rays, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(launch_raycast(s->stack, s->d, s->d_rays, n, tmin, tmax, s->d_prim, s->d_t, 0));
    CUDA_TRY(cudaMemcpy(prim_out, s->d_prim, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(t_out, s->d_t, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return B200RT_OK;
}

int b200rt_debug_camera_rays(const B200rtCamera *cam, const uint32_t *pixels_xy, const uint32_t *rnd, int64_t n, double *rays_out,
                             int device) {
    CameraParams C{};
    if (int rc = fill_camera(cam, C)) return rc;
    if (n < 0 || (n && (!pixels_xy || !rnd || !rays_out))) return fail(B200RT_EINVAL, "bad debug_camera_rays buffers");
    if (device_count_quiet() == 0) return fail(B200RT_ENODEVICE, "no CUDA device: libb200rt has no CPU path");
    if (n == 0) return B200RT_OK;
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; } }
    DeviceGuard g(device);
    uint32_t *d_pix = nullptr, *d_rnd = nullptr;
    double *d_rays = nullptr;
    cudaError_t e = dev_alloc(&d_pix, (size_t)n * 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = dev_alloc(&d_rnd, (size_t)n * 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = dev_alloc(&d_rays, (size_t)n * 6 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_pix, pixels_xy, (size_t)n * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_rnd, rnd, (size_t)n * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_debug_camera(C, d_pix, d_rnd, n, d_rays, 0);
    if (e == cudaSuccess) e = cudaMemcpy(rays_out, d_rays, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost);
    cudaStreamSynchronize(0);
    dev_free(d_pix); dev_free(d_rnd); dev_free(d_rays);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200RT_ECUDA, std::string("debug_camera_rays: ") + cudaGetErrorString(e)); }
    return B200RT_OK;
}

int b200rt_debug_shade(void *scene, const double *rays, const uint32_t *rnd, int64_t n, double tmin, double tmax,
                       B200rtShadeRecord *records_out) {
    static_assert(sizeof(B200rtShadeRecord) == 88, "record layout is shared with kernels.cu");
    SceneImpl *s = as_scene(scene);
    if (!s) return fail(B200RT_EINVAL, "bad scene handle");
    if (n < 0 || (n && (!rays || !rnd || !records_out))) return fail(B200RT_EINVAL, "bad debug_shade buffers");
    if (n == 0) return B200RT_OK;
    DeviceGuard g(s->device);
    double *d_rays = nullptr;
    uint32_t *d_rnd = nullptr;
    void *d_rec = nullptr;
    cudaError_t e = dev_alloc(&d_rays, (size_t)n * 6 * sizeof(double));
    if (e == cudaSuccess) e = dev_alloc(&d_rnd, (size_t)n * 4 * sizeof(uint32_t));
    if (e == cudaSuccess) e = dev_alloc(&d_rec, (size_t)n * sizeof(B200rtShadeRecord));
    if (e == cudaSuccess) e = cudaMemcpy(d_rays, rays, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_rnd, rnd, (size_t)n * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_debug_shade(s->stack, s->d, d_rays, d_rnd, n, tmin, tmax, d_rec, 0);
    if (e == cudaSuccess) e = cudaMemcpy(records_out, d_rec, (size_t)n * sizeof(B200rtShadeRecord), cudaMemcpyDeviceToHost);
    cudaStreamSynchronize(0);
    dev_free(d_rays); dev_free(d_rnd); dev_free(d_rec);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200RT_ECUDA, std::string("debug_shade: ") + cudaGetErrorString(e)); }
    return B200RT_OK;
}

int b200rt_render_device(void *scene, const B200rtCamera *cam, const B200rtRenderOpts *opts, float *out_rgb_device,
                         void *stream, B200rtStats *stats) {
    SceneImpl *s = as_scene(scene);
    if (!s) return fail(B200RT_EINVAL, "bad scene handle");
    if (!out_rgb_device) return fail(B200RT_EINVAL, "output pointer is NULL");
    DeviceGuard g(s->device);
    return render_on_device(s, cam, opts, out_rgb_device, static_cast<cudaStream_t>(stream), stats, stats != nullptr);
}

int b200rt_render(void *scene, const B200rtCamera *cam, const B200rtRenderOpts *opts, float *out_rgb, B200rtStats *stats) {
    SceneImpl *s = as_scene(scene);
    if (!s) return fail(B200RT_EINVAL, "bad scene handle");
    if (!cam || !out_rgb) return fail(B200RT_EINVAL, "camera or output pointer is NULL");
    DeviceGuard g(s->device);
    const double t0 = now_ms();
    const size_t floats = (size_t)cam->image_w * cam->image_h * 3;
    if (floats > s->frame_floats) {
        if (s->d_frame) { dev_free(s->d_frame); s->d_frame = nullptr; s->frame_floats = 0; }
        CUDA_TRY(dev_alloc(&s->d_frame, floats * sizeof(float)));
        s->frame_floats = floats;
    }
    B200rtRenderOpts o{};
    if (opts) o = *opts;
    o.flags &= ~(uint32_t)B200RT_FLAG_ACCUMULATE;   // host entry always overwrites
    B200rtStats local{};
    if (int rc = render_on_device(s, cam, &o, s->d_frame, 0, &local, true)) return rc;
    const double t1 = now_ms();
    CUDA_TRY(cudaMemcpy(out_rgb, s->d_frame, floats * sizeof(float), cudaMemcpyDeviceToHost));
    const double t2 = now_ms();
    local.d2h_ms = t2 - t1;
    local.d2h_bytes = floats * sizeof(float);
    local.total_ms = t2 - t0;
    if (stats) *stats = local;
    return B200RT_OK;
}

int b200rt_render_scene(const B200rtSceneDesc *desc, const B200rtCamera *cam, const B200rtRenderOpts *opts,
                        const B200rtBuildOpts *bopts, float *out_rgb, B200rtStats *stats, B200rtSceneInfo *info) {
    const double t0 = now_ms();
    void *scene = nullptr;
    if (int rc = b200rt_scene_create(desc, bopts, &scene)) return rc;
    B200rtStats local{};
    int rc = b200rt_render(scene, cam, opts, out_rgb, &local);
    SceneImpl *s = as_scene(scene);
    if (!rc) {
        local.h2d_ms = s->info.upload_ms;
        local.h2d_bytes = s->info.device_bytes;
        if (info) *info = s->info;
    }
    b200rt_scene_destroy(scene);
    local.total_ms = now_ms() - t0;
    if (!rc && stats) *stats = local;
    return rc;
}

int b200rt_tonemap_device(const float *hdr_device, int64_t n_pixels, int32_t *out_device, int clamp, int device, void *stream) {
    if (n_pixels < 0 || (n_pixels && (!hdr_device || !out_device))) return fail(B200RT_EINVAL, "bad tonemap buffers");
    if (device_count_quiet() == 0) return fail(B200RT_ENODEVICE, "no CUDA device: libb200rt has no CPU path");
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; } }
    DeviceGuard g(device);
    CUDA_TRY(launch_tonemap(hdr_device, n_pixels, out_device, clamp, static_cast<cudaStream_t>(stream)));
    return B200RT_OK;
}

int b200rt_finalize_device(float *frame_device, int64_t n_pixels, double scale, int32_t *ldr_device_or_null, int clamp,
                           int device, void *stream) {
    if (n_pixels < 0 || (n_pixels && !frame_device)) return fail(B200RT_EINVAL, "bad frame buffer");
    if (device_count_quiet() == 0) return fail(B200RT_ENODEVICE, "no CUDA device: libb200rt has no CPU path");
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; } }
    DeviceGuard g(device);
    CUDA_TRY(launch_finalize(frame_device, n_pixels, (float)scale, ldr_device_or_null, clamp, static_cast<cudaStream_t>(stream)));
    return B200RT_OK;
}

int b200rt_finalize_peers_device(const float *const *peer_frames, int n_peers, int rank, int64_t n_pixels, double scale,
                                 float *root_hdr, int32_t *root_ldr_or_null, int clamp, int device, void *stream) {
    if (n_peers < 1 || n_peers > kMaxPeers || rank < 0 || rank >= n_peers) return fail(B200RT_EINVAL, "bad peer count / rank");
    if (n_pixels < 0 || !peer_frames || (n_pixels && !root_hdr)) return fail(B200RT_EINVAL, "bad frame buffers");
    PeerFrames in{};
    for (int r = 0; r < n_peers; ++r) {
        if (n_pixels && (!peer_frames[r] || (reinterpret_cast<uintptr_t>(peer_frames[r]) & 15)))
            return fail(B200RT_EINVAL, "peer frame pointers must be non-null and 16-byte aligned");
        in.p[r] = peer_frames[r];
    }
    if ((reinterpret_cast<uintptr_t>(root_hdr) & 15) || (reinterpret_cast<uintptr_t>(root_ldr_or_null) & 15))
        return fail(B200RT_EINVAL, "root buffers must be 16-byte aligned");
    if (device_count_quiet() == 0) return fail(B200RT_ENODEVICE, "no CUDA device: libb200rt has no CPU path");
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; } }
    DeviceGuard g(device);
    CUDA_TRY(launch_reduce_finalize_peers(in, n_peers, rank, n_pixels, (float)scale, root_hdr, root_ldr_or_null, clamp,
                                          static_cast<cudaStream_t>(stream)));
    return B200RT_OK;
}

int b200rt_tonemap(const float *hdr, int64_t n_pixels, int32_t *out, int clamp) {
    if (n_pixels < 0 || (n_pixels && (!hdr || !out))) return fail(B200RT_EINVAL, "bad tonemap buffers");
    if (device_count_quiet() == 0) return fail(B200RT_ENODEVICE, "no CUDA device: libb200rt has no CPU path");
    if (n_pixels == 0) return B200RT_OK;
    float *d_in = nullptr;
    int32_t *d_out = nullptr;
    CUDA_TRY(dev_alloc(&d_in, (size_t)n_pixels * 3 * sizeof(float)));
    cudaError_t e = dev_alloc(&d_out, (size_t)n_pixels * 3 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemcpy(d_in, hdr, (size_t)n_pixels * 3 * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_tonemap(d_in, n_pixels, d_out, clamp, 0);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n_pixels * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaStreamSynchronize(0);
    dev_free(d_in);
    dev_free(d_out);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200RT_ECUDA, std::string("tonemap: ") + cudaGetErrorString(e)); }
    return B200RT_OK;
}

// ---- CPU-side self test of the host builder (no GPU needed; used by tests -m "not gpu") -------
// Builds the BVH for `desc` and checks its structural invariants.  Fills depth / node count.
int b200rt_selftest_bvh(const B200rtSceneDesc *desc, const B200rtBuildOpts *opts, B200rtSceneInfo *info) {
    if (int rc = validate_desc(desc)) return rc;
    std::vector<Box3> boxes;
    compute_boxes(desc, boxes, host_threads(opts));
    BuiltBVH bvh;
    const char *err = nullptr;
    const double t0 = now_ms();
    if (!build_bvh4(boxes, desc->n_spheres, desc->n_quads, build_params(opts), bvh, &err))
        return fail(B200RT_EINVAL, std::string("BVH build failed: ") + (err ? err : "?"));
    const double t1 = now_ms();
    if (!validate_bvh4(bvh, boxes, desc->n_spheres, desc->n_quads, &err))
        return fail(B200RT_EINTERNAL, std::string("BVH invariant violated: ") + (err ? err : "?"));
    if (info) {
        std::memset(info, 0, sizeof *info);
        info->n_prims = desc->n_spheres + desc->n_quads;
        info->n_spheres = desc->n_spheres; info->n_quads = desc->n_quads; info->n_materials = desc->n_materials;
        info->n_nodes = bvh.nodes.size();
        info->tree_depth = bvh.depth;
        info->stack_entries = 3 * bvh.depth;
        info->build_ms = t1 - t0;
    }
    return B200RT_OK;
}

}  // extern "C"
