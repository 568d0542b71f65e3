/* b200rt.h -- C ABI of the B200-native path tracer (libb200rt.so).
 *
 * This is the drop-in boundary for ONE path of DeltaPavonis/cpp_raytracer: the render call
 *     Image Camera::render(const Scene &world)                 (reference include/base/camera.h:301-303)
 *     template<T> auto Camera::render(const T &world)          (reference include/base/camera.h:264-297)
 * and the closest-hit query underneath it
 *     BVH::hit_by / Scene::hit_by                              (reference include/acceleration/bvh.h:585-715,
 *                                                               include/base/scene.h:59-75).
 * The reference has no FFI; a maintainer would bind these entry points from Camera::render
 * (see INTEGRATION.md for the stub).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions: every function returns 0 on success and a negative B200RT_E* code on failure
 * (never exit()s, unlike the reference's std::exit(-1) in image.h:39-42); the message is
 * available from b200rt_last_error() on the calling thread.  The caller owns every host
 * buffer; the library owns device memory behind the opaque scene handle.  Calls block until
 * their result is in the caller's buffer unless stated otherwise.  One handle may be used
 * from one host thread at a time.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with B200RT_ENODEVICE.
 */
#ifndef B200RT_H
#define B200RT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RT_VERSION 1

enum {
    B200RT_OK = 0,
    B200RT_EINVAL = -1,      /* bad argument / malformed scene (unknown material kind, index out of range ...) */
    B200RT_ENODEVICE = -2,   /* no usable CUDA device: there is no CPU path */
    B200RT_ECUDA = -3,       /* CUDA runtime error; text in b200rt_last_error() */
    B200RT_ENOMEM = -4,
    B200RT_EINTERNAL = -5
};

/* Material kinds: the closed set of reference materials (include/base/material.h). */
enum {
    B200RT_MAT_LAMBERTIAN = 0,   /* rgb = intrinsic colour                (material.h:58-95)   */
    B200RT_MAT_METAL = 1,        /* rgb = colour, param = fuzz (<= 1)     (material.h:105-152) */
    B200RT_MAT_DIELECTRIC = 2,   /* param = refractive index              (material.h:164-227) */
    B200RT_MAT_LIGHT = 3         /* rgb = colour, param = intensity       (material.h:231-275) */
};

/* Flat scene description.  `prim` is the canonical primitive index: the position of the
 * primitive in Scene::get_primitive_components() order (scene.h:85-105; a Box contributes its
 * six faces in constructor order, box.h:53-84).  b200rt_raycast reports hits by this index. */
#pragma pack(push, 1)
typedef struct B200rtMaterial { uint32_t kind, pad; double rgb[3]; double param; } B200rtMaterial;      /* 40 B */
typedef struct B200rtSphere { double c[3]; double r; uint32_t mat, prim; } B200rtSphere;                /* 40 B, sphere.h:16-22 */
typedef struct B200rtQuad { double v[3], s1[3], s2[3]; uint32_t mat, prim; } B200rtQuad;                /* 80 B, parallelogram.h:14-48 */

/* Camera.  The first block is what the reference's fluent setters store (camera.h:308-406);
 * the second block is what Camera::init() derives from it (camera.h:87-157).
 * b200rt_camera_init() fills the second block from the first, in double precision on the
 * host, operation for operation as the reference does.  Angles are radians; vfov/hfov < 0
 * means "not given" (exactly one of them must be given). */
typedef struct B200rtCamera {
    uint64_t image_w, image_h, spp, max_depth;
    double center[3], dir[3], up[3];
    double focus_dist;            /* < 0: default to |dir| (camera.h:101-103) */
    double defocus_angle;         /* radians; <= 0 turns blur off (camera.h:186) */
    double vfov, hfov;
    double background[3];
    /* derived */
    double pixel00[3], delta_x[3], delta_y[3], disk_x[3], disk_y[3];
} B200rtCamera;
#pragma pack(pop)

typedef struct B200rtSceneDesc {
    uint64_t n_materials, n_spheres, n_quads;
    const B200rtMaterial *materials;
    const B200rtSphere *spheres;
    const B200rtQuad *quads;
} B200rtSceneDesc;

/* Acceleration-structure build knobs (all optional; pass NULL for defaults).  The reference's
 * knobs are BVH(world, num_buckets = 32, max_primitives_in_node = 12) (bvh.h:754-756); the
 * tree here is a different design (4-wide, FP32 conservative boxes), results are tree
 * independent. */
typedef struct B200rtBuildOpts {
    int32_t device;          /* CUDA device ordinal; -1 = current device */
    int32_t max_leaf_prims;  /* 1..8, 0 = default */
    int32_t sah_bins;        /* 4..64, 0 = default */
    int32_t build_threads;   /* host threads for the builder, 0 = all */
    int32_t builder;         /* B200RT_BUILDER_* */
} B200rtBuildOpts;

enum {
    B200RT_BUILDER_AUTO = 0,       /* HOST_SAH below 65536 primitives, GPU_LBVH from there up */
    B200RT_BUILDER_HOST_SAH = 1,   /* parallel binned SAH on the host: best trees, ~0.3 us per primitive */
    B200RT_BUILDER_GPU_LBVH = 2    /* Morton-order LBVH built on the GPU: tens of ms for millions of primitives,
                                      somewhat more node visits per ray; falls back to HOST_SAH if too deep */
};

typedef struct B200rtSceneInfo {
    uint64_t n_prims, n_spheres, n_quads, n_materials;
    uint64_t n_nodes;             /* 4-wide nodes, 128 B each */
    uint64_t device_bytes;        /* total device footprint of the scene */
    uint32_t tree_depth;          /* depth of the 4-wide tree */
    uint32_t stack_entries;       /* traversal stack entries the kernels are instantiated with */
    double build_ms, upload_ms;
} B200rtSceneInfo;

enum {
    B200RT_VARIANT_MEGAKERNEL = 0,          /* one thread per pixel, path regeneration (default) */
    B200RT_VARIANT_WAVEFRONT = 1,           /* per-material ray queues */
    B200RT_VARIANT_MEGAKERNEL_VOTED = 2     /* experiment: warp-voted node/leaf/shade step scheduling */
};
enum {
    B200RT_FLAG_SUM = 1,          /* write the per-pixel SUM over the samples of this call instead of the mean */
    B200RT_FLAG_ACCUMULATE = 2,   /* device entry only: add into the output buffer instead of overwriting */
    B200RT_FLAG_COUNTERS = 4      /* also count node visits and primitive tests (slower; for roofline accounting) */
};

typedef struct B200rtRenderOpts {
    uint64_t seed;             /* RNG key; streams are a pure function of (seed, pixel, sample, bounce) */
    uint64_t sample_offset;    /* first sample index of this call (sample split across GPUs) */
    uint64_t sample_count;     /* samples per pixel rendered by this call; 0 = camera.spp */
    int32_t variant;           /* B200RT_VARIANT_* */
    uint32_t flags;            /* B200RT_FLAG_* */
} B200rtRenderOpts;

typedef struct B200rtStats {
    double kernel_ms;          /* CUDA-event time of the path kernel(s) on their stream */
    double h2d_ms, d2h_ms, total_ms;
    uint64_t paths, rays;      /* pixel-samples and hit_by-equivalent ray casts */
    uint64_t node_visits, prim_tests;   /* only with B200RT_FLAG_COUNTERS */
    uint64_t kernel_launches;
    uint64_t h2d_bytes, d2h_bytes;
} B200rtStats;

/* ---- lifetime ------------------------------------------------------------------------- */
int b200rt_device_count(void);
const char *b200rt_last_error(void);
int b200rt_version(void);

/* Replaces: BVH::BVH(world) + the pointer graph it keeps (bvh.h:754-776).  Flattens the
 * primitives into SoA device arrays, builds the wide BVH and uploads everything. */
int b200rt_scene_create(const B200rtSceneDesc *desc, const B200rtBuildOpts *opts, void **scene_out);
int b200rt_scene_info(void *scene, B200rtSceneInfo *info);
void b200rt_scene_destroy(void *scene);

/* Replaces: Camera::init() (camera.h:87-157).  Host, double precision. */
int b200rt_camera_init(B200rtCamera *cam);

/* ---- closest hit (deterministic ray-cast harness) -------------------------------------- */
/* Replaces: world.hit_by(ray, Interval(tmin, tmax)) for n rays (bvh.h:585-715 /
 * scene.h:59-75).  rays = n x {ox,oy,oz,dx,dy,dz} doubles (directions NOT normalised, as in
 * the reference).  prim_out[i] = canonical primitive index of the closest hit with
 * tmin < t < tmax, or -1; t_out[i] = its hit time (0 on miss).  Exact ties in t resolve to the
 * lowest canonical index, which is what Scene::hit_by returns. */
int b200rt_raycast(void *scene, const double *rays, int64_t n, double tmin, double tmax,
                   int32_t *prim_out, double *t_out);

/* Test hook: ONE surface interaction per ray, performed by the same device function the path kernels
 * call -- closest hit (as b200rt_raycast), hit point, the hit_info face rule (hittable.h:46-71), then
 * Lambertian / Metal / Dielectric / DiffuseLight scatter + emit (material.h:64-263, vec3d.h:144-200) on a
 * unit throughput -- but with the caller's four 32-bit random words per ray in place of the kernel's
 * Philox stream, so each branch can be checked deterministically.  Word use: Lambertian and Metal
 * sample the unit sphere from words 0,1 (u = (w >> 8) / 2^24; z = 1 - 2 u0, azimuth 2 pi u1); Dielectric
 * compares u(word 2) with the Schlick reflectance (no word is used under total internal reflection). */
typedef struct B200rtShadeRecord {
    double scattered[6];   /* origin (= hit point) and unnormalised direction of the scattered ray; zeros if none */
    double t;              /* hit time */
    float atten[3];        /* attenuation applied (1,1,1 for Dielectric); zeros if the path ended */
    float emit[3];         /* emitted radiance (DiffuseLight: intensity * colour, both faces) */
    int32_t prim;          /* canonical primitive index, -1 = miss */
    int32_t flags;         /* bit 0: scattered ray present */
} B200rtShadeRecord;
int b200rt_debug_shade(void *scene, const double *rays, const uint32_t *rnd, int64_t n, double tmin, double tmax,
                       B200rtShadeRecord *records_out);

/* Test hook: the kernels' primary-ray construction (Camera::random_ray_through_pixel, camera.h:184-200, defocus
 * disk camera.h:160-168) for n pixels (pixels_xy = n x {col, row}) with the caller's four random words per ray
 * (u = (w >> 8) / 2^24): words 0,1 -> the U(-0.5,0.5) factors of delta_x, delta_y; words 2,3 -> the point in the
 * unit disk (radius sqrt(u2), azimuth 2 pi u3), used only when defocus_angle > 0.  rays_out = n x 6 doubles. */
int b200rt_debug_camera_rays(const B200rtCamera *cam, const uint32_t *pixels_xy, const uint32_t *rnd, int64_t n,
                             double *rays_out, int device);

/* ---- render ---------------------------------------------------------------------------- */
/* Replaces: Camera::render<BVH>(bvh) (camera.h:264-297): for every pixel, the mean (or sum)
 * over the requested samples of ray_color (camera.h:205-258).  out_rgb = image_h x image_w x 3
 * floats, row-major, row 0 at the top, linear HDR (what the reference stores in Image before
 * RGB::as_string tone-maps it). */
int b200rt_render(void *scene, const B200rtCamera *cam, const B200rtRenderOpts *opts,
                  float *out_rgb, B200rtStats *stats);

/* Same, but out_rgb_device is a DEVICE pointer on the scene's device and `stream` a
 * cudaStream_t (NULL = the legacy default stream).  Returns after enqueueing when stats is
 * NULL; with stats it synchronises the stream to read the counters back. */
int b200rt_render_device(void *scene, const B200rtCamera *cam, const B200rtRenderOpts *opts,
                         float *out_rgb_device, void *stream, B200rtStats *stats);

/* Replaces: Camera::render(const Scene &) end to end (camera.h:301-303): build + upload +
 * render + read back + free, all inside one call on host buffers. */
int b200rt_render_scene(const B200rtSceneDesc *desc, const B200rtCamera *cam,
                        const B200rtRenderOpts *opts, const B200rtBuildOpts *bopts,
                        float *out_rgb, B200rtStats *stats, B200rtSceneInfo *info);

/* ---- tone map --------------------------------------------------------------------------- */
/* Replaces: RGB::as_string() defaults (rgb.h:90-113): Reinhard by luminance, gamma 2,
 * int(255.999999 * v), NO clamp unless clamp != 0 (the reference does not clamp, so a
 * saturated channel can exceed 255).  hdr = n_pixels x 3 floats; out = n_pixels x 3 int32. */
int b200rt_tonemap(const float *hdr, int64_t n_pixels, int32_t *out, int clamp);
int b200rt_tonemap_device(const float *hdr_device, int64_t n_pixels, int32_t *out_device,
                          int clamp, int device, void *stream);

/* Frame epilogue for the multi-GPU sample split, on the rank that holds the reduced SUM frame:
 * frame *= scale in place (pixel_color /= samples_per_pixel, camera.h:290) and, when
 * ldr_device_or_null is given, the tone-mapped integers of the same pixels in the same pass. */
int b200rt_finalize_device(float *frame_device, int64_t n_pixels, double scale, int32_t *ldr_device_or_null,
                           int clamp, int device, void *stream);

/* The same epilogue with the exchange fused in, over peer-mapped memory (NVLink / NVSwitch), in place
 * of "reduce to rank 0, then b200rt_finalize_device": one launch per rank, each rank owning one slice
 * of the frame.  peer_frames[r] is rank r's n_pixels x 3 FP32 SUM buffer as addressable FROM THIS
 * device (CUDA peer / symmetric-memory mapping; peer_frames[rank] is the local one); the kernel adds
 * the n_peers buffers in rank order over its slice, scales, and writes the mean frame into root_hdr
 * and (when given) the tone-mapped integers into root_ldr_or_null -- both pointers into the ROOT
 * rank's memory as addressable from this device.  root_hdr may be peer_frames[0] itself (in place).
 * All pointers must be 16-byte aligned.  The caller orders it: every rank's render must be complete
 * and visible before any rank launches this (a device barrier across ranks), and a second barrier
 * must pass before the root reads the result or any rank reuses its sum buffer.  Slices:
 * chunk = ceil(n_pixels / n_peers) rounded up to 1024 pixels; rank k owns
 * [min(n, k*chunk), min(n, (k+1)*chunk)).  1 <= n_peers <= 16.
 * Replaces: the serial accumulation of camera.h:286-290 across GPUs (no reference counterpart for
 * the exchange itself). */
int b200rt_finalize_peers_device(const float *const *peer_frames, int n_peers, int rank, int64_t n_pixels,
                                 double scale, float *root_hdr, int32_t *root_ldr_or_null, int clamp,
                                 int device, void *stream);

/* ---- host-only self test ---------------------------------------------------------------------- */
/* Builds the acceleration structure for `desc` on the host and checks its invariants (every
 * primitive in exactly one leaf, boxes nested, reported depth exact).  Needs no GPU; fills
 * n_nodes / tree_depth / stack_entries / build_ms of `info`. */
int b200rt_selftest_bvh(const B200rtSceneDesc *desc, const B200rtBuildOpts *opts, B200rtSceneInfo *info);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H */
