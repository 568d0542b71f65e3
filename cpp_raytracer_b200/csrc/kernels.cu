// kernels.cu -- the CUDA kernels behind the C ABI (sm_100a only; compiled with -fmad=false so
// FP64 primitive tests reproduce the reference's unfused arithmetic bit for bit).
//
//   raycast_kernel        one thread per ray, closest hit -> (canonical prim index, t)
//                         [BVH::hit_by / Scene::hit_by, reference bvh.h:585-715, scene.h:59-75]
//   path_megakernel       one thread per pixel, all of that pixel's samples, with path regeneration:
//                         a lane whose path ended starts its next sample right away
//                         [Camera::render<T> + ray_color, reference camera.h:205-297]
//   tonemap_kernel        Reinhard + gamma 2 + int(255.999999 v)   [RGB::as_string, rgb.h:90-113]
#include "kernels.h"
#include "shade.cuh"

namespace b200rt {

// ------------------------------------------------------------------------------------------
template <int STACK>
__global__ void __launch_bounds__(128) raycast_kernel(DeviceScene S, const double *__restrict__ rays, long long n,
                                                      double tmin, double tmax, int32_t *__restrict__ prim_out,
                                                      double *__restrict__ t_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *r = rays + i * 6;
    const Hit h = closest_hit<STACK, false>(S, r[0], r[1], r[2], r[3], r[4], r[5], tmin, tmax, nullptr);
    if (h.ref == kNoHit) {
        prim_out[i] = -1;
        t_out[i] = 0.0;
    } else {
        prim_out[i] = (int32_t)canonical_prim(S, h.ref);
        t_out[i] = h.t;
    }
}

// One surface interaction per ray with the CALLER's random words: closest hit, then shade_hit -- the very
// device function the path kernels call -- on a unit throughput.  Test hook (b200rt_debug_shade): makes the
// hit_info face rule and the four materials checkable ray by ray against the oracle's reflected / refracted /
// reflectance, which the path kernels' own Philox streams do not allow.
struct ShadeRecord {
    double scattered[6];   // origin, direction of the scattered ray (zeros if the path ended)
    double t;
    float atten[3];        // throughput after the interaction (the material's attenuation)
    float emit[3];         // emitted radiance picked up at the hit
    int32_t prim;          // canonical primitive index or -1
    int32_t flags;         // bit 0: a scattered ray continues
};
static_assert(sizeof(ShadeRecord) == 88, "ShadeRecord is mirrored by numpy in capi.py");

// Primary rays for given pixels with the CALLER's random words (test hook b200rt_debug_camera_rays): the
// kernels' camera_ray, checkable against Camera::random_ray_through_pixel draw for draw.
__global__ void debug_camera_kernel(const __grid_constant__ CameraParams C, const uint32_t *__restrict__ pixels,
                                    const uint32_t *__restrict__ rnd, long long n, double *__restrict__ rays_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox4 w{rnd[i * 4 + 0], rnd[i * 4 + 1], rnd[i * 4 + 2], rnd[i * 4 + 3]};
    Ray r;
    PathState p;
    camera_ray(C, pixels[i * 2 + 0], pixels[i * 2 + 1], w, r, p);
    double *o = rays_out + i * 6;
    o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.dx; o[4] = r.dy; o[5] = r.dz;
}

// Test hooks (b200rt_debug_philox / b200rt_debug_samplers): the generator and the direction samplers exactly as the
// path kernels call them, on caller-supplied inputs.
__global__ void debug_philox_kernel(const uint32_t *__restrict__ in, long long n, uint32_t *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Philox4 r = philox4x32_10(in[i * 6 + 0], in[i * 6 + 1], in[i * 6 + 2], in[i * 6 + 3], in[i * 6 + 4], in[i * 6 + 5]);
    out[i * 4 + 0] = r.x; out[i * 4 + 1] = r.y; out[i * 4 + 2] = r.z; out[i * 4 + 3] = r.w;
}
__global__ void debug_samplers_kernel(const uint32_t *__restrict__ rnd, long long n, double *__restrict__ sphere_out,
                                      double *__restrict__ disk_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y, z;
    sample_unit_sphere(u01(rnd[i * 2 + 0]), u01(rnd[i * 2 + 1]), x, y, z);
    sphere_out[i * 3 + 0] = x; sphere_out[i * 3 + 1] = y; sphere_out[i * 3 + 2] = z;
    double ax, ay;
    sample_unit_disk(u01(rnd[i * 2 + 0]), u01(rnd[i * 2 + 1]), ax, ay);
    disk_out[i * 2 + 0] = ax; disk_out[i * 2 + 1] = ay;
}

template <int STACK>
__global__ void __launch_bounds__(128) debug_shade_kernel(DeviceScene S, const double *__restrict__ rays,
                                                          const uint32_t *__restrict__ rnd, long long n, double tmin, double tmax,
                                                          ShadeRecord *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r{rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2], rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]};
    const Hit h = closest_hit<STACK, false>(S, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, tmin, tmax, nullptr);
    ShadeRecord rec{};
    rec.prim = -1;
    if (h.ref != kNoHit) {
        const Philox4 w{rnd[i * 4 + 0], rnd[i * 4 + 1], rnd[i * 4 + 2], rnd[i * 4 + 3]};
        PathState p{1.0f, 1.0f, 1.0f};
        float er = 0.0f, eg = 0.0f, eb = 0.0f;
        const bool cont = shade_hit(S, h, w, r, p, er, eg, eb);
        rec.prim = (int32_t)canonical_prim(S, h.ref);
        rec.t = h.t;
        rec.flags = cont ? 1 : 0;
        if (cont) {
            rec.scattered[0] = r.ox; rec.scattered[1] = r.oy; rec.scattered[2] = r.oz;
            rec.scattered[3] = r.dx; rec.scattered[4] = r.dy; rec.scattered[5] = r.dz;
            rec.atten[0] = p.tr; rec.atten[1] = p.tg; rec.atten[2] = p.tb;
        }
        rec.emit[0] = er; rec.emit[1] = eg; rec.emit[2] = eb;
    }
    out[i] = rec;
}

// ------------------------------------------------------------------------------------------
// Thread -> pixel: a block is an 8 x (kPathBlock / 8) pixel tile, each warp an 8x4 sub-tile, so the 32 primary
// rays of a warp are neighbours on the image plane.
struct PixelMap {
    uint32_t px, py, pixel;
    bool valid;
};
__device__ __forceinline__ PixelMap map_pixel(const CameraParams &C) {
    const uint32_t tiles_x = (C.w + (uint32_t)kPathTileW - 1u) / (uint32_t)kPathTileW;
    const uint32_t tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    PixelMap m;
    m.px = tile_x * (uint32_t)kPathTileW + (threadIdx.x % (uint32_t)kPathTileW);
    m.py = tile_y * (uint32_t)kPathTileH + (threadIdx.x / (uint32_t)kPathTileW);
    m.valid = m.px < C.w && m.py < C.h;
    m.pixel = m.py * C.w + m.px;
    return m;
}

__device__ __forceinline__ void write_pixel_and_counters(const RenderParams &P, const PixelMap &m, float sum_r, float sum_g,
                                                         float sum_b, uint32_t lane_rays, bool count,
                                                         const TraversalCounters &ctr) {
    unsigned long long rays = lane_rays;
    if (m.valid && B200RT_CHECK(P.scene, m.pixel < P.cam.w * P.cam.h, 3)) {
        float *o = P.out + (size_t)m.pixel * 3;
        const float k = P.scale;
        if (P.flags & kRenderAccumulate) { o[0] += sum_r * k; o[1] += sum_g * k; o[2] += sum_b * k; }
        else { o[0] = sum_r * k; o[1] = sum_g * k; o[2] = sum_b * k; }
    }
    // counters: one atomic per warp
    for (int off = 16; off; off >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, off);
    if ((threadIdx.x & 31) == 0 && rays) atomicAdd(&P.counters[0], rays);
    if (count) {
        unsigned long long a = ctr.nodes, b = ctr.prims, c = ctr.quads;
        for (int off = 16; off; off >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
            c += __shfl_down_sync(0xffffffffu, c, off);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&P.counters[1], a); atomicAdd(&P.counters[2], b); atomicAdd(&P.counters[3], c); }
    }
}

// One path segment's worth of "shade + continue or regenerate": consumes the finished traversal
// in T (if any), and leaves either a new ray in `ray` (returns true) or the lane finished.
struct LaneState {
    PathState p;
    Ray ray;
    uint32_t s = 0, bounce = 0;
    bool has_path = false;
    float sum_r = 0.f, sum_g = 0.f, sum_b = 0.f;   // every contribution (throughput x emission / background) is added here directly
    uint32_t rays = 0;
};
__device__ __forceinline__ bool shade_and_advance(const RenderParams &P, const PixelMap &m, const Hit &best, LaneState &L) {
    const CameraParams &C = P.cam;
    const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
    if (L.has_path) {
        ++L.rays;
        bool cont;
        if (best.ref == kNoHit) {
            L.sum_r += L.p.tr * C.background[0]; L.sum_g += L.p.tg * C.background[1]; L.sum_b += L.p.tb * C.background[2];
            cont = false;   // camera.h:248
        } else {
            const Philox4 rnd = philox4x32_10(m.pixel, P.sample_begin + L.s, L.bounce + 1u, 0u, k0, k1);
            cont = shade_hit(P.scene, best, rnd, L.ray, L.p, L.sum_r, L.sum_g, L.sum_b);
            // ray_color(scattered, depth_left - 1): contributes nothing once depth_left hits 0 (camera.h:211-213)
            if (cont && ++L.bounce == C.max_depth) cont = false;
        }
        if (!cont) {
            L.has_path = false;
            ++L.s;
        }
    }
    if (!L.has_path) {
        if (L.s == P.sample_count) return false;
        const Philox4 rnd = philox4x32_10(m.pixel, P.sample_begin + L.s, 0u, 0u, k0, k1);
        camera_ray(C, m.px, m.py, rnd, L.ray, L.p);
        L.bounce = 0;
        L.has_path = true;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// Tile work pool.  A block owns an 8 x 32 pixel tile (kPathTileW x kPathTileH) but its THREADS do not own pixels: the tile's (pixel, sample) items
// -- sample-major, so that the lanes of a warp work on neighbouring pixels of the same sample at any time -- are handed
// out through a shared-memory counter, and a lane whose path ends takes the next item whatever pixel it belongs to.
// With thread-owns-pixel a lane on a cheap pixel (sky: one ray per path) ran out of samples long before its neighbours
// on glass (five rays per path) and idled for the rest of the warp's life: 12-22 % of all lane-slots of the node and
// leaf steps (profiles/r2_lane_accounting.md).
//   Radiance is summed per pixel in shared memory as 64-bit FIXED POINT (2^-30 units) with integer atomics: integer
// addition is associative, so the sum -- and with it the image -- is bit-deterministic and does not depend on which
// lane rendered which sample, in which order.  One contribution per path (emission at a light, or the background on a
// miss), so three atomics per path at most.  Resolution 9.3e-10, range +-8.6e9 per pixel per launch.
struct TilePool {
    unsigned int next;                 // next (pixel, sample) item of this tile
    long long acc[kPathBlock * 3];     // per tile pixel: sum of radiance x 2^30
};
constexpr float kFixScale = 1073741824.0f;            // 2^30
constexpr double kFixInv = 1.0 / 1073741824.0;

struct TileMap {
    uint32_t x0, y0, vw, n_valid;      // tile origin, valid width, valid pixel count (edge tiles are ragged)
};
__device__ __forceinline__ TileMap map_tile(const CameraParams &C) {
    const uint32_t tiles_x = (C.w + (uint32_t)kPathTileW - 1u) / (uint32_t)kPathTileW;
    const uint32_t tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    TileMap t;
    t.x0 = tile_x * (uint32_t)kPathTileW;
    t.y0 = tile_y * (uint32_t)kPathTileH;
    t.vw = min((uint32_t)kPathTileW, C.w - t.x0);
    t.n_valid = t.vw * min((uint32_t)kPathTileH, C.h - t.y0);
    return t;
}

struct PoolLane {
    PathState p;
    uint32_t pixel = 0, lp = 0, s = 0, bounce = 0;    // image pixel index, pixel within the tile, sample, bounce
    bool has_path = false;
    uint32_t rays = 0;
};

// (The 64-bit shared-memory atomicAdd compiles to a compare-and-swap loop, ATOMS.CAST.SPIN.64.  Splitting it into two native
//  32-bit atomics with an exact carry -- the thread whose addition wrapped the low word carries into the high one -- was
//  measured 2-5 % SLOWER on all six scenes in a same-box A/B, so the plain form stays.)
__device__ __forceinline__ void pool_contribute(TilePool &tp, uint32_t lp, float r, float g, float b) {
    if (r != 0.0f) atomicAdd(reinterpret_cast<unsigned long long *>(&tp.acc[lp * 3 + 0]), (unsigned long long)__float2ll_rn(r * kFixScale));
    if (g != 0.0f) atomicAdd(reinterpret_cast<unsigned long long *>(&tp.acc[lp * 3 + 1]), (unsigned long long)__float2ll_rn(g * kFixScale));
    if (b != 0.0f) atomicAdd(reinterpret_cast<unsigned long long *>(&tp.acc[lp * 3 + 2]), (unsigned long long)__float2ll_rn(b * kFixScale));
}

// Takes the tile's next (pixel, sample) item into L (false: none left).  Item order: group_shift = 0: sample-major
// (consecutive items = neighbouring pixels of one sample).  group_shift = g: consecutive items = 2^g samples of the SAME
// pixel, so the primary rays a warp starts together are near-identical and traverse in lockstep: C2 +3 %, C4 +5 %,
// C4b +15 %, C5 +5 % -- when the background is black.  With a bright background every escaping path adds to its pixel's
// accumulator and lanes on the same pixel serialise on it (C1 -3 %; warp-aggregating the sums first costs more than it
// saves), so the host picks g = 3 for a black background and 0 otherwise (profiles/r2_tile_shape.md).  The image is the same
// bit for bit under every order.
__device__ __forceinline__ bool pool_fetch(const RenderParams &P, const TileMap &tm, TilePool &tp, uint32_t total_items, PoolLane &L,
                                           uint32_t &px, uint32_t &py) {
    if (total_items == 0u) return false;
    const uint32_t gs = P.group_shift;
    if (gs == 0u) {
        const uint32_t item = atomicAdd(&tp.next, 1u);
        if (item >= total_items) return false;
        if (tm.n_valid == (uint32_t)kPathBlock) { L.s = item / (uint32_t)kPathBlock; L.lp = item % (uint32_t)kPathBlock; }
        else { L.s = item / tm.n_valid; L.lp = item - L.s * tm.n_valid; }
    } else {
        const uint32_t per_block = tm.n_valid << gs;                             // items per block of 2^g samples
        const uint32_t limit = per_block * ((P.sample_count + (1u << gs) - 1u) >> gs);
        while (true) {
            const uint32_t item = atomicAdd(&tp.next, 1u);
            if (item >= limit) return false;
            const uint32_t s_hi = item / per_block, rem = item - s_hi * per_block;
            L.lp = rem >> gs;
            L.s = (s_hi << gs) + (rem & ((1u << gs) - 1u));
            if (L.s < P.sample_count) break;                                      // ragged last block of samples
        }
    }
    const uint32_t lx = tm.vw == (uint32_t)kPathTileW ? L.lp % (uint32_t)kPathTileW : L.lp % tm.vw;
    const uint32_t ly = tm.vw == (uint32_t)kPathTileW ? L.lp / (uint32_t)kPathTileW : L.lp / tm.vw;
    px = tm.x0 + lx;
    py = tm.y0 + ly;
    L.pixel = py * P.cam.w + px;
    L.bounce = 0;
    return true;
}

// shade_and_advance with the tile pool: same path semantics, same Philox keys (pixel, sample, bounce).
__device__ __forceinline__ bool pool_shade_and_advance(const RenderParams &P, const TileMap &tm, TilePool &tp, uint32_t total_items,
                                                       const Hit &best, PoolLane &L, Ray &ray) {
    const CameraParams &C = P.cam;
    const uint32_t k0 = (uint32_t)P.seed, k1 = (uint32_t)(P.seed >> 32);
    if (L.has_path) {
        ++L.rays;
        bool cont;
        float er = 0.0f, eg = 0.0f, eb = 0.0f;
        if (best.ref == kNoHit) {
            er = L.p.tr * C.background[0]; eg = L.p.tg * C.background[1]; eb = L.p.tb * C.background[2];
            cont = false;   // camera.h:248
        } else {
            const Philox4 rnd = philox4x32_10(L.pixel, P.sample_begin + L.s, L.bounce + 1u, 0u, k0, k1);
            cont = shade_hit(P.scene, best, rnd, ray, L.p, er, eg, eb);
            // ray_color(scattered, depth_left - 1): contributes nothing once depth_left hits 0 (camera.h:211-213)
            if (cont && ++L.bounce == C.max_depth) cont = false;
        }
        pool_contribute(tp, L.lp, er, eg, eb);
        if (!cont) L.has_path = false;
    }
    if (!L.has_path) {
        uint32_t px, py;
        if (!pool_fetch(P, tm, tp, total_items, L, px, py)) return false;
        const Philox4 rnd = philox4x32_10(L.pixel, P.sample_begin + L.s, 0u, 0u, k0, k1);
        camera_ray(C, px, py, rnd, ray, L.p);
        L.has_path = true;
    }
    return true;
}

__device__ __forceinline__ void pool_init(TilePool &tp) {
    if (threadIdx.x == 0) tp.next = 0u;
    for (int i = threadIdx.x; i < kPathBlock * 3; i += kPathBlock) tp.acc[i] = 0;
    __syncthreads();
}

// After the block's last path: thread t writes tile pixel t.
__device__ __forceinline__ void pool_write_tile(const RenderParams &P, const TileMap &tm, TilePool &tp) {
    __syncthreads();
    const uint32_t lp = threadIdx.x;
    if (lp < tm.n_valid) {
        const uint32_t px = tm.x0 + lp % tm.vw, py = tm.y0 + lp / tm.vw;
        const uint32_t pixel = py * P.cam.w + px;
        if (B200RT_CHECK(P.scene, pixel < P.cam.w * P.cam.h, 3)) {
            const float sum_r = (float)((double)tp.acc[lp * 3 + 0] * kFixInv), sum_g = (float)((double)tp.acc[lp * 3 + 1] * kFixInv),
                        sum_b = (float)((double)tp.acc[lp * 3 + 2] * kFixInv);
            float *o = P.out + (size_t)pixel * 3;
            const float k = P.scale;
            if (P.flags & kRenderAccumulate) { o[0] += sum_r * k; o[1] += sum_g * k; o[2] += sum_b * k; }
            else { o[0] = sum_r * k; o[1] = sum_g * k; o[2] = sum_b * k; }
        }
    }
}

__device__ __forceinline__ void write_counters(const RenderParams &P, uint32_t lane_rays, bool count, const TraversalCounters &ctr) {
    unsigned long long rays = lane_rays;
    for (int off = 16; off; off >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, off);
    if ((threadIdx.x & 31) == 0 && rays) atomicAdd(&P.counters[0], rays);
    if (count) {
        unsigned long long a = ctr.nodes, b = ctr.prims, c = ctr.quads;
        for (int off = 16; off; off >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, off);
            b += __shfl_down_sync(0xffffffffu, b, off);
            c += __shfl_down_sync(0xffffffffu, c, off);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(&P.counters[1], a); atomicAdd(&P.counters[2], b); atomicAdd(&P.counters[3], c); }
    }
}

// path_megakernel_pool (variant 0, default): every lane runs traverse-then-shade in a loop and takes the tile's next
// (pixel, sample) item as soon as its path ends.
template <int STACK, bool COUNT, bool LEAF_FIRST = false, int MINB = kPathMinBlocks>
__global__ void __launch_bounds__(kPathBlock, MINB) path_megakernel_pool(const __grid_constant__ RenderParams P) {
    __shared__ TilePool tp;
    const CameraParams &C = P.cam;
    const TileMap tm = map_tile(C);
    pool_init(tp);
    const uint32_t total_items = C.max_depth > 0 ? tm.n_valid * P.sample_count : 0u;
    PoolLane L;
    Ray ray;
    TraversalCounters ctr;
    Hit best{0.0, kNoHit};
    while (pool_shade_and_advance(P, tm, tp, total_items, best, L, ray)) {
        // world.hit_by(ray, Interval::with_min(0.00001))  (camera.h:217)
        best = closest_hit<STACK, COUNT, LEAF_FIRST>(P.scene, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, 0.00001,
                                                  __longlong_as_double(0x7ff0000000000000LL), &ctr);
    }
    pool_write_tile(P, tm, tp);
    write_counters(P, L.rays, COUNT, ctr);
}

// ------------------------------------------------------------------------------------------
// path_megakernel (thread-owns-pixel form, kept for A/B: B200RT_FLAG_THREAD_PIXELS): every lane runs traverse-then-shade
// in a loop over ITS pixel's samples and starts the next sample as soon as a path ends.  64 registers (16 blocks = 32 warps per SM) measured
// fastest on B200: 8/12/16/20/24 blocks per SM gave 2707/3171/3263/2630/2307 Mpaths/s on C2.
template <int STACK, bool COUNT, bool LEAF_FIRST = false, int MINB = kPathMinBlocks>
__global__ void __launch_bounds__(kPathBlock, MINB) path_megakernel(const __grid_constant__ RenderParams P) {
    const CameraParams &C = P.cam;
    const PixelMap m = map_pixel(C);
    LaneState L;
    TraversalCounters ctr;
    Hit best{0.0, kNoHit};
    if (m.valid && C.max_depth > 0 && P.sample_count > 0) {
        while (shade_and_advance(P, m, best, L)) {
            // world.hit_by(ray, Interval::with_min(0.00001))  (camera.h:217)
            best = closest_hit<STACK, COUNT, LEAF_FIRST>(P.scene, L.ray.ox, L.ray.oy, L.ray.oz, L.ray.dx, L.ray.dy, L.ray.dz, 0.00001,
                                             __longlong_as_double(0x7ff0000000000000LL), &ctr);
        }
    }
    write_pixel_and_counters(P, m, L.sum_r, L.sum_g, L.sum_b, L.rays, COUNT, ctr);
}

// ------------------------------------------------------------------------------------------
// path_lanes_kernel: MEASUREMENT ONLY (b200rt_debug_lane_accounting).  The default kernel's schedule -- per
// round every lane shades / regenerates, then the warp iterates "node step for the lanes at a node, leaf step for the
// lanes at a leaf" (or the reverse order, as the launcher picks) until its slowest lane is done -- replayed warp-synchronously so that each warp can count, per executed step,
// how many of its 32 lanes took part and what the others were doing:
//   acc[0] shade executions           acc[1] lanes taking part
//   acc[2] node-step executions       acc[3] lanes at a node     acc[4] idle: at a leaf   acc[5] idle: traversal done,
//                                                                waiting for the round    acc[6] idle: pixel out of samples
//   acc[7] leaf-step executions       acc[8] lanes at a leaf     acc[9] idle: at a node   acc[10] idle: done   acc[11] idle: finished
//   acc[12] primitive-test executions (a leaf step loops over its primitives)            acc[13] lanes taking part
//   acc[14] warps   acc[15] lane-rounds (shade calls that started a ray)
// Same Philox keys and the same arithmetic as path_megakernel, so it also writes the same image.
template <int STACK, bool POOL>
__global__ void __launch_bounds__(kPathBlock) path_lanes_kernel(const __grid_constant__ RenderParams P, unsigned long long *__restrict__ acc, bool leaf_first) {
    __shared__ TilePool tp;
    const CameraParams &C = P.cam;
    const PixelMap m = map_pixel(C);
    const TileMap tm = map_tile(C);
    if (POOL) pool_init(tp);
    const uint32_t total_items = C.max_depth > 0 ? tm.n_valid * P.sample_count : 0u;
    LaneState L;
    PoolLane Lp;
    Trav T;
    uint2 stack[STACK];
    T.cur = kTravDone;
    T.best.t = 0.0; T.best.ref = kNoHit;
    bool finished = POOL ? false : !(m.valid && C.max_depth > 0 && P.sample_count > 0);
    unsigned long long a[16] = {};
    const unsigned full = 0xffffffffu;
    while (true) {
        const unsigned ms = __ballot_sync(full, !finished);
        if (!ms) break;
        a[0]++; a[1] += __popc(ms);
        if (!finished) {
            bool go;
            if (POOL) {
                go = pool_shade_and_advance(P, tm, tp, total_items, T.best, Lp, T.ray);
                if (go) trav_setup(T, 0.00001, __longlong_as_double(0x7ff0000000000000LL));
            } else {
                go = shade_and_advance(P, m, T.best, L);
                if (go) trav_init(T, L.ray.ox, L.ray.oy, L.ray.oz, L.ray.dx, L.ray.dy, L.ray.dz, 0.00001, __longlong_as_double(0x7ff0000000000000LL));
            }
            if (go) a[15]++;
            else { finished = true; T.cur = kTravDone; }
        }
        while (true) {   // one iteration of closest_hit's loop: both kinds of step, in the product's order
            if (!__ballot_sync(full, !finished && !trav_done(T))) break;
#pragma unroll
            for (int phase = 0; phase < 2; ++phase) {
                const bool node_phase = (phase == 0) != leaf_first;
                const bool node = !finished && trav_at_node(T), leaf = !finished && trav_at_leaf(T);
                const unsigned mn = __ballot_sync(full, node), ml = __ballot_sync(full, leaf), mf = __ballot_sync(full, finished);
                const unsigned md = ~(mn | ml | mf);
                if (node_phase) {
                    if (mn) { a[2]++; a[3] += __popc(mn); a[4] += __popc(ml); a[5] += __popc(md); a[6] += __popc(mf); }
                    if (node) trav_node_step(P.scene, T, stack);
                } else {
                    if (ml) {
                        a[7]++; a[8] += __popc(ml); a[9] += __popc(mn); a[10] += __popc(md); a[11] += __popc(mf);
                        const uint32_t cnt = leaf ? ((T.cur >> 26) & 0xFu) : 0u;
                        for (uint32_t i = 0; i < 8; ++i) {
                            const unsigned mi = __ballot_sync(full, cnt > i);
                            if (!mi) break;
                            a[12]++; a[13] += __popc(mi);
                        }
                    }
                    if (leaf) trav_leaf_step(P.scene, T, stack);
                }
            }
        }
    }
    TraversalCounters ctr;
    if (POOL) { pool_write_tile(P, tm, tp); write_counters(P, Lp.rays, false, ctr); }
    else write_pixel_and_counters(P, m, L.sum_r, L.sum_g, L.sum_b, L.rays, false, ctr);
    a[14] = 1;
    unsigned long long rounds = a[15];
    for (int off = 16; off; off >>= 1) rounds += __shfl_down_sync(full, rounds, off);
    a[15] = rounds;
    if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 16; ++i) atomicAdd(&acc[i], a[i]);
}

// ------------------------------------------------------------------------------------------
__global__ void tonemap_kernel(const float *__restrict__ hdr, long long n_pixels, int32_t *__restrict__ out, int clamp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const double r = hdr[i * 3 + 0], g = hdr[i * 3 + 1], b = hdr[i * 3 + 2];
    const double L = 0.2126 * r + 0.7152 * g + 0.0722 * b;          // rgb.h:28-30
    const double scale = 255 + 0.999999;                             // rgb.h:106
    const double v[3] = {r / (1 + L), g / (1 + L), b / (1 + L)};     // rgb.h:99-104
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int q = (int)(scale * sqrt(v[c]));                           // pow(v, 1/2), truncation (rgb.h:107-111)
        if (clamp) q = q < 0 ? 0 : (q > 255 ? 255 : q);
        out[i * 3 + c] = q;
    }
}

// Frame epilogue on the reducing rank: sum -> mean in place (pixel_color /= spp, camera.h:290) and,
// optionally, the tone-mapped integers of the same pixels in the same pass.
__global__ void finalize_kernel(float *__restrict__ frame, long long n_pixels, float scale, int32_t *__restrict__ ldr, int clamp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    const float rf = frame[i * 3 + 0] * scale, gf = frame[i * 3 + 1] * scale, bf = frame[i * 3 + 2] * scale;
    frame[i * 3 + 0] = rf; frame[i * 3 + 1] = gf; frame[i * 3 + 2] = bf;
    if (ldr) {
        const double r = rf, g = gf, b = bf;
        const double L = 0.2126 * r + 0.7152 * g + 0.0722 * b;
        const double s255 = 255 + 0.999999;
        const double v[3] = {r / (1 + L), g / (1 + L), b / (1 + L)};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int q = (int)(s255 * sqrt(v[c]));
            if (clamp) q = q < 0 ? 0 : (q > 255 ? 255 : q);
            ldr[i * 3 + c] = q;
        }
    }
}

// Multi-GPU frame exchange as ONE kernel over peer memory (NVLink / NVSwitch): this rank's slice of
// "sum the per-rank sample sums, /= spp (camera.h:290), tone-map (rgb.h:90-113)" with the result
// written straight into the root rank's buffers.  Every rank runs it on a different slice, so the
// N frames cross the links once, spread over all of them (reduce-scatter + epilogue + gather in one
// pass), instead of funnelling into rank 0.  Peers are added in rank order 0..N-1 whichever rank owns
// the slice, so the result does not depend on the slicing.
//   Block = kPeerTilePx pixels: coalesced float4 loads from each peer -> sum -> scale -> coalesced
//   float4 store to the root HDR frame (may alias peer 0's buffer: each element is read and written by
//   the same thread) -> shared memory -> per-pixel tone map -> coalesced store of the integers.
constexpr int kPeerTilePx = 1024, kPeerThreads = 256;
__global__ void __launch_bounds__(kPeerThreads) reduce_finalize_peers_kernel(const __grid_constant__ PeerFrames in, int n_peers, long long px_lo,
                                                                             long long px_hi, float scale, float *root_hdr,
                                                                             int32_t *root_ldr, int clamp) {
    __shared__ __align__(16) float tile[kPeerTilePx * 3];
    const long long tile_px = px_lo + (long long)blockIdx.x * kPeerTilePx;
    const long long base = tile_px * 3;                                                     // first float of the tile
    const long long n_fl = ((px_hi - tile_px < kPeerTilePx) ? (px_hi - tile_px) : kPeerTilePx) * 3;  // floats in this tile
#pragma unroll
    for (int k = 0; k < kPeerTilePx * 3 / 4 / kPeerThreads; ++k) {
        const int j = (threadIdx.x + k * kPeerThreads) * 4;
        if (j + 3 < n_fl) {
            float4 a = *reinterpret_cast<const float4 *>(in.p[0] + base + j);
            for (int r = 1; r < n_peers; ++r) {
                const float4 b = *reinterpret_cast<const float4 *>(in.p[r] + base + j);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
            *reinterpret_cast<float4 *>(root_hdr + base + j) = a;
            *reinterpret_cast<float4 *>(tile + j) = a;
        } else {
            for (int e = j; e < n_fl && e < j + 4; ++e) {                                    // ragged end of the frame
                float a = in.p[0][base + e];
                for (int r = 1; r < n_peers; ++r) a += in.p[r][base + e];
                a *= scale;
                root_hdr[base + e] = a;
                tile[e] = a;
            }
        }
    }
    if (!root_ldr) return;
    __syncthreads();
    int q[kPeerTilePx / kPeerThreads][3];
#pragma unroll
    for (int k = 0; k < kPeerTilePx / kPeerThreads; ++k) {
        const int px = threadIdx.x + k * kPeerThreads;
        const double r = tile[px * 3 + 0], g = tile[px * 3 + 1], b = tile[px * 3 + 2];
        const double L = 0.2126 * r + 0.7152 * g + 0.0722 * b;
        const double s255 = 255 + 0.999999;
        const double v[3] = {r / (1 + L), g / (1 + L), b / (1 + L)};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int t = (px * 3 < n_fl) ? (int)(s255 * sqrt(v[c])) : 0;
            if (clamp) t = t < 0 ? 0 : (t > 255 ? 255 : t);
            q[k][c] = t;
        }
    }
    __syncthreads();
    int32_t *itile = reinterpret_cast<int32_t *>(tile);
#pragma unroll
    for (int k = 0; k < kPeerTilePx / kPeerThreads; ++k) {
        const int px = threadIdx.x + k * kPeerThreads;
        itile[px * 3 + 0] = q[k][0]; itile[px * 3 + 1] = q[k][1]; itile[px * 3 + 2] = q[k][2];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPeerTilePx * 3 / 4 / kPeerThreads; ++k) {
        const int j = (threadIdx.x + k * kPeerThreads) * 4;
        if (j + 3 < n_fl) *reinterpret_cast<int4 *>(root_ldr + base + j) = *reinterpret_cast<const int4 *>(itile + j);
        else for (int e = j; e < n_fl && e < j + 4; ++e) root_ldr[base + e] = itile[e];
    }
}

// ------------------------------------------------------------------------------------------
template <int STACK>
static cudaError_t launch_raycast_t(const DeviceScene &S, const double *rays, long long n, double tmin, double tmax,
                                    int32_t *prim, double *t, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int block = 128;
    raycast_kernel<STACK><<<(unsigned)((n + block - 1) / block), block, 0, st>>>(S, rays, n, tmin, tmax, prim, t);
    return cudaGetLastError();
}

cudaError_t launch_raycast(int stack, const DeviceScene &S, const double *rays, long long n, double tmin, double tmax,
                           int32_t *prim, double *t, cudaStream_t st) {
    if (stack <= 32) return launch_raycast_t<32>(S, rays, n, tmin, tmax, prim, t, st);
    if (stack <= 64) return launch_raycast_t<64>(S, rays, n, tmin, tmax, prim, t, st);
    return launch_raycast_t<128>(S, rays, n, tmin, tmax, prim, t, st);
}

cudaError_t launch_debug_camera(const CameraParams &C, const uint32_t *pixels, const uint32_t *rnd, long long n, double *rays_out,
                                cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    debug_camera_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(C, pixels, rnd, n, rays_out);
    return cudaGetLastError();
}

cudaError_t launch_debug_philox(const uint32_t *in, long long n, uint32_t *out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    debug_philox_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(in, n, out);
    return cudaGetLastError();
}
cudaError_t launch_debug_samplers(const uint32_t *rnd, long long n, double *sphere_out, double *disk_out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    debug_samplers_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(rnd, n, sphere_out, disk_out);
    return cudaGetLastError();
}

cudaError_t launch_debug_shade(int stack, const DeviceScene &S, const double *rays, const uint32_t *rnd, long long n, double tmin,
                               double tmax, void *records, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 127) / 128);
    ShadeRecord *out = static_cast<ShadeRecord *>(records);
    if (stack <= 32) debug_shade_kernel<32><<<blocks, 128, 0, st>>>(S, rays, rnd, n, tmin, tmax, out);
    else if (stack <= 64) debug_shade_kernel<64><<<blocks, 128, 0, st>>>(S, rays, rnd, n, tmin, tmax, out);
    else debug_shade_kernel<128><<<blocks, 128, 0, st>>>(S, rays, rnd, n, tmin, tmax, out);
    return cudaGetLastError();
}

template <int STACK>
static cudaError_t launch_path_t(const RenderParams &P, bool count, bool leaf_first, cudaStream_t st) {
    const uint32_t tiles = ((P.cam.w + (uint32_t)kPathTileW - 1u) / (uint32_t)kPathTileW) * ((P.cam.h + (uint32_t)kPathTileH - 1u) / (uint32_t)kPathTileH);
    if (tiles == 0) return cudaSuccess;
    if (!(P.flags & kRenderThreadPixels)) {
        if (leaf_first) {
            if (count) path_megakernel_pool<STACK, true, true><<<tiles, kPathBlock, 0, st>>>(P);
            else path_megakernel_pool<STACK, false, true><<<tiles, kPathBlock, 0, st>>>(P);
        } else if (count) path_megakernel_pool<STACK, true><<<tiles, kPathBlock, 0, st>>>(P);
        else path_megakernel_pool<STACK, false><<<tiles, kPathBlock, 0, st>>>(P);
        return cudaGetLastError();
    }
    if (count) path_megakernel<STACK, true><<<tiles, kPathBlock, 0, st>>>(P);      // round-1 work distribution: node-then-leaf only
    else path_megakernel<STACK, false><<<tiles, kPathBlock, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_path_megakernel(int stack, const RenderParams &P, bool count, bool leaf_first, cudaStream_t st) {
    if (stack <= 32) return launch_path_t<32>(P, count, leaf_first, st);
    if (stack <= 64) return launch_path_t<64>(P, count, leaf_first, st);
    return launch_path_t<128>(P, count, leaf_first, st);
}

cudaError_t launch_path_lanes(int stack, const RenderParams &P, unsigned long long *acc, bool leaf_first, cudaStream_t st) {
    const uint32_t tiles = ((P.cam.w + (uint32_t)kPathTileW - 1u) / (uint32_t)kPathTileW) * ((P.cam.h + (uint32_t)kPathTileH - 1u) / (uint32_t)kPathTileH);
    if (tiles == 0) return cudaSuccess;
    const bool pool = !(P.flags & kRenderThreadPixels);
    if (stack <= 32) { if (pool) path_lanes_kernel<32, true><<<tiles, kPathBlock, 0, st>>>(P, acc, leaf_first); else path_lanes_kernel<32, false><<<tiles, kPathBlock, 0, st>>>(P, acc, leaf_first); }
    else if (stack <= 64) { if (pool) path_lanes_kernel<64, true><<<tiles, kPathBlock, 0, st>>>(P, acc, leaf_first); else path_lanes_kernel<64, false><<<tiles, kPathBlock, 0, st>>>(P, acc, leaf_first); }
    else { if (pool) path_lanes_kernel<128, true><<<tiles, kPathBlock, 0, st>>>(P, acc, leaf_first); else path_lanes_kernel<128, false><<<tiles, kPathBlock, 0, st>>>(P, acc, leaf_first); }
    return cudaGetLastError();
}

cudaError_t launch_tonemap(const float *hdr, long long n_pixels, int32_t *out, int clamp, cudaStream_t st) {
    if (n_pixels == 0) return cudaSuccess;
    tonemap_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, st>>>(hdr, n_pixels, out, clamp);
    return cudaGetLastError();
}

cudaError_t launch_finalize(float *frame, long long n_pixels, float scale, int32_t *ldr, int clamp, cudaStream_t st) {
    if (n_pixels == 0) return cudaSuccess;
    finalize_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, st>>>(frame, n_pixels, scale, ldr, clamp);
    return cudaGetLastError();
}

__global__ void add_frame_kernel(float *__restrict__ frame, const float *__restrict__ other, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) frame[i] += other[i];
}
cudaError_t launch_add_frame(float *frame, const float *other, long long n_floats, cudaStream_t st) {
    if (n_floats == 0) return cudaSuccess;
    add_frame_kernel<<<(unsigned)((n_floats + 255) / 256), 256, 0, st>>>(frame, other, n_floats);
    return cudaGetLastError();
}

void peer_slice(long long n_pixels, int n_peers, int rank, long long *lo, long long *hi) {
    long long chunk = (n_pixels + n_peers - 1) / n_peers;
    chunk = (chunk + kPeerTilePx - 1) / kPeerTilePx * kPeerTilePx;        // whole tiles: slices start 16-byte aligned
    const long long a = chunk * rank, b = a + chunk;
    *lo = a < n_pixels ? a : n_pixels;
    *hi = b < n_pixels ? b : n_pixels;
}

cudaError_t launch_reduce_finalize_peers(const PeerFrames &in, int n_peers, int rank, long long n_pixels, float scale,
                                         float *root_hdr, int32_t *root_ldr, int clamp, cudaStream_t st) {
    long long lo, hi;
    peer_slice(n_pixels, n_peers, rank, &lo, &hi);
    if (hi <= lo) return cudaSuccess;
    const unsigned blocks = (unsigned)((hi - lo + kPeerTilePx - 1) / kPeerTilePx);
    reduce_finalize_peers_kernel<<<blocks, kPeerThreads, 0, st>>>(in, n_peers, lo, hi, scale, root_hdr, root_ldr, clamp);
    return cudaGetLastError();
}

}  // namespace b200rt
