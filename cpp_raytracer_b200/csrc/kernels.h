// kernels.h -- launch interface between the C ABI (api.cu) and the kernels (kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "shade.cuh"

namespace b200rt {

// Threads per block of the path kernels = pixels per tile of the work pool.  Measured with the pool (B200, 128 spp,
// gpurun_out/r2_variants*.txt -> profiles/r2_tile_shape.md): 8 x 8 / 8 x 16 / 8 x 32 / 8 x 64 / 8 x 128 tiles give
// C2 4020 / 4104 / 4135 / 4134 / 4171 and C4b 3243 / 3384 / 3580 / 3576 / 3629 Mpaths/s; 16- and 32-wide tiles are
// slower than 8-wide ones of the same size (the 32 consecutive items a warp holds form an 8 x 4 pixel patch).
#ifndef B200RT_PATH_BLOCK
#define B200RT_PATH_BLOCK 256
#endif
#ifndef B200RT_TILE_W
#define B200RT_TILE_W 8
#endif
constexpr int kPathBlock = B200RT_PATH_BLOCK;       // threads per block of the path kernels = pixels per tile
constexpr int kPathTileW = B200RT_TILE_W;           // tile = kPathTileW x (kPathBlock / kPathTileW) pixels
constexpr int kPathTileH = kPathBlock / kPathTileW;
static_assert(kPathBlock % kPathTileW == 0 && kPathBlock % 32 == 0, "tile shape");
#ifdef B200RT_PATH_MINB
constexpr int kPathMinBlocks = B200RT_PATH_MINB;    // occupancy experiments
#else
constexpr int kPathMinBlocks = 1024 / kPathBlock;   // 1024 threads = 32 warps per SM at 64 registers
#endif
constexpr uint32_t kRenderAccumulate = 2u;    // == B200RT_FLAG_ACCUMULATE
constexpr uint32_t kRenderThreadPixels = 16u; // == B200RT_FLAG_THREAD_PIXELS
constexpr uint32_t kMaxSamplesPerLaunch = 1u << 22;   // tile items (up to 1024 pixels x samples, + one group) are counted in 32 bits

struct RenderParams {
    DeviceScene scene;
    CameraParams cam;
    unsigned long long seed;
    uint32_t sample_begin, sample_count;
    float *out;                        // image_h x image_w x 3
    float scale;                       // 1 (sum) or 1/sample_count (mean)
    uint32_t flags;
    unsigned long long *counters;      // [0] rays, [1] node visits, [2] primitive tests, [3] of which quad tests
    uint32_t group_shift;              // work-pool item order: 2^group_shift consecutive items are samples of one pixel
};

// Path-slot pool of the wavefront variant (wavefront.cu); all pointers are device memory.
// One slot = one 128-byte line, read and written with 16-byte vector accesses.
struct alignas(128) PathSlot {
    double ox, oy, oz, dx, dy, dz;   //  0: the ray (direction not normalised)
    double hit_t;                    // 48: closest hit of the last trace ...
    uint32_t hit_ref;                // 56: ... and its primitive reference (kNoHit on a miss)
    uint32_t pixel;                  // 60: pixel index; 0xFFFFFFFF = dead slot
    float tr, tg, tb;                // 64: path throughput
    uint32_t sample;                 // 76
    uint32_t bounce;                 // 80
    uint32_t cls;                    // 84: material class of the last hit (queue it was sorted into)
    uint32_t pad[10];
};
static_assert(sizeof(PathSlot) == 128, "PathSlot must be one 128-byte line");

struct WavefrontPool {
    uint32_t n_slots;
    unsigned long long total_items;     // image_w * image_h * samples of this call
    PathSlot *slots;
    uint32_t *queue;                    // 5 class queues of n_slots entries each
    uint32_t *queue_count;              // [5]
    uint32_t *alive;
    unsigned long long *next_item;
};
size_t wavefront_pool_alloc_bytes(uint32_t n_slots);
void wavefront_pool_layout(void *base, uint32_t n_slots, WavefrontPool &W);
cudaError_t run_wavefront(int stack, const RenderParams &P, const WavefrontPool &W, bool count, cudaStream_t st, int sm_count,
                          unsigned long long *launches);

// `stack` = traversal stack entries the scene needs (3 x depth of the 4-wide tree); the
// launchers pick the smallest instantiation that fits (32 / 64 / 128).
cudaError_t launch_raycast(int stack, const DeviceScene &S, const double *rays, long long n, double tmin, double tmax,
                           int32_t *prim, double *t, cudaStream_t st);
// leaf_first: order of the two steps inside one iteration of the traversal loop (see closest_hit)
cudaError_t launch_path_megakernel(int stack, const RenderParams &P, bool count, bool leaf_first, cudaStream_t st);
// measurement only: 16 lane-accounting counters (see path_lanes_kernel)
cudaError_t launch_path_lanes(int stack, const RenderParams &P, unsigned long long *acc, bool leaf_first, cudaStream_t st);
cudaError_t launch_finalize(float *frame, long long n_pixels, float scale, int32_t *ldr, int clamp, cudaStream_t st);
cudaError_t launch_tonemap(const float *hdr, long long n_pixels, int32_t *out, int clamp, cudaStream_t st);
cudaError_t launch_debug_camera(const CameraParams &C, const uint32_t *pixels, const uint32_t *rnd, long long n, double *rays_out,
                                cudaStream_t st);
cudaError_t launch_debug_philox(const uint32_t *in, long long n, uint32_t *out, cudaStream_t st);
cudaError_t launch_debug_samplers(const uint32_t *rnd, long long n, double *sphere_out, double *disk_out, cudaStream_t st);
// records: n x 88-byte ShadeRecord (kernels.cu)
cudaError_t launch_debug_shade(int stack, const DeviceScene &S, const double *rays, const uint32_t *rnd, long long n, double tmin,
                               double tmax, void *records, cudaStream_t st);

// Peer-memory frame exchange (multi-GPU): per-rank sum buffers addressed from this GPU.
constexpr int kMaxPeers = 16;
struct PeerFrames { const float *p[kMaxPeers]; };
void peer_slice(long long n_pixels, int n_peers, int rank, long long *lo, long long *hi);
// frame += other (multi-GPU exchange without peer mapping: the root adds the copied per-device frames in order)
cudaError_t launch_add_frame(float *frame, const float *other, long long n_floats, cudaStream_t st);
cudaError_t launch_reduce_finalize_peers(const PeerFrames &in, int n_peers, int rank, long long n_pixels, float scale,
                                         float *root_hdr, int32_t *root_ldr, int clamp, cudaStream_t st);

}  // namespace b200rt
