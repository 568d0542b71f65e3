// kernels.h -- launch interface between the C ABI (api.cu) and the kernels (kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "shade.cuh"

namespace b200rt {

constexpr int kPathBlock = 64;
constexpr uint32_t kRenderAccumulate = 2u;   // == B200RT_FLAG_ACCUMULATE

struct RenderParams {
    DeviceScene scene;
    CameraParams cam;
    unsigned long long seed;
    uint32_t sample_begin, sample_count;
    float *out;                        // image_h x image_w x 3
    float scale;                       // 1 (sum) or 1/sample_count (mean)
    uint32_t flags;
    unsigned long long *counters;      // [0] rays, [1] node visits, [2] primitive tests
};

// `stack` = traversal stack entries the scene needs (3 x depth of the 4-wide tree); the
// launchers pick the smallest instantiation that fits (32 / 64 / 128).
cudaError_t launch_raycast(int stack, const DeviceScene &S, const double *rays, long long n, double tmin, double tmax,
                           int32_t *prim, double *t, cudaStream_t st);
cudaError_t launch_path_megakernel(int stack, const RenderParams &P, bool count, bool voted, cudaStream_t st);
cudaError_t launch_finalize(float *frame, long long n_pixels, float scale, int32_t *ldr, int clamp, cudaStream_t st);
cudaError_t launch_tonemap(const float *hdr, long long n_pixels, int32_t *out, int clamp, cudaStream_t st);

}  // namespace b200rt
