// devmem.h -- device memory and host<->device staging shared by the C ABI (api.cu) and the GPU builder (lbvh.cu).
//
//  * Device memory comes from a PRIVATE stream-ordered pool per device (cudaMemPoolCreate), so that the
//    create / render / destroy cycle of the drop-in call (Camera::render(const Scene&), one per frame) reuses
//    memory instead of paying cudaMalloc / cudaFree and their device-wide synchronisation each time -- without
//    touching the attributes of the process-wide default pool, which belongs to the host application.
//  * The caller's arrays (pageable memory) reach the device, and frames come back, through a small set of
//    pinned staging buffers: chunk k is copied into pinned memory by a few host threads while chunk k-1 is in
//    flight on the copy engine (a blocking cudaMemcpy from pageable memory stages through ONE driver thread and
//    was ~50 of the 57-90 ms it took to make the multi-million-primitive scenes resident).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace b200rt {

constexpr int kMaxDevices = 64;

// Allocation on the CURRENT device, ordered on `st`: usable by work enqueued on `st` afterwards; any other stream
// must first wait for an event recorded on `st` (or the caller synchronises `st`).
cudaError_t dev_alloc_async(void **p, size_t bytes, cudaStream_t st);
// Allocation + cudaStreamSynchronize(0): usable from any stream on return.
cudaError_t dev_alloc(void **p, size_t bytes);
template <typename T>
cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc(reinterpret_cast<void **>(p), bytes); }
template <typename T>
cudaError_t dev_alloc_async(T **p, size_t bytes, cudaStream_t st) { return dev_alloc_async(reinterpret_cast<void **>(p), bytes, st); }
void dev_free(void *p);                        // cudaFreeAsync on the legacy default stream of the CURRENT device
void dev_free_on(void *p, cudaStream_t st);
cudaError_t dev_pool_trim_all();

// Buffers that OTHER devices touch -- the frames the fused multi-GPU exchange dereferences, and the scene arrays that
// are copied device to device -- are plain cudaMalloc memory, which cudaDeviceEnablePeerAccess maps into every peer
// (stream-ordered pool memory is not mapped unless the whole pool is opened with cudaMemPoolSetAccess, and then peer
// copies of it fall back to staging otherwise).  They are cached per device across calls, because the one-call render
// entry creates and destroys its scene every frame; b200rt_trim() frees the idle ones.  `dev` must be the current device.
cudaError_t peer_buffer_acquire(int dev, size_t bytes, void **out);
void peer_buffer_release(int dev, void *p);
void peer_buffer_trim();

// Pageable host memory -> device, pipelined through pinned staging chunks; returns after the LAST chunk has been
// enqueued on `st` (the host source has been fully read by then; the device copy completes in stream order).
cudaError_t staged_upload(void *dst_device, const void *src_host, size_t bytes, cudaStream_t st);
// Device -> pageable host memory, pipelined the same way; returns when dst_host holds the data.
cudaError_t staged_download(void *dst_host, const void *src_device, size_t bytes, cudaStream_t st);

}  // namespace b200rt
