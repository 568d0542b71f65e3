// devmem.cu -- private per-device memory pools and pinned staging transfers (see devmem.h).
#include "devmem.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "thread_pool.h"

namespace b200rt {

namespace {

struct DevicePools {
    std::mutex mu;
    cudaMemPool_t pool[kMaxDevices] = {};
    bool ready[kMaxDevices] = {};
};
DevicePools &pools() {
    static DevicePools p;
    return p;
}

cudaError_t device_pool(int dev, cudaMemPool_t *out) {
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    DevicePools &P = pools();
    std::lock_guard<std::mutex> lk(P.mu);
    if (!P.ready[dev]) {
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaError_t e = cudaMemPoolCreate(&P.pool[dev], &props);
        if (e != cudaSuccess) return e;
        unsigned long long keep = 4096ull << 20;
        if (const char *env = std::getenv("B200RT_POOL_KEEP_MB")) keep = (unsigned long long)std::max(0ll, std::atoll(env)) << 20;
        cudaMemPoolSetAttribute(P.pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        P.ready[dev] = true;
    }
    *out = P.pool[dev];
    return cudaSuccess;
}

// ---- peer-shared frames --------------------------------------------------------------------
struct SharedFrame { void *p; size_t bytes; bool busy; };
struct SharedFrames {
    std::mutex mu;
    std::vector<SharedFrame> per_dev[kMaxDevices];
};
SharedFrames &shared_frames() {
    static SharedFrames f;
    return f;
}

// ---- pinned staging -------------------------------------------------------------------------
constexpr size_t kChunk = 8u << 20;   // bytes per staging buffer
constexpr int kBufs = 3;
constexpr size_t kDirectBelow = 256u << 10;   // tiny transfers: let the driver stage them

struct Staging {
    std::mutex mu;   // one pipelined transfer at a time (the buffers are shared by every scene of the process)
    char *buf[kBufs] = {};
    std::unique_ptr<ThreadPool> workers;
    cudaError_t ensure() {
        if (buf[0]) return cudaSuccess;
        for (int k = 0; k < kBufs; ++k) {
            cudaError_t e = cudaHostAlloc((void **)&buf[k], kChunk, cudaHostAllocPortable);
            if (e != cudaSuccess) {
                for (int j = 0; j < k; ++j) { cudaFreeHost(buf[j]); buf[j] = nullptr; }
                return e;
            }
        }
        workers = std::make_unique<ThreadPool>(std::min(8, ThreadPool::hardware_threads()));
        return cudaSuccess;
    }
    // memcpy split over the worker threads (1 MiB pieces)
    void copy(char *dst, const char *src, size_t n) {
        constexpr size_t piece = 1u << 20;
        const int pieces = (int)((n + piece - 1) / piece);
        if (pieces <= 1) { std::memcpy(dst, src, n); return; }
        workers->parallel_for(pieces, [&](int i) {
            const size_t lo = (size_t)i * piece, len = std::min(piece, n - lo);
            std::memcpy(dst + lo, src + lo, len);
        });
    }
};
Staging &staging() {
    static Staging s;
    return s;
}

struct Events {
    cudaEvent_t ev[kBufs] = {};
    cudaError_t create() {
        for (int k = 0; k < kBufs; ++k) {
            cudaError_t e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    ~Events() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
};

}  // namespace

cudaError_t dev_alloc_async(void **p, size_t bytes, cudaStream_t st) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaMemPool_t pool;
    e = device_pool(dev, &pool);
    if (e != cudaSuccess) return e;
    return cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, st);
}

cudaError_t dev_alloc(void **p, size_t bytes) {
    cudaError_t e = dev_alloc_async(p, bytes, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    return e;
}

void dev_free(void *p) {
    if (p) cudaFreeAsync(p, 0);
}
void dev_free_on(void *p, cudaStream_t st) {
    if (p) cudaFreeAsync(p, st);
}

cudaError_t dev_pool_trim_all() {
    DevicePools &P = pools();
    std::lock_guard<std::mutex> lk(P.mu);
    cudaError_t first = cudaSuccess;
    for (int d = 0; d < kMaxDevices; ++d)
        if (P.ready[d]) {
            cudaError_t e = cudaMemPoolTrimTo(P.pool[d], 0);
            if (e != cudaSuccess && first == cudaSuccess) first = e;
        }
    return first;
}

cudaError_t peer_buffer_acquire(int dev, size_t bytes, void **out) {
    if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
    SharedFrames &F = shared_frames();
    std::lock_guard<std::mutex> lk(F.mu);
    std::vector<SharedFrame> &v = F.per_dev[dev];
    int best = -1;
    for (int i = 0; i < (int)v.size(); ++i)
        if (!v[i].busy && v[i].bytes >= bytes && v[i].bytes <= 2 * bytes + (1u << 20) && (best < 0 || v[i].bytes < v[best].bytes)) best = i;
    if (best < 0) {
        // a miss means the working set changed (another scene, another resolution): keep at most 1 GiB of idle
        // buffers per device around, dropping the largest first
        size_t idle = 0;
        for (const SharedFrame &f : v) if (!f.busy) idle += f.bytes;
        while (idle > (1ull << 30)) {
            int big = -1;
            for (int i = 0; i < (int)v.size(); ++i)
                if (!v[i].busy && (big < 0 || v[i].bytes > v[big].bytes)) big = i;
            if (big < 0) break;
            idle -= v[big].bytes;
            cudaFree(v[big].p);
            v.erase(v.begin() + big);
        }
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaErrorMemoryAllocation) {   // make room: drop every idle buffer of this device, then retry once
            cudaGetLastError();
            for (int i = (int)v.size() - 1; i >= 0; --i)
                if (!v[i].busy) { cudaFree(v[i].p); v.erase(v.begin() + i); }
            e = cudaMalloc(&p, bytes ? bytes : 1);
        }
        if (e != cudaSuccess) return e;
        v.push_back(SharedFrame{p, bytes, false});
        best = (int)v.size() - 1;
    }
    v[best].busy = true;
    *out = v[best].p;
    return cudaSuccess;
}

void peer_buffer_release(int dev, void *p) {
    if (!p || dev < 0 || dev >= kMaxDevices) return;
    SharedFrames &F = shared_frames();
    std::lock_guard<std::mutex> lk(F.mu);
    for (SharedFrame &f : F.per_dev[dev])
        if (f.p == p) f.busy = false;
}

void peer_buffer_trim() {
    SharedFrames &F = shared_frames();
    std::lock_guard<std::mutex> lk(F.mu);
    int prev = -1;
    cudaGetDevice(&prev);
    for (int d = 0; d < kMaxDevices; ++d) {
        std::vector<SharedFrame> &v = F.per_dev[d];
        for (int i = (int)v.size() - 1; i >= 0; --i)
            if (!v[i].busy) {
                cudaSetDevice(d);
                cudaFree(v[i].p);
                v.erase(v.begin() + i);
            }
    }
    if (prev >= 0) cudaSetDevice(prev);
}

cudaError_t staged_upload(void *dst_device, const void *src_host, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    if (bytes < kDirectBelow) {
        cudaError_t e = cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, st);
        return e == cudaSuccess ? cudaStreamSynchronize(st) : e;   // pageable source: do not return before it has been read
    }
    Staging &S = staging();
    std::lock_guard<std::mutex> lk(S.mu);
    cudaError_t e = S.ensure();
    if (e != cudaSuccess) return e;
    Events E;
    if ((e = E.create()) != cudaSuccess) return e;
    const char *src = static_cast<const char *>(src_host);
    char *dst = static_cast<char *>(dst_device);
    int k = 0;
    for (size_t off = 0; off < bytes; off += kChunk, ++k) {
        const int b = k % kBufs;
        const size_t len = std::min(kChunk, bytes - off);
        if (k >= kBufs && (e = cudaEventSynchronize(E.ev[b])) != cudaSuccess) return e;   // buffer b has left the host
        S.copy(S.buf[b], src + off, len);
        if ((e = cudaMemcpyAsync(dst + off, S.buf[b], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(E.ev[b], st)) != cudaSuccess) return e;
    }
    for (int b = 0; b < std::min(k, kBufs); ++b)   // the staging buffers are shared: drain before releasing them
        if ((e = cudaEventSynchronize(E.ev[b])) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t staged_download(void *dst_host, const void *src_device, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return cudaSuccess;
    if (bytes < kDirectBelow) {
        cudaError_t e = cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, st);
        return e == cudaSuccess ? cudaStreamSynchronize(st) : e;
    }
    Staging &S = staging();
    std::lock_guard<std::mutex> lk(S.mu);
    cudaError_t e = S.ensure();
    if (e != cudaSuccess) return e;
    Events E;
    if ((e = E.create()) != cudaSuccess) return e;
    const char *src = static_cast<const char *>(src_device);
    char *dst = static_cast<char *>(dst_host);
    const int n_chunks = (int)((bytes + kChunk - 1) / kChunk);
    auto drain = [&](int j) -> cudaError_t {   // chunk j: wait for its copy, then hand it to the caller's buffer
        cudaError_t e2 = cudaEventSynchronize(E.ev[j % kBufs]);
        if (e2 != cudaSuccess) return e2;
        const size_t off = (size_t)j * kChunk;
        S.copy(dst + off, S.buf[j % kBufs], std::min(kChunk, bytes - off));
        return cudaSuccess;
    };
    for (int k = 0; k < n_chunks; ++k) {
        if (k >= kBufs && (e = drain(k - kBufs)) != cudaSuccess) return e;   // frees buffer k % kBufs
        const size_t off = (size_t)k * kChunk, len = std::min(kChunk, bytes - off);
        if ((e = cudaMemcpyAsync(S.buf[k % kBufs], src + off, len, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(E.ev[k % kBufs], st)) != cudaSuccess) return e;
    }
    for (int j = std::max(0, n_chunks - kBufs); j < n_chunks; ++j)
        if ((e = drain(j)) != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace b200rt
