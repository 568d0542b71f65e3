// lbvh.cu -- acceleration-structure construction ON THE GPU (B200RT_BUILDER_GPU_LBVH).
//
// Stands in for BVH::build_bvh_tree / flatten_bvh_tree (reference include/acceleration/bvh.h:183-550,
// single-threaded: 5.5-7.2 s for the 2.2 M / 3.1 M primitive scenes on the B200 box's host) when
// time-to-first-pixel matters more than tree quality -- the multi-million-primitive scenes, where
// even the parallel host SAH builder (bvh_builder.cpp, ~0.8 s) costs more than the render at 8 GPUs.
//
// Pipeline, all device-side (the host only uploads the caller's flat arrays as they are):
//   1. prim_setup      index validation of the description; primitive bounds exactly as the reference's constructors compute them
//                      (sphere.h:112-123, parallelogram.h:281-295) in double, rounded outward to FP32;
//                      centroid bounds by warp shuffle + ordered-int atomics
//   2. morton_codes    63-bit Morton code of the centroid (21 bits per axis)
//   3. cub radix sort  (key, primitive) pairs                    [library call: a plain sort]
//   4. radix_tree      Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees and
//                      k-d trees": one thread per internal node, duplicates broken by index
//   5. fit_boxes       bottom-up union of child boxes, second arrival continues (atomic counters)
//   6. collapse_level  binary -> 4-wide, breadth first, one launch per level: each task pulls up the
//                      larger-area grandchildren until four slots are used, reserves its Node4 with an
//                      atomic counter, patches its parent's child slot and queues its interior children
//   7. reorder         spheres / quads / meta into leaf (= Morton) order; the quad constants
//                      (unit normal, n/|n|^2) are computed in double exactly as on the host path
// Closest-hit results do not depend on the tree, so the parity tests run against this builder too.
#include <cstdio>
#include <cub/cub.cuh>

#include "../../include/b200rt.h"
#include "bvh_builder.h"
#include "devmem.h"
#include "kernels.h"

namespace b200rt {

namespace {

#define LB_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

__device__ __forceinline__ int float_to_ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct PrimBoxes {
    float *lox, *loy, *loz, *hix, *hiy, *hiz;   // outward-rounded FP32 bounds per primitive
    float *cx, *cy, *cz;                         // centroid
};

__global__ void prim_setup(const B200rtSphere *sph, uint32_t n_sph, const B200rtQuad *quads, uint32_t n_quad, PrimBoxes B,
                           int *cbounds /* [6]: ordered-int min xyz, max xyz; [6]: bad-index flag */, uint32_t n_materials) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = n_sph + n_quad;
    float c[3] = {0.f, 0.f, 0.f};
    const bool valid = i < n;
    if (valid) {
        double lo[3], hi[3];
        // the scene description's index checks (material / canonical primitive index in range) happen here for the
        // multi-million-primitive scenes instead of in a serial host loop over the caller's arrays
        const uint32_t mat = i < n_sph ? sph[i].mat : quads[i - n_sph].mat, prim = i < n_sph ? sph[i].prim : quads[i - n_sph].prim;
        if (mat >= n_materials || prim >= n) cbounds[6] = 1;
        if (i < n_sph) {                                   // sphere.h:112-123
            const B200rtSphere s = sph[i];
            for (int a = 0; a < 3; ++a) {
                const double p = s.c[a] - s.r, q = s.c[a] + s.r;
                lo[a] = fmin(p, q); hi[a] = fmax(p, q);
            }
        } else {                                           // parallelogram.h:281-295 + ensure_min_axis_length(1e-4)
            const B200rtQuad q = quads[i - n_sph];
            for (int a = 0; a < 3; ++a) {
                const double p0 = q.v[a], p1 = q.v[a] + q.s1[a], p2 = q.v[a] + q.s2[a], p3 = (q.v[a] + q.s1[a]) + q.s2[a];
                lo[a] = fmin(fmin(p0, p1), fmin(p2, p3));
                hi[a] = fmax(fmax(p0, p1), fmax(p2, p3));
                const double size = hi[a] - lo[a];
                if (size < 1e-4) { const double pad = (1e-4 - size) / 2; lo[a] -= pad; hi[a] += pad; }
            }
        }
        B.lox[i] = __double2float_rd(lo[0]); B.loy[i] = __double2float_rd(lo[1]); B.loz[i] = __double2float_rd(lo[2]);
        B.hix[i] = __double2float_ru(hi[0]); B.hiy[i] = __double2float_ru(hi[1]); B.hiz[i] = __double2float_ru(hi[2]);
        for (int a = 0; a < 3; ++a) c[a] = (float)(0.5 * lo[a] + 0.5 * hi[a]);
        B.cx[i] = c[0]; B.cy[i] = c[1]; B.cz[i] = c[2];
    }
    // centroid bounds: warp reduce, then one atomic per warp and axis
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = valid ? c[a] : inf, mx = valid ? c[a] : -inf;
        for (int off = 16; off; off >>= 1) {
            mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, off));
            mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, off));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&cbounds[a], float_to_ordered(mn));
            atomicMax(&cbounds[3 + a], float_to_ordered(mx));
        }
    }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long x) {   // 21 bits -> every third bit
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_codes(PrimBoxes B, uint32_t n, const int *cbounds, unsigned long long *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long code = 0;
    const float *c[3] = {B.cx, B.cy, B.cz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ordered_to_float(cbounds[a]), hi = ordered_to_float(cbounds[3 + a]);
        const float ext = hi - lo;
        float u = ext > 0.f ? (c[a][i] - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        const unsigned long long q = (unsigned long long)fminf(u * 2097152.0f, 2097151.0f);
        code |= spread21(q) << a;
    }
    keys[i] = code;
    vals[i] = i;
}

// Common-prefix length of sorted keys i and j (-1 out of range); equal keys are told apart by index.
__device__ __forceinline__ int delta(const unsigned long long *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

struct BinTree {
    int *left, *right;        // child: >= 0 internal node index; < 0 leaf ~position (position = ~child)
    int *parent;              // parent of internal node i at [i], of leaf p at [n_internal + p]
    float *lox, *loy, *loz, *hix, *hiy, *hiz;   // internal node boxes
    int *visits;
};

__global__ void radix_tree(const unsigned long long *keys, int n, BinTree T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? ~gamma : gamma;
    const int rc = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    T.left[i] = lc;
    T.right[i] = rc;
    T.parent[lc >= 0 ? lc : (n - 1) + ~lc] = i;
    T.parent[rc >= 0 ? rc : (n - 1) + ~rc] = i;
    if (i == 0) T.parent[0] = -1;
}

__global__ void fit_boxes(PrimBoxes B, const uint32_t *vals, int n, BinTree T) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int cur = T.parent[(n - 1) + p];
    while (cur >= 0) {
        __threadfence();
        if (atomicAdd(&T.visits[cur], 1) == 0) return;   // the second arrival finds both children done
        float lo[3], hi[3];
        const int ch[2] = {T.left[cur], T.right[cur]};
        for (int a = 0; a < 3; ++a) { lo[a] = __int_as_float(0x7f800000); hi[a] = __int_as_float(0xff800000); }
        for (int k = 0; k < 2; ++k) {
            if (ch[k] >= 0) {
                const int c = ch[k];
                lo[0] = fminf(lo[0], __ldcg(&T.lox[c])); lo[1] = fminf(lo[1], __ldcg(&T.loy[c])); lo[2] = fminf(lo[2], __ldcg(&T.loz[c]));
                hi[0] = fmaxf(hi[0], __ldcg(&T.hix[c])); hi[1] = fmaxf(hi[1], __ldcg(&T.hiy[c])); hi[2] = fmaxf(hi[2], __ldcg(&T.hiz[c]));
            } else {
                const uint32_t prim = vals[~ch[k]];
                lo[0] = fminf(lo[0], B.lox[prim]); lo[1] = fminf(lo[1], B.loy[prim]); lo[2] = fminf(lo[2], B.loz[prim]);
                hi[0] = fmaxf(hi[0], B.hix[prim]); hi[1] = fmaxf(hi[1], B.hiy[prim]); hi[2] = fmaxf(hi[2], B.hiz[prim]);
            }
        }
        T.lox[cur] = lo[0]; T.loy[cur] = lo[1]; T.loz[cur] = lo[2];
        T.hix[cur] = hi[0]; T.hiy[cur] = hi[1]; T.hiz[cur] = hi[2];
        cur = T.parent[cur];
    }
}

__global__ void quad_flags(const uint32_t *vals, uint32_t n, uint32_t n_sph, uint32_t *flags) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) flags[p] = vals[p] >= n_sph ? 1u : 0u;
}

struct CollapseTask { int bnode, parent4, slot; };

__device__ __forceinline__ void child_box(const BinTree &T, const PrimBoxes &B, const uint32_t *vals, int c, float lo[3], float hi[3]) {
    if (c >= 0) {
        lo[0] = T.lox[c]; lo[1] = T.loy[c]; lo[2] = T.loz[c]; hi[0] = T.hix[c]; hi[1] = T.hiy[c]; hi[2] = T.hiz[c];
    } else {
        const uint32_t prim = vals[~c];
        lo[0] = B.lox[prim]; lo[1] = B.loy[prim]; lo[2] = B.loz[prim]; hi[0] = B.hix[prim]; hi[1] = B.hiy[prim]; hi[2] = B.hiz[prim];
    }
}

__global__ void collapse_level(BinTree T, PrimBoxes B, const uint32_t *vals, const uint32_t *quad_rank, uint32_t n_sph,
                               const CollapseTask *tasks, uint32_t n_tasks, CollapseTask *next, uint32_t *next_count,
                               Node4 *nodes4, uint32_t *node_count) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const CollapseTask task = tasks[t];
    int kids[4];
    int nk = 0;
    kids[nk++] = T.left[task.bnode];
    kids[nk++] = T.right[task.bnode];
    while (nk < 4) {                       // pull up the interior child with the largest box
        int pick = -1;
        float best = -1.f;
        for (int k = 0; k < nk; ++k) {
            if (kids[k] < 0) continue;
            const int c = kids[k];
            const float ex = T.hix[c] - T.lox[c], ey = T.hiy[c] - T.loy[c], ez = T.hiz[c] - T.loz[c];
            float area = ex * ey + ey * ez + ez * ex;
            if (!(area >= 0.f)) area = 3.4e38f;            // inf/NaN extents: treat as the largest
            if (area > best) { best = area; pick = k; }
        }
        if (pick < 0) break;
        const int c = kids[pick];
        kids[pick] = T.left[c];
        kids[nk++] = T.right[c];
    }
    const uint32_t me = atomicAdd(node_count, 1u);
    if (task.parent4 >= 0) nodes4[task.parent4].child[task.slot] = (int)me;
    Node4 nd;
    const float inf = __int_as_float(0x7f800000);
    for (int k = 0; k < 4; ++k) {
        nd.lox[k] = nd.loy[k] = nd.loz[k] = inf;
        nd.hix[k] = nd.hiy[k] = nd.hiz[k] = -inf;
        nd.child[k] = kEmptyChild;
        nd.pad[k] = 0;
    }
    for (int k = 0; k < nk; ++k) {
        float lo[3], hi[3];
        child_box(T, B, vals, kids[k], lo, hi);
        nd.lox[k] = lo[0]; nd.loy[k] = lo[1]; nd.loz[k] = lo[2];
        nd.hix[k] = hi[0]; nd.hiy[k] = hi[1]; nd.hiz[k] = hi[2];
        if (kids[k] < 0) {
            const uint32_t p = (uint32_t)~kids[k];
            const bool quad = vals[p] >= n_sph;
            const uint32_t typed = quad ? quad_rank[p] : p - quad_rank[p];
            nd.child[k] = (int)(kLeafFlag | (quad ? kQuadFlag : 0u) | (1u << kLeafCountShift) | typed);
        } else {
            nd.child[k] = 0;   // patched by the child's own task
            const uint32_t q = atomicAdd(next_count, 1u);
            next[q] = CollapseTask{kids[k], (int)me, k};
        }
    }
    // the child slots of interior kids are written later by other threads: store everything but those
    Node4 *dst = nodes4 + me;
    for (int k = 0; k < 4; ++k) {
        dst->lox[k] = nd.lox[k]; dst->loy[k] = nd.loy[k]; dst->loz[k] = nd.loz[k];
        dst->hix[k] = nd.hix[k]; dst->hiy[k] = nd.hiy[k]; dst->hiz[k] = nd.hiz[k];
        dst->pad[k] = 0;
        if (k >= nk || kids[k] < 0) dst->child[k] = nd.child[k];
    }
}

// spheres / quads / meta into leaf (Morton) order
__global__ void reorder_prims(const B200rtSphere *sph, uint32_t n_sph, const B200rtQuad *quads, const uint32_t *vals,
                              const uint32_t *quad_rank, uint32_t n, double2 *out_sph, uint2 *out_sph_meta, double2 *out_quads,
                              uint2 *out_quad_meta) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t prim = vals[p];
    if (prim < n_sph) {
        const B200rtSphere s = sph[prim];
        const uint32_t i = p - quad_rank[p];
        out_sph[2 * (size_t)i] = make_double2(s.c[0], s.c[1]);
        out_sph[2 * (size_t)i + 1] = make_double2(s.c[2], s.r);
        out_sph_meta[i] = make_uint2(s.prim, s.mat);
    } else {
        const B200rtQuad q = quads[prim - n_sph];
        const uint32_t i = quad_rank[p];
        // parallelogram.h:275-279: n = cross(s1, s2); unit = n * (1 / |n|); w = n * (1 / |n|^2)
        const double nx = q.s1[1] * q.s2[2] - q.s1[2] * q.s2[1], ny = q.s1[2] * q.s2[0] - q.s1[0] * q.s2[2],
                     nz = q.s1[0] * q.s2[1] - q.s1[1] * q.s2[0];
        const double m2 = nx * nx + ny * ny + nz * nz;
        const double im = 1 / sqrt(m2), im2 = 1 / m2;
        const double f[16] = {nx * im, ny * im, nz * im, q.v[0], q.v[1], q.v[2], nx * im2, ny * im2, nz * im2,
                              q.s1[0], q.s1[1], q.s1[2], q.s2[0], q.s2[1], q.s2[2], 0.0};
        for (int k = 0; k < 8; ++k) out_quads[8 * (size_t)i + k] = make_double2(f[2 * k], f[2 * k + 1]);
        out_quad_meta[i] = make_uint2(q.prim, q.mat);
    }
}

// B200rtMaterial (the caller's 40-byte records) -> DeviceMaterial, as api.cu's device_materials() does on the host for
// small scenes: lights pre-multiplied by their intensity (material.h:261-263), metal fuzz clamped to 1
// (material.h:150-151).  The reference's big scenes carry one material PER PRIMITIVE (millions), so this runs here.
__global__ void convert_materials(const B200rtMaterial *__restrict__ raw, uint32_t n, DeviceMaterial *__restrict__ out, int *bad_kind) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const B200rtMaterial m = raw[i];
    if (m.kind > B200RT_MAT_LIGHT) *bad_kind = 1;
    DeviceMaterial dm;
    const double k = m.kind == B200RT_MAT_LIGHT ? m.param : 1.0;
    dm.r = (float)(k * m.rgb[0]); dm.g = (float)(k * m.rgb[1]); dm.b = (float)(k * m.rgb[2]);
    dm.kind = m.kind;
    dm.param = m.kind == B200RT_MAT_METAL ? fmin(m.param, 1.0) : m.param;
    dm.pad = 0.0;
    out[i] = dm;
}

template <typename T>
cudaError_t tmp_alloc(std::vector<void *> &owned, T **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = dev_alloc_async(&q, (count ? count : 1) * sizeof(T), 0);
    if (e != cudaSuccess) return e;
    owned.push_back(q);
    *p = static_cast<T *>(q);
    return cudaSuccess;
}

}  // namespace

cudaError_t convert_materials_device(const B200rtMaterial *d_raw, uint32_t n, DeviceMaterial *d_out, int *bad_kind_out) {
    *bad_kind_out = 0;
    if (n == 0) return cudaSuccess;
    int *flag = nullptr;
    cudaError_t e = dev_alloc_async(&flag, sizeof(int), 0);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(flag, 0, sizeof(int), 0);
    if (e == cudaSuccess) {
        convert_materials<<<(n + 255) / 256, 256>>>(d_raw, n, d_out, flag);
        e = cudaMemcpyAsync(bad_kind_out, flag, sizeof(int), cudaMemcpyDeviceToHost, 0);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    dev_free(flag);
    return e;
}

// Builds the scene's device arrays on the GPU.  d_sph / d_quads are the caller's flat structs already
// on the device.  Outputs (allocated by the caller with the sizes known up front): leaf-ordered
// sphere / quad / meta arrays; the Node4 array is allocated here (its size is only known at the end)
// with cudaMallocAsync and handed back through nodes_out.
cudaError_t build_lbvh_device(const B200rtSphere *d_sph, uint32_t n_sph, const B200rtQuad *d_quads, uint32_t n_quad,
                              uint32_t n_materials, double2 *out_sph, uint2 *out_sph_meta, double2 *out_quads, uint2 *out_quad_meta,
                              float4 **nodes_out, uint32_t *n_nodes_out, uint32_t *depth_out, int *bad_index_out) {
    const uint32_t n = n_sph + n_quad;
    // Encoding limits: a leaf reference holds a 26-bit index into its per-type array (bvh_builder.h), and the radix
    // tree's index arithmetic ((n - 1) + ~child, lmax * d) must stay inside int.  The host builder rejects the same
    // scenes (bvh_builder.cpp); api.cu checks before calling, this is the backstop.
    if (n < 2 || n_sph > kLeafIndexMask || n_quad > kLeafIndexMask || n > (1u << 29)) return cudaErrorInvalidValue;
    std::vector<void *> owned;
    Node4 *nodes4 = nullptr;
    auto cleanup = [&]() { for (void *p : owned) dev_free(p); };
    auto fail = [&](cudaError_t e) { cleanup(); dev_free(nodes4); return e; };
    const int B = 256;
    const unsigned gn = (n + B - 1) / B;
#define LB(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(e_); } while (0)
    PrimBoxes pb;
    LB(tmp_alloc(owned, &pb.lox, n)); LB(tmp_alloc(owned, &pb.loy, n)); LB(tmp_alloc(owned, &pb.loz, n));
    LB(tmp_alloc(owned, &pb.hix, n)); LB(tmp_alloc(owned, &pb.hiy, n)); LB(tmp_alloc(owned, &pb.hiz, n));
    LB(tmp_alloc(owned, &pb.cx, n)); LB(tmp_alloc(owned, &pb.cy, n)); LB(tmp_alloc(owned, &pb.cz, n));
    int *cbounds;
    LB(tmp_alloc(owned, &cbounds, 7));
    const int init[7] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000, 0};
    LB(cudaMemcpyAsync(cbounds, init, sizeof init, cudaMemcpyHostToDevice, 0));
    prim_setup<<<gn, B>>>(d_sph, n_sph, d_quads, n_quad, pb, cbounds, n_materials);
    int bad_index = 0;
    LB(cudaMemcpyAsync(&bad_index, cbounds + 6, sizeof(int), cudaMemcpyDeviceToHost, 0));   // read at the first sync below
    unsigned long long *keys, *keys_alt;
    uint32_t *vals, *vals_alt;
    LB(tmp_alloc(owned, &keys, n)); LB(tmp_alloc(owned, &keys_alt, n));
    LB(tmp_alloc(owned, &vals, n)); LB(tmp_alloc(owned, &vals_alt, n));
    morton_codes<<<gn, B>>>(pb, n, cbounds, keys, vals);
    {
        cub::DoubleBuffer<unsigned long long> dk(keys, keys_alt);
        cub::DoubleBuffer<uint32_t> dv(vals, vals_alt);
        size_t tmp_bytes = 0;
        LB(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, 63));
        char *tmp;
        LB(tmp_alloc(owned, &tmp, tmp_bytes));
        LB(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dv, (int)n, 0, 63));
        keys = dk.Current();
        vals = dv.Current();
    }
    BinTree T;
    LB(tmp_alloc(owned, &T.left, n)); LB(tmp_alloc(owned, &T.right, n)); LB(tmp_alloc(owned, &T.parent, 2 * (size_t)n));
    LB(tmp_alloc(owned, &T.lox, n)); LB(tmp_alloc(owned, &T.loy, n)); LB(tmp_alloc(owned, &T.loz, n));
    LB(tmp_alloc(owned, &T.hix, n)); LB(tmp_alloc(owned, &T.hiy, n)); LB(tmp_alloc(owned, &T.hiz, n));
    LB(tmp_alloc(owned, &T.visits, n));
    LB(cudaMemsetAsync(T.visits, 0, n * sizeof(int), 0));
    radix_tree<<<gn, B>>>(keys, (int)n, T);
    fit_boxes<<<gn, B>>>(pb, vals, (int)n, T);
    // typed (per primitive kind) positions in leaf order
    uint32_t *flags, *quad_rank;
    LB(tmp_alloc(owned, &flags, n)); LB(tmp_alloc(owned, &quad_rank, n));
    quad_flags<<<gn, B>>>(vals, n, n_sph, flags);
    {
        size_t tmp_bytes = 0;
        LB(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flags, quad_rank, (int)n));
        char *tmp;
        LB(tmp_alloc(owned, &tmp, tmp_bytes));
        LB(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flags, quad_rank, (int)n));
    }
    reorder_prims<<<gn, B>>>(d_sph, n_sph, d_quads, vals, quad_rank, n, out_sph, out_sph_meta, out_quads, out_quad_meta);
    // breadth-first collapse to 4-wide nodes (at most n - 1 of them)
    LB(dev_alloc_async(&nodes4, (size_t)(n - 1) * sizeof(Node4), 0));
    CollapseTask *qa, *qb;
    uint32_t *counters;   // [0] next queue size, [1] nodes emitted
    LB(tmp_alloc(owned, &qa, n)); LB(tmp_alloc(owned, &qb, n)); LB(tmp_alloc(owned, &counters, 2));
    LB(cudaMemsetAsync(counters, 0, 2 * sizeof(uint32_t), 0));
    const CollapseTask root{0, -1, 0};
    LB(cudaMemcpyAsync(qa, &root, sizeof root, cudaMemcpyHostToDevice, 0));
    uint32_t n_tasks = 1, depth = 0;
    while (n_tasks) {
        ++depth;
        collapse_level<<<(n_tasks + B - 1) / B, B>>>(T, pb, vals, quad_rank, n_sph, qa, n_tasks, qb, counters, nodes4, counters + 1);
        LB(cudaMemcpyAsync(&n_tasks, counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, 0));
        LB(cudaMemsetAsync(counters, 0, sizeof(uint32_t), 0));
        LB(cudaStreamSynchronize(0));
        std::swap(qa, qb);
        if (depth > 512) return fail(cudaErrorUnknown);
    }
    uint32_t n_nodes = 0;
    LB(cudaMemcpyAsync(&n_nodes, counters + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, 0));
    LB(cudaStreamSynchronize(0));
    LB(cudaGetLastError());
    cleanup();
    *bad_index_out = bad_index;
    *nodes_out = reinterpret_cast<float4 *>(nodes4);
    *n_nodes_out = n_nodes;
    *depth_out = depth;
    return cudaSuccess;
#undef LB
}

}  // namespace b200rt
