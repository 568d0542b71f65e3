// bvh_builder.h -- host-side builder of the 4-wide BVH the device kernels traverse.
//
// Takes the place of the reference's BVH::build_bvh_tree / flatten_bvh_tree
// (reference include/acceleration/bvh.h:183-550) but is a different design: a binned-SAH
// binary tree built in parallel on the host, collapsed to 4-wide nodes whose four child boxes
// are stored SoA in one 128-byte line with FP32 bounds rounded OUTWARD (conservative), and
// leaves that reference contiguous runs of a per-type primitive array in leaf order.
// Closest-hit results do not depend on the tree (only on the primitives), so the tree is free
// to differ from the reference's; see DESIGN.md "Acceleration structure".
#pragma once
#include <cstdint>
#include <vector>

namespace b200rt {

struct Box3 {
    double lo[3], hi[3];
};

// One 4-wide node = one 128-byte cache line = eight 16-byte vector loads.
// Child i's box is component i of each float4.  child[i] encodes:
//   >= 0            : index of an interior Node4
//   bit 31 set      : leaf; bit 30 = primitive type (0 sphere, 1 quad); bits 26..29 = count
//                     (0 = empty slot); bits 0..25 = first index in the per-type leaf-ordered array
struct alignas(128) Node4 {
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int32_t child[4];
    int32_t pad[4];
};
static_assert(sizeof(Node4) == 128, "Node4 must be exactly one 128-byte line");

constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr uint32_t kQuadFlag = 0x40000000u;
constexpr int kLeafCountShift = 26;
constexpr uint32_t kLeafIndexMask = (1u << kLeafCountShift) - 1;
constexpr int kMaxLeafPrims = 8;
constexpr int32_t kEmptyChild = (int32_t)kLeafFlag;   // leaf with count 0

struct BuildParams {
    int max_leaf_prims = 2;
    int sah_bins = 32;
    int threads = 0;          // 0 = all
    int max_binary_depth = 40;
    double cost_traversal = 1.0;   // relative cost of one child-box test
    double cost_sphere = 2.0;      // relative cost of one FP64 sphere test
    double cost_quad = 3.0;        // relative cost of one FP64 quad test
};

struct BuiltBVH {
    std::vector<Node4> nodes;              // nodes[0] is the root
    std::vector<uint32_t> sphere_order;    // leaf-ordered position -> input sphere index
    std::vector<uint32_t> quad_order;      // leaf-ordered position -> input quad index
    uint32_t depth = 0;                    // depth of the 4-wide tree (root = 1)
    uint32_t binary_depth = 0;
    uint64_t n_leaves = 0;
};

// prim_boxes: n_spheres sphere boxes followed by n_quads quad boxes (double precision, as the
// reference computes them: sphere.h:112-123, parallelogram.h:281-295).
// Returns false (with a message in err) if the scene exceeds the node encoding limits.
bool build_bvh4(const std::vector<Box3> &prim_boxes, uint64_t n_spheres, uint64_t n_quads,
                const BuildParams &params, BuiltBVH &out, const char **err);

// Checks the structural invariants of a built tree (every primitive in exactly one leaf, every
// primitive's box inside every ancestor slot box, leaf runs contiguous, depth as reported).
// Used by the CPU test-suite through b200rt_selftest_bvh().
bool validate_bvh4(const BuiltBVH &bvh, const std::vector<Box3> &prim_boxes, uint64_t n_spheres,
                   uint64_t n_quads, const char **err);

}  // namespace b200rt
