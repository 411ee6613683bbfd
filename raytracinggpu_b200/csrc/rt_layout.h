/*
 * rt_layout.h — device-resident scene layout (HBM) shared by the host uploader and the kernels.
 *
 * One contiguous blob per scene, so that a scene built on rank 0 can be broadcast to the other GPUs in
 * one collective:
 *
 *   [ SceneHeader, padded to RT_HEADER_BYTES ]
 *   [ inner nodes : n_inner  x 64 B ]   both children's boxes + both child references in one record
 *   [ triangles   : n_tris   x 64 B ]   A, e1 = B-A, e2 = C-A, N = e1 x e2 (12 floats) + N/|N| (3 floats) + first
 *                                       triangle of the reference's leaf (tie-break key): one 64-B-aligned record,
 *                                       a 256-bit + a 128-bit load per test (LDG.E.256 on sm_100a)
 *   [ leaf table  : n_leaves x 32 B ]   box of the reference leaf (6 floats), leaf code (first triangle << 2 | count - 1),
 *                                       first triangle of the reference leaf: what the anchored-ray bins list
 *                                       (rt_bins.cuh); a reference leaf larger than RT_LEAF_MAX appears once per chunk
 *   [ wide nodes  : n_wide   x 128 B ]  the search index of wf_traverse: up to four (box, reference) children of 32 B
 *                                       each, one LDG.E.256 per child (see "Wide index" below)
 *
 * A child reference is one int: >= 0 inner-node index; < 0 a leaf, -1 - ref = (first triangle << 2) | (count - 1)
 * (triangle indices stay below 2^24 because the reference stores them in floats). Leaves hold at most RT_LEAF_MAX
 * triangles: a larger leaf of the reference BVH (the cat has one of 73) is hung under a small
 * tree of "virtual" inner nodes whose child boxes all equal the leaf's own box. The box test of a virtual node
 * repeats the computation that already succeeded for its parent, so exactly the reference's triangles are still
 * tested, while every traversal task stays small and uniform.
 *
 * Wide index. A child box of the reference BVH lies inside its parent's box (compute_bbox takes the min/max over a
 * sub-range of the parent's triangles, optimized.cu:466-474), and every operation of BoundingBox::intersect
 * (optimized.cu:173-184: subtract, divide, swap, min/max, compare) is monotone under IEEE rounding, so "the child's box
 * passes the slab test" implies "the parent's box passes it". The reference's traversal therefore tests exactly the
 * triangles of the leaves whose OWN box passes the slab test (and whose ray does not divide 0 by 0, which needs a zero
 * direction component: those rays take the binary records). Which inner boxes are consulted on the way is free. The
 * wide nodes are the two-child records collapsed two-to-three levels at a time (always opening the child with the
 * largest box): a third of the node visits and less than half the levels, i.e. less than half the dependent steps
 * of a single expensive ray. Inner children may be opened conservatively; leaf children are decided exactly.
 * A wide reference is (wide node index << 2) | (children - 1).
 *
 * versus the reference interchange format (what rt_scene_set_mesh receives and optimized.cu:814-826 uploads):
 * 40-B nodes read as 10 scalar loads with every child node read twice (optimized.cu:223-238, 255-261), and
 * a 40-B index record + 3 dependent 12-B vertex gathers per triangle test (:271). Here an inner-node visit is
 * two 32-B loads of one 64-B record, a triangle test a 32-B and a 16-B load of one 64-B record, no indirection.
 *
 * e1, e2, N and N/|N| depend only on the mesh, so they are precomputed once on the device with the same
 * unfused IEEE operations moller_trumbore (optimized.cu:209-211) and N.normalize() (:282) apply per ray:
 * bit-identical values, 15 fewer FLOP per triangle test.
 */
#pragma once
#include <stdint.h>

#define RT_MAX_SPHERES 16
#define RT_HEADER_BYTES 1024
#define RT_NODE_BYTES 64
#define RT_LEAF_MAX 4
#define RT_TRI_BYTES 64
#define RT_STACK_CAP 64 /* traversal stack entries; rt_scene_set_mesh rejects deeper trees */
#define RT_WIDE 4
#define RT_WNODE_BYTES 128
#define RT_BLOB_MAGIC 0x52544233u /* "RTB3" */
#define RT_LEAFREC_BYTES 32
#define RT_LAYOUT_VERSION 4

struct DevSphere {
    float cx, cy, cz, R;
    float RR; /* RN(R*R), the product Sphere::intersect recomputes per ray (optimized.cu:124) */
    float ax, ay, az;
    int32_t mirror;
    float n_in, n_out;
    int32_t id;
};

struct SceneHeader {
    uint32_t magic;
    uint32_t layout_version;
    int32_t n_spheres;
    int32_t has_mesh;
    int32_t n_inner;   /* inner-node records, virtual ones included */
    int32_t n_leaves;  /* leaves of the packed tree (<= RT_LEAF_MAX triangles each) */
    int32_t n_tris;
    int32_t max_depth; /* levels of the packed tree (virtual levels included): bound on the traversal stacks */
    int32_t mesh_id;
    int32_t mesh_mirror;
    float mesh_n_in, mesh_n_out;
    float mesh_albedo[3];
    float root_mn[3], root_mx[3];
    float box_abs[3]; /* largest |coordinate| over all node boxes, per axis (bound used by the certified slab test) */
    int32_t root_ref;
    int32_t wroot_ref; /* wide reference of the root's wide node (meaningful when root_ref >= 0) */
    float L[3];
    float intensity;
    uint64_t off_nodes, off_tris, total_bytes; /* byte offsets inside the blob */
    uint64_t off_wide;
    uint64_t off_leaves;
    int32_t n_wide;
    int32_t wide_depth; /* levels of the wide index */
    int32_t max_leaf;   /* largest leaf of the reference BVH (triangles): bounds the in-leaf offset of the tie-break rank */
    int32_t pad_;
    DevSphere spheres[RT_MAX_SPHERES];                    /* ascending id */
};

static_assert(sizeof(DevSphere) == 48, "DevSphere layout");
static_assert(sizeof(SceneHeader) <= RT_HEADER_BYTES, "SceneHeader must fit its slot");
