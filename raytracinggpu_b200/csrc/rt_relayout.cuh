/*
 * rt_relayout.cuh — the node relayout of rt_scene_set_mesh on the DEVICE (rt_scene_set_mesh_device).
 *
 * Input: the reference's interchange arrays already resident on the device (what rt_mesh_build_bvh_gpu leaves there, or any
 * caller's device copies): arr_bvh float[n_nodes * 10] (pre-order, optimized.cu:512-534), TriangleIndices int[nt * 10], vertices.
 * Output: the packed scene blob of rt_layout.h — 64-B two-child records, the leaf table, the per-triangle leaf key — built by a
 * handful of kernels instead of the host passes of rt_scene_set_mesh (0.5-1.2 s at 10 M triangles, plus 1.1 GB over PCIe).
 *
 * Same semantics as the host relayout, different numbering (results do not depend on it: a record's index is only a name):
 *   real inner node k of the reference tree  -> record inner_rank[k] (rank among the inner nodes in pre-order)
 *   a reference leaf with more than RT_LEAF_MAX triangles -> chunks of RT_LEAF_MAX under a heap-numbered binary tree of "virtual" records
 *     that repeat the leaf's own box (rt_layout.h): internal i in [1, chunks) has children 2i and 2i + 1, a child c >= chunks is chunk
 *     c - chunks; depth <= ceil(log2(chunks)) + 1
 * The wide index is not built here (n_wide = 0): scenes uploaded this way use the two-child records for every tree search.
 */
#pragma once
#include "rt_layout.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtrelayout {

struct Totals {
    int n_real;      /* inner nodes of the reference tree */
    int n_virtual;   /* virtual records of all chunked leaves */
    int n_leafrecs;  /* leaf-table entries (chunks) */
    int max_leaf;    /* largest reference leaf */
    int max_chunks;  /* largest chunk count of a leaf */
    int error;       /* != 0: malformed arr_bvh (child index / triangle range out of bounds) */
    unsigned box_abs[3]; /* float bits of the largest |coordinate| per axis (non-negative floats order like their bits) */
    int max_depth;   /* levels of the reference tree */
};

/* per node: 1 if inner; chunk and virtual-record counts of leaves; structure checks */
__global__ void classify(const float* __restrict__ arr, int n_nodes, int nt, int* __restrict__ is_inner, int* __restrict__ n_chunks, int* __restrict__ n_virt, Totals* tot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes) return;
    const float* a = arr + (size_t)k * 10;
    const int l = (int)a[0], r = (int)a[1];
    int inner = 0, chunks = 0, virt = 0;
    if ((l == -1) != (r == -1)) atomicExch(&tot->error, 1);
    if (l != -1) {
        inner = 1;
        if (l <= k || r <= k || l >= n_nodes || r >= n_nodes || l == r) atomicExch(&tot->error, 2);
    } else {
        const int ts = (int)a[8], te = (int)a[9];
        if (ts < 0 || te < ts || te > nt) atomicExch(&tot->error, 3);
        const int cnt = te - ts;
        atomicMax(&tot->max_leaf, cnt);
        chunks = cnt > 0 ? (cnt + RT_LEAF_MAX - 1) / RT_LEAF_MAX : 0; /* an empty leaf has no leaf-table entry */
        virt = chunks > 1 ? chunks - 1 : 0;
        atomicMax(&tot->max_chunks, chunks);
    }
    is_inner[k] = inner;
    n_chunks[k] = chunks;
    n_virt[k] = virt;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float m = fmaxf(fabsf(a[2 + c]), fabsf(a[5 + c]));
        atomicMax(&tot->box_abs[c], __float_as_uint(m));
    }
}

/* levels of the reference tree, one launch per level: nodes at depth `level` hand level + 1 to their children */
__global__ void depth_step(const float* __restrict__ arr, int n_nodes, int level, int* __restrict__ depth, Totals* tot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes || depth[k] != level) return;
    const float* a = arr + (size_t)k * 10;
    const int l = (int)a[0], r = (int)a[1];
    if (l != -1 && l > k && r > k && l < n_nodes && r < n_nodes) {
        depth[l] = level + 1;
        depth[r] = level + 1;
        atomicMax(&tot->max_depth, level + 1);
    }
}

/* the reference of a child of the packed tree: >= 0 a record, < 0 a leaf code (rt_layout.h) */
__device__ __forceinline__ int child_ref(const float* __restrict__ arr, int node, int nt, const int* __restrict__ inner_rank, const int* __restrict__ n_chunks,
                                         const int* __restrict__ virt_base, int n_real) {
    const float* a = arr + (size_t)node * 10;
    if ((int)a[0] != -1) return inner_rank[node];
    const int ts = (int)a[8], cnt = (int)a[9] - ts;
    if (cnt <= 0) return -1 - ((nt << 2) | 0); /* empty leaf: points past the last triangle, tests nothing */
    if (n_chunks[node] == 1) return -1 - ((ts << 2) | (cnt - 1));
    return n_real + virt_base[node]; /* root (heap index 1) of the leaf's chunk tree */
}

/* one 64-B record per real inner node: both children's boxes + references */
__global__ void write_inner(const float* __restrict__ arr, int n_nodes, int nt, const int* __restrict__ is_inner, const int* __restrict__ inner_rank,
                            const int* __restrict__ n_chunks, const int* __restrict__ virt_base, int n_real, float4* __restrict__ records) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes || !is_inner[k]) return;
    const float* a = arr + (size_t)k * 10;
    const int l = (int)a[0], r = (int)a[1];
    const float* L = arr + (size_t)l * 10;
    const float* R = arr + (size_t)r * 10;
    const int rl = child_ref(arr, l, nt, inner_rank, n_chunks, virt_base, n_real), rr = child_ref(arr, r, nt, inner_rank, n_chunks, virt_base, n_real);
    float4* o = records + 4 * (size_t)inner_rank[k];
    o[0] = make_float4(L[2], L[3], L[4], L[5]);
    o[1] = make_float4(L[6], L[7], R[2], R[3]);
    o[2] = make_float4(R[4], R[5], R[6], R[7]);
    o[3] = make_float4(__int_as_float(rl), __int_as_float(rr), __int_as_float(0), __int_as_float(0));
}

/* per reference leaf: its leaf-table entries (one per chunk), the virtual records of its chunk tree, the per-triangle leaf key */
__global__ void write_leaves(const float* __restrict__ arr, int n_nodes, int nt, const int* __restrict__ is_inner, const int* __restrict__ n_chunks,
                             const int* __restrict__ virt_base, const int* __restrict__ leaf_base, int n_real, float4* __restrict__ records,
                             float4* __restrict__ leaf_table, int* __restrict__ leaf_start_of_tri) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_nodes || is_inner[k]) return;
    const float* a = arr + (size_t)k * 10;
    const int ts = (int)a[8], te = (int)a[9], chunks = n_chunks[k];
    if (chunks <= 0 || ts < 0 || te > nt) return;
    const float4 b0 = make_float4(a[2], a[3], a[4], a[5]);
    for (int i = ts; i < te; i++) leaf_start_of_tri[i] = ts; /* first triangle of the REFERENCE leaf: the tie-break key */
    const int lb = leaf_base[k];
    for (int j = 0; j < chunks; j++) {
        const int st = ts + j * RT_LEAF_MAX, cnt = min(RT_LEAF_MAX, te - st);
        const int code = (st << 2) | (cnt - 1);
        leaf_table[2 * (size_t)(lb + j)] = b0;
        leaf_table[2 * (size_t)(lb + j) + 1] = make_float4(a[6], a[7], __int_as_float(code), __int_as_float(ts));
    }
    if (chunks > 1) {
        const int vb = n_real + virt_base[k];
        for (int i = 1; i < chunks; i++) { /* heap-numbered internal node i: children 2i, 2i + 1 */
            int ref[2];
            for (int c = 0; c < 2; c++) {
                const int ch = 2 * i + c;
                if (ch < chunks) {
                    ref[c] = vb + (ch - 1);
                } else {
                    const int j = ch - chunks, st = ts + j * RT_LEAF_MAX, cnt = min(RT_LEAF_MAX, te - st);
                    ref[c] = -1 - ((st << 2) | (cnt - 1));
                }
            }
            float4* o = records + 4 * (size_t)(vb + (i - 1));
            o[0] = b0;
            o[1] = make_float4(a[6], a[7], a[2], a[3]);
            o[2] = make_float4(a[4], a[5], a[6], a[7]);
            o[3] = make_float4(__int_as_float(ref[0]), __int_as_float(ref[1]), __int_as_float(1) /* virtual */, __int_as_float(0));
        }
    }
}

/* gather of whole triangle records by the builder's permutation: out[i] = in[perm[i]] */
__global__ void gather_records(const int* __restrict__ in, const int* __restrict__ perm, int nt, int* __restrict__ out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nt * 10) return;
    const int i = (int)(t / 10), w = (int)(t % 10);
    out[t] = in[(size_t)perm[i] * 10 + w];
}

/* vertex indices (nt x 3) out of whole records (nt x 10), for the builder */
__global__ void records_to_idx3(const int* __restrict__ recs, int nt, int* __restrict__ idx3) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)nt * 3) return;
    idx3[t] = recs[(t / 3) * 10 + (t % 3)];
}

/* vertex indices in range? */
__global__ void check_indices(const int* __restrict__ recs, int nt, int nv, Totals* tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    for (int k = 0; k < 3; k++) {
        const int v = recs[(size_t)i * 10 + k];
        if (v < 0 || v >= nv) atomicExch(&tot->error, 4);
    }
}

} // namespace rtrelayout
