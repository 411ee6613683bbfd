/*
 * rt_bvh_build.cuh — the reference's BVH builder (compute_bbox / buildBVH / bvhTreeToArray, optimized.cu:466-534) on the
 * device, producing the IDENTICAL tree and triangle order. The reference builds top-down, one node at a
 * time, single-threaded (seconds at 10 M triangles, inside its "Rendering time"); global_launcher.cu:298-331 runs the
 * same recursion in ONE device thread. Here a whole tree level is built at once:
 *
 *   bbox      min/max over the vertices of every node's triangle range (float min/max are exact; atomics on an
 *             order-preserving integer key, one per warp where the warp lies in one node)          :466-474
 *   split     longest axis (ties x, then y), split = (mn + mx) / 2                                  :485-494
 *   flags     centroid (a + b + c) / 3 < split, in the reference's operation order                  :496
 *   partition the reference's in-place loop `if (left) swap(T[i], T[pivot++])` (:495-501) is NOT a stable partition: the
 *             left elements keep their order, but every swap sends the front element of the block of right elements
 *             to its back. In closed form: with L[q] = position of the q-th left element, the right element at
 *             position r ends at f(r), f(q) = q if q >= nL else f(L[q]) — a forest, resolved for all elements of all
 *             nodes of the level together by pointer doubling (log2(largest node) rounds).
 *   leaf rule pivot <= start || pivot >= end - 1 || n < 5 (:503), decided AFTER the partition: leaves are reordered too.
 *   numbering nodes are created level by level; subtree sizes (bottom-up) give the pre-order index bvhTreeToArray
 *             assigns (left child = parent + 1, right child = parent + 1 + size(left subtree))      :512-534
 *
 * The result (arr_bvh, permutation of the triangle records) is compared element by element with the host builder in
 * tests/test_gpu_build.py. One difference is possible and invisible: a box bound that is a zero can come out as -0.0
 * where the sequential std::min/max kept +0.0 (or vice versa); every consumer compares or subtracts it, where the two are equal.
 */
#pragma once
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include "rt_relayout.cuh"

#include <vector>

namespace rtbuild {

__device__ __forceinline__ unsigned fkey(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

struct Nodes { /* structure of arrays, one entry per node in creation (level) order */
    int* start;
    int* end;
    unsigned* kmn; /* 3 per node */
    unsigned* kmx;
    int* left;     /* node ids, -1 for a leaf */
    int* right;
    int* axis;
    float* split;
    int* nL;       /* triangles left of the split */
    int* size;     /* nodes in the subtree */
    int* pre;      /* pre-order index */
};

__global__ void k_init(int nt, int* perm, int* seg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) {
        perm[i] = i;
        seg[i] = 0;
    }
}

__global__ void k_level_reset(Nodes n, int base, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    for (int c = 0; c < 3; c++) {
        n.kmn[3 * (base + k) + c] = 0xffffffffu;
        n.kmx[3 * (base + k) + c] = 0u;
    }
    n.nL[base + k] = 0;
    n.left[base + k] = -1;
    n.right[base + k] = -1;
}

__global__ void k_bbox(int nt, const int* __restrict__ perm, const int* __restrict__ seg, const float* __restrict__ V, const int* __restrict__ I, Nodes n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = i < nt ? seg[i] : -1;
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    if (s >= 0) {
        const int t = perm[i];
        for (int v = 0; v < 3; v++) {
            const int vi = I[3 * t + v];
            for (int c = 0; c < 3; c++) {
                const unsigned k = fkey(V[3 * (size_t)vi + c]);
                mn[c] = min(mn[c], k);
                mx[c] = max(mx[c], k);
            }
        }
    }
    const int s0 = __shfl_sync(0xffffffffu, s, 0);
    if (__all_sync(0xffffffffu, s == s0)) { /* the whole warp lies in one node (or in none) */
        if (s0 < 0) return;
        for (int c = 0; c < 3; c++) {
            const unsigned a = __reduce_min_sync(0xffffffffu, mn[c]), b = __reduce_max_sync(0xffffffffu, mx[c]);
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&n.kmn[3 * s0 + c], a);
                atomicMax(&n.kmx[3 * s0 + c], b);
            }
        }
    } else if (s >= 0) {
        for (int c = 0; c < 3; c++) {
            atomicMin(&n.kmn[3 * s + c], mn[c]);
            atomicMax(&n.kmx[3 * s + c], mx[c]);
        }
    }
}

__global__ void k_split(Nodes n, int base, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int id = base + k;
    float mn[3], mx[3];
    for (int c = 0; c < 3; c++) {
        mn[c] = fkey_inv(n.kmn[3 * id + c]);
        mx[c] = fkey_inv(n.kmx[3 * id + c]);
    }
    const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    const int axis = (dx >= dy && dx >= dz) ? 0 : ((dy >= dx && dy >= dz) ? 1 : 2); /* optimized.cu:485-491 */
    n.axis[id] = axis;
    n.split[id] = (mn[axis] + mx[axis]) / 2; /* :494 */
}

__global__ void k_flags(int nt, const int* __restrict__ perm, const int* __restrict__ seg, const float* __restrict__ V, const int* __restrict__ I, Nodes n,
                        int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = i < nt ? seg[i] : -1;
    int f = 0;
    if (s >= 0) {
        const int t = perm[i], axis = n.axis[s];
        const float a = V[3 * (size_t)I[3 * t] + axis], b = V[3 * (size_t)I[3 * t + 1] + axis], c = V[3 * (size_t)I[3 * t + 2] + axis];
        const float cen = (a + b + c) / 3; /* :496 */
        f = cen < n.split[s] ? 1 : 0;
    }
    if (i < nt) flag[i] = f;
    const int s0 = __shfl_sync(0xffffffffu, s, 0);
    if (__all_sync(0xffffffffu, s == s0)) {
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (s0 >= 0 && m && (threadIdx.x & 31) == 0) atomicAdd(&n.nL[s0], __popc(m));
    } else if (f) {
        atomicAdd(&n.nL[s], 1);
    }
}

/* next[] of the pointer forest: the q-th left element's position for q < nL (relative to the node), itself otherwise */
__global__ void k_next_init(int nt, int* next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) next[i] = i;
}
__global__ void k_next_link(int nt, const int* __restrict__ seg, const int* __restrict__ flag, const int* __restrict__ S, Nodes n, int* next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int s = seg[i];
    if (s < 0 || !flag[i]) return;
    const int st = n.start[s];
    next[st + (S[i] - S[st])] = i;
}
__global__ void k_jump(int nt, const int* __restrict__ in, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) out[i] = in[in[i]];
}

/* leaf rule, children: cnt[k] = 0 (leaf) or 2; statistics */
__global__ void k_decide(Nodes n, int base, int count, int* __restrict__ cnt, int* __restrict__ stats /* [0] leaves, [1] largest leaf, [2] largest child range */) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int id = base + k;
    const int st = n.start[id], en = n.end[id], pivot = st + n.nL[id];
    const bool leaf = pivot <= st || pivot >= en - 1 || en - st < 5; /* :503 */
    cnt[k] = leaf ? 0 : 2;
    if (leaf) {
        atomicAdd(&stats[0], 1);
        atomicMax(&stats[1], en - st);
    } else {
        atomicMax(&stats[2], max(pivot - st, en - pivot));
    }
}
__global__ void k_children(Nodes n, int base, int count, const int* __restrict__ off, int next_base) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int id = base + k;
    if (off[k + 1] == off[k]) return; /* leaf */
    const int l = next_base + off[k], r = l + 1;
    const int st = n.start[id], en = n.end[id], pivot = st + n.nL[id];
    n.left[id] = l;
    n.right[id] = r;
    n.start[l] = st;
    n.end[l] = pivot;
    n.start[r] = pivot;
    n.end[r] = en;
}
/* apply the level's partitions; seg2 = the child a position now belongs to, -1 inside a leaf */
__global__ void k_scatter(int nt, const int* __restrict__ perm, const int* __restrict__ seg, const int* __restrict__ flag, const int* __restrict__ S,
                          const int* __restrict__ dest, Nodes n, int* __restrict__ perm2, int* __restrict__ seg2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int s = seg[i];
    if (s < 0) {
        perm2[i] = perm[i];
        seg2[i] = -1;
        return;
    }
    const int st = n.start[s];
    const int d = flag[i] ? st + (S[i] - S[st]) : dest[i];
    perm2[d] = perm[i];
    const int l = n.left[s];
    seg2[d] = l < 0 ? -1 : (d < st + n.nL[s] ? l : l + 1);
}

__global__ void k_sizes(Nodes n, int base, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int id = base + k, l = n.left[id];
    n.size[id] = l < 0 ? 1 : 1 + n.size[l] + n.size[n.right[id]];
}
__global__ void k_preorder(Nodes n, int base, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int id = base + k, l = n.left[id];
    if (l < 0) return;
    n.pre[l] = n.pre[id] + 1;
    n.pre[n.right[id]] = n.pre[id] + 1 + n.size[l];
}
__global__ void k_emit(Nodes n, int total, float* __restrict__ arr) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    float* o = arr + (size_t)n.pre[id] * 10;
    const int l = n.left[id];
    o[0] = l < 0 ? -1.f : (float)n.pre[l];
    o[1] = l < 0 ? -1.f : (float)n.pre[n.right[id]];
    for (int c = 0; c < 3; c++) {
        o[2 + c] = fkey_inv(n.kmn[3 * id + c]);
        o[5 + c] = fkey_inv(n.kmx[3 * id + c]);
    }
    o[8] = (float)n.start[id];
    o[9] = (float)n.end[id];
}

#define RTB_TRY(x)                                   \
    do {                                             \
        cudaError_t e_ = (x);                        \
        if (e_ != cudaSuccess) {                     \
            err = e_;                                \
            goto done;                               \
        }                                            \
    } while (0)

/* info: [0] nodes, [1] leaves, [2] depth, [3] largest leaf. Returns a cudaError_t (0 = success). */
/* What a build leaves on the device for rt_scene_set_mesh_device: the interchange arrays in their final (post-build) order. */
struct DeviceMesh {
    int device = -1;
    float* V = nullptr;    /* nv * 3 */
    int* recs = nullptr;   /* nt * 10, permuted */
    float* arr = nullptr;  /* nn * 10 */
    int nv = 0, nt = 0, nn = 0;
};

/* h_recs: whole triangle records (nt x 10). keep != NULL: the results STAY on the device (keep->V / recs / arr, owned by the caller from
 * then on) and nothing is copied back (perm_out / arr_out stay empty): the host mirror is fetched on demand (download). */
inline int build(int device, const float* h_vertices, int nv, const int32_t* h_recs, int nt, std::vector<int32_t>& perm_out, std::vector<float>& arr_out,
                 int32_t info[4], double* build_ms, DeviceMesh* keep = nullptr) {
    cudaError_t err = cudaSuccess;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return (int)cudaErrorInvalidDevice;
    const int max_nodes = 2 * nt;
    const int T = 256, B = (nt + T - 1) / T;
    float *V = nullptr, *arr = nullptr;
    int *R_in = nullptr, *R_out = nullptr;
    int *I = nullptr, *perm = nullptr, *perm2 = nullptr, *seg = nullptr, *seg2 = nullptr, *flag = nullptr, *S = nullptr, *nx = nullptr, *nx2 = nullptr, *cnt = nullptr, *off = nullptr,
        *stats = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    Nodes n = {};
    std::vector<int> level_base, level_count;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int h_stats[3] = {0, 0, 0};
    int total = 0;
    {
        size_t a = 0, b = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, a, (int*)nullptr, (int*)nullptr, nt + 1);
        cub::DeviceScan::ExclusiveSum(nullptr, b, (int*)nullptr, (int*)nullptr, max_nodes / 2 + 2);
        tmp_bytes = std::max(a, b) + 256;
    }
    RTB_TRY(cudaMalloc(&V, (size_t)nv * 3 * sizeof(float)));
    RTB_TRY(cudaMalloc(&I, (size_t)nt * 3 * sizeof(int)));
    RTB_TRY(cudaMalloc(&perm, (size_t)nt * sizeof(int)));
    RTB_TRY(cudaMalloc(&perm2, (size_t)nt * sizeof(int)));
    RTB_TRY(cudaMalloc(&seg, (size_t)nt * sizeof(int)));
    RTB_TRY(cudaMalloc(&seg2, (size_t)nt * sizeof(int)));
    RTB_TRY(cudaMalloc(&flag, ((size_t)nt + 1) * sizeof(int)));
    RTB_TRY(cudaMalloc(&S, ((size_t)nt + 1) * sizeof(int)));
    RTB_TRY(cudaMalloc(&nx, (size_t)nt * sizeof(int)));
    RTB_TRY(cudaMalloc(&nx2, (size_t)nt * sizeof(int)));
    RTB_TRY(cudaMalloc(&cnt, ((size_t)nt + 2) * sizeof(int)));
    RTB_TRY(cudaMalloc(&off, ((size_t)nt + 2) * sizeof(int)));
    RTB_TRY(cudaMalloc(&stats, 4 * sizeof(int)));
    RTB_TRY(cudaMalloc(&tmp, tmp_bytes));
    RTB_TRY(cudaMalloc(&n.start, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.end, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.kmn, (size_t)max_nodes * 3 * sizeof(unsigned)));
    RTB_TRY(cudaMalloc(&n.kmx, (size_t)max_nodes * 3 * sizeof(unsigned)));
    RTB_TRY(cudaMalloc(&n.left, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.right, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.axis, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.split, (size_t)max_nodes * sizeof(float)));
    RTB_TRY(cudaMalloc(&n.nL, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.size, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMalloc(&n.pre, (size_t)max_nodes * sizeof(int)));
    RTB_TRY(cudaMemcpy(V, h_vertices, (size_t)nv * 3 * sizeof(float), cudaMemcpyHostToDevice));
    RTB_TRY(cudaMalloc(&R_in, (size_t)nt * 10 * sizeof(int)));
    RTB_TRY(cudaMemcpy(R_in, h_recs, (size_t)nt * 10 * sizeof(int), cudaMemcpyHostToDevice));
    rtrelayout::records_to_idx3<<<(unsigned)(((size_t)nt * 3 + T - 1) / T), T>>>(R_in, nt, I);
    RTB_TRY(cudaEventCreate(&e0));
    RTB_TRY(cudaEventCreate(&e1));
    RTB_TRY(cudaEventRecord(e0));
    RTB_TRY(cudaMemset(stats, 0, 4 * sizeof(int)));
    RTB_TRY(cudaMemset(flag + nt, 0, sizeof(int)));
    k_init<<<B, T>>>(nt, perm, seg);
    {
        const int zero = 0;
        RTB_TRY(cudaMemcpy(n.start, &zero, sizeof(int), cudaMemcpyHostToDevice));
        RTB_TRY(cudaMemcpy(n.end, &nt, sizeof(int), cudaMemcpyHostToDevice));
    }
    {
        int base = 0, count = 1, largest = nt;
        while (count > 0) {
            level_base.push_back(base);
            level_count.push_back(count);
            const int NB = (count + T - 1) / T;
            k_level_reset<<<NB, T>>>(n, base, count);
            k_bbox<<<B, T>>>(nt, perm, seg, V, I, n);
            k_split<<<NB, T>>>(n, base, count);
            k_flags<<<B, T>>>(nt, perm, seg, V, I, n, flag);
            RTB_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, flag, S, nt + 1));
            k_next_init<<<B, T>>>(nt, nx);
            k_next_link<<<B, T>>>(nt, seg, flag, S, n, nx);
            int rounds = 1;
            while ((1 << rounds) < largest) rounds++;
            for (int r = 0; r < rounds; r++) {
                k_jump<<<B, T>>>(nt, nx, nx2);
                std::swap(nx, nx2);
            }
            RTB_TRY(cudaMemset(stats + 2, 0, sizeof(int)));
            k_decide<<<NB, T>>>(n, base, count, cnt, stats);
            RTB_TRY(cudaMemset(cnt + count, 0, sizeof(int)));
            RTB_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, off, count + 1));
            int n_children = 0;
            RTB_TRY(cudaMemcpy(&n_children, off + count, sizeof(int), cudaMemcpyDeviceToHost));
            RTB_TRY(cudaMemcpy(h_stats, stats, 3 * sizeof(int), cudaMemcpyDeviceToHost));
            if (base + count + n_children > max_nodes) {
                err = cudaErrorMemoryAllocation;
                goto done;
            }
            k_children<<<NB, T>>>(n, base, count, off, base + count);
            k_scatter<<<B, T>>>(nt, perm, seg, flag, S, nx, n, perm2, seg2);
            std::swap(perm, perm2);
            std::swap(seg, seg2);
            base += count;
            count = n_children;
            largest = std::max(h_stats[2], 1);
        }
        total = base;
    }
    for (int l = (int)level_base.size() - 1; l >= 0; l--) k_sizes<<<(level_count[l] + T - 1) / T, T>>>(n, level_base[l], level_count[l]);
    RTB_TRY(cudaMemset(n.pre, 0, sizeof(int)));
    for (size_t l = 0; l < level_base.size(); l++) k_preorder<<<(level_count[l] + T - 1) / T, T>>>(n, level_base[l], level_count[l]);
    RTB_TRY(cudaMalloc(&arr, (size_t)total * 10 * sizeof(float)));
    k_emit<<<(total + T - 1) / T, T>>>(n, total, arr);
    RTB_TRY(cudaEventRecord(e1));
    RTB_TRY(cudaDeviceSynchronize());
    RTB_TRY(cudaGetLastError());
    {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (build_ms) *build_ms = ms;
    }
    if (keep) {
        RTB_TRY(cudaMalloc(&R_out, (size_t)nt * 10 * sizeof(int)));
        rtrelayout::gather_records<<<(unsigned)(((size_t)nt * 10 + T - 1) / T), T>>>(R_in, perm, nt, R_out);
        RTB_TRY(cudaDeviceSynchronize());
        RTB_TRY(cudaGetLastError());
        keep->device = device;
        keep->V = V;
        keep->recs = R_out;
        keep->arr = arr;
        keep->nv = nv;
        keep->nt = nt;
        keep->nn = total;
        V = nullptr; /* handed over */
        R_out = nullptr;
        arr = nullptr;
    } else {
        perm_out.resize(nt);
        arr_out.resize((size_t)total * 10);
        RTB_TRY(cudaMemcpy(perm_out.data(), perm, (size_t)nt * sizeof(int), cudaMemcpyDeviceToHost));
        RTB_TRY(cudaMemcpy(arr_out.data(), arr, (size_t)total * 10 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    info[0] = total;
    info[1] = h_stats[0];
    info[2] = (int32_t)level_base.size();
    info[3] = h_stats[1];
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    void* frees[] = {R_in, R_out, V, arr, I, perm, perm2, seg, seg2, flag, S, nx, nx2, cnt, off, stats, tmp, n.start, n.end, n.kmn, n.kmx, n.left, n.right, n.axis, n.split, n.nL, n.size, n.pre};
    for (void* p : frees)
        if (p) cudaFree(p);
    if (prev >= 0) cudaSetDevice(prev);
    return (int)err;
}

} // namespace rtbuild
