/*
 * rt_bins.cuh — candidate lists for rays whose line passes through a fixed point (an "anchor").
 *
 * Two families of rays of the reference's path are anchored: the camera rays (all start at Scene camera C,
 * optimized.cu:747-759) and the shadow rays (all aim at the light L, optimized.cu:616-618) — every ray of the
 * primary + shadow configuration, and (1 + B) of the 2 B rays of a B-segment path.
 *
 * What the reference's mesh query computes for a ray is fixed by the LEAVES whose own box passes
 * BoundingBox::intersect (rt_layout.h, "Wide index": a child box that passes implies its parent passes, because the
 * test is monotone under IEEE rounding): every triangle of those leaves is tested, the strict minimum of the accepted t
 * wins. The inner boxes only serve to find those leaves. For anchored rays they can be found without a tree: the
 * directions around the anchor are cut into 3 faces (dominant axis x, y or z; d and -d share a cell because the
 * slab test knows no t >= 0) of R x R cells of the two ratios d_a / d_k, d_b / d_k in [-1, 1], and every cell lists the
 * leaves whose box — inflated by eps — meets a line through the anchor with a direction of that cell. A ray looks up the cell of its direction, applies the reference's own slab test (certified
 * fast path + exact fallback, rt_math.cuh) to the listed boxes, and emits one (ray, leaf) task per box that passes;
 * wf_leaves then tests the triangles. Versus the tree search: 2.3 box tests per leaf found instead of 9, no task pools,
 * no dependent chain of tree levels, and the triangle work arrives as uniform independent tasks.
 *
 * Superset property (why no leaf is lost). If the slab test of ray (O, u) passes for box [mn, mx], every pair of the
 * six computed distances satisfies t1_j > t0_k; each is the correctly rounded quotient of a correctly rounded difference,
 * relative error < 2^-23, so the true per-axis parameter intervals of the box inflated by 2^-21 (S + |O|) (S = largest
 * box coordinate) intersect pairwise, hence (intervals on a line) have a common point: the true line meets the inflated
 * box. A camera ray's line passes through the anchor exactly. A shadow ray starts at P' and has the direction
 * fl((L - P') / |L - P'|), within 2^-21 of the true one: over the distances involved (D = |L - P'| plus the extent of
 * the scene) its line stays within 2^-20 (D + S + |L|) of the line through P' and L. With eps = 2^-12 (S + |anchor|) and
 * the guard D <= 2^7 (S + |anchor|) on the ray side, both are covered with a factor > 2 to spare; rays beyond the guard,
 * and rays with a zero, subnormal or non-finite direction component (where a slab distance may be NaN and the
 * implication child => parent fails), are answered by the exact two-child traversal instead (wf_exact_query).
 * The cell of a ray is computed with a rounding-level error in the ratios (a few 2^-23): the ratio bounds of a leaf are widened
 * by 2^-20, still inside what the inflation leaves unused (eps / distance >= 2^-13 in ratio units).
 * tools/proto_bins.py checks the property on every ray of a frame on the CPU; the GPU parity tests check the results.
 *
 * Leaves whose inflated box contains the anchor have no bounded set of cells; a scene that has one keeps the tree
 * search for that anchor's rays (the builder reports it).
 */
#pragma once
#include "rt_layout.h"
#include "rt_math.cuh"

namespace rtk {

/* Only the part of every face that the mesh's root box can project to has cells (the "window": for a light or camera some
 * way off the mesh a few per cent of the face), so that building the lists — done again whenever an anchor moves — touches
 * kilobytes, not the whole direction space. A direction outside the window has no candidates. */
struct BinsWindow {
    int ca0, cb0, wa, wb; /* first cell and extent of the window of a face (wa == 0: the mesh cannot be met on this face) */
    int base;             /* index of the window's first cell in cell_start */
};

struct BinsView {       /* by value in kernel arguments */
    float ax, ay, az;   /* the anchor */
    float eps;          /* inflation of the leaf boxes */
    float max_D2;       /* guard for rays that do not start at the anchor: (distance origin -> anchor)^2 must not exceed it */
    int R;              /* cells per face side */
    BinsWindow win[3];
    const int* cell_start; /* (cells of the three windows) + 1 */
    const int* items;      /* leaf-table indices */
    int items_cap;         /* entries of `items`; a list that ends beyond it was cut short by the build (exact search instead) */
    const int* status;     /* [0] != 0: a leaf box contains the anchor, the lists are not complete (exact search instead) */
};

/* cell of a direction (from or towards the anchor: the sign cancels in the ratios); -1 outside the windows */
__device__ __forceinline__ int bins_cell(const BinsView& b, F3 d) {
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int k;
    float dk, da, db;
    if (ax >= ay && ax >= az) { k = 0; dk = d.x; da = d.y; db = d.z; }
    else if (ay >= az)        { k = 1; dk = d.y; da = d.z; db = d.x; }
    else                      { k = 2; dk = d.z; da = d.x; db = d.y; }
    const float r = rcp_approx(dk);
    const float h = 0.5f * (float)b.R;
    const int ca = min(max((int)floorf(__fmaf_rn(da * r, h, h)), 0), b.R - 1);
    const int cb = min(max((int)floorf(__fmaf_rn(db * r, h, h)), 0), b.R - 1);
    const BinsWindow w = b.win[k];
    const unsigned ia = (unsigned)(ca - w.ca0), ib = (unsigned)(cb - w.cb0);
    if (ia >= (unsigned)w.wa || ib >= (unsigned)w.wb) return -1;
    return w.base + (int)ib * w.wa + (int)ia;
}

/* The rectangle of cells of face k that a box (v0, v1: its inflated corners relative to the anchor) can project to through
 * the anchor, for the v_k > 0 side (part 0) or the v_k < 0 side (part 1); interval arithmetic, inclusion-monotone: the
 * rectangle of a box inside another box lies inside the other's rectangle. false: no cell. */
__host__ __device__ inline bool bins_face_rect(const float v0[3], const float v1[3], int k, int part, int R, int& ca0, int& ca1, int& cb0, int& cb1) {
    const int a = (k + 1) % 3, b = (k + 2) % 3;
    const float mina = (v0[a] <= 0.f && v1[a] >= 0.f) ? 0.f : fminf(fabsf(v0[a]), fabsf(v1[a]));
    const float minb = (v0[b] <= 0.f && v1[b] >= 0.f) ? 0.f : fminf(fabsf(v0[b]), fabsf(v1[b]));
    /* on face k the direction's k component dominates: |v_k| >= |v_a|, |v_b| >= m. (m == 0 with v_k straddling 0 would be a
     * box around the anchor, excluded by the caller up to the degenerate flat cases, which the tiny bound handles
     * conservatively: the ratios then cover the whole face.) */
    const float m = fmaxf(fmaxf(mina, minb), 1e-30f);
    float k0, k1;
    if (part == 0) { k0 = fmaxf(v0[k], m); k1 = v1[k]; }   /* v_k > 0 side */
    else           { k0 = v0[k]; k1 = fminf(v1[k], -m); }  /* v_k < 0 side */
    if (!(k0 <= k1)) return false;
    float ra0 = 1e30f, ra1 = -1e30f, rb0 = 1e30f, rb1 = -1e30f;
    for (int c = 0; c < 4; c++) {
        const float kk = (c & 1) ? k1 : k0;
        const float qa = ((c & 2) ? v1[a] : v0[a]) / kk, qb = ((c & 2) ? v1[b] : v0[b]) / kk;
        ra0 = fminf(ra0, qa); ra1 = fmaxf(ra1, qa);
        rb0 = fminf(rb0, qb); rb1 = fmaxf(rb1, qb);
    }
    ra0 = fmaxf(ra0, -1.f); ra1 = fminf(ra1, 1.f);
    rb0 = fmaxf(rb0, -1.f); rb1 = fminf(rb1, 1.f);
    if (!(ra0 <= ra1) || !(rb0 <= rb1)) return false;
    /* RATIO_SLACK covers the roundings of these bounds and of the ray's own cell computation (a few 2^-23, relative, on
     * values <= 1): far below what the inflation by eps leaves unused (>= 2^-15 in ratio units, see the header) */
    const float RATIO_SLACK = 9.5367431640625e-07f; /* 2^-20 */
    const float h = 0.5f * (float)R;
    ca0 = (int)floorf(fmaf(ra0 - RATIO_SLACK, h, h)); ca1 = (int)floorf(fmaf(ra1 + RATIO_SLACK, h, h));
    cb0 = (int)floorf(fmaf(rb0 - RATIO_SLACK, h, h)); cb1 = (int)floorf(fmaf(rb1 + RATIO_SLACK, h, h));
    ca0 = ca0 < 0 ? 0 : ca0; cb0 = cb0 < 0 ? 0 : cb0;
    ca1 = ca1 > R - 1 ? R - 1 : ca1; cb1 = cb1 > R - 1 ? R - 1 : cb1;
    return ca0 <= ca1 && cb0 <= cb1;
}

/* The cells a leaf must be listed in: calls f(index in cell_start) for each (a cell may be visited twice; harmless). The
 * cells are dealt out to the `n_workers` threads that share the leaf (worker = this thread's index among them).
 * Returns false when the inflated box contains the anchor (no bounded set of cells). */
template <typename F>
__device__ __forceinline__ bool bins_leaf_cells(const float4 q0, const float4 q1, const BinsView& bv, int worker, int n_workers, F f) {
    const float v0[3] = {q0.x - bv.eps - bv.ax, q0.y - bv.eps - bv.ay, q0.z - bv.eps - bv.az};
    const float v1[3] = {q0.w + bv.eps - bv.ax, q1.x + bv.eps - bv.ay, q1.y + bv.eps - bv.az};
    if (v0[0] <= 0.f && v1[0] >= 0.f && v0[1] <= 0.f && v1[1] >= 0.f && v0[2] <= 0.f && v1[2] >= 0.f) return false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const BinsWindow w = bv.win[k];
#pragma unroll
        for (int part = 0; part < 2; part++) {
            int ca0, ca1, cb0, cb1;
            if (!bins_face_rect(v0, v1, k, part, bv.R, ca0, ca1, cb0, cb1)) continue;
            /* inside the window by construction (a leaf box lies in the root box); clipped all the same */
            ca0 = max(ca0, w.ca0); ca1 = min(ca1, w.ca0 + w.wa - 1);
            cb0 = max(cb0, w.cb0); cb1 = min(cb1, w.cb0 + w.wb - 1);
            if (ca0 > ca1 || cb0 > cb1) continue;
            const int wa = ca1 - ca0 + 1, n_cells = wa * (cb1 - cb0 + 1);
            for (int t = worker; t < n_cells; t += n_workers) f(w.base + (cb0 - w.cb0 + t / wa) * w.wa + (ca0 - w.ca0 + t % wa));
        }
    }
    return true;
}

/* both passes: BINS_GROUP threads per leaf (a leaf near the anchor covers thousands of cells, and the fill pass waits for
 * every atomic's return value) */
#define BINS_GROUP 256
/* pass 1: how many leaves does every cell list? status[0] = 1 when a leaf contains the anchor */
__global__ void bins_count(const float4* __restrict__ leaves, int n_leaves, const BinsView bv, int* __restrict__ counts, int* __restrict__ status) {
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) / BINS_GROUP;
    if (l >= n_leaves) return;
    const float4 q0 = leaves[2 * (size_t)l], q1 = leaves[2 * (size_t)l + 1];
    if (!bins_leaf_cells(q0, q1, bv, threadIdx.x % BINS_GROUP, BINS_GROUP, [&](int cell) { atomicAdd(counts + cell, 1); })) status[0] = 1;
}

/* pass 2 (after the exclusive scan of counts into cell_start): fill the lists. A leaf takes the slot
 * cell_start[cell] + (what is left of the cell's count) - 1, counting the counts back down to zero: the count array needs
 * no clearing before the next build. */
__global__ void bins_fill(const float4* __restrict__ leaves, int n_leaves, const BinsView bv, const int* __restrict__ cell_start, int* __restrict__ counts,
                          int* __restrict__ items, int items_cap, int* __restrict__ status) {
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) / BINS_GROUP;
    if (l >= n_leaves) return;
    const float4 q0 = leaves[2 * (size_t)l], q1 = leaves[2 * (size_t)l + 1];
    bins_leaf_cells(q0, q1, bv, threadIdx.x % BINS_GROUP, BINS_GROUP, [&](int cell) {
        const int at = cell_start[cell] + atomicSub(counts + cell, 1) - 1;
        if ((unsigned)at < (unsigned)items_cap) items[at] = l; /* unsigned: a scan that overflowed 2^31 gives negative offsets */
        else status[1] = 1; /* the lists outgrew the buffer (the anchor moved): rays of the cut lists take the exact search, the host enlarges it */
    });
}

} // namespace rtk
