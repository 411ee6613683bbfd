/*
 * rt_wavefront.cuh — the production render path: the per-pixel path of the reference (optimized.cu:746-771,
 * 539-661) split into streaming "owner" kernels and one persistent traversal kernel that drains global ray
 * queues resident in HBM/L2.
 *
 * Measured motivation (profiles/r01_notes.md): for BASELINE.json config 2, 9 % of the pixels hold 79 % of the BVH
 * steps and single rays through the cat's head take 600-870 steps against a mean of 12. One thread per pixel
 * (render_mega) ran as long as its slowest warp with the SMs 38 % busy; a per-tile shared-memory queue
 * (render_wave) still left whole tiles as the unit of imbalance (SMs 55 % busy). The reference algorithm fixes
 * WHICH boxes and triangles a ray tests (no pruning is possible: the winner is the strictly smallest COMPUTED t,
 * and the computed t of a grazing triangle is not bounded by its box) — but not who tests them, nor when. So:
 *
 *   wf_generate  one thread per pixel: ray generation, the six sphere tests, the root-box test; pixels whose ray
 *                misses the mesh's root box are shaded on the spot (and may post a shadow query); the others
 *                post a closest-hit query.                                     [streaming, full occupancy]
 *   wf_traverse  persistent warps drain the round's queues: an idle lane takes the next queued ray, and when
 *                the queue is empty it takes over a pending subtree from the stack of a busy lane of its warp
 *                (ballot-matched thief/donor pairs) — subtrees of one ray are independent searches whose
 *                results merge with atomicMin on (t bits, tie-break rank), the reference's strict-minimum +
 *                first-visited rule made order-free. Node steps and triangle steps run in separate
 *                warp-uniform phases. Shadow queries stop at the first blocker and paint the pixel black.
 *   wf_shade     one thread per answered closest-hit query: merge with the spheres, shade, bounce or post the
 *                shadow query of the next round.                                [streaming]
 *
 * A frame with S path segments is 2 + 2 S launches (4 for primary + shadow). Queue entries are 48 B
 * (3 x float4); a 1080p frame moves about 60 MB through L2.
 */
#pragma once
#include "rt_kernels.cuh"

namespace rtk {

#define WF_MODE_CLOSEST 1
#define WF_MODE_ANY 2
#define WF_MAX_ROUNDS 12
#define WF_THREADS 128
#define WF_NOHIT 0xffffffffffffffffull

struct __align__(16) QEntry {
    float ox, oy, oz, aux; /* aux: closest -> t of the closest sphere of this segment; any -> |L - P'|^2 */
    float ux, uy, uz;
    int pixel;             /* compact pixel index (row_in_call * W + column) */
    float n_ray;           /* refraction index the ray travels in (Ray::refraction_index) */
    int packed;            /* sphere slot (8 bits, 0xff none) | depth << 8 | mode << 24 */
    unsigned long long res; /* closest: (t bits << 32) | tie-break rank, WF_NOHIT = none; any: != 0 = blocked */
};
static_assert(sizeof(QEntry) == 48, "QEntry is three float4");

struct WfCounters {
    unsigned long long stats[8]; /* rays, node visits, triangle tests, max stack, slab fallbacks, exact triangle evals */
    int nA[WF_MAX_ROUNDS + 2];   /* closest-hit queries posted for round r */
    int nS[WF_MAX_ROUNDS + 2];   /* shadow queries posted for round r */
    int head[WF_MAX_ROUNDS + 2]; /* traversal fetch cursor of round r */
};

struct WfArgs {
    RenderArgs a;
    QEntry* qA[2]; /* closest-hit queue of round r lives in qA[r & 1] */
    QEntry* qS;    /* shadow queue of the current round */
    WfCounters* c;
    int round;     /* round whose queues this launch consumes (traverse, shade) */
};

__device__ __forceinline__ unsigned tie_rank(int i, int leaf_start, int n_tris, int push_order, int off_bits) {
    /* smaller wins at equal t. push_order 1 (L popped first, optimized.cu:265-266): ascending triangle index.
     * push_order 0 (R popped first, cpu_launcher.cpp:291-292): descending leaf order, ascending index inside a
     * leaf -> ((n_tris - leaf_start) << off_bits) | (i - leaf_start). */
    return push_order == 1 ? (unsigned)i : (((unsigned)(n_tris - leaf_start)) << off_bits) | (unsigned)(i - leaf_start);
}
__device__ __forceinline__ int rank_to_tri(unsigned rank, int n_tris, int push_order, int off_bits) {
    if (push_order == 1) return (int)rank;
    const int leaf_start = n_tris - (int)(rank >> off_bits);
    return leaf_start + (int)(rank & ((1u << off_bits) - 1u));
}

/* append one entry to a global queue; the lanes that reach this point together share one atomic */
__device__ __forceinline__ int queue_slot(int* counter) {
    const unsigned m = __activemask();
    const int leader = __ffs(m) - 1;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void store_entry(QEntry* q, int slot, F3 O, F3 u, float aux, int pixel, float n_ray, int packed, unsigned long long res) {
    float4* p = reinterpret_cast<float4*>(q + slot);
    p[0] = make_float4(O.x, O.y, O.z, aux);
    p[1] = make_float4(u.x, u.y, u.z, __int_as_float(pixel));
    p[2] = make_float4(n_ray, __int_as_float(packed), __uint_as_float((unsigned)(res & 0xffffffffull)), __uint_as_float((unsigned)(res >> 32)));
}

/* sample average + transfer function + store (optimized.cu:762-771) */
__device__ __forceinline__ void write_pixel(const RenderArgs& a, int px, F3 color) {
    if (!a.rgb) return;
    F3 total = f3(0.f, 0.f, 0.f);
    for (int s = 0; s < a.num_rays; s++) total = total + color;
    const F3 avg = a.num_rays == 1 ? total : total / (float)a.num_rays; /* x / 1.0f == x */
    const float* T = a.gamma_tab + a.gamma_mode * 256;
    a.rgb[(size_t)px * 3 + 0] = (uint8_t)quantise(avg.x, T);
    a.rgb[(size_t)px * 3 + 1] = (uint8_t)quantise(avg.y, T);
    a.rgb[(size_t)px * 3 + 2] = (uint8_t)quantise(avg.z, T);
}

__device__ __forceinline__ void closest_sphere(const SceneHeader& h, F3 O, F3 u, float& ts, int& sidx) {
    ts = RTK_INF;
    sidx = -1;
    for (int k = 0; k < h.n_spheres; k++) { /* ascending id, strict < (optimized.cu:543-554) */
        float t;
        if (sphere_t(h.spheres[k], O, u, t) && t < ts) {
            ts = t;
            sidx = k;
        }
    }
}

/* Advance one pixel's path until it ends or needs the mesh. have_hit: (t_hit, sidx, tri) already hold the
 * answer of intersect_all for the current ray (wf_shade); otherwise the segment starts here. */
template <bool COUNT>
__device__ __forceinline__ void path_advance(const SceneHeader& h, const WfArgs& g, const float4* __restrict__ nhat, int post_round, int px, F3 O, F3 u,
                                             float n_ray, int depth, bool have_hit, float t_hit, int sidx, int tri, Work& w) {
    const RenderArgs& a = g.a;
    const F3 Lp = f3(h.L[0], h.L[1], h.L[2]);
    const float eps = a.eps_surface;
    for (;;) {
        if (!have_hit) {
            if (depth >= a.segments) { /* path budget used up without a diffuse hit: colour 0 (fold of optimized.cu:653-660) */
                write_pixel(a, px, f3(0.f, 0.f, 0.f));
                return;
            }
            w.rays++;
            closest_sphere(h, O, u, t_hit, sidx);
            tri = -1;
            if (h.has_mesh) {
                const RayCtx ctx = make_ray_ctx(O, u, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                float tn;
                if (slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tn, w.slab_fallbacks)) {
                    const int slot = queue_slot(&g.c->nA[post_round]);
                    store_entry(g.qA[post_round & 1], slot, O, u, t_hit, px, n_ray, (sidx & 0xff) | (depth << 8) | (WF_MODE_CLOSEST << 24), WF_NOHIT);
                    return;
                }
            }
        }
        have_hit = false;
        /* ---- the hit of this segment is known: shade it (optimized.cu:571-650) ---------------------------------- */
        const int obj = tri >= 0 ? h.mesh_id : (sidx >= 0 ? h.spheres[sidx].id : -1);
        if (depth == 0) {
            if (a.hit_obj) a.hit_obj[px] = obj;
            if (a.hit_tri) a.hit_tri[px] = tri;
            if (a.hit_t) a.hit_t[px] = t_hit;
        }
        depth++;
        if (obj < 0) { /* the ray left the scene */
            write_pixel(a, px, f3(0.f, 0.f, 0.f));
            return;
        }
        const F3 P = O + t_hit * u; /* :555 */
        F3 N, albedo;
        int mirror;
        float n_in, n_out;
        if (tri < 0) {
            const DevSphere& s = h.spheres[sidx];
            N = normalized(P - f3(s.cx, s.cy, s.cz)); /* :132-133 */
            albedo = f3(s.ax, s.ay, s.az);
            mirror = s.mirror;
            n_in = s.n_in;
            n_out = s.n_out;
        } else {
            const float4 nh = __ldg(nhat + tri); /* N.normalize() :282, precomputed */
            N = f3(nh.x, nh.y, nh.z);
            albedo = f3(h.mesh_albedo[0], h.mesh_albedo[1], h.mesh_albedo[2]);
            mirror = h.mesh_mirror;
            n_in = h.mesh_n_in;
            n_out = h.mesh_n_out;
        }
        if (mirror) { /* :572-579 */
            const F3 nu = u - (2 * dot(u, N)) * N;
            O = P + eps * N;
            u = nu;
        } else if (n_in != n_out) { /* :580-609 */
            float ratio;
            const bool out2in = n_ray == n_out;
            if (out2in) {
                ratio = n_out / n_in;
            } else {
                ratio = n_in / n_out;
                N = -N;
            }
            const float un = dot(u, N);
            if (((out2in && n_ray > n_in) || (!out2in && n_ray > n_out)) && (ratio * ratio) * (1 - un * un) > 1) {
                const F3 nu = u - (2 * un) * N; /* total internal reflection :596-600 */
                O = P + eps * N;
                u = nu;
            } else {
                const F3 Ncomp = (-sqrtf(1 - (ratio * ratio) * (1 - un * un))) * N;
                const F3 Tcomp = ratio * (u - un * N);
                O = P - eps * N;
                u = Ncomp + Tcomp;
                n_ray = out2in ? n_in : n_out;
            }
        } else { /* diffuse :610-650 — the deterministic path ends here */
            const F3 Padj = P + eps * N;
            const F3 toL = Lp - Padj;
            const float D2 = norm2(toL);
            const F3 su = toL / sqrtf(D2); /* NORMED_VEC :618 */
            /* shadow ray: the spheres here, the mesh through the queue (see mesh_query for the equivalence) */
            w.rays++;
            bool blocked = false;
            for (int s = 0; s < h.n_spheres && !blocked; s++) {
                float t;
                if (sphere_t(h.spheres[s], Padj, su, t) && t < RTK_INF && blocks_light(Padj, su, t, D2)) blocked = true;
            }
            if (blocked) {
                write_pixel(a, px, f3(0.f, 0.f, 0.f));
                if (a.shadow) a.shadow[px] = 1;
                return;
            }
            const F3 PL = Lp - P;
            const F3 wl = normalized(PL);
            const float ndl = dot(N, wl);
            const float lambert = (ndl < 0.f) ? 0.f : ndl; /* std::max(dot, 0.f) */
            const float l = (float)((double)h.intensity / (12.566370614359172 * (double)norm2(PL)) * (double)lambert); /* :628, in double */
            write_pixel(a, px, (l * albedo) / 3.14159274f);                                                      /* :629 */
            if (a.shadow) a.shadow[px] = 0;
            if (h.has_mesh) {
                const RayCtx ctx = make_ray_ctx(Padj, su, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                float tn;
                if (slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tn, w.slab_fallbacks)) {
                    /* the pixel keeps its lit colour unless the traversal finds a blocker, which paints it black */
                    const int slot = queue_slot(&g.c->nS[post_round]);
                    store_entry(g.qS, slot, Padj, su, D2, px, 1.f, 0xff | (depth << 8) | (WF_MODE_ANY << 24), 0ull);
                }
            }
            return;
        }
    }
}

__device__ __forceinline__ void flush_work(const Work& w, WfCounters* c, bool count) {
    const int lane = threadIdx.x & 31;
    const unsigned rays = __reduce_add_sync(0xffffffffu, w.rays);
    if (lane == 0 && rays) atomicAdd(&c->stats[0], (unsigned long long)rays);
    if (count) {
        const unsigned nn = __reduce_add_sync(0xffffffffu, w.nodes), tt = __reduce_add_sync(0xffffffffu, w.tris);
        const unsigned ms = __reduce_max_sync(0xffffffffu, w.max_stack);
        const unsigned sf = __reduce_add_sync(0xffffffffu, w.slab_fallbacks), te = __reduce_add_sync(0xffffffffu, w.tri_exact);
        if (lane == 0) {
            if (nn) atomicAdd(&c->stats[1], (unsigned long long)nn);
            if (tt) atomicAdd(&c->stats[2], (unsigned long long)tt);
            atomicMax(&c->stats[3], (unsigned long long)ms);
            if (sf) atomicAdd(&c->stats[4], (unsigned long long)sf);
            if (te) atomicAdd(&c->stats[5], (unsigned long long)te);
        }
    }
}

/* ---- wf_generate: one thread per pixel (a warp covers an 8x4 tile) ------------------------------------------------ */
template <bool COUNT>
__global__ void __launch_bounds__(WF_THREADS) wf_generate(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                         const __grid_constant__ WfArgs g) {
    const RenderArgs& a = g.a;
    const float4* nhat = reinterpret_cast<const float4*>(blob + h.off_nhat);
    const int lane = threadIdx.x & 31;
    const int wt = blockIdx.x * (WF_THREADS / 32) + (threadIdx.x >> 5);
    const int tiles_x = (a.W + 7) >> 3;
    const int j = (wt % tiles_x) * 8 + (lane & 7);
    const int kr = (wt / tiles_x) * 4 + (lane >> 3);
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    if (j < a.W && kr < a.rows) {
        const int px = kr * a.W + j;
        const int i = a.row_begin + kr * a.row_step;
        const F3 uc = f3((float)j - (float)a.W / 2 + 0.5f, (float)a.H / 2 - (float)i - 0.5f, a.z); /* optimized.cu:751, exact in float */
        const F3 u0 = normalized(uc); /* sigma == 0: the jitter terms of :758 are exactly 0 */
        if (a.hit_obj) a.hit_obj[px] = -1;
        if (a.hit_tri) a.hit_tri[px] = -1;
        if (a.hit_t) a.hit_t[px] = RTK_INF;
        if (a.shadow) a.shadow[px] = 2;
        path_advance<COUNT>(h, g, nhat, 0, px, f3(a.camx, a.camy, a.camz), u0, 1.f, 0, false, 0.f, -1, -1, w);
    }
    flush_work(w, g.c, COUNT);
}

/* ---- wf_shade: one thread per answered closest-hit query of round g.round ------------------------------------------ */
template <bool COUNT>
__global__ void __launch_bounds__(WF_THREADS) wf_shade(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                      const __grid_constant__ WfArgs g) {
    const RenderArgs& a = g.a;
    const float4* nhat = reinterpret_cast<const float4*>(blob + h.off_nhat);
    const int n = g.c->nA[g.round];
    const QEntry* q = g.qA[g.round & 1];
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    const int n_round = (n + 31) & ~31; /* whole warps stay in the loop so that flush_work sees 32 lanes */
    for (int e = blockIdx.x * WF_THREADS + threadIdx.x; e < n_round; e += gridDim.x * WF_THREADS) {
        if (e >= n) continue;
        const float4* p = reinterpret_cast<const float4*>(q + e);
        const float4 p0 = p[0], p1 = p[1], p2 = p[2];
        const F3 O = f3(p0.x, p0.y, p0.z), u = f3(p1.x, p1.y, p1.z);
        float t_hit = p0.w;
        const int px = __float_as_int(p1.w);
        const int packed = __float_as_int(p2.y);
        int sidx = packed & 0xff;
        if (sidx == 0xff) sidx = -1;
        const int depth = (packed >> 8) & 0xffff;
        const unsigned long long key = ((unsigned long long)__float_as_uint(p2.w) << 32) | __float_as_uint(p2.z);
        int tri = -1;
        if (key != WF_NOHIT) { /* mesh vs spheres: ascending id, strict < (optimized.cu:549) */
            const float tm = __uint_as_float((unsigned)(key >> 32));
            const int sid = sidx >= 0 ? h.spheres[sidx].id : -1;
            if (tm < t_hit || (tm == t_hit && h.mesh_id < sid)) {
                t_hit = tm;
                sidx = -1;
                tri = rank_to_tri((unsigned)key, h.n_tris, a.push_order, a.rank_off_bits);
            }
        }
        path_advance<COUNT>(h, g, nhat, g.round + 1, px, O, u, p2.x, depth, true, t_hit, sidx, tri, w);
    }
    flush_work(w, g.c, COUNT);
}

/* ---- wf_traverse: persistent warps drain the closest-hit and shadow queues of round g.round ------------------------ */
struct LaneTask {
    RayCtx ctx;
    float t_best;       /* closest: best t of THIS lane's subtrees; any: t limit */
    float d2;
    unsigned rank_best;
    int entry;          /* index into the round's queues (closest entries first, then shadow entries) */
    int pixel;
    int mode;
    int2 cur;
    int ti;             /* next triangle of the current leaf */
    int sp;
    bool busy;
    bool shared;        /* other lanes of the warp work on subtrees of the same ray */
};

template <bool COUNT>
__global__ void __launch_bounds__(WF_THREADS) wf_traverse(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                         const __grid_constant__ WfArgs g) {
    const RenderArgs& a = g.a;
    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int nA = g.c->nA[g.round], nS = g.c->nS[g.round];
    const int total = nA + nS;
    QEntry* qA = g.qA[g.round & 1];
    QEntry* qS = g.qS;
    int* head = &g.c->head[g.round];
    const int2 root = make_int2(h.root_a, h.root_b);
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;

    LaneTask k;
    k.busy = false;
    k.shared = false;
    k.sp = 0;
    k.entry = 0;
    k.pixel = 0;
    k.mode = 0;
    k.cur = root;
    k.ti = 0;
    int2 stack[RT_STACK_CAP];
    int rs_next = 0, rs_end = 0; /* this warp's reserved range of queue indices (warp-uniform) */
    bool exhausted = total == 0;

    auto entry_ptr = [&](int e) -> QEntry* { return e < nA ? (qA + e) : (qS + (e - nA)); };
    auto begin = [&](int e, int2 ref, bool shared) {
        const float4* p = reinterpret_cast<const float4*>(entry_ptr(e));
        const float4 p0 = __ldcg(p), p1 = __ldcg(p + 1);
        k.entry = e;
        k.mode = e < nA ? WF_MODE_CLOSEST : WF_MODE_ANY;
        k.pixel = __float_as_int(p1.w);
        k.ctx = make_ray_ctx(f3(p0.x, p0.y, p0.z), f3(p1.x, p1.y, p1.z), h.box_abs[0], h.box_abs[1], h.box_abs[2]);
        k.d2 = p0.w;
        /* any: a hit with t > 1.001 sqrt(D2) cannot satisfy the shadow predicate (|t u| ~ t) */
        k.t_best = (k.mode == WF_MODE_ANY) ? (sqrtf(k.d2) * 1.001f + 1e-3f) : RTK_INF;
        k.rank_best = 0xffffffffu;
        k.cur = ref;
        k.ti = ref.x;
        k.sp = 0;
        k.busy = true;
        k.shared = shared;
    };
    auto finish = [&]() { /* publish this lane's result for its ray and become idle */
        if (k.mode == WF_MODE_CLOSEST && k.rank_best != 0xffffffffu) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(k.t_best) << 32) | k.rank_best;
            atomicMin(&entry_ptr(k.entry)->res, key);
        }
        k.busy = false;
    };
    auto pop = [&]() { /* next pending subtree of this lane, or the end of its task */
        if (k.sp == 0 || (k.mode == WF_MODE_ANY && k.shared && *((volatile unsigned long long*)&entry_ptr(k.entry)->res) != 0ull)) {
            finish();
            return;
        }
        k.cur = stack[--k.sp];
        k.ti = k.cur.x;
    };

    for (;;) {
        /* ---- work distribution -------------------------------------------------------------------------------- */
        const unsigned idle = __ballot_sync(FULL, !k.busy);
        if (idle) {
            if (rs_next >= rs_end && !exhausted) { /* reserve the next 32 queue indices for this warp */
                int base = 0;
                if (lane == 0) base = atomicAdd(head, 32);
                base = __shfl_sync(FULL, base, 0);
                rs_next = base;
                rs_end = min(base + 32, total);
                if (base >= total) {
                    exhausted = true;
                    rs_next = rs_end = 0;
                }
            }
            const int avail = rs_end - rs_next;
            if (avail > 0) {
                const int r = __popc(idle & lt_mask);
                if (!k.busy && r < avail) begin(rs_next + r, root, false);
                rs_next += min(avail, __popc(idle));
            }
            /* nothing left to fetch: idle lanes take a pending subtree from busy lanes of this warp */
            const unsigned thieves = __ballot_sync(FULL, !k.busy);
            const unsigned donors = __ballot_sync(FULL, k.busy && k.sp > 0);
            if (thieves && donors) {
                const int pairs = min(__popc(thieves), __popc(donors));
                const int my_rank = __popc((k.busy ? donors : thieves) & lt_mask);
                const bool donate = k.busy && k.sp > 0 && my_rank < pairs;
                const bool steal = !k.busy && my_rank < pairs;
                int2 give = make_int2(0, 0);
                if (donate) {
                    give = stack[--k.sp];
                    k.shared = true;
                }
                const int src = steal ? (int)__fns(donors, 0, my_rank + 1) : lane;
                const int g_entry = __shfl_sync(FULL, k.entry, src);
                const int gx = __shfl_sync(FULL, give.x, src);
                const int gy = __shfl_sync(FULL, give.y, src);
                if (steal) begin(g_entry, make_int2(gx, gy), true);
            }
        }
        if (!__any_sync(FULL, k.busy)) {
            if (exhausted) break;
            continue;
        }

        /* ---- node phase: every lane whose current reference is an inner node steps it ----------------------------- */
        while (__any_sync(FULL, k.busy && k.cur.y < 0)) {
            if (k.busy && k.cur.y < 0) {
                const float4* n = nodes + 4 * (size_t)k.cur.x;
                const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2);
                const int4 q3 = __ldg(reinterpret_cast<const int4*>(n + 3));
                if (COUNT) w.nodes++;
                float tnL, tnR;
                const bool okL = slab_fast(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, k.ctx, tnL, w.slab_fallbacks);
                const bool okR = slab_fast(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, k.ctx, tnR, w.slab_fallbacks);
                int2 cl = make_int2(q3.x, q3.y), cr = make_int2(q3.z, q3.w);
                if (k.mode == WF_MODE_ANY && okL && okR && tnR < tnL) { /* nearest first: blockers are found sooner */
                    const int2 s = cl;
                    cl = cr;
                    cr = s;
                }
                if (okL) {
                    k.cur = cl;
                    k.ti = cl.x;
                    if (okR) {
                        stack[k.sp++] = cr;
                        if (COUNT) w.max_stack = max(w.max_stack, (unsigned)k.sp);
                    }
                } else if (okR) {
                    k.cur = cr;
                    k.ti = cr.x;
                } else {
                    pop();
                }
            }
        }
        /* ---- triangle phase: one triangle per lane per step ---------------------------------------------------------- */
        while (__any_sync(FULL, k.busy && k.cur.y >= 0)) {
            if (k.busy && k.cur.y >= 0) {
                if (k.ti < k.cur.y) {
                    const int i = k.ti++;
                    if (COUNT) w.tris++;
                    float t;
                    if (tri_fast(tris + 3 * (size_t)i, k.ctx.O, k.ctx.u, k.t_best, t, w.tri_exact) && t > a.eps_tri) {
                        if (k.mode == WF_MODE_ANY) {
                            if (blocks_light(k.ctx.O, k.ctx.u, t, k.d2)) { /* the light is blocked: the pixel is black (:620-622) */
                                if (a.rgb) {
                                    a.rgb[(size_t)k.pixel * 3 + 0] = 0;
                                    a.rgb[(size_t)k.pixel * 3 + 1] = 0;
                                    a.rgb[(size_t)k.pixel * 3 + 2] = 0;
                                }
                                if (a.shadow) a.shadow[k.pixel] = 1;
                                if (k.shared) *((volatile unsigned long long*)&entry_ptr(k.entry)->res) = 1ull;
                                k.sp = 0;
                                k.ti = k.cur.y; /* leaves the leaf; pop() ends the task */
                            }
                        } else {
                            const unsigned rank = tie_rank(i, k.cur.x, h.n_tris, a.push_order, a.rank_off_bits);
                            if (t < k.t_best || (t == k.t_best && rank < k.rank_best)) {
                                k.t_best = t;
                                k.rank_best = rank;
                            }
                        }
                    }
                } else {
                    pop();
                }
            }
            /* leave the phase early when inner-node work is waiting and few lanes still have triangles */
            const unsigned tri_lanes = __ballot_sync(FULL, k.busy && k.cur.y >= 0);
            const unsigned node_lanes = __ballot_sync(FULL, k.busy && k.cur.y < 0);
            if (node_lanes && __popc(tri_lanes) < 12) break;
        }
    }
    flush_work(w, g.c, COUNT);
}

} // namespace rtk
