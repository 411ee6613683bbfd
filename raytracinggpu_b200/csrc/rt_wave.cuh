/*
 * rt_wave.cuh — the production render kernel: persistent thread blocks, a shared-memory ray queue per pixel
 * tile, and a warp-ballot work distributor for the BVH traversal.
 *
 * Why not one thread per pixel start to finish (render_mega)? Measured on the B200 for BASELINE.json config 2
 * (profiles/r01_notes.md): 9 % of the pixels hold 79 % of the traversal steps, single rays through the cat's
 * head take 600-870 steps against a mean of 12, a warp runs as long as its slowest lane and the whole launch
 * as long as its slowest warp — the SMs were busy 38 % of the launch. The reference algorithm fixes WHICH
 * boxes and triangles a ray must test (no pruning: the winner is the strictly smallest computed t, and the
 * computed t of a grazing triangle is not bounded by its box), but not WHO tests them nor in which order. So:
 *
 *   - a block owns an 8x4 pixel tile at a time (tiles are handed out by an atomic counter: persistent
 *     blocks, natural load balance across the 148 SMs); warp 0 owns the 32 pixels, all four warps traverse, so
 *     a heavy tile gets four lanes per pixel;
 *   - per round every pixel's owner thread does the cheap, uniform part of the path (ray generation,
 *     six sphere tests, root-box test, shading; optimized.cu:746-760, 539-559, 561-661) and, if its ray
 *     enters the mesh's root box, posts a query into the tile's shared-memory queue;
 *   - then all warps of the block drain the queue together: an idle lane takes the next queued ray (ballot +
 *     one atomic per warp), and when the queue is empty it takes over a pending subtree from the traversal
 *     stack of a busy lane of its warp (ballot-matched thief/donor pairs). Subtrees of one ray are
 *     independent searches whose results merge with an atomicMin on (t bits, tie-break rank) — the
 *     reference's strict-minimum + first-visited rule made order-free — so a 800-step ray ends
 *     up spread over the 32 lanes of a warp instead of holding 31 of them idle;
 *   - node steps and triangle steps run in separate warp-uniform phases (while-while), so the lanes of a warp
 *     execute the same code on different nodes.
 * Shadow queries stop at the first blocker (see mesh_query in rt_kernels.cuh for why that is the reference's
 * result); a blocker found by one lane releases the other lanes working on the same ray.
 */
#pragma once
#include "rt_kernels.cuh"

namespace rtk {

#define WV_THREADS 128
#define WV_WARPS (WV_THREADS / 32)
#define WV_TILE_W 8
#define WV_TILE_H 4
#define WV_TILE_PIX 32 /* one owner warp per tile; all WV_WARPS warps drain its queue (4 lanes per pixel) */
#define WV_NOHIT 0xffffffffffffffffull

enum { WV_DONE = 0, WV_TRACE = 1, WV_WAIT_HIT = 2, WV_WAIT_SHADOW = 3 };
enum { WV_Q_CLOSEST = 1, WV_Q_ANY = 2 };

struct WaveSmem {
    float gamma[256];
    /* the pixel's current ray; it doubles as the query ray of slot == owner thread */
    float ox[WV_TILE_PIX], oy[WV_TILE_PIX], oz[WV_TILE_PIX];
    float ux[WV_TILE_PIX], uy[WV_TILE_PIX], uz[WV_TILE_PIX];
    float ts[WV_TILE_PIX];                 /* closest sphere t of the current segment */
    float d2[WV_TILE_PIX];                 /* shadow queries: |L - P'|^2 */
    unsigned long long res[WV_TILE_PIX];   /* closest: (t bits << 32) | rank, WV_NOHIT if none; any: 1 = blocked */
    int sidx[WV_TILE_PIX];
    int queue[WV_TILE_PIX];
    unsigned char qmode[WV_TILE_PIX];
    int q_count, q_head, tile;
};

/* tie-break rank in the low 32 bits of the closest-hit key (smaller wins at equal t):
 * push_order 1 (L popped first, optimized.cu:265-266): ascending triangle index -> rank = i.
 * push_order 0 (R popped first, cpu_launcher.cpp:291-292): leaves in descending order, ascending index inside a
 * leaf -> rank = ((n_tris - leaf_start) << off_bits) | (i - leaf_start); off_bits is chosen by the host so both
 * fields fit (RenderArgs::rank_off_bits; the host falls back to render_mega when a leaf is too large). */
__device__ __forceinline__ unsigned tie_rank(int i, int leaf_start, int n_tris, int push_order, int off_bits) {
    return push_order == 1 ? (unsigned)i : (((unsigned)(n_tris - leaf_start)) << off_bits) | (unsigned)(i - leaf_start);
}
__device__ __forceinline__ int rank_to_tri(unsigned rank, int n_tris, int push_order, int off_bits) {
    if (push_order == 1) return (int)rank;
    const int leaf_start = n_tris - (int)(rank >> off_bits);
    return leaf_start + (int)(rank & ((1u << off_bits) - 1u));
}

struct LaneTask {
    RayCtx ctx;
    float t_best;   /* closest: best t so far of THIS lane's subtrees; any: t limit */
    float d2;
    unsigned rank_best;
    int slot;
    int mode;
    int2 cur;
    int ti; /* next triangle of the current leaf */
    int sp;
    bool busy;
};

__device__ __forceinline__ void task_begin(LaneTask& k, const WaveSmem& sm, const SceneHeader& h, int slot, int2 ref) {
    k.slot = slot;
    k.mode = sm.qmode[slot];
    const F3 O = f3(sm.ox[slot], sm.oy[slot], sm.oz[slot]);
    const F3 u = f3(sm.ux[slot], sm.uy[slot], sm.uz[slot]);
    k.ctx = make_ray_ctx(O, u, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
    k.d2 = sm.d2[slot];
    /* any: a hit with t > 1.001 sqrt(D2) cannot satisfy the shadow predicate (|t u| ~ t) */
    k.t_best = (k.mode == WV_Q_ANY) ? (sqrtf(k.d2) * 1.001f + 1e-3f) : RTK_INF;
    k.rank_best = 0xffffffffu;
    k.cur = ref;
    k.ti = ref.x;
    k.sp = 0;
    k.busy = true;
}

/* publish this lane's result for its ray and become idle */
__device__ __forceinline__ void task_end(LaneTask& k, WaveSmem& sm) {
    if (k.mode == WV_Q_CLOSEST && k.rank_best != 0xffffffffu) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(k.t_best) << 32) | k.rank_best;
        atomicMin(&sm.res[k.slot], key);
    }
    k.busy = false;
}

/* next pending subtree of this lane, or the end of its task */
__device__ __forceinline__ void task_pop(LaneTask& k, const int2* stack, WaveSmem& sm) {
    if (k.sp == 0 || (k.mode == WV_Q_ANY && *((volatile unsigned long long*)&sm.res[k.slot]) != 0ull)) {
        task_end(k, sm);
        return;
    }
    k.cur = stack[--k.sp];
    k.ti = k.cur.x;
}

template <bool COUNT>
__device__ __forceinline__ void drain_queue(const SceneHeader& h, const float4* __restrict__ nodes, const float4* __restrict__ tris, WaveSmem& sm,
                                            const RenderArgs& a, Work& w) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int qn = sm.q_count;
    const int share = max(1, (qn + WV_WARPS - 1) / WV_WARPS);
    const int2 root = make_int2(h.root_a, h.root_b);
    LaneTask k;
    k.busy = false;
    k.sp = 0;
    k.slot = 0;
    k.mode = 0;
    k.cur = root;
    k.ti = 0;
    int2 stack[RT_STACK_CAP];

    for (;;) {
        /* ---- work distribution ---------------------------------------------------------------------------- */
        const unsigned idle = __ballot_sync(FULL, !k.busy);
        if (idle) {
            /* a warp takes its share of the queue, not all of it: the other warps of the block drain the same tile */
            const int take = min(__popc(idle), share);
            int base = qn;
            if (lane == 0 && *((volatile int*)&sm.q_head) < qn) base = atomicAdd(&sm.q_head, take);
            base = __shfl_sync(FULL, base, 0);
            if (!k.busy) {
                const int r = __popc(idle & lt_mask);
                const int my = base + r;
                if (r < take && my < qn) task_begin(k, sm, h, sm.queue[my], root);
            }
            /* queue empty: idle lanes take a pending subtree from busy lanes of this warp */
            const unsigned thieves = __ballot_sync(FULL, !k.busy);
            const unsigned donors = __ballot_sync(FULL, k.busy && k.sp > 0);
            if (thieves && donors) {
                const int pairs = min(__popc(thieves), __popc(donors));
                const int my_rank = __popc((k.busy ? donors : thieves) & lt_mask);
                const bool donate = k.busy && k.sp > 0 && my_rank < pairs;
                const bool steal = !k.busy && my_rank < pairs;
                int2 give = make_int2(0, 0);
                if (donate) give = stack[--k.sp];
                const int src = steal ? (int)__fns(donors, 0, my_rank + 1) : lane;
                const int g_slot = __shfl_sync(FULL, k.slot, src);
                const int gx = __shfl_sync(FULL, give.x, src);
                const int gy = __shfl_sync(FULL, give.y, src);
                if (steal) task_begin(k, sm, h, g_slot, make_int2(gx, gy));
            }
        }
        if (!__any_sync(FULL, k.busy)) break;

        /* ---- node phase: every lane whose current reference is an inner node steps it ----------------------- */
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
                if (k.mode == WV_Q_ANY && okL && okR && tnR < tnL) { /* nearest first: blockers are found sooner */
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
                    task_pop(k, stack, sm);
                }
            }
        }
        /* ---- triangle phase: one triangle per lane per step ------------------------------------------------- */
        while (__any_sync(FULL, k.busy && k.cur.y >= 0)) {
            if (k.busy && k.cur.y >= 0) {
                if (k.ti < k.cur.y) {
                    const int i = k.ti++;
                    if (COUNT) w.tris++;
                    float t;
                    if (tri_fast(tris + 3 * (size_t)i, k.ctx.O, k.ctx.u, k.t_best, t, w.tri_exact) && t > a.eps_tri) {
                        if (k.mode == WV_Q_ANY) {
                            if (blocks_light(k.ctx.O, k.ctx.u, t, k.d2)) {
                                sm.res[k.slot] = 1ull;
                                k.sp = 0;
                                k.ti = k.cur.y; /* leaves the leaf; task_pop ends the task */
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
                    task_pop(k, stack, sm);
                    if (k.busy && k.cur.y < 0) {
                        /* next reference is an inner node: wait for the node phase */
                    }
                }
            }
            /* leave the phase early when inner-node work is waiting and few lanes still have triangles */
            const unsigned tri_lanes = __ballot_sync(FULL, k.busy && k.cur.y >= 0);
            const unsigned node_lanes = __ballot_sync(FULL, k.busy && k.cur.y < 0);
            if (node_lanes && __popc(tri_lanes) < 12) break;
        }
    }
}

/* the owner thread's share of Scene::intersect_all for the current ray: the spheres (ascending id, strict <) */
__device__ __forceinline__ void closest_sphere(const SceneHeader& h, F3 O, F3 u, float& ts, int& sidx) {
    ts = RTK_INF;
    sidx = -1;
    for (int k = 0; k < h.n_spheres; k++) {
        float t;
        if (sphere_t(h.spheres[k], O, u, t) && t < ts) {
            ts = t;
            sidx = k;
        }
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(WV_THREADS) render_wave(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob, const RenderArgs a,
                                                         int* __restrict__ tile_counter) {
    __shared__ WaveSmem sm;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < 256; k += WV_THREADS) sm.gamma[k] = a.gamma_tab[a.gamma_mode * 256 + k];

    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const float4* nhat = reinterpret_cast<const float4*>(blob + h.off_nhat);
    const int tiles_x = (a.W + WV_TILE_W - 1) / WV_TILE_W;
    const int n_tiles = tiles_x * ((a.rows + WV_TILE_H - 1) / WV_TILE_H);
    const F3 Lp = f3(h.L[0], h.L[1], h.L[2]);
    const float eps = a.eps_surface;
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;

    for (;;) {
        __syncthreads(); /* everyone is done with the previous tile's shared state */
        if (tid == 0) sm.tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile = sm.tile;
        if (tile >= n_tiles) break;
        const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
        const int j = tile_x * WV_TILE_W + (lane & 7);
        const int kr = tile_y * WV_TILE_H + (lane >> 3);
        const bool live = warp == 0 && j < a.W && kr < a.rows; /* warp 0 owns the tile's 32 pixels */
        const size_t px = (size_t)kr * a.W + j;

        /* per-pixel path state (registers of the owner) */
        int state = live ? WV_TRACE : WV_DONE;
        int depth = 0;
        float n_ray = 1.f;
        F3 color = f3(0.f, 0.f, 0.f); /* colour the pixel gets if its pending shadow query finds no blocker */
        if (live) {
            const int i = a.row_begin + kr * a.row_step;
            const F3 uc = f3((float)j - (float)a.W / 2 + 0.5f, (float)a.H / 2 - (float)i - 0.5f, a.z); /* optimized.cu:751 */
            const F3 u0 = normalized(uc);
            sm.ox[tid] = a.camx;
            sm.oy[tid] = a.camy;
            sm.oz[tid] = a.camz;
            sm.ux[tid] = u0.x;
            sm.uy[tid] = u0.y;
            sm.uz[tid] = u0.z;
            if (a.hit_obj) a.hit_obj[px] = -1;
            if (a.hit_tri) a.hit_tri[px] = -1;
            if (a.hit_t) a.hit_t[px] = RTK_INF;
            if (a.shadow) a.shadow[px] = 2;
        }

        for (;;) { /* rounds: owner step, then the block drains the queue */
            if (tid == 0) {
                sm.q_count = 0;
                sm.q_head = 0;
            }
            __syncthreads();
            bool post = false;
            /* ---- owner step: advance this pixel's path until it needs a mesh query or ends ------------------ */
            while (state != WV_DONE) {
                const F3 O = f3(sm.ox[tid], sm.oy[tid], sm.oz[tid]);
                const F3 u = f3(sm.ux[tid], sm.uy[tid], sm.uz[tid]);
                if (state == WV_WAIT_SHADOW) {
                    if (sm.res[tid] != 0ull) color = f3(0.f, 0.f, 0.f);
                    if (a.shadow) a.shadow[px] = sm.res[tid] != 0ull ? 1 : 0;
                    state = WV_DONE;
                    break;
                }
                float t_hit;
                int sidx, tri = -1;
                if (state == WV_TRACE) {
                    if (depth >= a.segments) {
                        state = WV_DONE;
                        break;
                    }
                    w.rays++;
                    closest_sphere(h, O, u, t_hit, sidx);
                    if (h.has_mesh) {
                        RayCtx ctx = make_ray_ctx(O, u, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                        float tn;
                        if (slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tn, w.slab_fallbacks)) {
                            sm.ts[tid] = t_hit;
                            sm.sidx[tid] = sidx;
                            sm.qmode[tid] = WV_Q_CLOSEST;
                            sm.res[tid] = WV_NOHIT;
                            sm.d2[tid] = 0.f;
                            state = WV_WAIT_HIT;
                            post = true;
                            break;
                        }
                    }
                } else { /* WV_WAIT_HIT: merge the mesh result with the spheres (ascending id, strict <) */
                    t_hit = sm.ts[tid];
                    sidx = sm.sidx[tid];
                    const unsigned long long key = sm.res[tid];
                    if (key != WV_NOHIT) {
                        const float tm = __uint_as_float((unsigned)(key >> 32));
                        const int sid = sidx >= 0 ? h.spheres[sidx].id : -1;
                        if (tm < t_hit || (tm == t_hit && h.mesh_id < sid)) {
                            t_hit = tm;
                            sidx = -1;
                            tri = rank_to_tri((unsigned)key, h.n_tris, a.push_order, a.rank_off_bits);
                        }
                    }
                }
                /* ---- shade the hit (optimized.cu:571-650) ----------------------------------------------------- */
                const int obj = tri >= 0 ? h.mesh_id : (sidx >= 0 ? h.spheres[sidx].id : -1);
                if (depth == 0) {
                    if (a.hit_obj) a.hit_obj[px] = obj;
                    if (a.hit_tri) a.hit_tri[px] = tri;
                    if (a.hit_t) a.hit_t[px] = t_hit;
                }
                depth++;
                state = WV_TRACE;
                if (obj < 0) {
                    state = WV_DONE;
                    break;
                }
                const F3 P = O + t_hit * u; /* :555 */
                F3 N, albedo;
                int mirror;
                float n_in, n_out;
                if (tri < 0) {
                    const DevSphere& s = h.spheres[sidx];
                    N = normalized(P - f3(s.cx, s.cy, s.cz));
                    albedo = f3(s.ax, s.ay, s.az);
                    mirror = s.mirror;
                    n_in = s.n_in;
                    n_out = s.n_out;
                } else {
                    const float4 nh = __ldg(nhat + tri);
                    N = f3(nh.x, nh.y, nh.z);
                    albedo = f3(h.mesh_albedo[0], h.mesh_albedo[1], h.mesh_albedo[2]);
                    mirror = h.mesh_mirror;
                    n_in = h.mesh_n_in;
                    n_out = h.mesh_n_out;
                }
                F3 nO, nu;
                if (mirror) { /* :572-579 */
                    nO = P + eps * N;
                    nu = u - (2 * dot(u, N)) * N;
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
                        nO = P + eps * N;
                        nu = u - (2 * un) * N;
                    } else {
                        nO = P - eps * N;
                        const F3 Ncomp = (-sqrtf(1 - (ratio * ratio) * (1 - un * un))) * N;
                        const F3 Tcomp = ratio * (u - un * N);
                        nu = Ncomp + Tcomp;
                        n_ray = out2in ? n_in : n_out;
                    }
                } else { /* diffuse :610-650 */
                    const F3 Padj = P + eps * N;
                    const F3 toL = Lp - Padj;
                    const float D2 = norm2(toL);
                    const F3 su = toL / sqrtf(D2);
                    const F3 PL = Lp - P;
                    const F3 wl = normalized(PL);
                    const float ndl = dot(N, wl);
                    const float lambert = (ndl < 0.f) ? 0.f : ndl;
                    const float l = (float)((double)h.intensity / (12.566370614359172 * (double)norm2(PL)) * (double)lambert); /* :628 */
                    color = (l * albedo) / 3.14159274f;                                                                  /* :629 */
                    /* shadow ray (:618-620): spheres first, then the mesh through the queue */
                    w.rays++;
                    bool blocked = false;
                    for (int s = 0; s < h.n_spheres && !blocked; s++) {
                        float t;
                        if (sphere_t(h.spheres[s], Padj, su, t) && t < RTK_INF && blocks_light(Padj, su, t, D2)) blocked = true;
                    }
                    state = WV_DONE;
                    if (blocked) {
                        color = f3(0.f, 0.f, 0.f);
                        if (a.shadow) a.shadow[px] = 1;
                    } else {
                        bool enter = false;
                        if (h.has_mesh) {
                            RayCtx ctx = make_ray_ctx(Padj, su, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                            float tn;
                            enter = slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tn, w.slab_fallbacks);
                        }
                        if (enter) {
                            sm.ox[tid] = Padj.x;
                            sm.oy[tid] = Padj.y;
                            sm.oz[tid] = Padj.z;
                            sm.ux[tid] = su.x;
                            sm.uy[tid] = su.y;
                            sm.uz[tid] = su.z;
                            sm.d2[tid] = D2;
                            sm.qmode[tid] = WV_Q_ANY;
                            sm.res[tid] = 0ull;
                            state = WV_WAIT_SHADOW;
                            post = true;
                        } else if (a.shadow) {
                            a.shadow[px] = 0;
                        }
                    }
                    break;
                }
                sm.ox[tid] = nO.x;
                sm.oy[tid] = nO.y;
                sm.oz[tid] = nO.z;
                sm.ux[tid] = nu.x;
                sm.uy[tid] = nu.y;
                sm.uz[tid] = nu.z;
            }
            /* ---- post the query: warp-aggregated append to the tile's queue ------------------------------------ */
            {
                const unsigned m = __ballot_sync(0xffffffffu, post);
                int base = 0;
                if (lane == 0 && m) base = atomicAdd(&sm.q_count, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (post) sm.queue[base + __popc(m & ((1u << lane) - 1u))] = tid;
            }
            __syncthreads();
            if (sm.q_count == 0) break; /* no pixel of the tile is waiting for the mesh: every path has ended */
            drain_queue<COUNT>(h, nodes, tris, sm, a, w);
            __syncthreads();
        }

        /* ---- sample average, transfer function, store (optimized.cu:762-771) ---------------------------------- */
        if (live) {
            F3 total = f3(0.f, 0.f, 0.f);
            for (int s = 0; s < a.num_rays; s++) total = total + color;
            const F3 avg = a.num_rays == 1 ? total : total / (float)a.num_rays; /* x / 1.0f == x */
            if (a.rgb) {
                a.rgb[px * 3 + 0] = (uint8_t)quantise(avg.x, sm.gamma);
                a.rgb[px * 3 + 1] = (uint8_t)quantise(avg.y, sm.gamma);
                a.rgb[px * 3 + 2] = (uint8_t)quantise(avg.z, sm.gamma);
            }
        }
    }

    unsigned int rays = __reduce_add_sync(0xffffffffu, w.rays);
    if (lane == 0 && rays) atomicAdd(a.counters + 0, (unsigned long long)rays);
    if (COUNT) {
        unsigned int nn = __reduce_add_sync(0xffffffffu, w.nodes);
        unsigned int tt = __reduce_add_sync(0xffffffffu, w.tris);
        unsigned int ms = __reduce_max_sync(0xffffffffu, w.max_stack);
        unsigned int sf = __reduce_add_sync(0xffffffffu, w.slab_fallbacks);
        unsigned int te = __reduce_add_sync(0xffffffffu, w.tri_exact);
        if (lane == 0) {
            atomicAdd(a.counters + 1, (unsigned long long)nn);
            atomicAdd(a.counters + 2, (unsigned long long)tt);
            atomicMax(a.counters + 3, (unsigned long long)ms);
            atomicAdd(a.counters + 4, (unsigned long long)sf);
            atomicAdd(a.counters + 5, (unsigned long long)te);
        }
    }
}

} // namespace rtk
