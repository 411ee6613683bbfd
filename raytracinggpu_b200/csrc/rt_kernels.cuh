/*
 * rt_kernels.cuh — sm_100a render kernels.
 *
 * render_mega: one thread per pixel, every stage of the reference's per-pixel path fused in one launch:
 *   ray generation            optimized.cu:746-760  (cpu_launcher.cpp:695-709)
 *   Scene::intersect_all      optimized.cu:539-559
 *   Sphere::intersect         optimized.cu:123-135
 *   TriangleMesh::intersect   optimized.cu:220-285  (BVH traversal; cpu_launcher.cpp:277-311, array_bvh.cu:231-307)
 *   BoundingBox::intersect    optimized.cu:173-184
 *   moller_trumbore           optimized.cu:208-218
 *   getColorIterative         optimized.cu:561-661  (deterministic subset: mirror / refraction chain, direct + shadow)
 *   gamma + 8-bit store       optimized.cu:764-771
 * A warp covers an 8x4 pixel tile (coherent rays: node and triangle loads of a warp mostly hit the same
 * 64-B / 48-B record and are served as one L1 wavefront).
 */
#pragma once
#include "rt_layout.h"
#include "rt_math.cuh"

namespace rtk {

struct RenderArgs {
    int32_t W, H, rows, row_begin, row_step;
    int32_t row0, group_shift; /* compact row k of this launch is compact row row0 + k of the call; groups of 1 << group_shift image rows */
    int32_t segments, num_rays;
    float camx, camy, camz, z;
    float eps_surface, eps_tri;
    int32_t push_order, gamma_mode;
    uint8_t* rgb;
    int32_t* hit_obj;
    int32_t* hit_tri;
    float* hit_t;
    uint8_t* shadow;
    unsigned long long* counters; /* [0] rays, [1] inner-node visits, [2] triangle tests, [3] max stack, [4] slab exact fallbacks, [5] exact triangle evaluations */
    const float* gamma_tab;       /* 2 x 256 thresholds */
    int32_t rank_off_bits;        /* render_wave: bits of the in-leaf offset in the tie-break rank (push_order 0) */
    /* viewer-derived features (realtime_render.cu), wavefront pipeline only; all zero = the launchers' behaviour */
    int32_t camera_mode;          /* 1: u_center = C + bz z + bx (j - W/2 + 0.5) + by (H/2 - i - 0.5), realtime_render.cu:1113 */
    float bx[3], by[3], bz[3];
    const float4* tri_normals;    /* 3 x float4 per triangle: the vertex normals Na, Nb, Nc (w of the first: 1 = present); NULL = geometric normals */
    float4* linear;               /* progressive accumulation: the frame's linear colour per compact pixel instead of the 8-bit store */
    int32_t debug_cost;           /* investigation aid: with COUNT, hit_t <- thread clocks, hit_tri <- nodes+tris, hit_obj <- smid */
};

/* image row of the launch's compact row k (rt_params::row_begin / row_step / row_group) */
__device__ __forceinline__ int image_row(const RenderArgs& a, int k) {
    const int kk = a.row0 + k;
    return a.row_begin + (kk >> a.group_shift) * a.row_step + (kk & ((1 << a.group_shift) - 1));
}

struct Work {
    unsigned int rays, nodes, tris, max_stack;
    unsigned int slab_fallbacks, tri_exact; /* certified fast paths that had to evaluate the exact code */
};

/* One 8-bit channel: trunc(min(c^(1/2.2), 255)) (optimized.cu:765 / cpu_launcher.cpp:714) evaluated EXACTLY
 * without pow: T[k] is the smallest float whose reference transfer value is >= k (built on the host with the
 * reference's libm expression, rt_device.cu:build_gamma_table), so the result is the number of thresholds
 * <= c. A two-MUFU estimate finds the neighbourhood, the table fixes it. */
__device__ __forceinline__ int quantise(float c, const float* __restrict__ T) {
    if (c < T[1]) return 0; /* below the first threshold: most channels of the scene (wall albedos have zero channels, shadows are black) */
    const float g = __powf(c, 0.45454545f);
    int k = (int)fminf(fmaxf(g, 0.f), 255.f);
    while (k > 0 && c < T[k]) --k;
    while (k < 255 && c >= T[k + 1]) ++k;
    return k;
}

/* Sphere::intersect, t only (the normal is a function of t and is formed for the winner only). */
__device__ __forceinline__ bool sphere_t(const DevSphere& s, F3 O, F3 u, float& t) {
    const F3 OC = f3(O.x - s.cx, O.y - s.cy, O.z - s.cz);
    const float b = dot(u, OC);
    const float delta = b * b - (norm2(OC) - s.RR);
    if (delta < 0) return false;
    const float sq = sqrtf(delta);
    const float bc = -b; /* dot(u, C-O) == -dot(u, O-C) bit for bit (negation commutes with RN) */
    const float t1 = bc - sq;
    const float t2 = bc + sq;
    if (t2 < 0) return false;
    t = t1 < 0 ? t2 : t1;
    return true;
}

/* moller_trumbore on the packed record (A, e1, e2, N precomputed exactly). Returns true with t when the
 * reference's function returns 1. */
template <bool V8 = true>
__device__ __forceinline__ bool tri_exact(const float4* __restrict__ rec, F3 O, F3 u, float& t) {
    float4 q0, q1;
    ld32B<V8>(rec, q0, q1);
    const float4 q2 = __ldg(rec + 2);
    const F3 A = f3(q0.x, q0.y, q0.z), e1 = f3(q0.w, q1.x, q1.y), e2 = f3(q1.z, q1.w, q2.x), N = f3(q2.y, q2.z, q2.w);
    const float d = dot(u, N);
    if (d == 0) return false;
    const F3 AO = A - O;
    const F3 c = cross(AO, u);
    const float beta = dot(e2, c) / d;
    const float gamma = -dot(e1, c) / d;
    if (!(0 <= beta && beta <= 1) || !(0 <= gamma && gamma <= 1)) return false;
    t = dot(AO, N) / d;
    return beta + gamma <= 1 && t > 0;
}

/* Certified triangle test: the numerators and the denominator are the reference's own (unfused) values; only
 * the three divisions are first approximated by a multiplication with rcp.approx(d) (relative distance from
 * the reference quotient < 2^-22). A triangle is rejected on the approximation only when the reference's
 * comparison is certain to fail (margins 2^-20 .. 2^-18 around 0, 1 and t_limit); everything else — every
 * accepted hit in particular — goes through the exact divisions, so accepted t values are the reference's bits.
 * t_limit: hits with t certainly greater than t_limit are of no interest to the caller (current closest hit). */
/* the test proper, on a record already in registers (q0, q1 = first 32 B, q2 = next 16 B) */
__device__ __forceinline__ bool tri_fast_regs(const float4 q0, const float4 q1, const float4 q2, F3 O, F3 u, float t_limit, float& t, unsigned int& exact_evals) {
    const F3 A = f3(q0.x, q0.y, q0.z), e1 = f3(q0.w, q1.x, q1.y), e2 = f3(q1.z, q1.w, q2.x), N = f3(q2.y, q2.z, q2.w);
    const float d = dot(u, N);
    if (d == 0) return false;
    const F3 AO = A - O;
    const F3 c = cross(AO, u);
    const float nb = dot(e2, c);
    const float ng = -dot(e1, c);
    const float nt = dot(AO, N);
    if (fabsf(d) >= 1e-30f) {
        const float rd = rcp_approx(d);
        const float b = nb * rd, g = ng * rd;
        if (b < -1e-30f || b > 1.000001f || g < -1e-30f || g > 1.000001f) return false;
        if (b + g > 1.000004f) return false;
        const float ta = nt * rd;
        if (ta < -1e-30f || ta * 0.999999f > t_limit) return false;
    }
    exact_evals++;
    const float beta = nb / d;
    const float gamma = ng / d;
    if (!(0 <= beta && beta <= 1) || !(0 <= gamma && gamma <= 1)) return false;
    t = nt / d;
    return beta + gamma <= 1 && t > 0;
}

__device__ __forceinline__ bool tri_fast(const float4* __restrict__ rec, F3 O, F3 u, float t_limit, float& t, unsigned int& exact_evals) {
    float4 q0, q1;
    ldg256(rec, q0, q1);
    const float4 q2 = __ldg(rec + 2);
    return tri_fast_regs(q0, q1, q2, O, u, t_limit, t, exact_evals);
}

/* The certified test split in two for latency: tri_screen is straight-line code (no branch), so the screens of
 * two triangles interleave in one instruction stream; tri_finish evaluates the reference's three divisions for
 * the few tests the screen could not reject. tri_screen + tri_finish == tri_fast_regs, decision for decision
 * (every comparison with a NaN is false in both, so an undecidable screen falls through to the exact code). */
struct TriScreen {
    float d, nb, ng, nt;
    bool maybe;
};
__device__ __forceinline__ TriScreen tri_screen(const float4 q0, const float4 q1, const float4 q2, F3 O, F3 u, float t_limit) {
    const F3 A = f3(q0.x, q0.y, q0.z), e1 = f3(q0.w, q1.x, q1.y), e2 = f3(q1.z, q1.w, q2.x), N = f3(q2.y, q2.z, q2.w);
    TriScreen s;
    s.d = dot(u, N);
    const F3 AO = A - O;
    const F3 c = cross(AO, u);
    s.nb = dot(e2, c);
    s.ng = -dot(e1, c);
    s.nt = dot(AO, N);
    const float rd = rcp_approx(s.d);
    const float b = s.nb * rd, g = s.ng * rd, ta = s.nt * rd;
    const bool rej = (b < -1e-30f) | (b > 1.000001f) | (g < -1e-30f) | (g > 1.000001f) | (b + g > 1.000004f) | (ta < -1e-30f) | (ta * 0.999999f > t_limit);
    s.maybe = (s.d != 0) & !((fabsf(s.d) >= 1e-30f) & rej);
    return s;
}
__device__ __forceinline__ bool tri_finish(const TriScreen& s, float& t) {
    /* the reference's three divisions by d (moller_trumbore, optimized.cu:213-216), bit for bit, with the reciprocal refined once (div3) */
    const F3 q = div3(f3(s.nb, s.ng, s.nt), s.d);
    const float beta = q.x, gamma = q.y;
    t = q.z;
    return (0 <= beta) & (beta <= 1) & (0 <= gamma) & (gamma <= 1) & (beta + gamma <= 1) & (t > 0);
}

/* TriangleMesh::intersect. The reference visits nodes in a fixed LIFO order without pruning and accepts
 * strictly closer hits, so the visiting order only decides exact-t ties: with
 * push_order 1 (optimized.cu:265-266) leaves are seen in ascending triangle order -> the smallest leaf start
 * wins a tie; with push_order 0 (cpu_launcher.cpp:291-292) in descending leaf order -> the largest leaf start
 * wins; inside a leaf the first (smallest) index wins either way. Applying that rule makes the outcome
 * independent of the order used here (left child first). */
/* The reference's shadow predicate (optimized.cu:618-620): the light is blocked iff the point the shadow ray's
 * closest hit reconstructs, P' + t u, is not farther from P' than the light. */
__device__ __forceinline__ bool blocks_light(F3 Padj, F3 su, float t, float D2) {
    const F3 Ps = Padj + t * su;
    return norm2(Ps - Padj) <= D2;
}

/* ANY = false: closest hit (t_best, tri_best). ANY = true: shadow query — returns as soon as one accepted
 * triangle hit satisfies blocks_light; tri_best != -1 then means "blocked". This is the reference's outcome:
 * f(t) = |fl(fl(P' + fl(t u)) - P')|^2 is non-decreasing in t (every rounding is monotone), so if any accepted
 * hit has f(t) <= D2 the closest accepted hit, which the reference uses, has too; and if none has, neither has
 * the closest. Children are visited nearest-first in ANY mode so blockers are found early. */
template <bool COUNT, bool FAST, bool ANY, typename Header, bool V8 = true>
__device__ __forceinline__ void mesh_query(const Header& h, const float4* __restrict__ nodes,
                                           const float4* __restrict__ tris, F3 O, F3 u, float eps_tri, int push_order, float D2, float t_limit, float& t_best, int& tri_best, Work& w) {
    t_best = ANY ? t_limit : RTK_INF;
    tri_best = -1;
    int leaf_best = -1;
    RayCtx ctx;
    float tnL = 0.f, tnR = 0.f;
    if (FAST) {
        ctx = make_ray_ctx(O, u, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
        if (!slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tnL, w.slab_fallbacks)) return;
    } else {
        if (!slab_exact(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], O, u)) return;
    }
    int stack[RT_STACK_CAP];
    int sp = 0;
    int cur = h.root_ref;
    for (;;) {
        if (cur >= 0) {
            const float4* n = nodes + 4 * (size_t)cur;
            float4 q0, q1, q2, q3f;
            ld32B<V8>(n, q0, q1);
            ld32B<V8>(n + 2, q2, q3f);
            const int4 q3 = make_int4(__float_as_int(q3f.x), __float_as_int(q3f.y), __float_as_int(q3f.z), __float_as_int(q3f.w));
            if (COUNT && q3.z == 0) w.nodes++; /* virtual nodes are not nodes of the reference BVH */
            bool okL, okR;
            if (FAST) {
                okL = slab_fast(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ctx, tnL, w.slab_fallbacks);
                okR = slab_fast(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ctx, tnR, w.slab_fallbacks);
            } else {
                okL = slab_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, O, u);
                okR = slab_exact(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, O, u);
            }
            int cl = q3.x, cr = q3.y;
            if (ANY && FAST && okL && okR && tnR < tnL) { /* nearest first; the order never changes the outcome */
                const int s = cl;
                cl = cr;
                cr = s;
            }
            if (okL) {
                cur = cl;
                if (okR) {
                    stack[sp++] = cr;
                    if (COUNT) w.max_stack = max(w.max_stack, (unsigned)sp);
                }
                continue;
            }
            if (okR) {
                cur = cr;
                continue;
            }
        } else {
            const int code = -1 - cur;
            const int i_begin = code >> 2, i_end = min(i_begin + (code & 3) + 1, h.n_tris);
            for (int i = i_begin; i < i_end; i++) {
                if (COUNT) w.tris++;
                float t;
                if (FAST) {
                    if (!tri_fast(tris + 4 * (size_t)i, O, u, t_best, t, w.tri_exact)) continue;
                } else {
                    if (!tri_exact<V8>(tris + 4 * (size_t)i, O, u, t)) continue;
                }
                if (!(t > eps_tri)) continue; /* optimized.cu:275 / cpu_launcher.cpp:301 */
                if (ANY) {
                    if (blocks_light(O, u, t, D2)) {
                        tri_best = i;
                        return;
                    }
                    continue;
                }
                /* exact tie: the first triangle of the reference's leaf is the key; inside one leaf the smallest index wins */
                const int leaf_id = (t == t_best) ? __float_as_int(__ldg(tris + 4 * (size_t)i + 3).w) : -2;
                const bool tie = (tri_best >= 0) && (t == t_best) && (leaf_best != leaf_id) && (push_order == 1 ? (leaf_id < leaf_best) : (leaf_id > leaf_best));
                if (t < t_best || tie || (tri_best >= 0 && t == t_best && leaf_best == leaf_id && i < tri_best)) {
                    t_best = t;
                    tri_best = i;
                    leaf_best = __float_as_int(__ldg(tris + 4 * (size_t)i + 3).w);
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}

struct SurfaceHit {
    float t;
    int obj;  /* object id, -1 = miss */
    int tri;  /* triangle index or -1 */
    int sidx; /* index into h.spheres or -1 */
};

/* Scene::intersect_all: ascending object id, strict t < t_min (lowest id wins exact ties). */
template <bool COUNT, bool FAST>
__device__ __forceinline__ SurfaceHit intersect_all(const SceneHeader& h, const float4* __restrict__ nodes,
                                                    const float4* __restrict__ tris, F3 O, F3 u, float eps_tri, int push_order, Work& w) {
    w.rays++;
    SurfaceHit r;
    r.t = RTK_INF;
    r.obj = -1;
    r.tri = -1;
    r.sidx = -1;
    for (int k = 0; k < h.n_spheres; k++) {
        float t;
        if (sphere_t(h.spheres[k], O, u, t) && t < r.t) {
            r.t = t;
            r.obj = h.spheres[k].id;
            r.sidx = k;
        }
    }
    if (h.has_mesh) {
        float tm;
        int tri;
        mesh_query<COUNT, FAST, false>(h, nodes, tris, O, u, eps_tri, push_order, 0.f, 0.f, tm, tri, w);
        if (tri >= 0 && (tm < r.t || (tm == r.t && h.mesh_id < r.obj))) {
            r.t = tm;
            r.obj = h.mesh_id;
            r.tri = tri;
            r.sidx = -1;
        }
    }
    return r;
}

/* Is the light blocked for the shadow ray (Padj, su)? The reference runs a full closest-hit intersect_all and
 * applies blocks_light to its result (optimized.cu:618-620). By the monotonicity argument at mesh_query this
 * equals "some object's reported hit satisfies blocks_light", which lets the spheres be checked first and the
 * mesh be left as soon as one blocker is found. */
template <bool COUNT, bool FAST>
__device__ __forceinline__ bool light_blocked(const SceneHeader& h, const float4* __restrict__ nodes,
                                              const float4* __restrict__ tris, F3 Padj, F3 su, float D2, float eps_tri, int push_order, Work& w) {
    w.rays++;
    for (int k = 0; k < h.n_spheres; k++) {
        float t;
        if (sphere_t(h.spheres[k], Padj, su, t) && t < RTK_INF && blocks_light(Padj, su, t, D2)) return true;
    }
    if (!h.has_mesh) return false;
    /* |t su| ~ t: a hit with t > 1.001 sqrt(D2) cannot satisfy the predicate, the certified triangle test may drop it */
    const float t_limit = sqrtf(D2) * 1.001f + 1e-3f;
    float tm;
    int tri;
    mesh_query<COUNT, FAST, true>(h, nodes, tris, Padj, su, eps_tri, push_order, D2, t_limit, tm, tri, w);
    return tri >= 0;
}

template <bool COUNT, bool FAST>
__global__ void __launch_bounds__(128) render_mega(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob, const RenderArgs a) {
    __shared__ float s_gamma[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_gamma[k] = a.gamma_tab[a.gamma_mode * 256 + k];
    __syncthreads();

    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);

    /* 16x8 pixel tile per block, 8x4 per warp */
    const int tiles_x = (a.W + 15) >> 4;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
    const int k = tile_y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool live = j < a.W && k < a.rows;

    const long long clk0 = clock64();
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    if (live) {
        const int i = image_row(a, k);
        /* optimized.cu:751 — half-integers, exact in float */
        const F3 uc = f3((float)j - (float)a.W / 2 + 0.5f, (float)a.H / 2 - (float)i - 0.5f, a.z);
        const F3 cam = f3(a.camx, a.camy, a.camz);
        const F3 Lp = f3(h.L[0], h.L[1], h.L[2]);
        F3 O = cam;
        F3 u = normalized(uc); /* sigma == 0: the jitter terms of :758 are exactly 0 */
        float n_ray = 1.f;
        F3 color = f3(0.f, 0.f, 0.f);
        int first_obj = -1, first_tri = -1, shadow = 2;
        float first_t = RTK_INF;
        const float eps = a.eps_surface;

        for (int depth = 0; depth < a.segments; depth++) {
            const SurfaceHit hit = intersect_all<COUNT, FAST>(h, nodes, tris, O, u, a.eps_tri, a.push_order, w);
            if (depth == 0) {
                first_obj = hit.obj;
                first_tri = hit.tri;
                first_t = hit.t;
            }
            if (hit.obj < 0) break;
            const F3 P = O + hit.t * u; /* :555 */
            F3 N, albedo;
            int mirror;
            float n_in, n_out;
            if (hit.sidx >= 0) {
                const DevSphere& s = h.spheres[hit.sidx];
                N = normalized(P - f3(s.cx, s.cy, s.cz)); /* :132-133 */
                albedo = f3(s.ax, s.ay, s.az);
                mirror = s.mirror;
                n_in = s.n_in;
                n_out = s.n_out;
            } else {
                const float4 nh = __ldg(tris + 4 * (size_t)hit.tri + 3); /* N.normalize() :282, precomputed */
                N = f3(nh.x, nh.y, nh.z);
                albedo = f3(h.mesh_albedo[0], h.mesh_albedo[1], h.mesh_albedo[2]);
                mirror = h.mesh_mirror;
                n_in = h.mesh_n_in;
                n_out = h.mesh_n_out;
            }
            if (mirror) { /* :572-579 */
                const F3 Padj = P + eps * N;
                const F3 dir = u - (2 * dot(u, N)) * N;
                O = Padj;
                u = dir;
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
                    const F3 Padj = P + eps * N;
                    const F3 dir = u - (2 * un) * N;
                    O = Padj;
                    u = dir;
                    continue;
                }
                const F3 Padj = P - eps * N;
                const F3 Ncomp = (-sqrtf(1 - (ratio * ratio) * (1 - un * un))) * N;
                const F3 Tcomp = ratio * (u - un * N);
                O = Padj;
                u = Ncomp + Tcomp;
                n_ray = out2in ? n_in : n_out;
            } else { /* diffuse :610-650 */
                const F3 Padj = P + eps * N;
                const F3 toL = Lp - Padj;
                const F3 su = toL / sqrtf(norm2(toL)); /* NORMED_VEC :618 */
                bool blocked;
                if (FAST) {
                    blocked = light_blocked<COUNT, FAST>(h, nodes, tris, Padj, su, norm2(toL), a.eps_tri, a.push_order, w);
                } else { /* literal reference: closest hit, then the predicate */
                    const SurfaceHit sh = intersect_all<COUNT, FAST>(h, nodes, tris, Padj, su, a.eps_tri, a.push_order, w);
                    blocked = blocks_light(Padj, su, sh.t, norm2(toL)); /* on a miss t = 1e9f, as the reference leaves it */
                }
                if (blocked) { /* :620 */
                    shadow = 1;
                } else {
                    shadow = 0;
                    const F3 PL = Lp - P;
                    const F3 wl = normalized(PL);
                    const float ndl = dot(N, wl);
                    const float lambert = (ndl < 0.f) ? 0.f : ndl; /* std::max(dot, 0.f) */
                    /* :628 in double, as the source promotes it */
                    const float l = (float)((double)h.intensity / (12.566370614359172 * (double)norm2(PL)) * (double)lambert);
                    color = (l * albedo) / 3.14159274f; /* :629, Vector / float(PI) */
                }
                break; /* deterministic mode: the path ends at the first diffuse hit */
            }
        }

        /* sample average :762-764 — the samples are identical in deterministic mode; keep the float sums */
        F3 total = f3(0.f, 0.f, 0.f);
        for (int s = 0; s < a.num_rays; s++) total = total + color;
        const F3 avg = total / (float)a.num_rays;
        const size_t px = (size_t)k * a.W + j;
        if (a.rgb) {
            a.rgb[px * 3 + 0] = (uint8_t)quantise(avg.x, s_gamma);
            a.rgb[px * 3 + 1] = (uint8_t)quantise(avg.y, s_gamma);
            a.rgb[px * 3 + 2] = (uint8_t)quantise(avg.z, s_gamma);
        }
        if (a.hit_obj) a.hit_obj[px] = first_obj;
        if (a.hit_tri) a.hit_tri[px] = first_tri;
        if (a.hit_t) a.hit_t[px] = first_t;
        if (a.shadow) a.shadow[px] = (uint8_t)shadow;
        if (COUNT && a.debug_cost) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            if (a.hit_t) a.hit_t[px] = (float)(clock64() - clk0);
            if (a.hit_tri) a.hit_tri[px] = (int)(w.nodes + w.tris);
            if (a.hit_obj) a.hit_obj[px] = (int)smid;
        }
    }

    /* ray / work counters: one atomic per warp */
    unsigned int rays = __reduce_add_sync(0xffffffffu, w.rays);
    if (lane == 0 && rays) atomicAdd(a.counters + 0, (unsigned long long)rays);
    if (COUNT) {
        unsigned int nn = __reduce_add_sync(0xffffffffu, w.nodes);
        unsigned int tt = __reduce_add_sync(0xffffffffu, w.tris);
        unsigned int ms = __reduce_max_sync(0xffffffffu, w.max_stack);
        if (lane == 0) {
            atomicAdd(a.counters + 1, (unsigned long long)nn);
            atomicAdd(a.counters + 2, (unsigned long long)tt);
            atomicMax(a.counters + 3, (unsigned long long)ms);
        }
        unsigned int sf = __reduce_add_sync(0xffffffffu, w.slab_fallbacks);
        unsigned int te = __reduce_add_sync(0xffffffffu, w.tri_exact);
        if (lane == 0) {
            atomicAdd(a.counters + 4, (unsigned long long)sf);
            atomicAdd(a.counters + 5, (unsigned long long)te);
        }
    }
}

/* Mesh repack: reference interchange arrays -> packed triangle records + unit normals (rt_layout.h).
 * e1, e2, N as moller_trumbore forms them (optimized.cu:209-211), N/|N| as :282 does. */
__global__ void repack_triangles(const float* __restrict__ vertices, const int32_t* __restrict__ recs, int nt, const int32_t* __restrict__ leaf_start,
                                 float4* __restrict__ tris) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int32_t* r = recs + (size_t)i * 10;
    const int ia = r[0], ib = r[1], ic = r[2];
    const F3 A = f3(vertices[3 * (size_t)ia], vertices[3 * (size_t)ia + 1], vertices[3 * (size_t)ia + 2]);
    const F3 B = f3(vertices[3 * (size_t)ib], vertices[3 * (size_t)ib + 1], vertices[3 * (size_t)ib + 2]);
    const F3 C = f3(vertices[3 * (size_t)ic], vertices[3 * (size_t)ic + 1], vertices[3 * (size_t)ic + 2]);
    const F3 e1 = B - A, e2 = C - A;
    const F3 N = cross(e1, e2);
    const F3 nh = normalized(N);
    tris[4 * (size_t)i + 0] = make_float4(A.x, A.y, A.z, e1.x);
    tris[4 * (size_t)i + 1] = make_float4(e1.y, e1.z, e2.x, e2.y);
    tris[4 * (size_t)i + 2] = make_float4(e2.z, N.x, N.y, N.z);
    tris[4 * (size_t)i + 3] = make_float4(nh.x, nh.y, nh.z, __int_as_float(leaf_start[i]));
}

/* Per-triangle vertex normals for smooth shading (get_smooth_normal reads normals[tid.ni / nj / nk], realtime_render.cu:239-241):
 * gathered once into 3 x float4 per triangle, in the order of the uploaded triangle records. */
__global__ void repack_normals(const float* __restrict__ normals, int n_normals, const int32_t* __restrict__ recs, int nt, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int32_t* r = recs + (size_t)i * 10;
    const int na = r[6], nb = r[7], nc = r[8];
    const bool ok = na >= 0 && nb >= 0 && nc >= 0 && na < n_normals && nb < n_normals && nc < n_normals;
    for (int k = 0; k < 3; k++) {
        const int j = ok ? r[6 + k] : 0;
        out[3 * (size_t)i + k] = ok ? make_float4(normals[3 * (size_t)j], normals[3 * (size_t)j + 1], normals[3 * (size_t)j + 2], 1.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

/* get_smooth_normal, realtime_render.cu:221-245: beta and gamma with moller_trumbore's own expressions on the packed record
 * (A, e1, e2, N = e1 x e2 are the reference's values bit for bit), N = normalize(alpha Na + beta Nb + gamma Nc). */
__device__ __forceinline__ F3 smooth_normal(const float4* __restrict__ tris, const float4* __restrict__ tri_normals, int tri, F3 O, F3 u, F3 flat) {
    const float4 na = __ldg(tri_normals + 3 * (size_t)tri);
    if (na.w == 0.f) return flat; /* the record carries no normal indices */
    const float4 nb = __ldg(tri_normals + 3 * (size_t)tri + 1), nc = __ldg(tri_normals + 3 * (size_t)tri + 2);
    const float4 q0 = __ldg(tris + 4 * (size_t)tri), q1 = __ldg(tris + 4 * (size_t)tri + 1), q2 = __ldg(tris + 4 * (size_t)tri + 2);
    const F3 A = f3(q0.x, q0.y, q0.z), e1 = f3(q0.w, q1.x, q1.y), e2 = f3(q1.z, q1.w, q2.x), Ng = f3(q2.y, q2.z, q2.w);
    const float beta = dot(e2, cross(A - O, u)) / dot(u, Ng);
    const float gamma = -dot(e1, cross(A - O, u)) / dot(u, Ng);
    const float alpha = 1 - beta - gamma;
    return normalized((alpha * f3(na.x, na.y, na.z) + beta * f3(nb.x, nb.y, nb.z)) + gamma * f3(nc.x, nc.y, nc.z));
}

/* Progressive accumulation, realtime_render.cu:1136-1140: accumbuffer += the frame's colour, the displayed 8-bit frame is
 * accumbuffer / framenumber through the transfer function. k == 1 starts a new accumulation. */
__global__ void accumulate_frame(float4* __restrict__ acc, const float4* __restrict__ linear, int npx, int k, uint8_t* __restrict__ rgb,
                                 const float* __restrict__ gamma_tab, int gamma_mode) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= npx) return;
    const float4 c = linear[px];
    float4 a = k == 1 ? make_float4(0.f, 0.f, 0.f, 0.f) : acc[px];
    a = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, 0.f);
    acc[px] = a;
    if (rgb) {
        const float* T = gamma_tab + gamma_mode * 256;
        const float kf = (float)k;
        rgb[(size_t)px * 3 + 0] = (uint8_t)quantise(a.x / kf, T);
        rgb[(size_t)px * 3 + 1] = (uint8_t)quantise(a.y / kf, T);
        rgb[(size_t)px * 3 + 2] = (uint8_t)quantise(a.z / kf, T);
    }
}

/* Device self-test of div_by_rcp against div.rn.f32 on pseudo-random operands in the range RaySafe admits.
 * out[0] += mismatches with 1 correction step, out[1] += mismatches with 2 steps, out[2] += pairs tested. */
__global__ void selftest_division(unsigned long long seed, int per_thread, unsigned long long* out) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned int bad1 = 0, bad2 = 0, n = 0;
    for (int it = 0; it < per_thread; it++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const unsigned int ra = (unsigned int)s, rb = (unsigned int)(s >> 32);
        /* exponents: a in 2^[-60,41), b in 2^[-40,2) ; random sign and mantissa; every 8th b has a
         * near-all-ones / near-zero mantissa (the hard cases for reciprocal-based division) */
        const unsigned int ea = 127 - 60 + (ra >> 24) % 101, eb = 127 - 40 + (rb >> 24) % 42;
        unsigned int ma = ra & 0x7fffffu, mb = rb & 0x7fffffu;
        if ((it & 7) == 0) mb = (it & 8) ? (0x7fffffu - (mb & 0xff)) : (mb & 0xff);
        if ((it & 7) == 1) ma = (it & 8) ? (0x7fffffu - (ma & 0xff)) : (ma & 0xff);
        const float x = __uint_as_float(((ra >> 23) & 1u) << 31 | ea << 23 | ma);
        const float y = __uint_as_float(((rb >> 23) & 1u) << 31 | eb << 23 | mb);
        const float ref = x / y;
        const float r = __frcp_rn(y);
        const float q1 = div_by_rcp<1>(x, y, r), q2 = div_by_rcp<2>(x, y, r);
        bad1 += __float_as_uint(q1) != __float_as_uint(ref);
        bad2 += __float_as_uint(q2) != __float_as_uint(ref);
        n++;
    }
    atomicAdd(out + 0, (unsigned long long)bad1);
    atomicAdd(out + 1, (unsigned long long)bad2);
    atomicAdd(out + 2, (unsigned long long)n);
}

/* Device self-test of div3 / div3_or_zero (rt_math.cuh) against three div.rn.f32: pseudo-random triples and divisors across and beyond
 * the guarded range (exponents 2^[-70, 70], zeros, hard mantissas), normalisation-shaped inputs (n = sqrtf(norm2)) and the constant pi.
 * out[0] += components that differ in any bit, out[1] += components tested. */
__global__ void selftest_division3(unsigned long long seed, int per_thread, unsigned long long* out) {
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned int bad = 0, n = 0;
    for (int it = 0; it < per_thread; it++) {
        float v[4];
        for (int k = 0; k < 4; k++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            const unsigned int r = (unsigned int)(s >> 16);
            const unsigned int e = 127 - 70 + (r >> 24) % 141;
            unsigned int m = r & 0x7fffffu;
            if (((it + k) & 7) == 0) m = (it & 8) ? (0x7fffffu - (m & 0xff)) : (m & 0xff);
            v[k] = __uint_as_float(((r >> 23) & 1u) << 31 | e << 23 | m);
            if ((r & 0x3f000000u) == 0x15000000u) v[k] = (r & 1) ? 0.f : -0.f; /* exact zeros of either sign */
        }
        F3 a = f3(v[0], v[1], v[2]);
        float d = v[3];
        const int shape = it & 3;
        if (shape == 1) { /* a normalisation: moderate components, the divisor is their norm */
            a = f3(v[0] * 1e-3f, (float)((int)(s & 0xffff) - 32768) * 0.5f, v[2] > 0 ? 0.57735026f : 940.f);
            d = sqrtf(norm2(a));
        } else if (shape == 2) {
            d = 3.14159274f;
        }
        const F3 ref = f3(a.x / d, a.y / d, a.z / d);
        const F3 q = shape == 2 ? div3_or_zero(a, d) : div3(a, d);
        bad += (__float_as_uint(q.x) != __float_as_uint(ref.x)) + (__float_as_uint(q.y) != __float_as_uint(ref.y)) + (__float_as_uint(q.z) != __float_as_uint(ref.z));
        n += 3;
    }
    atomicAdd(out + 0, (unsigned long long)bad);
    atomicAdd(out + 1, (unsigned long long)n);
}

} // namespace rtk
