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
#include "rt_bins.cuh"

namespace rtk {

/* Programmatic dependent launch (rt_device.cu: LaunchChain): first statement of every kernel of the pipeline. Waits until the kernel
 * this one was launched behind has completed and its writes are visible (a no-op for a plain launch), then lets the NEXT kernel of the
 * stream be placed as soon as all blocks of this one have started. Nothing upstream kernels produced may be read before the wait. */
__device__ __forceinline__ void pdl_wait_then_release() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

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
    unsigned long long stats[8]; /* rays, node visits, triangle tests, max stack, slab fallbacks, exact triangle evals, task buffer overflow, pool overflow */
    unsigned long long dbg[8];   /* COUNT only: N steps, N tasks, T steps, T tasks, admissions, idle iterations */
    int nA[WF_MAX_ROUNDS + 2];   /* closest-hit queries posted for round r */
    int nS[WF_MAX_ROUNDS + 2];   /* shadow queries posted for round r */
    int head[WF_MAX_ROUNDS + 2]; /* traversal fetch cursor of round r */
    int nTask[WF_MAX_ROUNDS + 2]; /* (ray, leaf) tasks posted for round r (anchored rays) */
    int nB[WF_MAX_ROUNDS + 2];    /* anchored closest-hit queries of round r: they fill the closest-hit queue from its END downwards,
                                   * so that wf_traverse (entries 0 .. nA) never sees them; wf_shade handles both ranges */
};

struct WfArgs {
    RenderArgs a;
    QEntry* qA[2]; /* closest-hit queue of round r lives in qA[r & 1] */
    QEntry* qS;    /* shadow queue of the current round */
    WfCounters* c;
    unsigned long long* sticky; /* [0] task buffer overflow, [1] node pool overflow: set by the kernels, NEVER cleared by rt_render (a frame
                                 * enqueued with RT_RENDER_NO_SYNC must not lose the flag to the next frame's counter reset); rt_scene_sync
                                 * reads and clears them */
    int round;     /* round whose queues this launch consumes (traverse, shade) */
    /* stochastic mode (rt_stochastic.cuh explains the stream): one wavefront pass per sample */
    int stoch;             /* 0: deterministic mode */
    int sample;            /* sample index of this pass; first-hit outputs come from sample 0 */
    int last_sample;       /* wf_fold of this pass writes the 8-bit pixel */
    int indirect;
    int libm;              /* 1: logf / cosf / sinf of the CUDA library instead of the double-evaluated canon (RtOptions::transcendentals) */
    float aa_sigma;
    int npx;               /* compact pixels of this strip: stride of rec */
    const uint4* rng_table; /* start state of every pixel of the W x H frame, 2 x uint4 per pixel (d, v0..v4, -, -) */
    const float2* jitter_tab; /* JITTER instantiations: the Box-Muller jitter of every pixel's first sample (jitter_table) */
    uint4* rng;            /* running state per compact pixel of the strip, 2 x uint4 */
    float4* total;         /* per compact pixel: colour summed over the samples (xyz), bit mask of the diffuse segments of this sample (w) */
    float4* rec;           /* [segment][compact pixel][2]: direct term, albedo of the diffuse hit that ended that segment */
    int run_shift;  /* log2 of the admission run length, -1 = chosen per launch from the queue length */
    int gss_factor; /* the last gss_factor * n_warps runs are admitted one at a time */
    int fair_share; /* 1: an admission takes at most the warp's even share of the queue (short queues) */
    const float4* top; /* TOPS instantiations of wf_traverse: the top levels of the tree as a table of n_top records in breadth-first order, */
    int n_top;         /* child references into the table rewritten to n_inner + slot (build_top_table); staged in shared memory per block */
    int* spill;     /* node-pool overflow area: spill_cap ints per traversal warp (global memory) */
    int spill_cap;
    int* dbg_warps; /* investigation aid (RT_DEBUG_WARPS, COUNT kernels only): 16 ints per traversal warp and round */
    /* anchored rays (rt_bins.cuh): camera rays look their candidate leaves up in bins[0], shadow rays in bins[1]; the
     * boxes that pass the slab test become (ray, leaf) tasks for wf_leaves. anchored == 0: everything goes through
     * wf_traverse. */
    int anchored;
    BinsView bins[2];
    int2* tasks;  /* x: queue entry | 0x80000000 for a shadow query, y: leaf-table index of the candidate */
    int task_cap;
    int qcap;     /* entries of each queue of this strip */
    float4 cam_sph[RT_MAX_SPHERES]; /* per sphere: O - C (xyz) and |O - C|^2 - R^2 (w) for O = the camera (closest_sphere_cam) */
#ifdef RT_TIMELINE
    unsigned long long *tl_min, *tl_max; /* investigation build (make EXTRA=-DRT_TIMELINE): first block start / last warp end of every launch */
    int tl_slot;
#endif
};

#ifdef RT_TIMELINE
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TL_BEGIN(g) do { if (threadIdx.x == 0 && (g).tl_min) atomicMin((g).tl_min + (g).tl_slot, tl_now()); } while (0)
#define TL_END(g) do { if ((threadIdx.x & 31) == 0 && (g).tl_max) atomicMax((g).tl_max + (g).tl_slot, tl_now()); } while (0)
#else
#define TL_BEGIN(g)
#define TL_END(g)
#endif

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

/* what a pixel's path needs next from the mesh (at most one query per path_advance call) */
struct Post {
    int kind; /* 0 nothing, WF_MODE_CLOSEST, WF_MODE_ANY */
    F3 O, u;
    float aux, n_ray;
    int packed;
    /* stochastic mode: a diffuse hit posts its shadow query AND the path goes on, so one call may need a second slot */
    int cand_start, cand_count; /* anchored query: its candidate leaves are items[cand_start .. + cand_count) of the anchor's bins */
    int kind2; /* 0 or WF_MODE_CLOSEST: the next segment, root box NOT tested yet (pixel sign bit set in the entry) */
    F3 O2, u2;
    float aux2;
    int packed2;
};
#define WF_ROOT_UNTESTED 0x80000000u /* in QEntry::pixel */

/* running XORWOW state of a pixel in global memory (rt_stochastic.cuh: XorwowState), 2 x uint4 */
struct RngRef {
    uint4* p;
    bool loaded;
    unsigned d, v0, v1, v2, v3, v4;
};
__device__ __forceinline__ float rng_uniform(RngRef& r) {
    if (!r.loaded) {
        const uint4 a = r.p[0], b = r.p[1];
        r.d = a.x; r.v0 = a.y; r.v1 = a.z; r.v2 = a.w; r.v3 = b.x; r.v4 = b.y;
        r.loaded = true;
    }
    const unsigned t = r.v0 ^ (r.v0 >> 2); /* curand(): xorwow + Weyl, curand_kernel.h:863-874 */
    r.v0 = r.v1; r.v1 = r.v2; r.v2 = r.v3; r.v3 = r.v4;
    r.v4 = (r.v4 ^ (r.v4 << 4)) ^ (t ^ (t << 1));
    r.d += 362437u;
    return (float)(r.v4 + r.d) * 2.3283064e-10f + (2.3283064e-10f / 2.0f); /* curand_uniform */
}
__device__ __forceinline__ void rng_store(const RngRef& r) {
    if (r.loaded) {
        r.p[0] = make_uint4(r.d, r.v0, r.v1, r.v2);
        r.p[1] = make_uint4(r.v3, r.v4, 0u, 0u);
    }
}

__device__ __forceinline__ void store_entry(QEntry* q, int slot, F3 O, F3 u, float aux, int pixel, float n_ray, int packed, unsigned long long res) {
    float4* p = reinterpret_cast<float4*>(q + slot);
    p[0] = make_float4(O.x, O.y, O.z, aux);
    p[1] = make_float4(u.x, u.y, u.z, __int_as_float(pixel));
    p[2] = make_float4(n_ray, __int_as_float(packed), __uint_as_float((unsigned)(res & 0xffffffffull)), __uint_as_float((unsigned)(res >> 32)));
}

/* sample average + transfer function + store (optimized.cu:762-771) */
__device__ __forceinline__ void write_pixel(const RenderArgs& a, int px, F3 color) {
    if (!a.rgb && !a.linear) return;
    F3 total = f3(0.f, 0.f, 0.f) + color; /* the first of num_rays additions (0 + x, as the reference's accumulation starts) */
    for (int s = 1; s < a.num_rays; s++) total = total + color;
    const F3 avg = a.num_rays == 1 ? total : total / (float)a.num_rays; /* x / 1.0f == x */
    if (a.linear) { /* progressive accumulation: accumulate_frame quantises after the frame */
        a.linear[px] = make_float4(avg.x, avg.y, avg.z, 0.f);
        return;
    }
    const float* T = a.gamma_tab + a.gamma_mode * 256;
    a.rgb[(size_t)px * 3 + 0] = (uint8_t)quantise(avg.x, T);
    a.rgb[(size_t)px * 3 + 1] = (uint8_t)quantise(avg.y, T);
    a.rgb[(size_t)px * 3 + 2] = (uint8_t)quantise(avg.z, T);
}

/* Is the ray outside the contract of the bins / the wide index (rt_bins.cuh)? A slab distance can only be NaN (0/0) when a
 * direction component is zero; flagged here is the superset "zero, subnormal, tiny, infinite or NaN component" — such rays
 * are answered by the exact tree search, which is right for every ray. */
__device__ __forceinline__ bool outside_contract(F3 u) {
    const float ax = fabsf(u.x), ay = fabsf(u.y), az = fabsf(u.z);
    return !(ax >= 1e-37f && ay >= 1e-37f && az >= 1e-37f && ax <= 3e38f && ay <= 3e38f && az <= 3e38f);
}

/* closest_sphere for a ray that starts at the camera: O - C and |O - C|^2 - R^2 of Sphere::intersect (optimized.cu:124-126)
 * do not depend on the pixel; cam_sph holds them (same operations, same order, evaluated once on the host). */
/* NS > 0: the host knows the scene has exactly NS spheres (the reference's room: six walls); the loops over them are unrolled completely and
 * every sphere constant becomes an operand from the constant bank instead of an indexed load */
template <int NS = 0>
__device__ __forceinline__ void closest_sphere_cam(const SceneHeader& h, const float4* __restrict__ cam_sph, F3 u, float& ts, int& sidx) {
    ts = RTK_INF;
    sidx = -1;
#pragma unroll(NS > 0 ? NS : 2)
    for (int k = 0; k < (NS > 0 ? NS : h.n_spheres); k++) {
        const float4 c = cam_sph[k];
        const float b = dot(u, f3(c.x, c.y, c.z));
        const float delta = b * b - c.w;
        if (delta < 0) continue;
        const float sq = sqrtf(delta);
        const float bc = -b;
        const float t1 = bc - sq, t2 = bc + sq;
        if (t2 < 0) continue;
        const float t = t1 < 0 ? t2 : t1;
        if (t < ts) {
            ts = t;
            sidx = k;
        }
    }
}

__device__ __forceinline__ void closest_sphere(const SceneHeader& h, F3 O, F3 u, float& ts, int& sidx) {
    ts = RTK_INF;
    sidx = -1;
#pragma unroll 2
    for (int k = 0; k < h.n_spheres; k++) { /* ascending id, strict < (optimized.cu:543-554) */
        float t;
        if (sphere_t(h.spheres[k], O, u, t) && t < ts) {
            ts = t;
            sidx = k;
        }
    }
}

/* The exact tree search of one ray by one lane (mesh_query: the reference's own slab and triangle arithmetic over the
 * two-child records, in the reference's visiting order): the way out for the rare rays the faster structures do not
 * cover. Returns the winning triangle (-1: none) and its t; ANY: a triangle that blocks the light. */
struct MeshRoot { /* the header fields mesh_query reads, by value: a kernel parameter cannot be handed to a real call by reference */
    float root_mn[3], root_mx[3], box_abs[3];
    int root_ref, n_tris;
};
__device__ __forceinline__ MeshRoot mesh_root(const SceneHeader& h) {
    MeshRoot r;
    for (int k = 0; k < 3; k++) {
        r.root_mn[k] = h.root_mn[k];
        r.root_mx[k] = h.root_mx[k];
        r.box_abs[k] = h.box_abs[k];
    }
    r.root_ref = h.root_ref;
    r.n_tris = h.n_tris;
    return r;
}
__device__ __noinline__ int2 exact_mesh_query_(const MeshRoot h, const float4* __restrict__ nodes, const float4* __restrict__ tris, float ox, float oy, float oz,
                                               float ux, float uy, float uz, float eps_tri, int push_order, float D2, int any) {
    const F3 O = f3(ox, oy, oz), u = f3(ux, uy, uz);
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    int tri;
    float tm;
    if (any) mesh_query<false, false, true, MeshRoot, false>(h, nodes, tris, O, u, eps_tri, push_order, D2, sqrtf(D2) * 1.001f + 1e-3f, tm, tri, w);
    else mesh_query<false, false, false, MeshRoot, false>(h, nodes, tris, O, u, eps_tri, push_order, 0.f, 0.f, tm, tri, w);
    return make_int2(tri, __float_as_int(tm));
}
__device__ __forceinline__ int exact_mesh_query(const MeshRoot h, const float4* __restrict__ nodes, const float4* __restrict__ tris, F3 O, F3 u, float eps_tri,
                                                int push_order, float D2, bool any, float& tm) {
    const int2 r = exact_mesh_query_(h, nodes, tris, O.x, O.y, O.z, u.x, u.y, u.z, eps_tri, push_order, D2, any ? 1 : 0);
    tm = __int_as_float(r.y);
    return r.x;
}

/* a shadow query found a blocker: the pixel is black (optimized.cu:620-622) */
template <bool STOCH>
__device__ __forceinline__ void light_is_blocked(const RenderArgs& a, const WfArgs& g, int px, int packed) {
    if (!STOCH) {
        if (a.linear) {
            a.linear[px] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (a.rgb) {
            a.rgb[(size_t)px * 3 + 0] = 0;
            a.rgb[(size_t)px * 3 + 1] = 0;
            a.rgb[(size_t)px * 3 + 2] = 0;
        }
        if (a.shadow) a.shadow[px] = 1;
    } else { /* stochastic mode: the direct term of that segment's record becomes 0 (direct_colors[ray_depth], :622) */
        const int seg = ((packed >> 8) & 0xffff) - 1;
        g.rec[((size_t)seg * g.npx + px) * 2] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (seg == 0 && g.sample == 0 && a.shadow) a.shadow[px] = 1;
    }
}

/* Advance one pixel's path until it ends or needs the mesh. have_hit: (t_hit, sidx, tri) already hold the
 * answer of intersect_all for the current ray (wf_shade); otherwise the segment starts here.
 * STOCH (stochastic mode, one pass per sample): nothing is written to the framebuffer here; every diffuse hit leaves a
 * (direct, albedo) record for wf_fold, draws the two uniforms of optimized.cu:633-634 from the pixel's stream and, with
 * the indirect bounce enabled, goes on along the cosine-weighted direction. */
/* DIFFUSE: the host knows that no object of the scene is a mirror or refractive (every path ends at its first hit): the
 * reflection / refraction code is left out of the kernel (67 KB of SASS otherwise; instruction fetch is wf_generate's top stall) */
/* LEAN (with DIFFUSE, deterministic mode, anchored-ray bins in use — the primary + shadow pipeline): every query is an
 * anchored one and every path ends at its first hit, so the tree-search branches (root-box tests with their exact
 * fallbacks, the general sphere search, the posting of tree-searched queries) are left out as well. */
template <bool COUNT, bool STOCH, bool DIFFUSE = false, bool LEAN = false, int NS = 0>
__device__ __forceinline__ void path_advance(const SceneHeader& h, const WfArgs& g, const float4* __restrict__ nodes, const float4* __restrict__ tris, int px, F3 O, F3 u, float n_ray,
                                             int depth, bool have_hit, float t_hit, int sidx, int tri, Work& w, Post& post) {
    const RenderArgs& a = g.a;
    const F3 Lp = f3(h.L[0], h.L[1], h.L[2]);
    const float eps = a.eps_surface;
    const bool first_sample = !STOCH || g.sample == 0;
    for (;;) {
        if (!have_hit) {
            if (depth >= a.segments) { /* path budget used up without a diffuse hit: colour 0 (fold of optimized.cu:653-660) */
                if (!STOCH) write_pixel(a, px, f3(0.f, 0.f, 0.f));
                return;
            }
            w.rays++;
            if (LEAN || depth == 0) closest_sphere_cam<NS>(h, g.cam_sph, u, t_hit, sidx); /* depth 0 without a hit: the camera ray (wf_generate) */
            else closest_sphere(h, O, u, t_hit, sidx);
            tri = -1;
            if (h.has_mesh && (LEAN || (g.anchored && depth == 0))) {
                /* a camera ray: its candidate leaves come from the camera's bins (rt_bins.cuh) */
                /* a zero / subnormal / non-finite component is outside the bins' contract: the query is posted without
                 * candidates (cand_count -1) and answered by the exact tree search at the end of the kernel (answer_exact) */
                const bool exact = outside_contract(u) || __ldg(g.bins[0].status) != 0;
                int c0 = 0, c1 = -1;
                if (!exact) {
                    const int cell = bins_cell(g.bins[0], u);
                    c1 = 0; /* outside the windows: no candidate */
                    if (cell >= 0) {
                        c0 = __ldg(g.bins[0].cell_start + cell);
                        c1 = __ldg(g.bins[0].cell_start + cell + 1);
                        if (c1 > g.bins[0].items_cap || c0 < 0 || c1 < c0) { /* a list the build had to cut short (or a scan that overflowed) */
                            c0 = 0;
                            c1 = -1;
                        }
                    }
                }
                {
                    if (c1 != c0) {
                        post.kind = WF_MODE_CLOSEST;
                        post.O = O;
                        post.u = u;
                        post.aux = t_hit;
                        post.n_ray = n_ray;
                        post.packed = (sidx & 0xff) | (depth << 8) | (WF_MODE_CLOSEST << 24);
                        post.cand_start = c0;
                        post.cand_count = c1 - c0;
                        return;
                    }
                }
            } else if (!LEAN && h.has_mesh) {
                const RayCtx ctx = make_ray_ctx(O, u, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                float tn;
                if (slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tn, w.slab_fallbacks)) {
                    post.kind = WF_MODE_CLOSEST;
                    post.O = O;
                    post.u = u;
                    post.aux = t_hit;
                    post.n_ray = n_ray;
                    post.packed = (sidx & 0xff) | (depth << 8) | (WF_MODE_CLOSEST << 24);
                    return;
                }
            }
        }
        have_hit = false;
        /* ---- the hit of this segment is known: shade it (optimized.cu:571-650) ---------------------------------- */
        const int obj = tri >= 0 ? h.mesh_id : (sidx >= 0 ? h.spheres[sidx].id : -1);
        if (depth == 0 && first_sample) {
            if (a.hit_obj) a.hit_obj[px] = obj;
            if (a.hit_tri) a.hit_tri[px] = tri;
            if (a.hit_t) a.hit_t[px] = t_hit;
        }
        depth++;
        if (obj < 0) { /* the ray left the scene */
            if (!STOCH) write_pixel(a, px, f3(0.f, 0.f, 0.f));
            return;
        }
        const F3 P = O + t_hit * u; /* :555 */
        F3 N, albedo;
        int mirror;
        float n_in, n_out;
        if (tri < 0) {
            const DevSphere& s = h.spheres[sidx];
            N = normalized3(P - f3(s.cx, s.cy, s.cz)); /* :132-133 */
            albedo = f3(s.ax, s.ay, s.az);
            mirror = s.mirror;
            n_in = s.n_in;
            n_out = s.n_out;
        } else {
            const float4 nh = __ldg(tris + 4 * (size_t)tri + 3); /* N.normalize() :282, precomputed */
            N = f3(nh.x, nh.y, nh.z);
            if (a.tri_normals) N = smooth_normal(tris, a.tri_normals, tri, O, u, N); /* realtime_render.cu:311 */
            albedo = f3(h.mesh_albedo[0], h.mesh_albedo[1], h.mesh_albedo[2]);
            mirror = h.mesh_mirror;
            n_in = h.mesh_n_in;
            n_out = h.mesh_n_out;
        }
        if (!DIFFUSE && mirror) { /* :572-579 */
            const F3 nu = u - (2 * dot(u, N)) * N;
            O = P + eps * N;
            u = nu;
        } else if (!DIFFUSE && n_in != n_out) { /* :580-609 */
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
        } else { /* diffuse :610-650 */
            const F3 Padj = P + eps * N;
            const F3 toL = Lp - Padj;
            const float D2 = norm2(toL);
            const float rootD = sqrtf(D2);
            const F3 su = div3(toL, rootD); /* NORMED_VEC :618 */
            /* shadow ray: the spheres here, the mesh through the queue (see mesh_query for the equivalence) */
            w.rays++;
            bool blocked = false;
            /* Certified screen of the sphere tests. A sphere blocks the light iff the reference's t (Sphere::intersect) satisfies
             * |fl(fl(P' + fl(t su)) - P')|^2 <= D2 (blocks_light). Every component of that reconstructed vector is within 2^-22 (A + t) of
             * t su_c (A = largest |coordinate| of P'), |su| is within 2^-22 of 1, the squared norm within 2^-22 relative: a t above
             * T_hi = (sqrt(D2) + 2^-18 A)(1 + 2^-17) cannot satisfy it. The discriminant is the reference's own; its square root is first taken
             * with sqrt.approx (relative error 2^-23), which puts the reference's t1 within E = 2^-20 (|b| + sqrt(delta)) of t1a: when
             * t1a - E > T_hi the near root is positive and beyond the light (not a blocker), when t2a + E < 0 the sphere is behind the ray
             * (no hit); only the remaining spheres — in the closed room: the one the point lies on — evaluate the IEEE square root and the
             * reference's predicate. Same decisions as the unscreened loop for every input (a NaN fails both screens). */
            const float A = fmaxf(fmaxf(fabsf(Padj.x), fabsf(Padj.y)), fabsf(Padj.z));
            const float T_hi = (rootD + A * 3.814697265625e-06f) * 1.00000762939453125f;
            /* no early exit: a sphere that blocks is rare, and without the exit the tests are independent of one another */
#pragma unroll(NS > 0 ? NS : 2)
            for (int s = 0; s < (NS > 0 ? NS : h.n_spheres); s++) {
                const DevSphere& sp = h.spheres[s];
                const F3 OC = f3(Padj.x - sp.cx, Padj.y - sp.cy, Padj.z - sp.cz);
                const float b = dot(su, OC);
                const float delta = b * b - (norm2(OC) - sp.RR); /* Sphere::intersect :124-126 */
                if (delta < 0) continue;
                const float bc = -b;
                const float sqa = sqrt_approx(delta);
                const float E = (fabsf(bc) + sqa) * 9.5367431640625e-07f;
                if ((bc - sqa) - E > T_hi || (bc + sqa) + E < 0) continue; /* certainly not a blocker */
                const float sq = sqrtf(delta);
                const float t1 = bc - sq, t2 = bc + sq;
                if (t2 < 0) continue;
                const float t = t1 < 0 ? t2 : t1;
                blocked = blocked || (t < RTK_INF && blocks_light(Padj, su, t, D2));
            }
            F3 dcol = f3(0.f, 0.f, 0.f);
            if (!blocked) {
                const F3 PL = Lp - P;
                const F3 wl = normalized3(PL);
                const float ndl = dot(N, wl);
                const float lambert = (ndl < 0.f) ? 0.f : ndl; /* std::max(dot, 0.f) */
                const float l = (float)((double)h.intensity / (12.566370614359172 * (double)norm2(PL)) * (double)lambert); /* :628, in double */
                dcol = div3_or_zero(l * albedo, 3.14159274f);                                                                       /* :629 */
            }
            const int seg = depth - 1;
            if (!STOCH) {
                write_pixel(a, px, dcol);
                if (a.shadow) a.shadow[px] = blocked ? 1 : 0;
            } else {
                float4* r = g.rec + ((size_t)seg * g.npx + px) * 2;
                r[0] = make_float4(dcol.x, dcol.y, dcol.z, 0.f);
                r[1] = (g.indirect ? make_float4(albedo.x, albedo.y, albedo.z, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f));
                int* types = reinterpret_cast<int*>(&g.total[px].w);
                *types |= 1 << seg;
                if (seg == 0 && first_sample && a.shadow) a.shadow[px] = blocked ? 1 : 0;
            }
            if (!blocked && h.has_mesh && (LEAN || g.anchored)) {
                /* a shadow ray: its candidate leaves come from the light's bins */
                const bool exact = outside_contract(su) || !(D2 <= g.bins[1].max_D2) || __ldg(g.bins[1].status) != 0; /* outside the bins' contract: see the camera rays above */
                int c0 = 0, c1 = -1;
                if (!exact) {
                    const int cell = bins_cell(g.bins[1], toL);
                    c1 = 0;
                    if (cell >= 0) {
                        c0 = __ldg(g.bins[1].cell_start + cell);
                        c1 = __ldg(g.bins[1].cell_start + cell + 1);
                        if (c1 > g.bins[1].items_cap || c0 < 0 || c1 < c0) {
                            c0 = 0;
                            c1 = -1;
                        }
                    }
                }
                {
                    if (c1 != c0) {
                        post.kind = WF_MODE_ANY;
                        post.O = Padj;
                        post.u = su;
                        post.aux = D2;
                        post.n_ray = 1.f;
                        post.packed = 0xff | (depth << 8) | (WF_MODE_ANY << 24);
                        post.cand_start = c0;
                        post.cand_count = c1 - c0;
                    }
                }
            } else if (!LEAN && !blocked && h.has_mesh) {
                const RayCtx ctx = make_ray_ctx(Padj, su, h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                float tn;
                if (slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], ctx, tn, w.slab_fallbacks)) {
                    /* the pixel (STOCH: the record) keeps its lit colour unless the traversal finds a blocker, which zeroes it */
                    post.kind = WF_MODE_ANY;
                    post.O = Padj;
                    post.u = su;
                    post.aux = D2;
                    post.n_ray = 1.f;
                    post.packed = 0xff | (depth << 8) | (WF_MODE_ANY << 24);
                }
            }
            if (!STOCH || !g.indirect) return; /* the deterministic path ends at the first diffuse hit */
            /* ---- :631-649 — two uniforms at every diffuse hit (the last segment included), cosine-weighted direction */
            RngRef rng;
            rng.p = g.rng + (size_t)px * 2;
            rng.loaded = false;
            const float q1 = rng_uniform(rng), q2 = rng_uniform(rng);
            rng_store(rng);
            if (depth >= a.segments) return;
            const float an = (float)(2 * 3.14159265358979323846 * (double)q1);
            const float sq = sqrtf(1 - q2);
            const float x = (g.libm ? cosf(an) : canon_cos(an)) * sq, y = (g.libm ? sinf(an) : canon_sin(an)) * sq, z = sqrtf(q2);
            /* one component of T1 is an exact zero: div.rn.f32 would take its slow path for it at every diffuse hit (div3_or_zero) */
            const F3 T1raw = (fabsf(N.y) != 0 && fabsf(N.x) != 0) ? f3(-N.y, N.x, 0.f) : f3(-N.z, 0.f, N.x);
            const F3 T1 = div3_or_zero(T1raw, sqrtf(norm2(T1raw)));
            const F3 T2 = cross(N, T1);
            u = (x * T1 + y * T2) + z * N;
            O = Padj;
            n_ray = 1.f;
            if (post.kind == WF_MODE_ANY) {
                /* this call already owes a shadow query: the next segment goes through the queue with its root-box test
                 * left to wf_traverse, so that no call ever posts more than one query of each kind */
                w.rays++;
                closest_sphere(h, O, u, t_hit, sidx);
                post.kind2 = WF_MODE_CLOSEST;
                post.O2 = O;
                post.u2 = u;
                post.aux2 = t_hit;
                post.packed2 = (sidx & 0xff) | (depth << 8) | (WF_MODE_CLOSEST << 24);
                return;
            }
        }
    }
}

/* Append the warp's queries to the global queues of round post_round: one atomic per warp and queue. Must be
 * called by all 32 lanes together (the offsets come from full-mask ballots). */
template <bool LEAN = false>
__device__ __forceinline__ int post_queries(const WfArgs& g, int post_round, const Post& post, int px) {
    int my_slot = -1; /* queue entry of this lane's first query (closest or shadow), for the anchored tasks */
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const bool wantB = post.kind == WF_MODE_CLOSEST && post.cand_count != 0; /* anchored closest-hit query (camera ray); -1: to be answered exactly */
    /* LEAN: no tree-searched query can be posted (every closest-hit query is a camera ray with candidates, no second query) */
    const bool wantA = !LEAN && post.kind == WF_MODE_CLOSEST && !wantB, wantS = post.kind == WF_MODE_ANY, want2 = !LEAN && post.kind2 == WF_MODE_CLOSEST;
    const unsigned mB = __ballot_sync(FULL, wantB);
    if (mB) {
        int baseB = 0;
        if (lane == 0) baseB = atomicAdd(&g.c->nB[post_round], __popc(mB));
        baseB = __shfl_sync(FULL, baseB, 0);
        if (wantB) {
            my_slot = g.qcap - 1 - (baseB + __popc(mB & lt));
            store_entry(g.qA[post_round & 1], my_slot, post.O, post.u, post.aux, px, post.n_ray, post.packed, WF_NOHIT);
        }
    }
    const unsigned mA = LEAN ? 0u : __ballot_sync(FULL, wantA);
    const unsigned mS = __ballot_sync(FULL, wantS);
    const unsigned m2 = LEAN ? 0u : __ballot_sync(FULL, want2);
    int baseA = 0, baseS = 0;
    if (lane == 0) {
        if (mA | m2) baseA = atomicAdd(&g.c->nA[post_round], __popc(mA) + __popc(m2));
        if (mS) baseS = atomicAdd(&g.c->nS[post_round], __popc(mS));
    }
    baseA = __shfl_sync(FULL, baseA, 0);
    baseS = __shfl_sync(FULL, baseS, 0);
    if (wantA) {
        my_slot = baseA + __popc(mA & lt);
        store_entry(g.qA[post_round & 1], my_slot, post.O, post.u, post.aux, px, post.n_ray, post.packed, WF_NOHIT);
    } else if (wantS) {
        my_slot = baseS + __popc(mS & lt);
        store_entry(g.qS, my_slot, post.O, post.u, post.aux, px, post.n_ray, post.packed, 0ull);
    }
    if (want2)
        store_entry(g.qA[post_round & 1], baseA + __popc(mA) + __popc(m2 & lt), post.O2, post.u2, post.aux2, (int)((unsigned)px | WF_ROOT_UNTESTED), 1.f, post.packed2, WF_NOHIT);
    return my_slot;
}

/* ---- anchored queries: one (ray, candidate leaf) task per entry of the ray's cell list ------------------------------------
 * Called by all 32 lanes after post_queries. Nothing is tested here: the lanes copy their candidate lists into the task
 * array (one warp scan, one atomic per call; the loads of the copy loop are independent, so a long list costs bandwidth,
 * not a chain of round trips). wf_leaves applies the slab test and, where it passes, the triangle tests. */
__device__ __forceinline__ void emit_tasks(const WfArgs& g, int post_round, const Post& post, int slot) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool mine = slot >= 0 && (post.kind == WF_MODE_CLOSEST || post.kind == WF_MODE_ANY) && post.cand_count > 0;
    if (!__any_sync(FULL, mine)) return; /* most warps of a frame never meet the mesh */
    const int n = mine ? post.cand_count : 0;
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(&g.c->nTask[post_round], total);
    base = __shfl_sync(FULL, base, 0);
    if (base < 0 || base > g.task_cap - total) { /* reported by rt_scene_sync: the frame is rendered again with a larger buffer */
        if (lane == 0) {
            atomicExch(&g.sticky[0], 1ull);
            /* keep the cursor just past the buffer: it must never climb towards 2^31 and wrap into valid indices */
            atomicMin(reinterpret_cast<unsigned*>(&g.c->nTask[post_round]), (unsigned)g.task_cap + 1u);
        }
        return;
    }
    if (n > 0) {
        const int* items = g.bins[post.kind == WF_MODE_ANY ? 1 : 0].items + post.cand_start;
        int2* out = g.tasks + base + (incl - n);
        const int etag = slot | (post.kind == WF_MODE_ANY ? (int)0x80000000u : 0);
#pragma unroll 4
        for (int i = 0; i < n; i++) out[i] = make_int2(etag, __ldg(items + i));
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

/* The wide index and the bins rest on "a child box that passes the slab test implies its parent box passes" (rt_layout.h),
 * which holds as long as no slab distance is NaN, i.e. as long as no direction component is zero (0/0). A ray with a
 * zero, subnormal or non-finite component (RayCtx::M is NaN), or a shadow ray from beyond the bins' distance guard, is
 * answered here, by one lane, with the reference's own arithmetic over the two-child records (exact_mesh_query). */
template <bool STOCH>
__device__ __forceinline__ void answer_exact(const SceneHeader& h, const unsigned char* __restrict__ blob, const WfArgs& g, QEntry* entry, bool any) {
    const RenderArgs& a = g.a;
    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const float4* p = reinterpret_cast<const float4*>(entry);
    const float4 p0 = p[0], p1 = p[1];
    const int px = (int)(__float_as_uint(p1.w) & ~WF_ROOT_UNTESTED);
    float tm;
    const int tri = exact_mesh_query(mesh_root(h), nodes, tris, f3(p0.x, p0.y, p0.z), f3(p1.x, p1.y, p1.z), a.eps_tri, a.push_order, p0.w, any, tm);
    if (any) {
        if (tri >= 0) light_is_blocked<STOCH>(a, g, px, entry->packed);
    } else if (tri >= 0) {
        const unsigned rank = tie_rank(tri, __float_as_int(__ldg(tris + 4 * (size_t)tri + 3).w), h.n_tris, a.push_order, a.rank_off_bits);
        entry->res = ((unsigned long long)__float_as_uint(tm) << 32) | rank;
    }
}
/* end of wf_generate / wf_shade: the lanes whose anchored query was posted without candidates answer it now */
template <bool STOCH>
__device__ __forceinline__ void answer_deferred(const SceneHeader& h, const unsigned char* __restrict__ blob, const WfArgs& g, int post_round, const Post& post, int slot) {
    const bool need = slot >= 0 && post.cand_count < 0 && (post.kind == WF_MODE_CLOSEST || post.kind == WF_MODE_ANY);
    if (!__any_sync(0xffffffffu, need)) return;
    if (need) answer_exact<STOCH>(h, blob, g, (post.kind == WF_MODE_ANY ? g.qS : g.qA[post_round & 1]) + slot, post.kind == WF_MODE_ANY);
}

/* ---- wf_generate: one thread per pixel (a warp covers an 8x4 tile) ------------------------------------------------ */
/* JITTER (deterministic instantiations only): the camera ray takes the Box-Muller jitter of optimized.cu:753-759 from the first two
 * uniforms of the pixel's stream; nothing else of the stochastic mode is needed when a frame is one sample of one segment. */
template <bool COUNT, bool STOCH, bool DIFFUSE = false, bool LEAN = false, bool JITTER = false, int NS = 0>
#ifndef RT_GEN_BLOCKS
#define RT_GEN_BLOCKS 12 /* resident blocks per SM the deterministic instantiations are compiled for: 40 registers, the same 44 B of spills as at 10 blocks /
                          * 48 registers, and 1-1.7 % faster frames (same-box A/B; 8 / 9 blocks at 56 registers without spills: slower, 14 / 16 at 32
                          * registers: no gain). The stochastic instantiations stay at 10 (at 12 they spill 250-280 B). */
#endif
__global__ void __launch_bounds__(WF_THREADS, STOCH ? 10 : RT_GEN_BLOCKS) wf_generate(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                         const __grid_constant__ WfArgs g) {
    pdl_wait_then_release();
    TL_BEGIN(g);
    const RenderArgs& a = g.a;
    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const int lane = threadIdx.x & 31;
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    /* one 8x4 tile per warp (a grid-stride loop over tiles, tried to make the kernel share the SMs with the LSU-bound
     * wf_leaves of the other row band, cost 16 registers and was slower in every launch shape: profiles/r01_notes.md) */
    /* grid: x = groups of four tiles along a row of tiles, y = the row of tiles (no division in the kernel) */
    const int tx = blockIdx.x * (WF_THREADS / 32) + (threadIdx.x >> 5);
    const int j = tx * 8 + (lane & 7);
    const int kr = blockIdx.y * 4 + (lane >> 3);
    Post post;
    post.kind = 0;
    post.kind2 = 0;
    post.cand_count = 0;
    const int px = kr * a.W + j;
    if (j < a.W && kr < a.rows) {
        const int i = image_row(a, kr);
        F3 uc = f3((float)j - (float)a.W / 2 + 0.5f, (float)a.H / 2 - (float)i - 0.5f, a.z); /* optimized.cu:751, exact in float */
        if (a.camera_mode == 1) /* the viewer's camera, realtime_render.cu:1113 (the camera position is part of the sum there) */
            uc = ((f3(a.camx, a.camy, a.camz) + a.z * f3(a.bz[0], a.bz[1], a.bz[2])) + uc.x * f3(a.bx[0], a.bx[1], a.bx[2])) + uc.y * f3(a.by[0], a.by[1], a.by[2]);
        F3 u0;
        if (!STOCH && JITTER) {
            /* the jitter of the pixel's FIRST sample, optimized.cu:756-758: a function of (seed, global pixel index, sigma) only, kept in a
             * table beside the stream start states (jitter_table below) */
            const float2 jt = __ldg(g.jitter_tab + (size_t)i * a.W + j);
            u0 = normalized3(uc + f3(jt.x, jt.y, 0.f)); /* :758-759 */
        } else if (!STOCH) {
            u0 = normalized3(uc); /* sigma == 0: the jitter terms of :758 are exactly 0 */
        } else {
            /* the pixel's stream: sample 0 starts from curand_init(seed, GLOBAL pixel index, 0) (optimized.cu:745), later
             * samples from where the previous sample's path left it */
            RngRef rng;
            rng.p = g.rng + (size_t)px * 2;
            rng.loaded = false;
            if (g.sample == 0) {
                const uint4* t = g.rng_table + ((size_t)i * a.W + j) * 2;
                const uint4 s0 = t[0], s1 = t[1];
                rng.d = s0.x; rng.v0 = s0.y; rng.v1 = s0.z; rng.v2 = s0.w; rng.v3 = s1.x; rng.v4 = s1.y;
                rng.loaded = true;
                g.total[px] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                g.total[px].w = 0.f; /* no diffuse segment recorded yet in this sample */
            }
            const float r1 = rng_uniform(rng), r2 = rng_uniform(rng); /* :756-757 */
            rng_store(rng);
            const float rad = g.aa_sigma * sqrtf(-2 * (g.libm ? logf(r1) : canon_log(r1)));
            const float ang = (float)(2 * 3.14159265358979323846 * (double)r2);
            const float ca = g.libm ? cosf(ang) : canon_cos(ang), sa = g.libm ? sinf(ang) : canon_sin(ang);
            u0 = normalized3(uc + f3(rad * ca, rad * sa, 0.f)); /* :758-759 */
        }
        if (!STOCH || g.sample == 0) {
            if (a.hit_obj) a.hit_obj[px] = -1;
            if (a.hit_tri) a.hit_tri[px] = -1;
            if (a.hit_t) a.hit_t[px] = RTK_INF;
            if (a.shadow) a.shadow[px] = 2;
        }
        path_advance<COUNT, STOCH, DIFFUSE, LEAN, NS>(h, g, nodes, tris, px, f3(a.camx, a.camy, a.camz), u0, 1.f, 0, false, 0.f, -1, -1, w, post);
    }
    const int slot = post_queries<LEAN>(g, 0, post, px);
    if (LEAN || g.anchored) {
        emit_tasks(g, 0, post, slot);
        answer_deferred<STOCH>(h, blob, g, 0, post, slot);
    }
    flush_work(w, g.c, COUNT);
    TL_END(g);
}

/* ---- wf_shade: one thread per answered closest-hit query of round g.round ------------------------------------------ */
template <bool COUNT, bool STOCH, bool DIFFUSE = false, bool LEAN = false, int NS = 0>
#ifndef RT_SHADE_BLOCKS
#define RT_SHADE_BLOCKS 8
#endif
__global__ void __launch_bounds__(WF_THREADS, RT_SHADE_BLOCKS) wf_shade(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                      const __grid_constant__ WfArgs g) {
    pdl_wait_then_release();
    TL_BEGIN(g);
    const RenderArgs& a = g.a;
    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const int nA = g.c->nA[g.round];
    const int n = nA + g.c->nB[g.round]; /* tree-searched queries from the front of the queue, anchored ones from its end */
    const QEntry* q = g.qA[g.round & 1];
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    const int n_round = (n + 31) & ~31; /* whole warps stay in the loop so that flush_work sees 32 lanes */
    for (int e = blockIdx.x * WF_THREADS + threadIdx.x; e < n_round; e += gridDim.x * WF_THREADS) {
        Post post;
        post.kind = 0;
        post.kind2 = 0;
        post.cand_count = 0;
        int px = 0;
        if (e < n) {
            const float4* p = reinterpret_cast<const float4*>(q + (e < nA ? e : g.qcap - 1 - (e - nA)));
            const float4 p0 = p[0], p1 = p[1], p2 = p[2];
            const F3 O = f3(p0.x, p0.y, p0.z), u = f3(p1.x, p1.y, p1.z);
            float t_hit = p0.w;
            px = (int)(__float_as_uint(p1.w) & ~WF_ROOT_UNTESTED);
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
#ifdef RT_TRACE
            printf("  shade e %d px %d key %llx t_hit %f sidx %d tri %d depth %d\n", e, px, key, t_hit, sidx, tri, depth);
#endif
            path_advance<COUNT, STOCH, DIFFUSE, LEAN, NS>(h, g, nodes, tris, px, O, u, p2.x, depth, true, t_hit, sidx, tri, w, post);
        }
        const int slot = post_queries<LEAN>(g, g.round + 1, post, px);
        if (LEAN || g.anchored) {
            emit_tasks(g, g.round + 1, post, slot);
            answer_deferred<STOCH>(h, blob, g, g.round + 1, post, slot);
        }
    }
    flush_work(w, g.c, COUNT);
    TL_END(g);
}

/* ---- wf_leaves: one thread per (ray, candidate leaf) task of round g.round ----------------------------------------------
 * The reference's mesh query for anchored rays: the leaf's box against the ray with the reference's slab test (certified fast
 * path, exact fallback) and, where it passes, the <= RT_LEAF_MAX triangles of the leaf, two at a time (both records
 * requested before either is tested). A closest-hit task merges
 * its accepted hits into the entry with a 64-bit atomicMin on (t bits, tie-break rank) — the reference's strict minimum
 * with its first-visited rule, order-free; a shadow task that finds a blocker marks the entry and paints
 * the pixel black. Tasks are uniform and independent: no pools, no tail. */
/* the triangles of one leaf against the ray of one queue entry (the T half of a task) */
template <bool STOCH>
__device__ __forceinline__ void leaf_triangles(const SceneHeader& h, const WfArgs& g, const float4* __restrict__ tris, QEntry* q, bool any, int code) {
    const RenderArgs& a = g.a;
    const float4* p = reinterpret_cast<const float4*>(q);
    const float4 p0 = __ldg(p), p1 = __ldg(p + 1); /* origin, direction: read-only in this kernel (L1-cached); res below is not */
    const F3 O = f3(p0.x, p0.y, p0.z), u = f3(p1.x, p1.y, p1.z);
    float t_limit;
    if (any) {
        if (__ldcg(&q->res) != 0ull) return; /* a blocker was already found */
        t_limit = sqrtf(p0.w) * 1.001f + 1e-3f; /* a hit with t > 1.001 sqrt(D2) cannot satisfy the shadow predicate (|t u| ~ t) */
    } else {
        /* the closest sphere of this segment (aux): a mesh hit certainly behind it loses the merge in wf_shade whatever its
         * rank, so the screen may drop it (it keeps equal t: the mesh wins a tie against a sphere of higher id) */
        t_limit = p0.w;
    }
    const int i_begin = code >> 2, i_end = min(i_begin + (code & 3) + 1, h.n_tris);
    for (int k = i_begin; k < i_end; k += 2) {
        const int k2 = min(k + 1, i_end - 1);
        float4 a0, a1, b0, b1;
        ldg256(tris + 4 * (size_t)k, a0, a1);
        ldg256(tris + 4 * (size_t)k2, b0, b1);
        const float4 a2 = __ldg(tris + 4 * (size_t)k + 2);
        const float4 b2 = __ldg(tris + 4 * (size_t)k2 + 2);
        const TriScreen s0 = tri_screen(a0, a1, a2, O, u, t_limit);
        TriScreen s1 = tri_screen(b0, b1, b2, O, u, t_limit);
        s1.maybe &= k2 != k;
        if (s0.maybe | s1.maybe) {
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const TriScreen& sj = j ? s1 : s0;
                if (!sj.maybe) continue;
                const int ij = j ? k2 : k;
                float t;
                if (!tri_finish(sj, t) || !(t > a.eps_tri)) continue;
                if (any) {
                    if (blocks_light(O, u, t, p0.w)) {
                        q->res = 1ull;
                        light_is_blocked<STOCH>(a, g, (int)(__float_as_uint(p1.w) & ~WF_ROOT_UNTESTED), q->packed);
                        return;
                    }
                } else {
                    unsigned rank = (unsigned)ij; /* push_order 1: ascending triangle index */
                    if (a.push_order != 1) rank = tie_rank(ij, __float_as_int(__ldg(tris + 4 * (size_t)ij + 3).w), h.n_tris, 0, a.rank_off_bits);
                    atomicMin(&q->res, ((unsigned long long)__float_as_uint(t) << 32) | rank);
                    if (t < t_limit) t_limit = t;
                }
            }
        }
    }
}

template <bool STOCH>
__global__ void __launch_bounds__(WF_THREADS) wf_leaves(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                       const __grid_constant__ WfArgs g) {
    /* Two phases per warp, both with full lanes: the box phase takes 32 tasks and keeps those whose box passes the slab
     * test (about half) in a small per-warp buffer; whenever the buffer holds 32 of them the triangle phase runs on 32. */
    pdl_wait_then_release();
    TL_BEGIN(g);
    __shared__ int2 hitbuf[WF_THREADS / 32][96]; /* < 32 left over + up to 64 new */
    const unsigned FULL = 0xffffffffu;
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const float4* leaves = reinterpret_cast<const float4*>(blob + h.off_leaves);
    const int n = max(0, min(g.c->nTask[g.round], g.task_cap));
    QEntry* const qA = g.qA[g.round & 1];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    int2* const buf = hitbuf[threadIdx.x >> 5];
    int fill = 0; /* warp-uniform */
    /* a lane takes TWO consecutive tasks per step (one 16-byte load): they usually belong to the same ray, whose slab
     * constants are then formed once, and the two leaf records are in flight together */
    const int n_round = (n + 63) & ~63;
    const int stride = gridDim.x * WF_THREADS * 2;
    int i = (blockIdx.x * WF_THREADS + threadIdx.x) * 2;
    const int4* const tasks2 = reinterpret_cast<const int4*>(g.tasks);
    int4 task = i < n ? __ldg(tasks2 + (i >> 1)) : make_int4(0, 0, 0, 0);
    for (; i < n_round; i += stride) {
        const int4 next = (i + stride < n) ? __ldg(tasks2 + ((i + stride) >> 1)) : make_int4(0, 0, 0, 0); /* requested one iteration ahead */
        bool hit0 = false, hit1 = false;
        int code0 = 0, code1 = 0;
        if (i < n) {
            const bool two = i + 1 < n;
            const QEntry* q = ((task.x < 0) ? g.qS : qA) + (task.x & 0x7fffffff);
            const float4* p = reinterpret_cast<const float4*>(q);
            const float4 p0 = __ldg(p), p1 = __ldg(p + 1);
            float4 l0, l1, m0, m1;
            ldg256(leaves + 2 * (size_t)task.y, l0, l1);
            ldg256(leaves + 2 * (size_t)(two ? task.w : task.y), m0, m1);
            RayCtx ctx = make_ray_ctx(f3(p0.x, p0.y, p0.z), f3(p1.x, p1.y, p1.z), h.box_abs[0], h.box_abs[1], h.box_abs[2]);
            float tn;
            unsigned fb = 0;
            hit0 = slab_fast(l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, ctx, tn, fb);
            code0 = __float_as_int(l1.z);
            if (two) {
                if (task.z != task.x) { /* the second task starts the next ray */
                    const float4* r = reinterpret_cast<const float4*>(((task.z < 0) ? g.qS : qA) + (task.z & 0x7fffffff));
                    const float4 r0 = __ldg(r), r1 = __ldg(r + 1);
                    ctx = make_ray_ctx(f3(r0.x, r0.y, r0.z), f3(r1.x, r1.y, r1.z), h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                }
                hit1 = slab_fast(m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, ctx, tn, fb);
                code1 = __float_as_int(m1.z);
            }
        }
        const unsigned b0 = __ballot_sync(FULL, hit0), b1 = __ballot_sync(FULL, hit1);
        if (hit0) buf[fill + __popc(b0 & lt)] = make_int2(task.x, code0);
        if (hit1) buf[fill + __popc(b0) + __popc(b1 & lt)] = make_int2(task.z, code1);
        fill += __popc(b0) + __popc(b1);
        __syncwarp();
        while (fill >= 32) {
            const int2 t = buf[fill - 32 + lane];
            fill -= 32;
            leaf_triangles<STOCH>(h, g, tris, ((t.x < 0) ? g.qS : qA) + (t.x & 0x7fffffff), t.x < 0, t.y);
            __syncwarp();
        }
        task = next;
    }
    if (lane < fill) {
        const int2 t = buf[lane];
        leaf_triangles<STOCH>(h, g, tris, ((t.x < 0) ? g.qS : qA) + (t.x & 0x7fffffff), t.x < 0, t.y);
    }
    __syncwarp();
    TL_END(g);
}

/* The Box-Muller jitter of the first sample of every pixel of the W x H frame (optimized.cu:756-758): the first two uniforms of the
 * pixel's stream curand_init(seed, pixel, 0), (sigma sqrt(-2 log r1) cos(2 pi r2), sigma sqrt(-2 log r1) sin(2 pi r2)). Depends on the
 * seed, the frame size, sigma and the transcendental canon only: built once beside the start-state table, 8 B per pixel. */
__global__ void __launch_bounds__(256) jitter_table(const uint4* __restrict__ rng_table, unsigned n, float sigma, int libm, float2* __restrict__ out) {
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint4 s0 = __ldg(rng_table + 2 * (size_t)p), s1 = __ldg(rng_table + 2 * (size_t)p + 1);
    RngRef rng;
    rng.p = nullptr;
    rng.d = s0.x; rng.v0 = s0.y; rng.v1 = s0.z; rng.v2 = s0.w; rng.v3 = s1.x; rng.v4 = s1.y;
    rng.loaded = true;
    const float r1 = rng_uniform(rng), r2 = rng_uniform(rng); /* :756-757 */
    const float rad = sigma * sqrtf(-2 * (libm ? logf(r1) : canon_log(r1)));
    const float ang = (float)(2 * 3.14159265358979323846 * (double)r2);
    const float ca = libm ? cosf(ang) : canon_cos(ang), sa = libm ? sinf(ang) : canon_sin(ang);
    out[p] = make_float2(rad * ca, rad * sa);
}

/* ---- wf_fold (stochastic mode): the end of one sample pass. Folds the pixel's diffuse records back to front,
 * c = albedo_i * c + direct_i (optimized.cu:653-660), adds the sample to the pixel's sum and, after the last sample,
 * writes the averaged 8-bit pixel (optimized.cu:764-771). Runs after the pass's last traversal, when every shadow query
 * has had its say about the direct terms. */
__global__ void __launch_bounds__(256) wf_fold(const __grid_constant__ WfArgs g) {
    pdl_wait_then_release();
    const RenderArgs& a = g.a;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= g.npx) return;
    float4 tot = g.total[px];
    const unsigned types = (unsigned)__float_as_int(tot.w);
    F3 ans = f3(0.f, 0.f, 0.f);
    for (int d = a.segments - 1; d >= 0; d--) {
        if ((types >> d) & 1u) {
            const float4* r = g.rec + ((size_t)d * g.npx + px) * 2;
            const float4 dc = r[0], al = r[1];
            ans = f3(al.x, al.y, al.z) * ans + f3(dc.x, dc.y, dc.z);
        }
    }
    const F3 sum = f3(tot.x, tot.y, tot.z) + ans; /* color_out = color_out + color, :762 */
    if (g.last_sample) {
        if (a.linear) {
            const F3 avg = sum / (float)a.num_rays;
            a.linear[px] = make_float4(avg.x, avg.y, avg.z, 0.f);
        } else if (a.rgb) {
            const F3 avg = sum / (float)a.num_rays; /* :764 */
            const float* T = a.gamma_tab + a.gamma_mode * 256;
            a.rgb[(size_t)px * 3 + 0] = (uint8_t)quantise(avg.x, T);
            a.rgb[(size_t)px * 3 + 1] = (uint8_t)quantise(avg.y, T);
            a.rgb[(size_t)px * 3 + 2] = (uint8_t)quantise(avg.z, T);
        }
    } else {
        g.total[px] = make_float4(sum.x, sum.y, sum.z, 0.f);
    }
}

/* ---- wf_traverse: persistent warps drain the closest-hit and shadow queues of round g.round ------------------------
 *
 * The reference's mesh query is an exhaustive search: every box on the way down whose slab test passes is
 * opened, every triangle of every reached leaf is tested, the strictly smallest t wins (ties by visiting order,
 * here by rank). Nothing orders or prunes it, so the search of one ray is a bag of independent (ray, node) and
 * (ray, leaf) tasks. A warp keeps two task pools in shared memory for the 2 x 32 rays it has in flight:
 *
 *   N step  the 32 lanes pop 32 node tasks (any mix of rays), each tests the two child boxes of ITS node against
 *           ITS ray (the ray's slab constants come from a shared-memory slot), and pushes the children that
 *           were hit — inner children to the node pool, leaves to the leaf pool — at ballot-computed offsets.
 *   T step  the lanes pop 32 leaf tasks and test their <= RT_LEAF_MAX triangles; an accepted hit updates the
 *           ray's best (t bits, rank) with a shared-memory atomicMin.
 *
 * All lanes run the same code on different tasks: no per-lane stack, no per-lane loop, no idle lane while a pool
 * holds 32 tasks — a 800-step ray is simply 800 tasks that spread over the lanes, where a thread-per-ray
 * traversal holds one lane for 800 iterations (profiles/r01_notes.md). Rays are admitted in batches of 32; two
 * batches overlap so the drain of one is covered by the other; a batch is complete when its count of
 * outstanding tasks (kept warp-uniform with ballots) reaches zero, then its results go back to the queue
 * entries. Shadow rays stop generating work once a blocker is found.
 */
#define WF_SLOTS 64
#define WF_TPOOL 160

struct WfWarpSmem { /* per warp; followed by the node pool (npool_cap ints) */
    float4 A[WF_SLOTS];  /* rx, ry, rz, M        (RN(1/u), certified-slab margin) */
    float4 B[WF_SLOTS];  /* nox, noy, noz, d2    (-RN(O r); shadow rays: |L - P'|^2, set to -1 once a blocker is found) */
    float4 C[WF_SLOTS];  /* ox, oy, oz, t_limit  (closest-hit rays: 1e9f, shadow rays: finite) */
    float4 D[WF_SLOTS];  /* ux, uy, uz, (int) pixel */
    unsigned long long best[WF_SLOTS]; /* closest: (t bits << 32) | rank, WF_NOHIT if none */
    int entry[WF_SLOTS];               /* queue index of the ray */
    int tpool[WF_TPOOL];
};

#ifndef RT_WIDE_LOADS
#define RT_WIDE_LOADS 4 /* child records requested together by a wide N step */
#endif
#ifndef RT_WIDE_BLOCKS
#define RT_WIDE_BLOCKS 6 /* resident blocks per SM the wide instantiations are compiled for: 80 registers, no spills (8: 64 registers, 86 B of spills, slower) */
#endif
/* ---- the top levels of the tree in shared memory (north_star's "shared-memory-staged top levels"; option "top_smem") ------------------
 * build_top_table copies the first WF_TOP_MAX inner records in breadth-first order (the root, its children, ...: four complete levels)
 * into a table and rewrites the child references that stay inside the table to n_inner + slot, an index no real record has. A TOPS
 * instantiation of wf_traverse stages the table in shared memory once per block and starts every ray at slot 0; an N step takes a record
 * from shared memory when its index is >= n_inner and from global memory otherwise. Same boxes, same references, same visiting order.
 * 15 records (960 B) is what fits: the pools take 26 KB per block and eight blocks share an SM's 228 KB. */
#define WF_TOP_MAX 15
__global__ void build_top_table(const float4* __restrict__ nodes, int n_inner, int root_ref, int limit, float4* __restrict__ top, int* __restrict__ n_top) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int idx[WF_TOP_MAX];
    int count = 0;
    if (root_ref >= 0 && limit > 0) {
        idx[0] = root_ref;
        count = 1;
    }
    for (int s = 0; s < count; s++) {
        const float4* n = nodes + 4 * (size_t)idx[s];
        float4 q3 = n[3];
        int rl = __float_as_int(q3.x), rr = __float_as_int(q3.y);
        if (rl >= 0 && count < limit) {
            idx[count] = rl;
            rl = n_inner + count++;
        }
        if (rr >= 0 && count < limit) {
            idx[count] = rr;
            rr = n_inner + count++;
        }
        q3.x = __int_as_float(rl);
        q3.y = __int_as_float(rr);
        top[4 * s + 0] = n[0];
        top[4 * s + 1] = n[1];
        top[4 * s + 2] = n[2];
        top[4 * s + 3] = q3;
    }
    *n_top = count;
}

template <bool COUNT, bool STOCH, bool WIDE, bool TOPS = false>
__global__ void __launch_bounds__(WF_THREADS, WIDE ? RT_WIDE_BLOCKS : 8) wf_traverse(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob,
                                                         const __grid_constant__ WfArgs g, const int npool_cap) {
    static_assert(!(TOPS && WIDE), "the top table holds two-child records");
    pdl_wait_then_release();
    TL_BEGIN(g);
    extern __shared__ __align__(16) unsigned char wf_smem[];
    __shared__ float4 top_sm[TOPS ? WF_TOP_MAX * 4 : 1];
    if (TOPS) {
        for (int i = threadIdx.x; i < g.n_top * 4; i += WF_THREADS) top_sm[i] = __ldg(g.top + i);
        __syncthreads();
    }
    const RenderArgs& a = g.a;
    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const float4* wnodes = reinterpret_cast<const float4*>(blob + h.off_wide);
    /* WIDE: an N step may push RT_WIDE children per popped task */
    constexpr int FAN = WIDE ? RT_WIDE : 2;
    constexpr int NROOM = 32 * (FAN - 1); /* free node-pool entries a full N step needs */
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const size_t warp_bytes = (sizeof(WfWarpSmem) + (size_t)npool_cap * sizeof(int) + 15) & ~(size_t)15;
    WfWarpSmem& sm = *reinterpret_cast<WfWarpSmem*>(wf_smem + warp * warp_bytes);
    int* npool = reinterpret_cast<int*>(wf_smem + warp * warp_bytes + sizeof(WfWarpSmem));

    const int nA = g.c->nA[g.round], nS = g.anchored ? 0 : g.c->nS[g.round]; /* anchored mode: shadow queries are wf_leaves' */
    const int total = nA + nS;
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;

    int nN = 0, nT = 0;               /* pool fill (warp-uniform) */
    int nSp = 0;                      /* node tasks parked in the warp's global overflow area (warp-uniform) */
    const int half = (npool_cap >> 1) & ~31;
    int out0 = 0, out1 = 0;           /* outstanding tasks of batch 0 / 1 (warp-uniform) */
    int cnt0 = 0, cnt1 = 0;           /* rays admitted in batch 0 / 1; 0 = batch free */
    unsigned vm0 = 0, vm1 = 0;        /* lanes (= slots) of batch 0 / 1 that hold a ray */
    /* Rays are admitted in runs of 2^rs consecutive queue entries. Long queues use runs of 8 (neighbouring pixels share
     * nodes and triangles); when there are few rays per warp (late bounce rounds, a small row shard) the runs shrink,
     * so that the expensive rays of a region are spread over many warps instead of piling up in one. */
    const int n_warps = gridDim.x * (WF_THREADS / 32);
    int rs = g.run_shift;
    if (rs < 0) {
        const int rpw = total / n_warps;
        rs = rpw >= 64 ? 3 : (rpw >= 16 ? 2 : (rpw >= 4 ? 1 : 0));
        /* a mesh of millions of leaves under long queues: no region of the image concentrates the expensive rays, longer runs only add
         * coherence (10 M triangles at 4K: 12.79 -> 12.47 ms; on the cat runs of 16 pile the head's rays up in few warps: 0.374 -> 0.476 ms) */
        if (rpw >= 512 && h.n_leaves > (1 << 20)) rs = 4;
    }
    const int n_runs = (total + (1 << rs) - 1) >> rs; /* runs of 2^rs consecutive queue entries */
    const int n_quarter = (n_runs + 3) >> 2;           /* the queue is consumed as four interleaved quarters */
    const int k_fair = g.fair_share ? max(1, (n_runs + n_warps - 1) / n_warps) : 32; /* runs per warp if the queue were dealt out evenly */
    int last_r0 = 0;                                   /* the queue cursor as this warp last saw it */
    bool exhausted = total == 0;
    bool failed = false;
    unsigned dbgNs = 0, dbgNt = 0, dbgTs = 0, dbgTt = 0, dbgAd = 0;
    unsigned long long dbg_t0 = 0;
    long long cycN = 0, cycT = 0, cycA = 0, cycR = 0, cyc_mark = 0;
    if (COUNT) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));

    for (;;) {
        /* Housekeeping — retiring finished batches, admitting rays, parking / fetching overflow, the exit test — only
         * matters when the pools run low (slots are wanted, or the work is over) or the node pool runs full; while a warp
         * has a comfortable backlog the loop is just the step (all of this is warp-uniform register arithmetic, but it was
         * 17 % of the instructions of a launch when evaluated at every step). */
        if (COUNT) cyc_mark = clock64();
        if (nN + nT < 48 || npool_cap - nN < NROOM) {
        /* ---- retire complete batches: results go back to the queue entries ---------------------------------------- */
        QEntry* const qA = g.qA[g.round & 1];
        QEntry* const qS = g.qS;
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int cnt = b ? cnt1 : cnt0, out = b ? out1 : out0;
            if (cnt > 0 && out == 0) {
                if (((b ? vm1 : vm0) >> lane) & 1u) {
                    const int slot = b * 32 + lane;
                    const int e = sm.entry[slot];
                    if (e < nA) {
                        (qA + e)->res = sm.best[slot];
                    } else if (sm.B[slot].w < 0.f) { /* the light is blocked: the pixel is black (optimized.cu:620-622) */
                        light_is_blocked<STOCH>(a, g, __float_as_int(sm.D[slot].w), STOCH ? (qS + (e - nA))->packed : 0);
                    }
                }
                __syncwarp();
                if (b) cnt1 = 0;
                else cnt0 = 0;
            }
        }
        if (COUNT) { const long long c = clock64(); cycR += c - cyc_mark; cyc_mark = c; }
        /* ---- admit a batch of 32 rays when a batch is free and the pools run low ----------------------------------- */
        if (!exhausted && (cnt0 == 0 || cnt1 == 0) && nN + nT < 48 && nSp == 0 && npool_cap - nN >= 32) { /* room for 32 root tasks */
            /* A batch is up to four runs of 8 consecutive queue entries taken a quarter of the queue apart: neighbouring
             * entries are neighbouring pixels, whose tasks address the same nodes and triangles (few cache lines per
             * step), while the expensive rays, which cluster in the image (the cat's head), are spread over four
             * times as many warps. */
            /* The fuller the pools, the fewer runs are admitted at once: a warp that sits on expensive rays (full pools)
             * takes on at most 8 more, so the expensive rays of a region end up spread over many warps. */
            const int fill = nN + nT;
            int k = max(1, (fill < 12 ? 32 : (fill < 28 ? 16 : 8)) >> rs);
            /* a short queue (a bounce round, a small row shard): no warp takes more than its share at once — with 21 rays per warp the
             * first two thirds of the warps used to leave with 32 each and the last third with none, and the round lasts as long as
             * its fullest warp */
            k = min(k, k_fair);
            /* guided self-scheduling: the last runs of the queue go out one at a time */
            if (4 * n_quarter - last_r0 < g.gss_factor * n_warps) k = 1;
            int r0 = 0;
            if (lane == 0) r0 = atomicAdd(&g.c->head[g.round], k);
            r0 = __shfl_sync(FULL, r0, 0);
            last_r0 = r0;
            if (r0 >= 4 * n_quarter) {
                exhausted = true;
            } else {
                const int b = cnt0 == 0 ? 0 : 1;
                if (COUNT) dbgAd++;
                const int r = r0 + (lane >> rs);
                const int run = (r & 3) * n_quarter + (r >> 2); /* consecutive run ids alternate between the four quarters of the queue */
                const int e = (run << rs) + (lane & ((1 << rs) - 1));
                bool valid = (lane >> rs) < k && r < 4 * n_quarter && run < n_runs && e < total;
                float4 p0, p1;
                RayCtx c;
                if ((STOCH || WIDE) && valid) {
                    QEntry* const qe = e < nA ? (qA + e) : (qS + (e - nA));
                    const float4* p = reinterpret_cast<const float4*>(qe);
                    p0 = __ldcg(p);
                    p1 = __ldcg(p + 1);
                    c = make_ray_ctx(f3(p0.x, p0.y, p0.z), f3(p1.x, p1.y, p1.z), h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                    if (WIDE && c.M != c.M) { /* a zero direction component: outside the wide index's contract */
                        answer_exact<STOCH>(h, blob, g, qe, e >= nA);
                        valid = false;
                    } else if (STOCH && (__float_as_uint(p1.w) & WF_ROOT_UNTESTED)) {
                        /* a bounce ray queued without its root-box test (path_advance); a ray that misses the root box needs
                         * nothing from the mesh — its entry keeps WF_NOHIT — and takes no slot */
                        float tn;
                        unsigned fb = 0;
                        valid = slab_fast(h.root_mn[0], h.root_mn[1], h.root_mn[2], h.root_mx[0], h.root_mx[1], h.root_mx[2], c, tn, fb);
                        p1.w = __uint_as_float(__float_as_uint(p1.w) & ~WF_ROOT_UNTESTED);
                    }
                }
                const unsigned vmask = __ballot_sync(FULL, valid);
                const int take = __popc(vmask);
                if (valid) {
                    if (!STOCH && !WIDE) {
                        const float4* p = reinterpret_cast<const float4*>(e < nA ? (qA + e) : (qS + (e - nA)));
                        p0 = __ldcg(p);
                        p1 = __ldcg(p + 1);
                        c = make_ray_ctx(f3(p0.x, p0.y, p0.z), f3(p1.x, p1.y, p1.z), h.box_abs[0], h.box_abs[1], h.box_abs[2]);
                    }
                    const int slot = b * 32 + lane;
                    const bool any = e >= nA;
                    sm.A[slot] = make_float4(c.rx, c.ry, c.rz, c.M);
                    sm.B[slot] = make_float4(c.nox, c.noy, c.noz, p0.w);
                    /* shadow rays: a hit with t > 1.001 sqrt(D2) cannot satisfy the shadow predicate (|t u| ~ t) */
                    sm.C[slot] = make_float4(p0.x, p0.y, p0.z, any ? (sqrtf(p0.w) * 1.001f + 1e-3f) : RTK_INF);
                    sm.D[slot] = p1;
                    sm.best[slot] = WF_NOHIT;
                    sm.entry[slot] = e;
                    const int task = (slot << 26) | (h.root_ref >= 0 ? (WIDE ? h.wroot_ref : (TOPS ? h.n_inner : h.root_ref)) : (-1 - h.root_ref));
                    const int pos = __popc(vmask & lt_mask);
                    if (h.root_ref >= 0) npool[nN + pos] = task;
                    else sm.tpool[nT + pos] = task;
                }
                if (h.root_ref >= 0) nN += take;
                else nT += take;
                if (b) {
                    cnt1 = take;
                    out1 = take;
                    vm1 = vmask;
                } else {
                    cnt0 = take;
                    out0 = take;
                    vm0 = vmask;
                }
                __syncwarp();
            }
        }
        if (COUNT) { const long long c = clock64(); cycA += c - cyc_mark; cyc_mark = c; }
        /* The node pool is a LIFO over all rays of the warp; nothing bounds it by the tree depth (a step expands up to 32
         * nodes of any mix of rays). When it runs full with no leaf work to drain, its older half is parked in global
         * memory and brought back when the pool has emptied. Rare (the cat: a few warps per frame). */
        if (nN == 0 && nSp > 0) {
            int* const spill = g.spill + (size_t)(blockIdx.x * (WF_THREADS / 32) + warp) * g.spill_cap;
            const int m = min(nSp, half);
            for (int i = lane; i < m; i += 32) npool[i] = spill[nSp - m + i];
            nSp -= m;
            nN = m;
            __syncwarp();
        } else if (npool_cap - nN < NROOM && nT == 0 && nN > half) {
            if (nSp + half > g.spill_cap) { /* overflow area exhausted: reported as an error by rt_scene_sync */
                failed = true;
                break;
            }
            int* const spill = g.spill + (size_t)(blockIdx.x * (WF_THREADS / 32) + warp) * g.spill_cap;
            for (int i = lane; i < half; i += 32) spill[nSp + i] = npool[i];
            __syncwarp();
            const int rest = nN - half;
            for (int i0 = 0; i0 < rest; i0 += 32) { /* slide the younger tasks down, 32 at a time (read, sync, write) */
                const bool mv = i0 + lane < rest;
                const int v = mv ? npool[half + i0 + lane] : 0;
                __syncwarp();
                if (mv) npool[i0 + lane] = v;
                __syncwarp();
            }
            nSp += half;
            nN = rest;
            __syncwarp();
        }
        if (nN == 0 && nT == 0) {
            if (cnt0 == 0 && cnt1 == 0 && exhausted) break;
            if (out0 != 0 || out1 != 0) { /* cannot happen: outstanding tasks with empty pools */
                failed = true;
                break;
            }
            continue;
        }
        } /* housekeeping */

        /* an N step pops cnt tasks and may push 2 cnt: it needs cnt free entries */
        const int room = npool_cap - nN;
        if (nT >= 32 || nN == 0 || (room < 8 * (FAN - 1) && nT > 0)) {
            /* ---- T step: one leaf (<= RT_LEAF_MAX triangles) per lane ---------------------------------------------- */
            const int cnt = min(nT, 32);
            if (COUNT) { dbgTs++; dbgTt += cnt; }
            const bool have = lane < cnt;
            int slot = 0;
            bool live = false;
            if (have) {
                const int task = sm.tpool[nT - 1 - lane];
                slot = (unsigned)task >> 26;
                const float d2 = sm.B[slot].w;
                live = !(d2 < 0.f);
                if (live) {
                    const int code = task & 0x3ffffff; /* first triangle << 2 | count - 1 */
                    const float4 c4 = sm.C[slot], d4 = sm.D[slot];
                    const F3 O = f3(c4.x, c4.y, c4.z), u = f3(d4.x, d4.y, d4.z);
                    const bool any = c4.w < RTK_INF;
                    const int i_begin = code >> 2, i_end = min(i_begin + (code & 3) + 1, h.n_tris);
                    float t_limit = c4.w;
                    if (!any) {
                        const unsigned long long cur = sm.best[slot];
                        if (cur != WF_NOHIT) t_limit = __uint_as_float((unsigned)(cur >> 32));
                    }
                    /* two triangles per iteration, both records requested before either is tested: one L2 round trip per
                     * pair instead of one per triangle (a lone warp's T step was four dependent round trips long) */
                    for (int i = i_begin; i < i_end; i += 2) {
                        const int i2 = min(i + 1, i_end - 1);
                        float4 p0, p1, r0, r1;
                        ldg256(tris + 4 * (size_t)i, p0, p1);
                        ldg256(tris + 4 * (size_t)i2, r0, r1);
                        const float4 p2 = __ldg(tris + 4 * (size_t)i + 2);
                        const float4 r2 = __ldg(tris + 4 * (size_t)i2 + 2);
                        const TriScreen s0 = tri_screen(p0, p1, p2, O, u, t_limit);
                        TriScreen s1 = tri_screen(r0, r1, r2, O, u, t_limit);
                        s1.maybe &= i2 != i;
                        if (COUNT) w.tris += 1 + (i2 != i);
                        bool stop = false;
                        if (s0.maybe | s1.maybe) { /* rare per lane (4 % of the tests) */
#pragma unroll
                            for (int k = 0; k < 2; k++) {
                                const TriScreen& sk = k ? s1 : s0;
                                if (!sk.maybe) continue;
                                const int ik = k ? i2 : i;
                                if (COUNT) w.tri_exact++;
                                float t;
                                if (!tri_finish(sk, t) || !(t > a.eps_tri)) continue;
                                if (any) {
                                    if (blocks_light(O, u, t, d2)) {
                                        sm.B[slot].w = -1.f; /* no more work for this ray */
                                        stop = true;
                                        break;
                                    }
                                } else {
                                    unsigned rank = (unsigned)ik; /* push_order 1: ascending triangle index */
                                    if (a.push_order != 1) rank = tie_rank(ik, __float_as_int(__ldg(tris + 4 * (size_t)ik + 3).w), h.n_tris, 0, a.rank_off_bits);
                                    atomicMin(&sm.best[slot], ((unsigned long long)__float_as_uint(t) << 32) | rank);
                                    if (t < t_limit) t_limit = t;
                                }
                            }
                        }
                        if (stop) break;
                    }
                }
            }
            const unsigned in1 = __ballot_sync(FULL, have && slot >= 32);
            out1 -= __popc(in1);
            out0 -= cnt - __popc(in1);
            nT -= cnt;
            if (COUNT) { __syncwarp(); cycT += clock64() - cyc_mark; }
        } else {
            /* ---- N step: one inner node (two child boxes) per lane ------------------------------------------------- */
            if (room < FAN - 1) { /* node pool full and no leaf work to drain: rt_scene_sync falls back to render_mega */
                failed = true;
                break;
            }
            const int cnt = min(min(nN, 32), room / (FAN - 1));
            if (COUNT) { dbgNs++; dbgNt += cnt; }
            const bool have = lane < cnt;
            int slot = 0;
            if (WIDE) {
                /* ---- wide N step: one wide node (<= RT_WIDE child boxes, one LDG.E.256 each) per lane ------------------ */
                unsigned mN = 0, mT = 0; /* bit k: child k is opened (inner) / queued (leaf) */
                int ref[RT_WIDE];
#pragma unroll
                for (int k = 0; k < RT_WIDE; k++) ref[k] = 0;
                if (have) {
                    const int task = npool[nN - 1 - lane];
                    slot = (unsigned)task >> 26;
                    const float4 a4 = sm.A[slot], b4 = sm.B[slot];
                    if (!(b4.w < 0.f)) {
                        const int payload = task & 0x3ffffff;
                        const unsigned child_mask = (2u << (payload & 3)) - 1u;
                        if (COUNT) w.nodes++;
                        const float4* n = wnodes + 8 * (size_t)(payload >> 2);
                        RayCtx c;
                        c.rx = a4.x;
                        c.ry = a4.y;
                        c.rz = a4.z;
                        c.M = a4.w;
                        c.nox = b4.x;
                        c.noy = b4.y;
                        c.noz = b4.z;
                        unsigned undecided = 0;
#pragma unroll
                        for (int k0 = 0; k0 < RT_WIDE; k0 += RT_WIDE_LOADS) {
                            float4 q[2 * RT_WIDE_LOADS];
#pragma unroll
                            for (int k = 0; k < RT_WIDE_LOADS; k++) ldg256(n + 2 * (k0 + k), q[2 * k], q[2 * k + 1]); /* one 128-B line; unused slots repeat child 0 */
#pragma unroll
                            for (int k = 0; k < RT_WIDE_LOADS; k++) {
                                float tn;
                                const int r = slab_certified(q[2 * k].x, q[2 * k].y, q[2 * k].z, q[2 * k].w, q[2 * k + 1].x, q[2 * k + 1].y, c, tn);
                                const int rk = __float_as_int(q[2 * k + 1].z);
                                ref[k0 + k] = rk;
                                /* undecided by the margin: an inner child is opened (conservative, rt_layout.h); a leaf is
                                 * decided by the reference's own slab test below */
                                const unsigned bit = 1u << (k0 + k);
                                if (rk >= 0) {
                                    if (r >= 0) mN |= bit;
                                } else {
                                    if (r > 0) mT |= bit;
                                    if (r == 0) undecided |= bit;
                                }
                            }
                        }
                        mN &= child_mask;
                        mT &= child_mask;
                        undecided &= child_mask;
                        if (undecided) { /* rare */
                            const float4 c4 = sm.C[slot], d4 = sm.D[slot];
#pragma unroll 1
                            for (int k = 0; k < RT_WIDE; k++)
                                if ((undecided >> k) & 1u) {
                                    float4 q0, q1;
                                    ldg256(n + 2 * k, q0, q1);
                                    if (COUNT) w.slab_fallbacks++;
                                    if (slab_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, f3(c4.x, c4.y, c4.z), f3(d4.x, d4.y, d4.z))) mT |= 1u << k;
                                }
                        }
                    }
                }
                nN -= cnt;
                const int cN = __popc(mN), cT = __popc(mT);
                /* inclusive warp scan of the (node, leaf) push counts, packed 16 + 16 bits */
                int incl = cN | (cT << 16);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, d);
                    if (lane >= d) incl += v;
                }
                const int tot = __shfl_sync(FULL, incl, 31);
                int pN = nN + (incl & 0xffff) - cN, pT = nT + (incl >> 16) - cT;
                const int tagged = slot << 26;
#pragma unroll
                for (int k = 0; k < RT_WIDE; k++) {
                    if ((mN >> k) & 1u) npool[pN++] = tagged | ref[k];
                    if ((mT >> k) & 1u) sm.tpool[pT++] = tagged | (-1 - ref[k]);
                }
                const unsigned in1 = __ballot_sync(FULL, have && slot >= 32);
                const int pushedN = tot & 0xffff, pushedT = tot >> 16;
                nN += pushedN;
                nT += pushedT;
                if (COUNT) w.max_stack = max(w.max_stack, (unsigned)nN);
                const int pushed1 = __reduce_add_sync(FULL, (have && slot >= 32) ? cN + cT : 0);
                out1 += pushed1 - __popc(in1);
                out0 += (pushedN + pushedT - pushed1) - (cnt - __popc(in1));
            } else {
            bool pNL = false, pNR = false, pTL = false, pTR = false;
            int refL = 0, refR = 0;
            if (have) {
                const int task = npool[nN - 1 - lane];
                slot = (unsigned)task >> 26;
                const float4 a4 = sm.A[slot], b4 = sm.B[slot];
                if (!(b4.w < 0.f)) {
                    const int node = task & 0x3ffffff;
                    float4 q0, q1, q2, q3f;
                    if (TOPS && node >= h.n_inner) { /* a record of the staged top levels */
                        const float4* n = top_sm + 4 * (node - h.n_inner);
                        q0 = n[0];
                        q1 = n[1];
                        q2 = n[2];
                        q3f = n[3];
                    } else {
                        const float4* n = nodes + 4 * (size_t)node;
                        ldg256(n, q0, q1);
                        ldg256(n + 2, q2, q3f);
                    }
                    const int2 q3 = make_int2(__float_as_int(q3f.x), __float_as_int(q3f.y));
                    if (COUNT && __float_as_int(q3f.z) == 0) w.nodes++;
                    RayCtx c;
                    c.rx = a4.x;
                    c.ry = a4.y;
                    c.rz = a4.z;
                    c.M = a4.w;
                    c.nox = b4.x;
                    c.noy = b4.y;
                    c.noz = b4.z;
                    float tn;
                    int rL = slab_certified(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, c, tn);
                    int rR = slab_certified(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, c, tn);
                    if (rL == 0 || rR == 0) { /* undecided by the margin: the reference's own slab test decides */
                        const float4 c4 = sm.C[slot], d4 = sm.D[slot];
                        const F3 O = f3(c4.x, c4.y, c4.z), u = f3(d4.x, d4.y, d4.z);
                        if (rL == 0) {
                            if (COUNT) w.slab_fallbacks++;
                            rL = slab_exact(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, O, u) ? 1 : -1;
                        }
                        if (rR == 0) {
                            if (COUNT) w.slab_fallbacks++;
                            rR = slab_exact(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, O, u) ? 1 : -1;
                        }
                    }
                    refL = q3.x;
                    refR = q3.y;
                    pNL = rL > 0 && refL >= 0;
                    pTL = rL > 0 && refL < 0;
                    pNR = rR > 0 && refR >= 0;
                    pTR = rR > 0 && refR < 0;
                }
            }
            nN -= cnt;
            const unsigned bNL = __ballot_sync(FULL, pNL), bNR = __ballot_sync(FULL, pNR);
            const unsigned bTL = __ballot_sync(FULL, pTL), bTR = __ballot_sync(FULL, pTR);
            const unsigned in1 = __ballot_sync(FULL, have && slot >= 32);
            const int tagged = slot << 26;
            if (pNL) npool[nN + __popc(bNL & lt_mask)] = tagged | refL;
            if (pNR) npool[nN + __popc(bNL) + __popc(bNR & lt_mask)] = tagged | refR;
            if (pTL) sm.tpool[nT + __popc(bTL & lt_mask)] = tagged | (-1 - refL);
            if (pTR) sm.tpool[nT + __popc(bTL) + __popc(bTR & lt_mask)] = tagged | (-1 - refR);
            nN += __popc(bNL) + __popc(bNR);
            nT += __popc(bTL) + __popc(bTR);
            if (COUNT) w.max_stack = max(w.max_stack, (unsigned)nN);
            const int mine = (int)pNL + (int)pNR + (int)pTL + (int)pTR;
            const int pushed1 = __reduce_add_sync(FULL, (have && slot >= 32) ? mine : 0);
            const int pushed = __popc(bNL) + __popc(bNR) + __popc(bTL) + __popc(bTR);
            out1 += pushed1 - __popc(in1);
            out0 += (pushed - pushed1) - (cnt - __popc(in1));
            }
            if (COUNT) { __syncwarp(); cycN += clock64() - cyc_mark; }
        }
        __syncwarp(); /* pool and slot writes of this step are visible to the next pop */
    }
    if (failed && lane == 0) atomicExch(&g.sticky[1], 1ull);
    if (COUNT && g.dbg_warps && lane == 0) {
        unsigned long long t1;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int* d = g.dbg_warps + ((size_t)g.round * gridDim.x * (WF_THREADS / 32) + blockIdx.x * (WF_THREADS / 32) + warp) * 16;
        d[0] = (int)(unsigned)dbg_t0;
        d[1] = (int)(unsigned)t1;
        d[2] = (int)dbgNs;
        d[3] = (int)dbgNt;
        d[4] = (int)dbgTs;
        d[5] = (int)dbgTt;
        d[6] = (int)dbgAd;
        d[7] = (int)smid;
        d[8] = (int)(cycN >> 4);
        d[9] = (int)(cycT >> 4);
        d[10] = (int)(cycA >> 4);
        d[11] = (int)(cycR >> 4);
    }
    if (COUNT && lane == 0) {
        atomicAdd(&g.c->dbg[0], (unsigned long long)dbgNs);
        atomicAdd(&g.c->dbg[1], (unsigned long long)dbgNt);
        atomicAdd(&g.c->dbg[2], (unsigned long long)dbgTs);
        atomicAdd(&g.c->dbg[3], (unsigned long long)dbgTt);
        atomicAdd(&g.c->dbg[4], (unsigned long long)dbgAd);
    }
    flush_work(w, g.c, COUNT);
    TL_END(g);
}

} // namespace rtk
