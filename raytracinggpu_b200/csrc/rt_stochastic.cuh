/*
 * rt_stochastic.cuh — the stochastic mode of the reference's per-pixel path: Box-Muller anti-aliasing jitter
 * (optimized.cu:753-760) and the cosine-weighted indirect bounce at every diffuse hit (optimized.cu:631-649), driven by
 * the reference's random stream: cuRAND XORWOW, curand_init(123456, pixel index, 0) + curand_uniform
 * (optimized.cu:32-37, 745).
 *
 * The number of uniforms a sample consumes depends on its path (two per diffuse hit), and the samples of a pixel share
 * one stream, so a pixel's samples are inherently sequential: this mode runs one thread per pixel over the whole
 * path (the traversal is render_mega's certified mesh_query), not the wavefront pipeline.
 *
 * What is NOT repeated from the reference: curand_init's skip-ahead to the pixel's subsequence (up to ~18 products of
 * a 160-bit state with a 160x160 bit matrix, the reference's most expensive per-pixel step, paid again at every
 * launch). The start states depend only on (seed, pixel index): xorwow_init_states computes them once per
 * (seed, W, H) with the library's own curand_init and render_stoch loads 24 B per pixel.
 *
 * Transcendentals: evaluated in double and rounded once to float (see oracle/rt_oracle.cpp, canon_log/cos/sin): the
 * reference's GPU build uses --use_fast_math intrinsics and its CPU build libm, so no two reference builds agree in
 * the last bits; double evaluation makes this path and the oracle agree.
 */
#pragma once
#include "rt_kernels.cuh"

#include <curand_kernel.h>

namespace rtk {

#define RT_STOCH_MAX_SEGMENTS 16

struct __align__(16) XorwowState { /* 32 B: two uint4 per pixel (the wavefront pipeline reads it that way) */
    unsigned int d, v[5], pad[2];
};

/* start state of every pixel of a W x H frame: the library's curand_init(seed, pixel, 0) */
__global__ void xorwow_init_states(unsigned long long seed, unsigned int npx, XorwowState* __restrict__ out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    curandStateXORWOW_t st;
    curand_init(seed, (unsigned long long)i, 0ull, &st);
    XorwowState o;
    o.d = st.d;
#pragma unroll
    for (int k = 0; k < 5; k++) o.v[k] = st.v[k];
    o.pad[0] = o.pad[1] = 0;
    out[i] = o;
}

/* Self-test probe: for each listed subsequence the library's start state and its first four curand_uniform values. */
__global__ void selftest_xorwow(unsigned long long seed, const unsigned int* __restrict__ subseq, int n, unsigned int* __restrict__ states6, float* __restrict__ uniforms4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    curandStateXORWOW_t st;
    curand_init(seed, (unsigned long long)subseq[i], 0ull, &st);
    states6[6 * i] = st.d;
    for (int k = 0; k < 5; k++) states6[6 * i + 1 + k] = st.v[k];
    for (int k = 0; k < 4; k++) uniforms4[4 * i + k] = curand_uniform(&st);
}

/* curand(): Marsaglia xorwow + Weyl sequence (curand_kernel.h:863-874); curand_uniform: x * 2^-32 + 2^-33 (the product is exact) */
__device__ __forceinline__ float xorwow_uniform(XorwowState& s) {
    const unsigned int t = s.v[0] ^ (s.v[0] >> 2);
    s.v[0] = s.v[1];
    s.v[1] = s.v[2];
    s.v[2] = s.v[3];
    s.v[3] = s.v[4];
    s.v[4] = (s.v[4] ^ (s.v[4] << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return (float)(s.v[4] + s.d) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

template <bool COUNT>
__global__ void __launch_bounds__(128) render_stoch(const __grid_constant__ SceneHeader h, const unsigned char* __restrict__ blob, const RenderArgs a,
                                                    const XorwowState* __restrict__ states, float aa_sigma, int indirect, int libm) {
    __shared__ float s_gamma[256];
    for (int k = threadIdx.x; k < 256; k += blockDim.x) s_gamma[k] = a.gamma_tab[a.gamma_mode * 256 + k];
    __syncthreads();
    const float4* nodes = reinterpret_cast<const float4*>(blob + h.off_nodes);
    const float4* tris = reinterpret_cast<const float4*>(blob + h.off_tris);
    const int tiles_x = (a.W + 15) >> 4;
    const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = tile_x * 16 + (warp & 1) * 8 + (lane & 7);
    const int k = tile_y * 8 + (warp >> 1) * 4 + (lane >> 3);
    Work w;
    w.rays = w.nodes = w.tris = w.max_stack = w.slab_fallbacks = w.tri_exact = 0;
    if (j < a.W && k < a.rows) {
        const int i = image_row(a, k);
        XorwowState rng = states[(size_t)i * a.W + j]; /* the GLOBAL pixel index keys the stream (optimized.cu:745) */
        const F3 uc = f3((float)j - (float)a.W / 2 + 0.5f, (float)a.H / 2 - (float)i - 0.5f, a.z);
        const F3 cam = f3(a.camx, a.camy, a.camz);
        const F3 Lp = f3(h.L[0], h.L[1], h.L[2]);
        const float eps = a.eps_surface;
        const int segments = min(a.segments, RT_STOCH_MAX_SEGMENTS);
        F3 total = f3(0.f, 0.f, 0.f);
        int first_obj = -1, first_tri = -1, first_shadow = 2;
        float first_t = RTK_INF;
        for (int s = 0; s < a.num_rays; s++) {
            const float r1 = xorwow_uniform(rng), r2 = xorwow_uniform(rng); /* :756-757 */
            const float rad = aa_sigma * sqrtf(-2 * (libm ? logf(r1) : canon_log(r1)));
            const float ang = (float)(2 * 3.14159265358979323846 * (double)r2);
            F3 O = cam;
            F3 u = normalized(uc + f3(rad * (libm ? cosf(ang) : canon_cos(ang)), rad * (libm ? sinf(ang) : canon_sin(ang)), 0.f)); /* :758-759 */
            float n_ray = 1.f;
            unsigned types = 0; /* bit d: segment d ended on a diffuse surface */
            F3 direct[RT_STOCH_MAX_SEGMENTS], albedo_of[RT_STOCH_MAX_SEGMENTS];
            for (int depth = 0; depth < segments; depth++) {
                const SurfaceHit hit = intersect_all<COUNT, true>(h, nodes, tris, O, u, a.eps_tri, a.push_order, w);
                if (s == 0 && depth == 0) {
                    first_obj = hit.obj;
                    first_tri = hit.tri;
                    first_t = hit.t;
                }
                if (hit.obj < 0) break;
                const F3 P = O + hit.t * u;
                F3 N, albedo;
                int mirror;
                float n_in, n_out;
                if (hit.sidx >= 0) {
                    const DevSphere& sp = h.spheres[hit.sidx];
                    N = normalized(P - f3(sp.cx, sp.cy, sp.cz));
                    albedo = f3(sp.ax, sp.ay, sp.az);
                    mirror = sp.mirror;
                    n_in = sp.n_in;
                    n_out = sp.n_out;
                } else {
                    const float4 nh = __ldg(tris + 4 * (size_t)hit.tri + 3);
                    N = f3(nh.x, nh.y, nh.z);
                    albedo = f3(h.mesh_albedo[0], h.mesh_albedo[1], h.mesh_albedo[2]);
                    mirror = h.mesh_mirror;
                    n_in = h.mesh_n_in;
                    n_out = h.mesh_n_out;
                }
                if (mirror) { /* :572-579 */
                    const F3 dir = u - (2 * dot(u, N)) * N;
                    O = P + eps * N;
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
                        const F3 dir = u - (2 * un) * N;
                        O = P + eps * N;
                        u = dir;
                        continue;
                    }
                    const F3 Ncomp = (-sqrtf(1 - (ratio * ratio) * (1 - un * un))) * N;
                    const F3 Tcomp = ratio * (u - un * N);
                    O = P - eps * N;
                    u = Ncomp + Tcomp;
                    n_ray = out2in ? n_in : n_out;
                } else { /* diffuse :610-650 */
                    const F3 Padj = P + eps * N;
                    const F3 toL = Lp - Padj;
                    const float D2 = norm2(toL);
                    const F3 su = toL / sqrtf(D2);
                    const bool blocked = light_blocked<COUNT, true>(h, nodes, tris, Padj, su, D2, a.eps_tri, a.push_order, w);
                    F3 dcol = f3(0.f, 0.f, 0.f);
                    if (!blocked) {
                        const F3 PL = Lp - P;
                        const F3 wl = normalized(PL);
                        const float ndl = dot(N, wl);
                        const float lambert = (ndl < 0.f) ? 0.f : ndl;
                        const float l = (float)((double)h.intensity / (12.566370614359172 * (double)norm2(PL)) * (double)lambert);
                        dcol = (l * albedo) / 3.14159274f;
                    }
                    if (s == 0 && depth == 0) first_shadow = blocked ? 1 : 0;
                    direct[depth] = dcol;
                    types |= 1u << depth;
                    if (!indirect) {
                        albedo_of[depth] = f3(0.f, 0.f, 0.f);
                        break;
                    }
                    const float q1 = xorwow_uniform(rng), q2 = xorwow_uniform(rng); /* :633-634, the last segment included */
                    const float an = (float)(2 * 3.14159265358979323846 * (double)q1);
                    const float sq = sqrtf(1 - q2);
                    const float x = (libm ? cosf(an) : canon_cos(an)) * sq, y = (libm ? sinf(an) : canon_sin(an)) * sq, z = sqrtf(q2);
                    const F3 T1 = normalized((fabsf(N.y) != 0 && fabsf(N.x) != 0) ? f3(-N.y, N.x, 0.f) : f3(-N.z, 0.f, N.x));
                    const F3 T2 = cross(N, T1);
                    u = (x * T1 + y * T2) + z * N;
                    O = Padj;
                    n_ray = 1.f;
                    albedo_of[depth] = albedo;
                }
            }
            F3 ans = f3(0.f, 0.f, 0.f); /* fold back to front :653-660 */
            for (int d = segments - 1; d >= 0; d--)
                if ((types >> d) & 1u) ans = albedo_of[d] * ans + direct[d];
            total = total + ans;
        }
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
        if (a.shadow) a.shadow[px] = (uint8_t)first_shadow;
    }
    const unsigned int rays = __reduce_add_sync(0xffffffffu, w.rays);
    if (lane == 0 && rays) atomicAdd(a.counters + 0, (unsigned long long)rays);
    if (COUNT) {
        const unsigned int nn = __reduce_add_sync(0xffffffffu, w.nodes), tt = __reduce_add_sync(0xffffffffu, w.tris);
        if (lane == 0) {
            atomicAdd(a.counters + 1, (unsigned long long)nn);
            atomicAdd(a.counters + 2, (unsigned long long)tt);
        }
    }
}

} // namespace rtk
