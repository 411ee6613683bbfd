/*
 * rt_device.cu — the device half of the C ABI (include/rt_b200.h): rt_scene_* and rt_render.
 *
 * Replaces the host steps of the reference launcher around KernelLaunch:
 *   cudaMalloc / cudaMemcpy H2D of arr_bvh, indices, vertices     optimized.cu:791-826
 *   per-block scene construction in shared memory                  optimized.cu:679-743 (here: one header, built once)
 *   KernelLaunch<<<>>> + cudaDeviceSynchronize + cudaMemcpy D2H    optimized.cu:828-856
 * There is no CPU path in this file: every entry point needs a CUDA device and fails with RT_ERR_CUDA otherwise.
 */
#include "host_common.h"
#include "rt_kernels.cuh"
#include "rt_layout.h"
#include "rt_rows.h"
#include <cub/cub.cuh>
#include <cuda.h> /* types of the stream memory operations only: the entry points come from cudaGetDriverEntryPoint */
#include "rt_wavefront.cuh"
#include "rt_stochastic.cuh"
#include "rt_bvh_build.cuh"
#include "rt_relayout.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <string>
#include <vector>

#define RT_NCOUNTERS 8
#define RT_RENDER_RETRY_ 0x40000000u /* internal: this call is itself a repeat */
#define RT_MAX_STRIPS 8
#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (expr);                                                                              \
        if (e__ != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

/* Tuning and cross-check options of a scene (rt_scene_set_option). Defaults are the production choices; every code path
 * they select gives the same results. Read once per rt_render call from the scene, never from the environment. */
struct RtOptions {
    int variant = 2;         /* 0 render_mega, exact arithmetic; 1 render_mega with the certified fast paths; 2 wavefront pipeline */
    int strips = 0;          /* row bands rendered concurrently on separate streams; 0: chosen per call */
    int anchored = -1;       /* anchored-ray bins (rt_bins.cuh): -1 by mesh size, 0 off (tree search), 1 on */
    int wide = -1;           /* 4-wide index for wf_traverse: -1 on for stochastic indirect bounces, 0 off, 1 on */
    int wide_count = 0;      /* instrumented renders walk the wide index (node_visits then counts wide nodes) */
    int bins_r = 0;          /* cells per face side of the bins; 0: 1024 (2048 beyond 200 k leaves) */
    int task_factor = 0;     /* (ray, leaf) task buffer entries per pixel; 0: 8, doubled after an overflow */
    int npool_cap = 0;       /* node pool entries per traversal warp; 0: by tree depth (a small pool forces the spill path) */
    int stoch_mega = 0;      /* stochastic mode through the thread-per-pixel kernel render_stoch (in-library cross-check) */
    int diffuse_kernels = 1; /* kernels without reflection / refraction code when no object needs it */
    int side_stream = 1;     /* tree search of a round beside wf_leaves on a side stream */
    int run_shift = -1;      /* log2 of wf_traverse's admission run length; -1: per launch */
    int gss = 2;             /* guided self-scheduling factor of wf_traverse */
    int leaves_blocks = 0;   /* resident blocks per SM for wf_leaves; 0: occupancy */
    int transcendentals = 1; /* stochastic mode, log / cos / sin of optimized.cu:756-758, 635-636: 1 CUDA's single-precision logf / cosf /
                              * sinf — what optimized.cu itself calls; frames equal those of its IEEE build bit for bit, and the oracle
                              * restates the functions for the CPU (rt_oracle.cpp: cuda_logf ...); 0 evaluated in double and rounded once
                              * (the oracle's other canon: what a correctly rounded libm would give; 3-4x the instructions) */
    int graph = 1;           /* replay a recorded CUDA graph when a frame repeats the previous call's plan */
    int fair_share = 1;      /* wf_traverse: an admission takes at most the warp's even share of a short queue */
    int top_smem = 0;        /* wf_traverse takes the top four levels of the tree from shared memory (rt_wavefront.cuh: build_top_table). Measured
                              * (profiles/r02_notes.md): see there; off by default */
    int split = 0;           /* two row bands: per cent of the rows in the first band; 0: equal halves */
    int six = 1;             /* kernels with the sphere loops unrolled for the reference's room of exactly six spheres (constants as direct operands) */
    int one_shot = 1;        /* stochastic frames of one sample and one segment through the deterministic pipeline with jittered camera rays */
    int pdl = 0;             /* programmatic dependent launch between consecutive kernels of a band (LaunchChain). Measured (profiles/r02_notes.md):
                              * no gain once a frame is a replayed graph (1/8 shard 0.299 vs 0.300 ms), +1 % on the full frame: off */
    int debug_times = 0, debug_pool = 0, debug_bins = 0, debug_cost = 0;
    char debug_warps[256] = {0}; /* RT_DEBUG_WARPS=<file> at scene creation: per-warp timeline of wf_traverse (count_work renders) */
};

struct RtOptionKey {
    const char* name;
    int RtOptions::*field;
    int lo, hi;
};
static const RtOptionKey kOptionKeys[] = {
    {"variant", &RtOptions::variant, 0, 2},
    {"strips", &RtOptions::strips, 0, 8},
    {"anchored", &RtOptions::anchored, -1, 1},
    {"wide", &RtOptions::wide, -1, 1},
    {"wide_count", &RtOptions::wide_count, 0, 1},
    {"bins_r", &RtOptions::bins_r, 0, 4096},
    {"task_factor", &RtOptions::task_factor, 0, 256},
    {"npool_cap", &RtOptions::npool_cap, 0, 4096},
    {"stoch_mega", &RtOptions::stoch_mega, 0, 1},
    {"diffuse_kernels", &RtOptions::diffuse_kernels, 0, 1},
    {"side_stream", &RtOptions::side_stream, 0, 1},
    {"run_shift", &RtOptions::run_shift, -1, 5},
    {"gss", &RtOptions::gss, 0, 64},
    {"leaves_blocks", &RtOptions::leaves_blocks, 0, 32},
    {"transcendentals", &RtOptions::transcendentals, 0, 1},
    {"graph", &RtOptions::graph, 0, 1},
    {"pdl", &RtOptions::pdl, 0, 1},
    {"one_shot", &RtOptions::one_shot, 0, 1},
    {"six", &RtOptions::six, 0, 1},
    {"split", &RtOptions::split, 0, 95},
    {"top_smem", &RtOptions::top_smem, 0, 1},
    {"fair_share", &RtOptions::fair_share, 0, 1},
    {"debug_times", &RtOptions::debug_times, 0, 1},
    {"debug_pool", &RtOptions::debug_pool, 0, 1},
    {"debug_bins", &RtOptions::debug_bins, 0, 1},
    {"debug_cost", &RtOptions::debug_cost, 0, 1},
};

struct rt_scene {
    int device = 0;
    RtOptions opt;
    bool plan_valid = false; /* unused marker kept for setters that invalidate recorded frames (the key comparison decides) */
    /* the recorded frame (rt_render); two of them, because frames with asynchronous host outputs alternate between two scratch sets */
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    std::vector<unsigned char> graph_key[2];   /* what each was recorded for */
    std::vector<unsigned char> last_key[2];    /* the previous call's key (per set): a frame is recorded when it repeats */
    int graph_launches[2] = {0, 0};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    SceneHeader header;         /* host copy, passed to the kernels by value */
    unsigned char* blob = nullptr;
    size_t blob_bytes = 0;
    bool header_dirty = true;

    float* gamma_tab = nullptr;              /* device, 2 x 256 */
    unsigned long long* counters = nullptr;  /* device, 4 */
    unsigned long long* h_counters = nullptr; /* pinned */

    /* scratch outputs for host-pointer callers */
    unsigned char* scratch[2][5] = {};
    size_t scratch_bytes[2][5] = {};
    /* host outputs of frames enqueued with RT_RENDER_NO_SYNC leave on a copy stream of their own, from two alternating sets of scratch
     * buffers: the copy-back of frame k overlaps the kernels of frame k + 1 (which write the other set) */
    int scratch_set = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_done[2] = {nullptr, nullptr}, frame_done = nullptr;
    bool copy_recorded[2] = {false, false};
    bool copies_pending = false;
    /* the pinned staging of rt_scene_set_mesh's fast path, two halves: the upload of step k + 1 does not wait for the frame of step k */
    unsigned char* pin_alt = nullptr;
    size_t pin_alt_bytes = 0;
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    bool pin_ev_recorded[2] = {false, false};
    int pin_set = 0;
    /* option top_smem: the top levels of the packed tree as a breadth-first table for the TOPS instantiations of wf_traverse */
    float4* top_nodes = nullptr;
    int* d_n_top = nullptr;
    int n_top = 0;
    uint64_t top_generation = ~0ull;
#ifdef RT_TIMELINE
    unsigned long long* tl_buf = nullptr; /* 2 x 64 words: min start / max end of the frame's launches (globaltimer, ns) */
    int tl_n = 0;
    const char* tl_name[64];
    int tl_band[64];
#endif
    /* ... and goes to the device on an upload stream of its own, into one of two small staging buffers: the copy runs while the frame before
     * renders instead of queueing behind that frame's copy-back (H2D behind D2H cost 0.09 ms per step, profiles/r02_notes.md) */
    cudaStream_t up_stream = nullptr;
    unsigned char* stage_up[2] = {nullptr, nullptr};
    size_t stage_up_bytes[2] = {0, 0};
    cudaEvent_t stage_free[2] = {nullptr, nullptr};
    bool stage_free_recorded[2] = {false, false};
    const int32_t* cur_recs = nullptr; /* the triangle records of the last upload on the device (rt_scene_set_mesh_normals reads words 6-8) */

    /* pending copy-back of a RT_RENDER_NO_SYNC call */
    bool pending = false;
    int pending_launches = 0;

    /* kernel variant: 0 = render_mega, plain exact arithmetic; 1 = render_mega with the certified fast paths;
     * 2 = wavefront pipeline (rt_wavefront.cuh: streaming owner kernels + persistent queue traversal with work
     * stealing). Same results. RT_VARIANT overrides. */
    int max_leaf = 0;        /* largest leaf of the uploaded BVH */
    int trav_blocks_per_sm = 0, sm_count = 0;
    size_t trav_smem = 0;
    rtk::QEntry* wf_queue = nullptr; /* 3 queues of wf_capacity entries: closest even / closest odd / shadow */
    size_t wf_capacity = 0;
    rtk::WfCounters* wf_counters = nullptr;
    rtk::WfCounters* h_wf_counters = nullptr; /* pinned */
    unsigned long long* sticky = nullptr;     /* device, 2: task buffer / node pool overflow since the last rt_scene_sync (never cleared by rt_render) */
    unsigned long long* h_sticky = nullptr;   /* pinned */
    bool last_was_wavefront = false;
    cudaStream_t strip_stream[RT_MAX_STRIPS] = {};
    cudaEvent_t strip_done[RT_MAX_STRIPS] = {};
    cudaEvent_t fork_ev = nullptr;
    /* a band's tree search (bounce rays) and its (ray, leaf) tasks of the same round are independent: side stream + events */
    cudaStream_t side_stream[RT_MAX_STRIPS] = {};
    cudaEvent_t side_fork[RT_MAX_STRIPS] = {}, side_join[RT_MAX_STRIPS] = {};
    rtk::XorwowState* rng_states = nullptr; /* stochastic mode: curand_init(seed, pixel, 0) of every pixel of a rng_W x rng_H frame */
    size_t rng_capacity = 0;
    float2* jitter_tab = nullptr;  /* first-sample jitter of every pixel (rt_wavefront.cuh: jitter_table), for one-sample one-segment frames */
    size_t jitter_capacity = 0;
    float jitter_sigma = 0.f;
    int jitter_libm = -1;
    bool jitter_valid = false;
    int rng_W = 0, rng_H = 0;
    unsigned long long rng_seed = 0;
    unsigned char* st_buf = nullptr; /* stochastic wavefront: per-pixel stream state, colour sum, diffuse records */
    size_t st_buf_bytes = 0;
    int* wf_spill = nullptr;  /* node-pool overflow area of wf_traverse */
    bool trav_wide = false;
    unsigned char* stage = nullptr; /* rt_scene_set_mesh: device staging of the interchange arrays */
    size_t stage_bytes = 0;
    unsigned char* pin = nullptr;   /* rt_scene_set_mesh: pinned host staging (meshes up to 64 MB) */
    size_t pin_bytes = 0;
    uint64_t mesh_generation = 0; /* bumped whenever the TREE part of the blob changes: invalidates the anchored-ray bins */
    std::vector<float> last_bvh;  /* host copy of the arr_bvh the packed tree was built from: a mesh uploaded again with the same tree
                                   * (the per-frame upload of a caller that owns the geometry) skips the node relayout, keeps the bins */
    int32_t last_nv = 0, last_nt = 0;
    size_t stage_r_off = 0;          /* where the uploaded triangle records sit in `stage` (rt_scene_set_mesh_normals reads words 6-8) */
    int32_t stage_nt = 0;
    float4* tri_normals = nullptr;   /* viewer feature: 3 x float4 per triangle (vertex normals), beside the blob */
    float* d_normals = nullptr;
    size_t tri_normals_cap = 0, d_normals_cap = 0;
    bool has_normals = false;
    float4* accum = nullptr;         /* viewer feature: progressive accumulation buffer + the frame's linear colour */
    float4* linear = nullptr;
    size_t accum_px = 0;
    /* anchored-ray bins (rt_bins.cuh): [0] camera, [1] light */
    struct AnchorBins {
        float A[3] = {0.f, 0.f, 0.f};
        int R = 0;
        uint64_t mesh_generation = ~0ull;
        bool built = false, usable = false;
        float eps = 0.f, max_D2 = 0.f;
        int* cell_start = nullptr; /* 3 R R + 1 */
        int* cursor = nullptr;
        size_t cells_cap = 0;
        int* items = nullptr;
        size_t items_cap = 0;
        int* status = nullptr;   /* device: [0] a leaf box contains the anchor, [1] the lists outgrew `items` */
        int* h_status = nullptr; /* pinned copy, refreshed by rt_scene_sync */
        bool grow = false;       /* the next build enlarges `items` */
        int n_leaves = -1;       /* leaves of the mesh the item buffer was sized for */
        rtk::BinsView view;      /* anchor, eps, windows (the pointers are filled in by bins_view) */
    } bins[2];
    void* scan_tmp = nullptr;
    size_t scan_tmp_bytes = 0;
    int2* wf_tasks = nullptr;
    size_t wf_tasks_cap = 0;
    int task_factor = 8;     /* task buffer entries per pixel; doubled after an overflow */
    bool last_was_anchored = false;
    int leaves_blocks_per_sm = 0;
    int bins_builds = 0;
    size_t wf_spill_ints = 0;
    int* dbg_warps = nullptr; /* RT_DEBUG_WARPS=<file>: per-warp timeline of wf_traverse (count_work renders) */
    size_t dbg_warps_ints = 0;
};

namespace {

/* Convenience for the tools/ scripts: RT_<KEY> in the environment presets an option ONCE, when a scene is created
 * (rt_scene_create). Nothing reads the environment afterwards; tests and callers use rt_scene_set_option. */
void options_from_env(RtOptions& o) {
    for (const RtOptionKey& k : kOptionKeys) {
        std::string name = "RT_";
        for (const char* c = k.name; *c; c++) name += (char)toupper((unsigned char)*c);
        if (const char* v = getenv(name.c_str())) o.*(k.field) = std::max(k.lo, std::min(atoi(v), k.hi));
    }
    if (const char* v = getenv("RT_ANCHOR")) o.anchored = std::max(-1, std::min(atoi(v), 1)); /* round-1 spelling */
    if (const char* v = getenv("RT_DEBUG_WARPS")) snprintf(o.debug_warps, sizeof o.debug_warps, "%s", v);
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

/* The 8-bit transfer function of the reference, evaluated on the host with its own libm expressions. */
int transfer(float c, int mode) {
    if (mode == 1) return (int)std::min((double)powf(c, (float)(1. / 2.2)), 255.); /* optimized.cu:765 */
    return (int)std::min(std::pow((double)c, 1. / 2.2), 255.);                     /* cpu_launcher.cpp:714 */
}

/* T[k] = smallest non-negative float c with transfer(c) >= k, by bisection over the (ordered) bit patterns. */
void build_gamma_table(float* T, int mode) {
    T[0] = 0.f;
    for (int k = 1; k < 256; k++) {
        uint32_t lo = 0, hi = 0x7f800000u; /* transfer(+0) = 0 < k <= transfer(+inf) = 255 */
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float c;
            memcpy(&c, &mid, 4);
            if (transfer(c, mode) >= k) hi = mid;
            else lo = mid;
        }
        memcpy(&T[k], &hi, 4);
    }
}

bool is_device_pointer(const void* p, int device) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) && (at.type == cudaMemoryTypeManaged || at.device == device);
}

int ensure_scratch(rt_scene* s, int set, int slot, size_t bytes) {
    if (s->scratch_bytes[set][slot] >= bytes) return RT_OK;
    if (s->copy_stream) CUDA_TRY(cudaStreamSynchronize(s->copy_stream)); /* a copy may still read the buffer that goes away */
    if (s->scratch[set][slot]) cudaFree(s->scratch[set][slot]);
    s->scratch[set][slot] = nullptr;
    s->scratch_bytes[set][slot] = 0;
    CUDA_TRY(cudaMalloc(&s->scratch[set][slot], bytes));
    s->scratch_bytes[set][slot] = bytes;
    return RT_OK;
}

int upload_header(rt_scene* s) {
    if (!s->blob) {
        /* spheres-only scene: the blob is just the header */
        CUDA_TRY(cudaMalloc(&s->blob, RT_HEADER_BYTES));
        s->blob_bytes = RT_HEADER_BYTES;
        s->header.off_nodes = s->header.off_tris = s->header.off_wide = s->header.off_leaves = RT_HEADER_BYTES;
        s->header.total_bytes = RT_HEADER_BYTES;
    }
    CUDA_TRY(cudaMemcpyAsync(s->blob, &s->header, sizeof(SceneHeader), cudaMemcpyHostToDevice, s->stream));
    s->header_dirty = false;
    return RT_OK;
}

void reset_mesh_fields(SceneHeader& h) {
    h.has_mesh = 0;
    h.n_inner = 0;
    h.n_tris = 0;
    h.max_depth = 0;
    h.root_ref = 0;
    h.wroot_ref = 0;
    h.n_wide = 0;
    h.wide_depth = 0;
    h.n_leaves = 0;
    h.max_leaf = 0;
    for (int k = 0; k < 3; k++) {
        h.root_mn[k] = 0.f;
        h.root_mx[k] = 0.f;
        h.box_abs[k] = 0.f;
    }
}

/* (Re)build the candidate lists of one anchor (rt_bins.cuh) when the anchor, the mesh or the resolution changed: three
 * small launches (count, scan, fill). The first build for a mesh reads the list total back to size the item buffer
 * (one synchronisation, next to the ones rt_scene_set_mesh already has); later builds — a moving light or camera — are
 * enqueued without any read-back: the buffer keeps a quarter of headroom, lists that still outgrow it are cut short and
 * their rays take the exact search (device-side check), and rt_scene_sync enlarges the buffer for the next build. A leaf
 * box around the anchor is flagged in the status word the kernels consult. */
int ensure_bins(rt_scene* s, int which, const float A[3]) {
    rt_scene::AnchorBins& b = s->bins[which];
    const SceneHeader& h = s->header;
    const int R = s->opt.bins_r > 0 ? std::min(std::max(s->opt.bins_r, 8), 4096) : (h.n_leaves > 200000 ? 2048 : 1024);
    /* the status word of the previous build is copied to pinned memory right behind the build (no synchronisation): by the
     * time the anchor moves again it tells whether the lists outgrew the item buffer */
    if (b.h_status && b.h_status[1]) {
        b.grow = true;
        b.h_status[1] = 0;
    }
    if (b.built && b.mesh_generation == s->mesh_generation && b.R == R && b.A[0] == A[0] && b.A[1] == A[1] && b.A[2] == A[2] && !b.grow) return RT_OK;
    /* the item buffer is sized by a read-back only when nothing is known about the lists: first build, or a mesh with a
     * different number of leaves; a mesh uploaded again (the per-frame upload of a caller that owns the geometry) keeps it */
    const bool first = b.R != R || b.items_cap == 0 || b.n_leaves != h.n_leaves;
    const bool dbg_bins = s->opt.debug_bins != 0;
    b.built = false;
    const float S = std::max(h.box_abs[0], std::max(h.box_abs[1], h.box_abs[2]));
    const float scale = S + std::max(std::fabs(A[0]), std::max(std::fabs(A[1]), std::fabs(A[2])));
    b.eps = scale * (1.f / 4096.f);
    b.max_D2 = (scale * 128.f) * (scale * 128.f);
    /* the windows: what the root box projects to on every face (bins_face_rect is inclusion-monotone, every leaf box lies in
     * the root box), one cell of margin, the whole face when the root box contains the anchor */
    rtk::BinsView& bv = b.view;
    bv.ax = A[0];
    bv.ay = A[1];
    bv.az = A[2];
    bv.eps = b.eps;
    bv.max_D2 = b.max_D2;
    bv.R = R;
    size_t n_cells = 0;
    {
        const float v0[3] = {h.root_mn[0] - b.eps - A[0], h.root_mn[1] - b.eps - A[1], h.root_mn[2] - b.eps - A[2]};
        const float v1[3] = {h.root_mx[0] + b.eps - A[0], h.root_mx[1] + b.eps - A[1], h.root_mx[2] + b.eps - A[2]};
        const bool inside = v0[0] <= 0.f && v1[0] >= 0.f && v0[1] <= 0.f && v1[1] >= 0.f && v0[2] <= 0.f && v1[2] >= 0.f;
        for (int k = 0; k < 3; k++) {
            int lo_a = R, hi_a = -1, lo_b = R, hi_b = -1;
            for (int part = 0; part < 2; part++) {
                int ca0, ca1, cb0, cb1;
                if (!rtk::bins_face_rect(v0, v1, k, part, R, ca0, ca1, cb0, cb1)) continue;
                lo_a = std::min(lo_a, ca0);
                hi_a = std::max(hi_a, ca1);
                lo_b = std::min(lo_b, cb0);
                hi_b = std::max(hi_b, cb1);
            }
            if (inside) {
                lo_a = lo_b = 0;
                hi_a = hi_b = R - 1;
            }
            rtk::BinsWindow& w = bv.win[k];
            if (hi_a < lo_a || hi_b < lo_b) {
                w.ca0 = w.cb0 = w.wa = w.wb = 0;
            } else {
                lo_a = std::max(lo_a - 1, 0);
                lo_b = std::max(lo_b - 1, 0);
                hi_a = std::min(hi_a + 1, R - 1);
                hi_b = std::min(hi_b + 1, R - 1);
                w.ca0 = lo_a;
                w.cb0 = lo_b;
                w.wa = hi_a - lo_a + 1;
                w.wb = hi_b - lo_b + 1;
            }
            w.base = (int)n_cells;
            n_cells += (size_t)w.wa * w.wb;
        }
    }
    if (b.cells_cap < n_cells + 1) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (b.cell_start) cudaFree(b.cell_start);
        if (b.cursor) cudaFree(b.cursor);
        b.cell_start = b.cursor = nullptr;
        b.cells_cap = 0;
        const size_t cap = 2 * (n_cells + 1); /* headroom: the windows change with the anchor */
        CUDA_TRY(cudaMalloc(&b.cell_start, cap * sizeof(int)));
        CUDA_TRY(cudaMalloc(&b.cursor, cap * sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(b.cursor, 0, cap * sizeof(int), s->stream)); /* the per-cell counts: every build leaves them at zero again */
        b.cells_cap = cap;
    }
    if (!b.status) {
        CUDA_TRY(cudaMalloc(&b.status, 4 * sizeof(int)));
        CUDA_TRY(cudaMallocHost(&b.h_status, 4 * sizeof(int)));
        memset(b.h_status, 0, 4 * sizeof(int));
    }
    if (dbg_bins) fprintf(stderr, "[bins %d] build #%d first %d grow %d cells %zu (cap %zu) items cap %zu\n", which, s->bins_builds, (int)first, (int)b.grow, n_cells, b.cells_cap, b.items_cap);
    if (b.grow) { /* the previous build outgrew the buffer (noticed by rt_scene_sync) */
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (b.items) cudaFree(b.items);
        b.items = nullptr;
        const size_t cap = b.items_cap * 2 + 1024;
        b.items_cap = 0;
        CUDA_TRY(cudaMalloc(&b.items, cap * sizeof(int)));
        b.items_cap = cap;
        b.grow = false;
    }
    const float4* leaves = reinterpret_cast<const float4*>(s->blob + h.off_leaves);
    const int threads = BINS_GROUP, blocks = h.n_leaves; /* one block per leaf */
    CUDA_TRY(cudaMemsetAsync(b.status, 0, 4 * sizeof(int), s->stream));
    rtk::bins_count<<<blocks, threads, 0, s->stream>>>(leaves, h.n_leaves, bv, b.cursor, b.status);
    CUDA_TRY(cudaGetLastError());
    size_t tmp_bytes = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, b.cursor, b.cell_start, (int)(n_cells + 1), s->stream));
    if (s->scan_tmp_bytes < tmp_bytes) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->scan_tmp) cudaFree(s->scan_tmp);
        s->scan_tmp = nullptr;
        s->scan_tmp_bytes = 0;
        CUDA_TRY(cudaMalloc(&s->scan_tmp, 2 * tmp_bytes + 4096)); /* the windows, and with them the scan's work space, change with the anchor */
        s->scan_tmp_bytes = 2 * tmp_bytes + 4096;
    }
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->scan_tmp, tmp_bytes, b.cursor, b.cell_start, (int)(n_cells + 1), s->stream));
    if (first) {
        int total = 0;
        CUDA_TRY(cudaMemcpyAsync(&total, b.cell_start + n_cells, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (total < 0) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_render: candidate lists of %d cells overflow 2^31 entries", (int)n_cells);
        const size_t cap = 3 * (size_t)total + 1024; /* room for the lists of a moving anchor */
        if (b.items_cap < cap) {
            if (b.items) cudaFree(b.items);
            b.items = nullptr;
            b.items_cap = 0;
            CUDA_TRY(cudaMalloc(&b.items, cap * sizeof(int)));
            b.items_cap = cap;
        }
    }
    rtk::bins_fill<<<blocks, threads, 0, s->stream>>>(leaves, h.n_leaves, bv, b.cell_start, b.cursor, b.items, (int)std::min<size_t>(b.items_cap, 0x7fffffff), b.status);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(b.h_status, b.status, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream)); /* read at the next build, never waited for */
    b.R = R;
    b.A[0] = A[0];
    b.A[1] = A[1];
    b.A[2] = A[2];
    b.mesh_generation = s->mesh_generation;
    b.n_leaves = h.n_leaves;
    b.built = true;
    b.usable = true;
    s->bins_builds++;
    return RT_OK;
}

rtk::BinsView bins_view(const rt_scene::AnchorBins& b) {
    rtk::BinsView v = b.view;
    v.cell_start = b.cell_start;
    v.items = b.items;
    v.items_cap = (int)std::min<size_t>(b.items_cap, 0x7fffffff);
    v.status = b.status;
    return v;
}

} // namespace

namespace rtb {
/* rt_comm.cu (multi-GPU entry points) reaches the scene through these three */
cudaStream_t scene_stream(rt_scene* s) { return s->stream; }
int scene_device(rt_scene* s) { return s->device; }
/* make the scene's blob at least `bytes` large and hand out its address: a broadcast is received in place and then adopted
 * by rt_scene_blob_import(s, that pointer, bytes) without a copy */
int scene_blob_reserve(rt_scene* s, size_t bytes, void** device_ptr) {
    DeviceGuard g(s->device);
    if (bytes < RT_HEADER_BYTES) return fail(RT_ERR_INVALID, "scene_blob_reserve: %zu bytes is smaller than a scene header", bytes);
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->blob_bytes < bytes) {
        if (s->blob) cudaFree(s->blob);
        s->blob = nullptr;
        s->blob_bytes = 0;
        CUDA_TRY(cudaMalloc(&s->blob, bytes));
        s->blob_bytes = bytes;
    }
    s->plan_valid = false;
    *device_ptr = s->blob;
    return RT_OK;
}
int bvh_build_device(int device, const float* vertices, int nv, const int32_t* recs10, int nt, std::vector<int32_t>* perm, std::vector<float>* arr, int32_t info[4],
                     double* build_ms, void** keep) {
    rtbuild::DeviceMesh* dm = keep ? new rtbuild::DeviceMesh() : nullptr;
    const int rc = rtbuild::build(device, vertices, nv, recs10, nt, *perm, *arr, info, build_ms, dm);
    if (keep) {
        if (rc != 0) {
            delete dm;
            dm = nullptr;
        }
        *keep = dm;
    }
    return rc;
}
int bvh_device_download(void* keep, int32_t* recs10_out, std::vector<float>* arr_out) {
    rtbuild::DeviceMesh* dm = static_cast<rtbuild::DeviceMesh*>(keep);
    DeviceGuard g(dm->device);
    arr_out->resize((size_t)dm->nn * 10);
    cudaError_t e = cudaMemcpy(recs10_out, dm->recs, (size_t)dm->nt * 10 * sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(arr_out->data(), dm->arr, (size_t)dm->nn * 10 * sizeof(float), cudaMemcpyDeviceToHost);
    return (int)e;
}
void bvh_device_free(void* keep) {
    rtbuild::DeviceMesh* dm = static_cast<rtbuild::DeviceMesh*>(keep);
    if (!dm) return;
    DeviceGuard g(dm->device);
    if (dm->V) cudaFree(dm->V);
    if (dm->recs) cudaFree(dm->recs);
    if (dm->arr) cudaFree(dm->arr);
    delete dm;
}
void bvh_device_arrays(void* keep, int* device, const float** vertices, const int32_t** recs10, const float** arr, int32_t* nn) {
    rtbuild::DeviceMesh* dm = static_cast<rtbuild::DeviceMesh*>(keep);
    *device = dm->device;
    *vertices = dm->V;
    *recs10 = dm->recs;
    *arr = dm->arr;
    *nn = dm->nn;
}
} // namespace rtb

extern "C" {

int rt_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (count) *count = (e == cudaSuccess) ? n : 0;
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return RT_OK;
}

int rt_scene_create(rt_scene** out, int device) {
    if (!out) return rtb::fail(RT_ERR_INVALID, "rt_scene_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return rtb::fail(RT_ERR_CUDA, "rt_scene_create: no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
    if (device < 0 || device >= n) return rtb::fail(RT_ERR_INVALID, "rt_scene_create: device %d out of range (0..%d)", device, n - 1);
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_scene_create: cudaSetDevice(%d) failed", device);
    rt_scene* s = new (std::nothrow) rt_scene();
    if (!s) return rtb::fail(RT_ERR_NOMEM, "rt_scene_create: out of memory");
    s->device = device;
    options_from_env(s->opt);
    if (s->opt.task_factor > 0) s->task_factor = s->opt.task_factor;
    memset(&s->header, 0, sizeof s->header);
    s->header.magic = RT_BLOB_MAGIC;
    s->header.layout_version = RT_LAYOUT_VERSION;
    s->header.mesh_id = -1;
    s->header.L[0] = -10.f; /* Scene::L / intensity defaults, optimized.cu:681-683 */
    s->header.L[1] = 20.f;
    s->header.L[2] = 40.f;
    s->header.intensity = 3e10f;
    reset_mesh_fields(s->header);
    std::vector<float> tab(512);
    build_gamma_table(tab.data(), 0);
    build_gamma_table(tab.data() + 256, 1);
    cudaError_t err = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    s->own_stream = (err == cudaSuccess);
    if (err == cudaSuccess) err = cudaEventCreate(&s->ev0);
    if (err == cudaSuccess) err = cudaEventCreate(&s->ev1);
    if (err == cudaSuccess) err = cudaMalloc(&s->gamma_tab, 512 * sizeof(float));
    if (err == cudaSuccess) err = cudaMalloc(&s->counters, RT_NCOUNTERS * sizeof(unsigned long long));
    if (err == cudaSuccess) err = cudaMallocHost(&s->h_counters, RT_NCOUNTERS * sizeof(unsigned long long));
    if (err == cudaSuccess) err = cudaMalloc(&s->sticky, 2 * sizeof(unsigned long long));
    if (err == cudaSuccess) err = cudaMemset(s->sticky, 0, 2 * sizeof(unsigned long long));
    if (err == cudaSuccess) err = cudaMallocHost(&s->h_sticky, 2 * sizeof(unsigned long long));
    if (err == cudaSuccess) s->h_sticky[0] = s->h_sticky[1] = 0;
    if (err == cudaSuccess) err = cudaMemcpy(s->gamma_tab, tab.data(), 512 * sizeof(float), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        rtb::fail(RT_ERR_CUDA, "rt_scene_create: %s", cudaGetErrorString(err));
        rt_scene_destroy(s);
        return RT_ERR_CUDA;
    }
    *out = s;
    return RT_OK;
}

void rt_scene_destroy(rt_scene* s) {
    if (!s) return;
    DeviceGuard g(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) {
        cudaStreamSynchronize(s->copy_stream);
        cudaStreamDestroy(s->copy_stream);
    }
    for (int k = 0; k < 2; k++) {
        if (s->copy_done[k]) cudaEventDestroy(s->copy_done[k]);
        if (s->pin_ev[k]) cudaEventDestroy(s->pin_ev[k]);
    }
    if (s->frame_done) cudaEventDestroy(s->frame_done);
    if (s->pin_alt) cudaFreeHost(s->pin_alt);
    if (s->top_nodes) cudaFree(s->top_nodes);
    if (s->d_n_top) cudaFree(s->d_n_top);
    if (s->up_stream) {
        cudaStreamSynchronize(s->up_stream);
        cudaStreamDestroy(s->up_stream);
    }
    for (int k = 0; k < 2; k++) {
        if (s->stage_up[k]) cudaFree(s->stage_up[k]);
        if (s->stage_free[k]) cudaEventDestroy(s->stage_free[k]);
    }
    for (int k = 0; k < 2; k++)
        if (s->graph_exec[k]) cudaGraphExecDestroy(s->graph_exec[k]);
    if (s->tri_normals) cudaFree(s->tri_normals);
    if (s->d_normals) cudaFree(s->d_normals);
    if (s->accum) cudaFree(s->accum);
    if (s->linear) cudaFree(s->linear);
    if (s->blob) cudaFree(s->blob);
    if (s->gamma_tab) cudaFree(s->gamma_tab);
    if (s->counters) cudaFree(s->counters);
    if (s->h_counters) cudaFreeHost(s->h_counters);
    if (s->sticky) cudaFree(s->sticky);
    if (s->h_sticky) cudaFreeHost(s->h_sticky);
    if (s->wf_queue) cudaFree(s->wf_queue);
    if (s->wf_counters) cudaFree(s->wf_counters);
    if (s->dbg_warps) cudaFree(s->dbg_warps);
    if (s->wf_spill) cudaFree(s->wf_spill);
    if (s->stage) cudaFree(s->stage);
    if (s->pin) cudaFreeHost(s->pin);
    for (int k = 0; k < 2; k++) {
        if (s->bins[k].cell_start) cudaFree(s->bins[k].cell_start);
        if (s->bins[k].cursor) cudaFree(s->bins[k].cursor);
        if (s->bins[k].items) cudaFree(s->bins[k].items);
    }
    for (int k = 0; k < 2; k++) {
        if (s->bins[k].status) cudaFree(s->bins[k].status);
        if (s->bins[k].h_status) cudaFreeHost(s->bins[k].h_status);
    }
    if (s->scan_tmp) cudaFree(s->scan_tmp);
    if (s->wf_tasks) cudaFree(s->wf_tasks);
    if (s->rng_states) cudaFree(s->rng_states);
    if (s->jitter_tab) cudaFree(s->jitter_tab);
    if (s->st_buf) cudaFree(s->st_buf);
    if (s->h_wf_counters) cudaFreeHost(s->h_wf_counters);
    for (int k = 0; k < 5; k++)
        for (int set = 0; set < 2; set++)
            if (s->scratch[set][k]) cudaFree(s->scratch[set][k]);
    for (int k = 0; k < RT_MAX_STRIPS; k++) {
        if (s->strip_stream[k]) cudaStreamDestroy(s->strip_stream[k]);
        if (s->strip_done[k]) cudaEventDestroy(s->strip_done[k]);
        if (s->side_stream[k]) cudaStreamDestroy(s->side_stream[k]);
        if (s->side_fork[k]) cudaEventDestroy(s->side_fork[k]);
        if (s->side_join[k]) cudaEventDestroy(s->side_join[k]);
    }
    if (s->fork_ev) cudaEventDestroy(s->fork_ev);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int rt_scene_set_stream(rt_scene* s, void* cuda_stream) {
    if (!s) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_stream: NULL scene");
    DeviceGuard g(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)cuda_stream;
    s->own_stream = false;
    s->last_key[0].clear();
    s->last_key[1].clear();
    return RT_OK;
}

int rt_scene_get_stream(rt_scene* s, void** cuda_stream) {
    if (!s || !cuda_stream) return rtb::fail(RT_ERR_INVALID, "rt_scene_get_stream: bad argument");
    *cuda_stream = (void*)s->stream;
    return RT_OK;
}

int rt_scene_set_option(rt_scene* s, const char* key, int64_t value) {
    if (!s || !key) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_option: bad argument");
    for (const RtOptionKey& k : kOptionKeys)
        if (strcmp(k.name, key) == 0) {
            if (value < k.lo || value > k.hi) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_option: %s = %lld outside [%d, %d]", key, (long long)value, k.lo, k.hi);
            s->opt.*(k.field) = (int)value;
            if (k.field == &RtOptions::task_factor && value > 0) s->task_factor = (int)value;
            s->plan_valid = false;
            return RT_OK;
        }
    return rtb::fail(RT_ERR_INVALID, "rt_scene_set_option: unknown option '%s'", key);
}

int rt_scene_get_option(rt_scene* s, const char* key, int64_t* value) {
    if (!s || !key || !value) return rtb::fail(RT_ERR_INVALID, "rt_scene_get_option: bad argument");
    for (const RtOptionKey& k : kOptionKeys)
        if (strcmp(k.name, key) == 0) {
            *value = (k.field == &RtOptions::task_factor) ? s->task_factor : s->opt.*(k.field);
            return RT_OK;
        }
    return rtb::fail(RT_ERR_INVALID, "rt_scene_get_option: unknown option '%s'", key);
}

int rt_scene_set_spheres(rt_scene* s, const rt_sphere* spheres, int32_t n) {
    if (!s || n < 0 || (n > 0 && !spheres)) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_spheres: bad argument");
    if (n > RT_MAX_SPHERES) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_scene_set_spheres: %d spheres, at most %d (the reference holds 10 objects, optimized.cu:663)", n, RT_MAX_SPHERES);
    std::vector<rt_sphere> v(spheres, spheres + n);
    std::stable_sort(v.begin(), v.end(), [](const rt_sphere& a, const rt_sphere& b) { return a.id < b.id; });
    for (int k = 0; k < n; k++) {
        if (v[k].id < 0 || (k > 0 && v[k].id == v[k - 1].id)) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_spheres: ids must be distinct and >= 0");
        DevSphere& d = s->header.spheres[k];
        d.cx = v[k].C[0];
        d.cy = v[k].C[1];
        d.cz = v[k].C[2];
        d.R = v[k].R;
        d.RR = v[k].R * v[k].R;
        d.ax = v[k].albedo[0];
        d.ay = v[k].albedo[1];
        d.az = v[k].albedo[2];
        d.mirror = v[k].mirror ? 1 : 0;
        d.n_in = v[k].n_in;
        d.n_out = v[k].n_out;
        d.id = v[k].id;
    }
    s->header.n_spheres = n;
    s->header_dirty = true;
    return RT_OK;
}

int rt_scene_set_light(rt_scene* s, const float L[3], float intensity) {
    if (!s || !L) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_light: bad argument");
    s->header.L[0] = L[0];
    s->header.L[1] = L[1];
    s->header.L[2] = L[2];
    s->header.intensity = intensity;
    s->header_dirty = true;
    return RT_OK;
}

int rt_scene_set_mesh(rt_scene* s, const float* vertices, int32_t nv, const int32_t* tri_records, int32_t nt, const float* arr_bvh,
                      int32_t n_nodes, const float albedo[3], int32_t mirror, float n_in, float n_out, int32_t id) {
    if (!s) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: NULL scene");
    DeviceGuard g(s->device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_scene_set_mesh: cudaSetDevice failed");
    SceneHeader& h = s->header;
    if (nt == 0) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->blob) cudaFree(s->blob);
        s->blob = nullptr;
        s->blob_bytes = 0;
        reset_mesh_fields(h);
        s->mesh_generation++;
        s->plan_valid = false;
        s->last_bvh.clear();
        s->last_nv = s->last_nt = 0;
        s->has_normals = false;
        s->stage_nt = 0;
        h.mesh_id = -1;
        s->header_dirty = true;
        return RT_OK;
    }
    if (nt < 0 || nv <= 0 || n_nodes <= 0 || !vertices || !tri_records || !arr_bvh || !albedo)
        return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: bad argument");
    if (id < 0) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: id must be >= 0");
    for (int64_t i = 0; i < nt; i++)
        for (int k = 0; k < 3; k++) {
            const int32_t v = tri_records[i * RT_TRI_RECORD_WORDS + k];
            if (v < 0 || v >= nv) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: triangle %lld references vertex %d (nv=%d)", (long long)i, v, nv);
        }

    /* ---- same tree as the last upload? Then only the triangles are new: the packed nodes, the wide index, the leaf table and the
     * per-triangle leaf keys (all functions of arr_bvh alone) stay on the device, the bins stay valid (they list leaf BOXES), and the
     * upload is vertices + records + the repack kernel. The interchange arrays still cross the bus: they are this call's input. */
    if (s->blob && h.has_mesh && s->last_nv == nv && s->last_nt == nt && (int32_t)(s->last_bvh.size() / RT_BVH_NODE_FLOATS) == n_nodes && s->pin &&
        memcmp(s->last_bvh.data(), arr_bvh, s->last_bvh.size() * sizeof(float)) == 0) {
        const size_t vbytes = (size_t)nv * 3 * sizeof(float), rbytes = (size_t)nt * RT_TRI_RECORD_WORDS * sizeof(int32_t);
        const size_t r_off = (vbytes + 255) & ~(size_t)255, l_off = (r_off + rbytes + 255) & ~(size_t)255;
        /* two pinned halves, each guarded by an event behind its last upload: the host never waits for the frames in flight, only
         * for the upload before last (long done) */
        const size_t up_bytes = r_off + rbytes;
        s->pin_set ^= 1;
        const int half = s->pin_set;
        if (half == 1 && s->pin_alt_bytes < up_bytes) {
            if (s->pin_ev_recorded[1]) CUDA_TRY(cudaEventSynchronize(s->pin_ev[1]));
            if (s->pin_alt) cudaFreeHost(s->pin_alt);
            s->pin_alt = nullptr;
            s->pin_alt_bytes = 0;
            CUDA_TRY(cudaMallocHost(&s->pin_alt, up_bytes + up_bytes / 4));
            s->pin_alt_bytes = up_bytes + up_bytes / 4;
        }
        if (!s->up_stream) CUDA_TRY(cudaStreamCreateWithFlags(&s->up_stream, cudaStreamNonBlocking));
        if (!s->pin_ev[half]) CUDA_TRY(cudaEventCreateWithFlags(&s->pin_ev[half], cudaEventDisableTiming));
        if (!s->stage_free[half]) CUDA_TRY(cudaEventCreateWithFlags(&s->stage_free[half], cudaEventDisableTiming));
        if (s->stage_up_bytes[half] < up_bytes) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (s->stage_up[half]) cudaFree(s->stage_up[half]);
            s->stage_up[half] = nullptr;
            s->stage_up_bytes[half] = 0;
            CUDA_TRY(cudaMalloc(&s->stage_up[half], up_bytes + up_bytes / 4));
            s->stage_up_bytes[half] = up_bytes + up_bytes / 4;
        }
        if (s->pin_ev_recorded[half]) CUDA_TRY(cudaEventSynchronize(s->pin_ev[half])); /* the upload before last has left this pinned half */
        unsigned char* const pin = half ? s->pin_alt : s->pin;
        unsigned char* const stage = s->stage_up[half];
        memcpy(pin, vertices, vbytes);
        memcpy(pin + r_off, tri_records, rbytes);
        if (s->stage_free_recorded[half]) CUDA_TRY(cudaStreamWaitEvent(s->up_stream, s->stage_free[half], 0)); /* its last repack has read it */
        CUDA_TRY(cudaMemcpyAsync(stage, pin, up_bytes, cudaMemcpyHostToDevice, s->up_stream));
        CUDA_TRY(cudaEventRecord(s->pin_ev[half], s->up_stream));
        s->pin_ev_recorded[half] = true;
        CUDA_TRY(cudaStreamWaitEvent(s->stream, s->pin_ev[half], 0));
        const int threads = 256, blocks = (nt + threads - 1) / threads;
        rtk::repack_triangles<<<blocks, threads, 0, s->stream>>>(reinterpret_cast<float*>(stage), reinterpret_cast<int32_t*>(stage + r_off), nt,
                                                                reinterpret_cast<int32_t*>(s->stage + l_off), reinterpret_cast<float4*>(s->blob + h.off_tris));
        CUDA_TRY(cudaEventRecord(s->stage_free[half], s->stream));
        s->stage_free_recorded[half] = true;
        s->cur_recs = reinterpret_cast<const int32_t*>(stage + r_off);
        CUDA_TRY(cudaGetLastError());
        s->has_normals = false; /* the records were uploaded again: rt_scene_set_mesh_normals must follow */
        const bool same_material = h.mesh_id == id && h.mesh_mirror == (mirror ? 1 : 0) && h.mesh_n_in == n_in && h.mesh_n_out == n_out &&
                                   memcmp(h.mesh_albedo, albedo, sizeof h.mesh_albedo) == 0;
        if (!same_material) {
            h.mesh_id = id;
            h.mesh_mirror = mirror ? 1 : 0;
            h.mesh_n_in = n_in;
            h.mesh_n_out = n_out;
            memcpy(h.mesh_albedo, albedo, sizeof h.mesh_albedo);
            s->header_dirty = true;
            s->plan_valid = false;
        }
        return RT_OK;
    }

    /* ---- node relayout on the host: 10-float pre-order nodes -> 64-B two-child records ----------------- */
    std::vector<int32_t> inner_index((size_t)n_nodes, -1);
    int32_t n_inner = 0, max_leaf = 0;
    for (int32_t k = 0; k < n_nodes; k++) {
        const float* a = arr_bvh + (size_t)k * RT_BVH_NODE_FLOATS;
        const int32_t l = (int32_t)a[0], r = (int32_t)a[1];
        if ((l == -1) != (r == -1)) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: node %d has exactly one child", k);
        if (l != -1) {
            if (l <= k || r <= k || l >= n_nodes || r >= n_nodes) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: node %d child index out of order/range", k);
            inner_index[k] = n_inner++;
        } else {
            const int32_t ts = (int32_t)a[8], te = (int32_t)a[9];
            if (ts < 0 || te < ts || te > nt) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: leaf %d triangle range [%d,%d) invalid", k, ts, te);
            max_leaf = std::max(max_leaf, te - ts);
        }
    }
    /* children have larger indices than their parent (pre-order), so one forward pass gives depths */
    std::vector<int32_t> depth((size_t)n_nodes, 0);
    depth[0] = 1;
    int32_t max_depth = 1;
    for (int32_t k = 0; k < n_nodes; k++) {
        const float* a = arr_bvh + (size_t)k * RT_BVH_NODE_FLOATS;
        const int32_t l = (int32_t)a[0], r = (int32_t)a[1];
        if (depth[k] == 0) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: node %d is unreachable", k);
        if (l != -1) {
            if (depth[l] != 0 || depth[r] != 0 || l == r) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh: node %d shares a child", k);
            depth[l] = depth[r] = depth[k] + 1;
            max_depth = std::max(max_depth, depth[k] + 1);
        }
    }
    if (max_depth > RT_STACK_CAP - 2)
        return rtb::fail(RT_ERR_UNSUPPORTED, "rt_scene_set_mesh: BVH depth %d exceeds the traversal stack (%d)", max_depth, RT_STACK_CAP - 2);

    float box_abs[3] = {0.f, 0.f, 0.f};
    for (int32_t k = 0; k < n_nodes; k++) {
        const float* a = arr_bvh + (size_t)k * RT_BVH_NODE_FLOATS;
        for (int c = 0; c < 3; c++) box_abs[c] = std::max(box_abs[c], std::max(std::fabs(a[2 + c]), std::fabs(a[5 + c])));
    }
    /* packed tree: 64-B records (two child boxes + two int references), a leaf table with at most RT_LEAF_MAX
     * triangles per entry; larger reference leaves hang under virtual nodes that repeat the leaf's own box */
    std::vector<float> packed((size_t)std::max(n_inner, 1) * 16);
    std::vector<int32_t> leaf_start_of_tri((size_t)nt, 0); /* first triangle of the reference's leaf: the tie-break key */
    int32_t n_leaves = 0, extra_levels = 0;
    std::vector<float> leaf_table; /* 8 words per packed leaf: box, leaf code, first triangle of the reference leaf */
    auto new_leaf = [&](int32_t start, int32_t count, int32_t orig_start, const float* bb) -> int32_t {
        n_leaves++;
        for (int32_t i = start; i < start + count; i++) leaf_start_of_tri[i] = orig_start;
        if (count > 0) {
            const size_t at = leaf_table.size();
            leaf_table.resize(at + 8);
            memcpy(&leaf_table[at], bb, 6 * sizeof(float));
            const int32_t code = (start << 2) | (count - 1);
            memcpy(&leaf_table[at + 6], &code, sizeof code);
            memcpy(&leaf_table[at + 7], &orig_start, sizeof orig_start);
        }
        /* an empty leaf (only a hand-made arr_bvh can hold one) points past the last triangle: the kernels clamp
         * the range to n_tris, so it tests nothing */
        if (count <= 0) return -1 - ((nt << 2) | 0);
        return -1 - ((start << 2) | (count - 1));
    };
    /* reference of the subtree holding chunks [lo, hi) of the reference leaf [ts, te) with box `bb` (6 floats) */
    std::function<int32_t(int32_t, int32_t, int32_t, int32_t, const float*, int32_t)> chunk_tree =
        [&](int32_t lo, int32_t hi, int32_t ts, int32_t te, const float* bb, int32_t level) -> int32_t {
        if (hi - lo == 1) {
            const int32_t st = ts + lo * RT_LEAF_MAX;
            return new_leaf(st, std::min(RT_LEAF_MAX, te - st), ts, bb);
        }
        extra_levels = std::max(extra_levels, level);
        const int32_t idx = (int32_t)(packed.size() / 16);
        packed.resize(packed.size() + 16);
        const int32_t mid = lo + (hi - lo + 1) / 2;
        const int32_t l = chunk_tree(lo, mid, ts, te, bb, level + 1);
        const int32_t r = chunk_tree(mid, hi, ts, te, bb, level + 1);
        float* o = &packed[(size_t)idx * 16];
        for (int c = 0; c < 6; c++) {
            o[c] = bb[c];
            o[6 + c] = bb[c];
        }
        const int32_t refs[4] = {l, r, 1 /* virtual: not a node of the reference BVH */, 0};
        memcpy(o + 12, refs, sizeof refs);
        return idx;
    };
    auto child_ref = [&](int32_t node) -> int32_t {
        const float* a = arr_bvh + (size_t)node * RT_BVH_NODE_FLOATS;
        if ((int32_t)a[0] != -1) return inner_index[node];
        const int32_t ts = (int32_t)a[8], te = (int32_t)a[9];
        const int32_t chunks = std::max(1, (te - ts + RT_LEAF_MAX - 1) / RT_LEAF_MAX);
        return chunk_tree(0, chunks, ts, te, a + 2, 1);
    };
    for (int32_t k = 0; k < n_nodes; k++) {
        if (inner_index[k] < 0) continue;
        const float* a = arr_bvh + (size_t)k * RT_BVH_NODE_FLOATS;
        const float* L = arr_bvh + (size_t)(int32_t)a[0] * RT_BVH_NODE_FLOATS;
        const float* R = arr_bvh + (size_t)(int32_t)a[1] * RT_BVH_NODE_FLOATS;
        const int32_t rl = child_ref((int32_t)a[0]), rr = child_ref((int32_t)a[1]); /* may grow `packed` */
        float* o = &packed[(size_t)inner_index[k] * 16];
        for (int c = 0; c < 6; c++) {
            o[c] = L[2 + c];
            o[6 + c] = R[2 + c];
        }
        const int32_t refs[4] = {rl, rr, 0, 0};
        memcpy(o + 12, refs, sizeof refs);
    }
    const int32_t root_ref = child_ref(0);
    n_inner = (int32_t)(packed.size() / 16);
    if (inner_index[0] < 0 && n_inner == 1 && root_ref < 0) n_inner = 0; /* the root is a small leaf: no node records */
    max_depth += extra_levels;
    if (max_depth > RT_STACK_CAP - 2)
        return rtb::fail(RT_ERR_UNSUPPORTED, "rt_scene_set_mesh: packed BVH depth %d exceeds the traversal stack (%d)", max_depth, RT_STACK_CAP - 2);

    /* ---- wide index (rt_layout.h): the two-child records collapsed into nodes of up to RT_WIDE children. A node starts
     * with the two children of a record and keeps opening the inner child with the largest box until it is full.
     * Built iteratively, parents before children. */
    std::vector<float> wide;
    int32_t n_wide = 0, wide_depth = 0, wroot_ref = 0;
    if (n_inner > 0 && root_ref >= 0) {
        struct Pending { int32_t record, wide_index, level; };
        std::vector<Pending> todo;
        std::vector<int32_t> wide_cnt;
        todo.push_back({root_ref, 0, 1});
        n_wide = 1;
        wide.resize(32);
        wide_cnt.push_back(0);
        auto half_area = [](const float* b) {
            const float dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
            return dx * dy + dy * dz + dz * dx;
        };
        while (!todo.empty()) {
            const Pending cur = todo.back();
            todo.pop_back();
            wide_depth = std::max(wide_depth, cur.level);
            float box[RT_WIDE][6];
            int32_t ref[RT_WIDE];
            int cnt = 0;
            auto add_children = [&](int32_t record) {
                const float* o = &packed[(size_t)record * 16];
                int32_t refs[2];
                memcpy(refs, o + 12, sizeof refs);
                for (int c = 0; c < 2; c++) {
                    memcpy(box[cnt], o + 6 * c, 6 * sizeof(float));
                    ref[cnt++] = refs[c];
                }
            };
            add_children(cur.record);
            while (cnt < RT_WIDE) {
                int best = -1;
                float best_area = -1.f;
                for (int c = 0; c < cnt; c++)
                    if (ref[c] >= 0) {
                        const float ar = half_area(box[c]);
                        if (best < 0 || ar > best_area) {
                            best = c;
                            best_area = ar;
                        }
                    }
                if (best < 0) break;
                const int32_t open = ref[best];
                for (int c = best; c + 1 < cnt; c++) { /* close the gap, keep the order */
                    memcpy(box[c], box[c + 1], sizeof box[c]);
                    ref[c] = ref[c + 1];
                }
                cnt--;
                add_children(open);
            }
            /* inner children become wide nodes of their own */
            int32_t child_wide[RT_WIDE];
            for (int c = 0; c < cnt; c++) {
                child_wide[c] = -1;
                if (ref[c] >= 0) {
                    child_wide[c] = n_wide++;
                    wide.resize((size_t)n_wide * 32);
                    wide_cnt.push_back(0);
                    todo.push_back({ref[c], child_wide[c], cur.level + 1});
                }
            }
            wide_cnt[cur.wide_index] = cnt;
            float* o = &wide[(size_t)cur.wide_index * 32];
            for (int c = 0; c < RT_WIDE; c++) {
                const int src = c < cnt ? c : 0; /* unused slots repeat child 0; the child count travels with every reference */
                memcpy(o + 8 * c, box[src], 6 * sizeof(float));
                const int32_t r = c < cnt ? (ref[c] >= 0 ? child_wide[c] /* patched below */ : ref[c]) : 0;
                memcpy(o + 8 * c + 6, &r, sizeof r);
                const int32_t aux = c < cnt ? 1 : 0;
                memcpy(o + 8 * c + 7, &aux, sizeof aux);
            }
        }
        /* inner references carry the child's own child count: (wide index << 2) | (count - 1) */
        for (int32_t k = 0; k < n_wide; k++)
            for (int c = 0; c < wide_cnt[k]; c++) {
                int32_t r;
                memcpy(&r, &wide[(size_t)k * 32 + 8 * c + 6], sizeof r);
                if (r >= 0) {
                    r = (r << 2) | (wide_cnt[r] - 1);
                    memcpy(&wide[(size_t)k * 32 + 8 * c + 6], &r, sizeof r);
                }
            }
        wroot_ref = (0 << 2) | (wide_cnt[0] - 1);
        if (n_wide >= (1 << 24)) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_scene_set_mesh: %d wide nodes exceed the task word", n_wide);
    }

    /* ---- blob ------------------------------------------------------------------------------------------ */
    const size_t off_nodes = RT_HEADER_BYTES;
    const size_t off_tris = (off_nodes + (size_t)n_inner * RT_NODE_BYTES + 63) & ~(size_t)63;
    const size_t off_wide = (off_tris + (size_t)nt * RT_TRI_BYTES + 127) & ~(size_t)127;
    const size_t off_leaves = off_wide + (size_t)n_wide * RT_WNODE_BYTES;
    const int32_t n_leafrecs = (int32_t)(leaf_table.size() / 8);
    const size_t total = off_leaves + (size_t)n_leafrecs * RT_LEAFREC_BYTES;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    if (s->blob_bytes < total || s->blob_bytes > 2 * total + (1u << 20)) {
        if (s->blob) cudaFree(s->blob);
        s->blob = nullptr;
        s->blob_bytes = 0;
        CUDA_TRY(cudaMalloc(&s->blob, total));
        s->blob_bytes = total;
    }
    /* staging of the interchange arrays: one grow-only device buffer kept with the scene (a mesh that is uploaded again
     * every frame pays no allocation) */
    const size_t vbytes = (size_t)nv * 3 * sizeof(float), rbytes = (size_t)nt * RT_TRI_RECORD_WORDS * sizeof(int32_t);
    const size_t v_off = 0, r_off = (vbytes + 255) & ~(size_t)255, l_off = (r_off + rbytes + 255) & ~(size_t)255;
    const size_t stage_need = l_off + (size_t)nt * sizeof(int32_t);
    if (s->stage_bytes < stage_need) {
        if (s->stage) cudaFree(s->stage);
        s->stage = nullptr;
        s->stage_bytes = 0;
        CUDA_TRY(cudaMalloc(&s->stage, stage_need + stage_need / 4));
        s->stage_bytes = stage_need + stage_need / 4;
    }
    float* d_vertices = reinterpret_cast<float*>(s->stage + v_off);
    int32_t* d_recs = reinterpret_cast<int32_t*>(s->stage + r_off);
    int32_t* d_leaf_start = reinterpret_cast<int32_t*>(s->stage + l_off);
    cudaError_t err = cudaSuccess;
    const size_t nodes_bytes = (size_t)n_inner * RT_NODE_BYTES, tail_bytes = (size_t)n_wide * RT_WNODE_BYTES + (size_t)n_leafrecs * RT_LEAFREC_BYTES;
    const size_t pin_nodes = (stage_need + 255) & ~(size_t)255, pin_tail = (pin_nodes + nodes_bytes + 255) & ~(size_t)255, pin_need = pin_tail + tail_bytes;
    bool need_sync = true;
    if (pin_need <= ((size_t)64 << 20)) {
        /* small and medium meshes (a mesh that is uploaded again every frame): everything is gathered in one pinned buffer kept
         * with the scene and goes over in three truly asynchronous copies; nothing on the host has to outlive this call */
        if (s->pin_bytes < pin_need) {
            if (s->pin) cudaFreeHost(s->pin);
            s->pin = nullptr;
            s->pin_bytes = 0;
            CUDA_TRY(cudaMallocHost(&s->pin, pin_need + pin_need / 4));
            s->pin_bytes = pin_need + pin_need / 4;
        }
        memcpy(s->pin + v_off, vertices, vbytes);
        memcpy(s->pin + r_off, tri_records, rbytes);
        memcpy(s->pin + l_off, leaf_start_of_tri.data(), (size_t)nt * sizeof(int32_t));
        if (nodes_bytes) memcpy(s->pin + pin_nodes, packed.data(), nodes_bytes);
        if (n_wide > 0) memcpy(s->pin + pin_tail, wide.data(), (size_t)n_wide * RT_WNODE_BYTES);
        if (n_leafrecs > 0) memcpy(s->pin + pin_tail + (size_t)n_wide * RT_WNODE_BYTES, leaf_table.data(), (size_t)n_leafrecs * RT_LEAFREC_BYTES);
        err = cudaMemcpyAsync(s->stage, s->pin, stage_need, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess && nodes_bytes) err = cudaMemcpyAsync(s->blob + off_nodes, s->pin + pin_nodes, nodes_bytes, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess && tail_bytes) err = cudaMemcpyAsync(s->blob + off_wide, s->pin + pin_tail, tail_bytes, cudaMemcpyHostToDevice, s->stream);
        need_sync = false; /* the next full upload waits for the stream before it touches the pinned buffer again; the fast path for this event */
        if (err == cudaSuccess && !s->pin_ev[0]) err = cudaEventCreateWithFlags(&s->pin_ev[0], cudaEventDisableTiming);
        if (err == cudaSuccess) err = cudaEventRecord(s->pin_ev[0], s->stream);
        s->pin_ev_recorded[0] = err == cudaSuccess;
        s->pin_set = 0;
    } else {
        if (err == cudaSuccess) err = cudaMemcpyAsync(d_vertices, vertices, vbytes, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess) err = cudaMemcpyAsync(d_recs, tri_records, rbytes, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess && n_inner > 0)
            err = cudaMemcpyAsync(s->blob + off_nodes, packed.data(), nodes_bytes, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess && n_wide > 0)
            err = cudaMemcpyAsync(s->blob + off_wide, wide.data(), (size_t)n_wide * RT_WNODE_BYTES, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess && n_leafrecs > 0)
            err = cudaMemcpyAsync(s->blob + off_leaves, leaf_table.data(), (size_t)n_leafrecs * RT_LEAFREC_BYTES, cudaMemcpyHostToDevice, s->stream);
        if (err == cudaSuccess) err = cudaMemcpyAsync(d_leaf_start, leaf_start_of_tri.data(), (size_t)nt * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream);
    }
    if (err == cudaSuccess) {
        const int threads = 256, blocks = (nt + threads - 1) / threads;
        rtk::repack_triangles<<<blocks, threads, 0, s->stream>>>(d_vertices, d_recs, nt, d_leaf_start, reinterpret_cast<float4*>(s->blob + off_tris));
        err = cudaGetLastError();
    }
    if (err == cudaSuccess && need_sync) err = cudaStreamSynchronize(s->stream); /* the host vectors above go out of scope */
    if (err != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_scene_set_mesh: %s", cudaGetErrorString(err));

    h.has_mesh = 1;
    h.n_inner = n_inner;
    h.n_leaves = n_leafrecs; /* records of the leaf table (empty hand-made leaves have none) */
    h.n_tris = nt;
    h.max_depth = max_depth;
    h.mesh_id = id;
    h.mesh_mirror = mirror ? 1 : 0;
    h.mesh_n_in = n_in;
    h.mesh_n_out = n_out;
    memcpy(h.mesh_albedo, albedo, sizeof h.mesh_albedo);
    for (int c = 0; c < 3; c++) {
        h.root_mn[c] = arr_bvh[2 + c];
        h.root_mx[c] = arr_bvh[5 + c];
    }
    h.root_ref = root_ref;
    memcpy(h.box_abs, box_abs, sizeof box_abs);
    s->max_leaf = max_leaf;
    h.max_leaf = max_leaf;
    h.off_nodes = off_nodes;
    h.off_tris = off_tris;
    h.off_wide = off_wide;
    h.off_leaves = off_leaves;
    s->mesh_generation++;
    s->plan_valid = false;
    s->last_bvh.assign(arr_bvh, arr_bvh + (size_t)n_nodes * RT_BVH_NODE_FLOATS);
    s->last_nv = pin_need <= ((size_t)64 << 20) ? nv : 0; /* the fast path needs the pinned staging layout */
    s->last_nt = nt;
    s->stage_r_off = r_off;
    s->cur_recs = reinterpret_cast<const int32_t*>(s->stage + r_off);
    s->stage_nt = nt;
    s->has_normals = false;
    h.n_wide = n_wide;
    h.wide_depth = wide_depth;
    h.wroot_ref = wroot_ref;
    h.total_bytes = total;
    s->header_dirty = true;
    return RT_OK;
}

/* rt_scene_set_mesh with the interchange arrays ALREADY ON THE DEVICE (rt_relayout.cuh): the node relayout, the leaf table and the
 * triangle repack run as kernels; nothing crosses the bus but a 64-byte summary. */
int rt_scene_set_mesh_device(rt_scene* s, const float* d_vertices, int32_t nv, const int32_t* d_tri_records, int32_t nt, const float* d_arr_bvh,
                             int32_t n_nodes, const float albedo[3], int32_t mirror, float n_in, float n_out, int32_t id) {
    if (!s) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh_device: NULL scene");
    if (nt <= 0 || nv <= 0 || n_nodes <= 0 || !d_vertices || !d_tri_records || !d_arr_bvh || !albedo || id < 0)
        return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh_device: bad argument");
    DeviceGuard g(s->device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_scene_set_mesh_device: cudaSetDevice failed");
    if (!is_device_pointer(d_vertices, s->device) || !is_device_pointer(d_tri_records, s->device) || !is_device_pointer(d_arr_bvh, s->device))
        return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh_device: the three arrays must be device memory of device %d (host arrays: rt_scene_set_mesh)", s->device);
    SceneHeader& h = s->header;
    cudaStream_t st = s->stream;
    CUDA_TRY(cudaStreamSynchronize(st));
    const int T = 256, NB = (n_nodes + T - 1) / T;
    /* scratch: 7 int arrays of n_nodes + 1, the totals, the scan's work space */
    int* scratch = nullptr;
    rtrelayout::Totals* d_tot = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (int*)nullptr, (int*)nullptr, n_nodes + 1);
    const size_t N1 = (size_t)n_nodes + 1;
    cudaError_t err = cudaMalloc(&scratch, 7 * N1 * sizeof(int));
    if (err == cudaSuccess) err = cudaMalloc(&d_tot, sizeof(rtrelayout::Totals));
    if (err == cudaSuccess) err = cudaMalloc(&tmp, tmp_bytes + 256);
    auto cleanup = [&]() {
        if (scratch) cudaFree(scratch);
        if (d_tot) cudaFree(d_tot);
        if (tmp) cudaFree(tmp);
    };
    if (err != cudaSuccess) {
        cleanup();
        return rtb::fail(RT_ERR_NOMEM, "rt_scene_set_mesh_device: %s", cudaGetErrorString(err));
    }
    int *is_inner = scratch, *n_chunks = scratch + N1, *n_virt = scratch + 2 * N1, *inner_rank = scratch + 3 * N1, *virt_base = scratch + 4 * N1,
        *leaf_base = scratch + 5 * N1, *depth = scratch + 6 * N1;
    cudaMemsetAsync(scratch, 0, 3 * N1 * sizeof(int), st);
    cudaMemsetAsync(depth, 0xff, N1 * sizeof(int), st);
    cudaMemsetAsync(depth, 0, sizeof(int), st);
    cudaMemsetAsync(d_tot, 0, sizeof(rtrelayout::Totals), st);
    rtrelayout::classify<<<NB, T, 0, st>>>(d_arr_bvh, n_nodes, nt, is_inner, n_chunks, n_virt, d_tot);
    rtrelayout::check_indices<<<(nt + T - 1) / T, T, 0, st>>>(d_tri_records, nt, nv, d_tot);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, is_inner, inner_rank, n_nodes + 1, st);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, n_virt, virt_base, n_nodes + 1, st);
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, n_chunks, leaf_base, n_nodes + 1, st);
    for (int level = 0; level < RT_STACK_CAP; level++) rtrelayout::depth_step<<<NB, T, 0, st>>>(d_arr_bvh, n_nodes, level, depth, d_tot);
    rtrelayout::Totals tot;
    int sums[3] = {0, 0, 0};
    float root[10];
    err = cudaMemcpyAsync(&tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&sums[0], inner_rank + n_nodes, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&sums[1], virt_base + n_nodes, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(&sums[2], leaf_base + n_nodes, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaMemcpyAsync(root, d_arr_bvh, sizeof root, cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    if (err != cudaSuccess) {
        cleanup();
        return rtb::fail(RT_ERR_CUDA, "rt_scene_set_mesh_device: %s", cudaGetErrorString(err));
    }
    if (tot.error) {
        cleanup();
        return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh_device: malformed arrays (%s)", tot.error == 1 ? "a node with exactly one child" : tot.error == 2 ? "child index out of order / range" :
                                                                                                 tot.error == 3 ? "leaf triangle range invalid" : "triangle references a vertex out of range");
    }
    const int n_real = sums[0], n_virtual = sums[1], n_leafrecs = sums[2];
    int extra_levels = 0;
    while ((1 << extra_levels) < tot.max_chunks) extra_levels++;
    if (tot.max_chunks > 1) extra_levels++;
    const int max_depth = tot.max_depth + 1 + extra_levels;
    if (max_depth > RT_STACK_CAP - 2) {
        cleanup();
        return rtb::fail(RT_ERR_UNSUPPORTED, "rt_scene_set_mesh_device: packed BVH depth %d exceeds the traversal stack (%d)", max_depth, RT_STACK_CAP - 2);
    }
    const int n_inner = n_real + n_virtual;
    const size_t off_nodes = RT_HEADER_BYTES;
    const size_t off_tris = (off_nodes + (size_t)n_inner * RT_NODE_BYTES + 63) & ~(size_t)63;
    const size_t off_wide = (off_tris + (size_t)nt * RT_TRI_BYTES + 127) & ~(size_t)127;
    const size_t off_leaves = off_wide; /* no wide index on this path */
    const size_t total = off_leaves + (size_t)n_leafrecs * RT_LEAFREC_BYTES;
    if (s->blob_bytes < total || s->blob_bytes > 2 * total + (1u << 20)) {
        if (s->blob) cudaFree(s->blob);
        s->blob = nullptr;
        s->blob_bytes = 0;
        err = cudaMalloc(&s->blob, total);
        if (err != cudaSuccess) {
            cleanup();
            return rtb::fail(RT_ERR_NOMEM, "rt_scene_set_mesh_device: %zu bytes for the scene blob: %s", total, cudaGetErrorString(err));
        }
        s->blob_bytes = total;
    }
    if (s->stage_bytes < (size_t)nt * sizeof(int32_t)) {
        if (s->stage) cudaFree(s->stage);
        s->stage = nullptr;
        s->stage_bytes = 0;
        err = cudaMalloc(&s->stage, (size_t)nt * sizeof(int32_t));
        if (err != cudaSuccess) {
            cleanup();
            return rtb::fail(RT_ERR_NOMEM, "rt_scene_set_mesh_device: %s", cudaGetErrorString(err));
        }
        s->stage_bytes = (size_t)nt * sizeof(int32_t);
    }
    int32_t* d_leaf_start = reinterpret_cast<int32_t*>(s->stage);
    float4* records = reinterpret_cast<float4*>(s->blob + off_nodes);
    rtrelayout::write_inner<<<NB, T, 0, st>>>(d_arr_bvh, n_nodes, nt, is_inner, inner_rank, n_chunks, virt_base, n_real, records);
    rtrelayout::write_leaves<<<NB, T, 0, st>>>(d_arr_bvh, n_nodes, nt, is_inner, n_chunks, virt_base, leaf_base, n_real, records,
                                               reinterpret_cast<float4*>(s->blob + off_leaves), d_leaf_start);
    rtk::repack_triangles<<<(nt + T - 1) / T, T, 0, st>>>(d_vertices, d_tri_records, nt, d_leaf_start, reinterpret_cast<float4*>(s->blob + off_tris));
    err = cudaGetLastError();
    if (err == cudaSuccess) err = cudaStreamSynchronize(st); /* the caller's arrays and the scratch may go away */
    /* root reference, as child_ref() of node 0 */
    int root_ref;
    {
        const int l = (int)root[0];
        const int ts = (int)root[8], cnt = (int)root[9] - ts;
        const int chunks = cnt > 0 ? (cnt + RT_LEAF_MAX - 1) / RT_LEAF_MAX : 0;
        if (l != -1) root_ref = 0; /* the root is the first inner node in pre-order */
        else if (cnt <= 0) root_ref = -1 - ((nt << 2) | 0);
        else if (chunks == 1) root_ref = -1 - ((ts << 2) | (cnt - 1));
        else root_ref = n_real + 0; /* virt_base of node 0 is 0 */
    }
    cleanup();
    if (err != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_scene_set_mesh_device: %s", cudaGetErrorString(err));
    h.has_mesh = 1;
    h.n_inner = n_inner;
    h.n_leaves = n_leafrecs;
    h.n_tris = nt;
    h.max_depth = max_depth;
    h.mesh_id = id;
    h.mesh_mirror = mirror ? 1 : 0;
    h.mesh_n_in = n_in;
    h.mesh_n_out = n_out;
    memcpy(h.mesh_albedo, albedo, sizeof h.mesh_albedo);
    for (int c = 0; c < 3; c++) {
        h.root_mn[c] = root[2 + c];
        h.root_mx[c] = root[5 + c];
        memcpy(&h.box_abs[c], &tot.box_abs[c], sizeof(float));
    }
    h.root_ref = root_ref;
    h.wroot_ref = 0;
    h.max_leaf = tot.max_leaf;
    s->max_leaf = tot.max_leaf;
    h.off_nodes = off_nodes;
    h.off_tris = off_tris;
    h.off_wide = off_wide;
    h.off_leaves = off_leaves;
    h.n_wide = 0;
    h.wide_depth = 0;
    h.total_bytes = total;
    s->mesh_generation++;
    s->last_bvh.clear();
    s->last_nv = s->last_nt = 0;
    s->stage_nt = 0; /* the records are the caller's: vertex normals need the host upload */
    s->has_normals = false;
    s->header_dirty = true;
    return RT_OK;
}

/* The mesh of an rt_mesh handle: straight from the device when rt_mesh_build_bvh_gpu left its arrays on this scene's device, through
 * the host interchange arrays otherwise. */
int rt_scene_set_mesh_from(rt_scene* s, rt_mesh* m, const float albedo[3], int32_t mirror, float n_in, float n_out, int32_t id) {
    if (!s || !m) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh_from: NULL argument");
    int32_t nv = 0, nt = 0, nn = 0;
    int rc = rt_mesh_counts(m, &nv, &nt, &nn);
    if (rc != RT_OK) return rc;
    if (nn <= 0) return rtb::fail(RT_ERR_STATE, "rt_scene_set_mesh_from: build the BVH first");
    void* keep = rtb::mesh_device_handle(m);
    if (keep) {
        int dev = -1;
        const float *V = nullptr, *arr = nullptr;
        const int32_t* recs = nullptr;
        int32_t dn = 0;
        rtb::bvh_device_arrays(keep, &dev, &V, &recs, &arr, &dn);
        if (dev == s->device) return rt_scene_set_mesh_device(s, V, nv, recs, nt, arr, dn, albedo, mirror, n_in, n_out, id);
    }
    return rt_scene_set_mesh(s, rt_mesh_vertices(m), nv, rt_mesh_tri_records(m), nt, rt_mesh_arr_bvh(m), nn, albedo, mirror, n_in, n_out, id);
}

int rt_scene_set_mesh_normals(rt_scene* s, const float* normals, int32_t n_normals) {
    if (!s || n_normals < 0 || (n_normals > 0 && !normals)) return rtb::fail(RT_ERR_INVALID, "rt_scene_set_mesh_normals: bad argument");
    DeviceGuard g(s->device);
    if (n_normals == 0) {
        s->has_normals = false;
        return RT_OK;
    }
    if (!s->header.has_mesh || s->stage_nt != s->header.n_tris || !s->stage)
        return rtb::fail(RT_ERR_STATE, "rt_scene_set_mesh_normals: call rt_scene_set_mesh first (the normal indices are words 6-8 of its triangle records)");
    const int nt = s->stage_nt;
    if (s->d_normals_cap < (size_t)n_normals * 3) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->d_normals) cudaFree(s->d_normals);
        s->d_normals = nullptr;
        s->d_normals_cap = 0;
        CUDA_TRY(cudaMalloc(&s->d_normals, (size_t)n_normals * 3 * sizeof(float)));
        s->d_normals_cap = (size_t)n_normals * 3;
    }
    if (s->tri_normals_cap < (size_t)nt * 3) {
        CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->tri_normals) cudaFree(s->tri_normals);
        s->tri_normals = nullptr;
        s->tri_normals_cap = 0;
        CUDA_TRY(cudaMalloc(&s->tri_normals, (size_t)nt * 3 * sizeof(float4)));
        s->tri_normals_cap = (size_t)nt * 3;
    }
    CUDA_TRY(cudaMemcpyAsync(s->d_normals, normals, (size_t)n_normals * 3 * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    rtk::repack_normals<<<(nt + 255) / 256, 256, 0, s->stream>>>(s->d_normals, n_normals, s->cur_recs, nt, s->tri_normals);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s->stream)); /* the caller's array may go away */
    s->has_normals = true;
    return RT_OK;
}

int rt_scene_blob_size(rt_scene* s, size_t* bytes) {
    if (!s || !bytes) return rtb::fail(RT_ERR_INVALID, "rt_scene_blob_size: bad argument");
    *bytes = s->blob ? (size_t)s->header.total_bytes : (size_t)RT_HEADER_BYTES;
    return RT_OK;
}

int rt_scene_blob_export(rt_scene* s, void** device_ptr, size_t* bytes) {
    if (!s || !device_ptr || !bytes) return rtb::fail(RT_ERR_INVALID, "rt_scene_blob_export: bad argument");
    DeviceGuard g(s->device);
    if (s->header_dirty || !s->blob) {
        int rc = upload_header(s);
        if (rc != RT_OK) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    *device_ptr = s->blob;
    *bytes = (size_t)s->header.total_bytes;
    return RT_OK;
}

int rt_scene_blob_copy_out(rt_scene* s, void* device_dst, size_t bytes) {
    void* src = nullptr;
    size_t n = 0;
    int rc = rt_scene_blob_export(s, &src, &n);
    if (rc != RT_OK) return rc;
    if (!device_dst || bytes < n) return rtb::fail(RT_ERR_INVALID, "rt_scene_blob_copy_out: destination of %zu bytes is smaller than the blob (%zu)", bytes, n);
    DeviceGuard g(s->device);
    CUDA_TRY(cudaMemcpyAsync(device_dst, src, n, cudaMemcpyDeviceToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return RT_OK;
}

int rt_scene_blob_import(rt_scene* s, const void* device_ptr, size_t bytes) {
    if (!s || !device_ptr || bytes < RT_HEADER_BYTES) return rtb::fail(RT_ERR_INVALID, "rt_scene_blob_import: bad argument");
    DeviceGuard g(s->device);
    SceneHeader h;
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    CUDA_TRY(cudaMemcpy(&h, device_ptr, sizeof h, cudaMemcpyDeviceToHost));
    if (h.magic != RT_BLOB_MAGIC || h.layout_version != RT_LAYOUT_VERSION || h.total_bytes != bytes)
        return rtb::fail(RT_ERR_INVALID, "rt_scene_blob_import: not a scene blob (magic %08x, %llu bytes declared, %llu given)", h.magic,
                         (unsigned long long)h.total_bytes, (unsigned long long)bytes);
    if (h.n_spheres < 0 || h.n_spheres > RT_MAX_SPHERES || h.off_tris + (uint64_t)h.n_tris * RT_TRI_BYTES > bytes ||
        h.n_wide < 0 || h.off_wide + (uint64_t)h.n_wide * RT_WNODE_BYTES > bytes || h.n_leaves < 0 ||
        h.off_leaves + (uint64_t)h.n_leaves * RT_LEAFREC_BYTES > bytes)
        return rtb::fail(RT_ERR_INVALID, "rt_scene_blob_import: inconsistent header");
    if ((const unsigned char*)device_ptr != s->blob) {
        if (s->blob_bytes < bytes) {
            if (s->blob) cudaFree(s->blob);
            s->blob = nullptr;
            s->blob_bytes = 0;
            CUDA_TRY(cudaMalloc(&s->blob, bytes));
            s->blob_bytes = bytes;
        }
        CUDA_TRY(cudaMemcpyAsync(s->blob, device_ptr, bytes, cudaMemcpyDeviceToDevice, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    s->header = h;
    s->max_leaf = h.max_leaf; /* host-side guard of the tie-break rank (push_order 0): travels with the blob */
    s->last_bvh.clear();
    s->last_nv = s->last_nt = 0;
    s->has_normals = false; /* vertex normals are not part of the blob */
    s->stage_nt = 0;
    s->header_dirty = false;
    s->mesh_generation++; /* the anchored-ray bins of the previous mesh are stale */
    s->plan_valid = false;
    return RT_OK;
}

int rt_scene_sync(rt_scene* s, rt_stats* stats) {
    if (!s) return rtb::fail(RT_ERR_INVALID, "rt_scene_sync: NULL scene");
    DeviceGuard g(s->device);
    if (s->pending) {
        if (s->last_was_wavefront && s->last_was_anchored)
            for (int k = 0; k < 2; k++)
                if (s->bins[k].status) CUDA_TRY(cudaMemcpyAsync(s->bins[k].h_status, s->bins[k].status, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaMemcpyAsync(s->h_sticky, s->sticky, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        if (s->last_was_wavefront)
            CUDA_TRY(cudaMemcpyAsync(s->h_wf_counters, s->wf_counters, RT_MAX_STRIPS * sizeof(rtk::WfCounters), cudaMemcpyDeviceToHost, s->stream));
        else
            CUDA_TRY(cudaMemcpyAsync(s->h_counters, s->counters, RT_NCOUNTERS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(s->stream));
#ifdef RT_TIMELINE
    if (s->tl_buf && s->tl_n > 0 && getenv("RT_TIMELINE_PRINT")) {
        unsigned long long hst[128];
        cudaMemcpy(hst, s->tl_buf, sizeof hst, cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (int k = 0; k < s->tl_n; k++) t0 = std::min(t0, hst[k]);
        fprintf(stderr, "[timeline us]");
        for (int k = 0; k < s->tl_n; k++)
            fprintf(stderr, " %s%d %.1f-%.1f", s->tl_name[k], s->tl_band[k], (double)(hst[k] - t0) * 1e-3, (double)(hst[64 + k] - t0) * 1e-3);
        fprintf(stderr, "\n");
    }
#endif
    if (s->copies_pending) {
        CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
        s->copies_pending = false;
    }
    if (s->pending && s->last_was_wavefront) { /* fold the strips' counters */
        for (int k = 0; k < RT_NCOUNTERS; k++) s->h_counters[k] = 0;
        for (int st = 0; st < RT_MAX_STRIPS; st++)
            for (int k = 0; k < RT_NCOUNTERS; k++)
                s->h_counters[k] = (k == 3) ? std::max<unsigned long long>(s->h_counters[k], s->h_wf_counters[st].stats[k]) : s->h_counters[k] + s->h_wf_counters[st].stats[k];
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        if (s->pending) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
            stats->kernel_ms = ms;
            stats->rays = s->h_counters[0];
            stats->node_visits = s->h_counters[1];
            stats->tri_tests = s->h_counters[2];
            stats->max_stack = (int32_t)s->h_counters[3];
            stats->slab_fallbacks = s->h_counters[4];
            stats->tri_exact = s->h_counters[5];
            stats->launches = s->pending_launches;
        }
    }
    /* overflow flags are sticky on the device: every frame enqueued since the last sync is covered, not just the last one */
    const bool failed = s->pending && s->h_sticky[1] != 0;
    const bool task_overflow = s->pending && s->h_sticky[0] != 0;
    if (failed || task_overflow) {
        s->h_sticky[0] = s->h_sticky[1] = 0;
        CUDA_TRY(cudaMemsetAsync(s->sticky, 0, 2 * sizeof(unsigned long long), s->stream));
    }
    if (s->pending && s->last_was_wavefront && s->opt.debug_pool) {
        unsigned long long d[8];
        cudaMemcpy(d, s->wf_counters->dbg, sizeof d, cudaMemcpyDeviceToHost);
        rtk::WfCounters hc;
        cudaMemcpy(&hc, s->wf_counters, sizeof hc, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[queues] nA %d %d %d %d  nS %d %d %d %d  head %d %d %d %d\n", hc.nA[0], hc.nA[1], hc.nA[2], hc.nA[3], hc.nS[0], hc.nS[1], hc.nS[2],
                hc.nS[3], hc.head[0], hc.head[1], hc.head[2], hc.head[3]);
        fprintf(stderr, "[pool] N steps %llu tasks %llu (%.1f/step)  T steps %llu tasks %llu (%.1f/step)  admissions %llu\n", d[0], d[1],
                d[0] ? (double)d[1] / d[0] : 0., d[2], d[3], d[2] ? (double)d[3] / d[2] : 0., d[4]);
    }
    if (s->pending && s->last_was_wavefront && s->dbg_warps && s->opt.debug_warps[0]) {
        std::vector<int> hst(s->dbg_warps_ints);
        cudaMemcpy(hst.data(), s->dbg_warps, hst.size() * sizeof(int), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(s->opt.debug_warps, "wb")) {
            fwrite(hst.data(), sizeof(int), hst.size(), f);
            fclose(f);
        }
    }
    if (s->pending && s->last_was_wavefront && s->last_was_anchored)
        for (int k = 0; k < 2; k++)
            if (s->bins[k].h_status && s->bins[k].h_status[1]) s->bins[k].grow = true; /* results were right (exact search for the cut lists); the next build gets room */
    s->pending = false;
    if (failed) return rtb::fail(RT_ERR_STATE, "rt_render: traversal task pool overflow (BVH deeper than the upload-time bound)");
    if (task_overflow) {
        /* more (ray, leaf) tasks than the buffer holds: the frame is incomplete. The next render gets a buffer twice the
         * size; a synchronous rt_render repeats the frame by itself. */
        s->task_factor = std::min(s->task_factor * 2, 256);
        if (s->opt.debug_pool) fprintf(stderr, "[tasks] buffer overflow, next frames get %d tasks per pixel\n", s->task_factor);
        return rtb::fail(RT_ERR_AGAIN, "rt_render: (ray, leaf) task buffer overflow; render the frame again (buffer doubled to %d tasks per pixel)", s->task_factor);
    }
    return RT_OK;
}

} /* extern "C" */

namespace {

/* What rt_render decides before it launches (plan_frame) and what the launches need (enqueue_frame). Plain data: two frames with
 * byte-equal plans (and scene state, see frame_key) are the same sequence of launches. */
struct FramePlan {
    rt_params p;
    uint32_t flags;
    int rows;
    size_t npx;
    void* user[5];
    void* dev[5];
    bool copy_back[5];
    size_t bytes[5];
    rtk::RenderArgs a;
    int variant;       /* 0 / 1 render_mega, 2 wavefront, 3 render_stoch */
    unsigned grid;     /* render_mega / render_stoch */
    bool stochastic, count;
    bool async_copy;   /* host outputs leave on the copy stream behind the frame (RT_RENDER_NO_SYNC), not inside it */
    int scratch_set;
    bool jitter;       /* one sample of one segment in stochastic mode: the deterministic pipeline with jittered camera rays */
    /* wavefront pipeline */
    bool wide, anchored, diffuse_only, trav_round0, dbg_times;
    bool tops;        /* wf_traverse reads the top levels from shared memory (option top_smem) */
    int n_top;
    int segments, npool_cap, n_strips, spill_cap;
    size_t trav_smem;
    unsigned pers_grid;
    int* dbg_ptr;
    size_t dbg_ints;
    size_t st_rng_off, st_total_off, st_rec_off, task_slack;
    int prelaunches;   /* one-off launches plan_frame enqueued itself (random-stream table) */
};

/* ---- rt_render, first half: everything that is decided, sized or (re)built before a frame is launched -------------------------
 * Validation, output buffers, kernel variant, queue / task / overflow buffers, the anchored-ray bins. May allocate and
 * synchronise; enqueues nothing of the frame itself (only one-off set-up work: the random-stream table, a bins build). */
int plan_frame(rt_scene* s, const rt_params* p, uint32_t flags, void* const user_in[5], FramePlan& P) {
    if (p->W <= 0 || p->H <= 0 || p->num_rays < 1 || p->num_bounce < 0) return rtb::fail(RT_ERR_INVALID, "rt_render: bad W/H/num_rays/num_bounce");
    const bool stochastic = p->aa_sigma != 0.f || p->indirect != 0;
    if (stochastic && p->num_bounce + (p->extra_segment ? 1 : 0) > RT_STOCH_MAX_SEGMENTS)
        return rtb::fail(RT_ERR_UNSUPPORTED, "rt_render: the stochastic mode supports at most %d path segments (the reference: MAX_RAY_DEPTH 10, optimized.cu:22)", RT_STOCH_MAX_SEGMENTS);
    if (p->gamma_mode != 0 && p->gamma_mode != 1) return rtb::fail(RT_ERR_INVALID, "rt_render: gamma_mode must be 0 or 1");
    if (p->push_order != 0 && p->push_order != 1) return rtb::fail(RT_ERR_INVALID, "rt_render: push_order must be 0 or 1");
    const int step = p->row_step > 0 ? p->row_step : 1;
    if (p->row_begin < 0 || p->row_begin >= p->H) return rtb::fail(RT_ERR_INVALID, "rt_render: row_begin out of range");
    const int group = p->row_group > 1 ? p->row_group : 1;
    if ((group & (group - 1)) != 0 || group > 64) return rtb::fail(RT_ERR_INVALID, "rt_render: row_group must be a power of two <= 64");
    if (group > 1 && step < group) return rtb::fail(RT_ERR_INVALID, "rt_render: row_step %d is smaller than row_group %d (groups would overlap)", step, group);
    int group_shift = 0;
    while ((1 << group_shift) < group) group_shift++;
    /* groups start at row_begin, row_begin + step, ...; only the last one may be cut short by the frame's end */
    const int n_groups = (p->H - p->row_begin + step - 1) / step;
    const int max_rows = (n_groups - 1) * group + std::min(group, p->H - (p->row_begin + (n_groups - 1) * step));
    const int rows = p->row_count > 0 ? p->row_count : max_rows;
    if (rows > max_rows) return rtb::fail(RT_ERR_INVALID, "rt_render: row_count %d exceeds the %d rows available", rows, max_rows);
    const SceneHeader& h = s->header;
    if (h.n_spheres == 0 && !h.has_mesh) return rtb::fail(RT_ERR_STATE, "rt_render: empty scene (set spheres and/or a mesh first)");
    /* object ids must be exactly 0..n-1, as Scene::objects indices are (optimized.cu:687-725) */
    {
        const int n_obj = h.n_spheres + (h.has_mesh ? 1 : 0);
        unsigned seen = 0;
        bool ok = true;
        for (int k = 0; k < h.n_spheres; k++) {
            const int id = h.spheres[k].id;
            if (id >= n_obj || (seen >> id) & 1u) ok = false;
            else seen |= 1u << id;
        }
        if (h.has_mesh && (h.mesh_id >= n_obj || ((seen >> h.mesh_id) & 1u))) ok = false;
        if (!ok) return rtb::fail(RT_ERR_INVALID, "rt_render: object ids must be a permutation of 0..%d", n_obj - 1);
    }
    /* a still-pending RT_RENDER_NO_SYNC call needs no wait: counters, scratch buffers and events are reused in
     * stream order, and the stats of the older call are simply superseded */
    /* the kernels take the header by value from the host copy: a changed light or sphere set needs no upload here
     * (the device copy inside the blob is refreshed by rt_scene_blob_export, its only reader) */
    if (!s->blob) {
        int rc = upload_header(s);
        if (rc != RT_OK) return rc;
        s->header_dirty = true;
    }

    const size_t npx = (size_t)rows * p->W;
    void* user[5] = {user_in[0], user_in[1], user_in[2], user_in[3], user_in[4]};
    const size_t bytes[5] = {npx * 3, npx * 4, npx * 4, npx * 4, npx};
    void* dev[5];
    bool copy_back[5];
    bool any_host = false;
    for (int k = 0; k < 5; k++) any_host = any_host || (user[k] && !is_device_pointer(user[k], s->device));
    /* host outputs of an enqueue-only call: the other scratch set, copied back on the copy stream behind the frame */
    const bool async_copy = any_host && (flags & RT_RENDER_NO_SYNC) != 0;
    if (async_copy) s->scratch_set ^= 1;
    const int set = async_copy ? s->scratch_set : 0;
    P.async_copy = async_copy;
    P.scratch_set = set;
    for (int k = 0; k < 5; k++) {
        dev[k] = nullptr;
        copy_back[k] = false;
        if (!user[k]) continue;
        if (is_device_pointer(user[k], s->device)) {
            dev[k] = user[k];
        } else {
            int rc = ensure_scratch(s, set, k, bytes[k]);
            if (rc != RT_OK) return rc;
            dev[k] = s->scratch[set][k];
            copy_back[k] = true;
        }
    }

    rtk::RenderArgs a;
    memset(&a, 0, sizeof a); /* padding too: the plan is compared byte-wise with the previous frame's */
    a.W = p->W;
    a.H = p->H;
    a.rows = rows;
    a.row_begin = p->row_begin;
    a.row_step = step;
    a.row0 = 0;
    a.group_shift = group_shift;
    a.segments = p->num_bounce + (p->extra_segment ? 1 : 0);
    a.num_rays = p->num_rays;
    a.camx = p->cam[0];
    a.camy = p->cam[1];
    a.camz = p->cam[2];
    a.z = p->z;
    a.eps_surface = p->eps_surface;
    a.eps_tri = p->eps_tri;
    a.push_order = p->push_order;
    a.gamma_mode = p->gamma_mode;
    a.rgb = (uint8_t*)dev[0];
    a.hit_obj = (int32_t*)dev[1];
    a.hit_tri = (int32_t*)dev[2];
    a.hit_t = (float*)dev[3];
    a.shadow = (uint8_t*)dev[4];
    a.counters = s->counters;
    a.gamma_tab = s->gamma_tab;
    a.debug_cost = s->opt.debug_cost;
    if (p->camera_mode != 0 && p->camera_mode != 1) return rtb::fail(RT_ERR_INVALID, "rt_render: camera_mode must be 0 or 1");
    a.camera_mode = p->camera_mode;
    for (int k = 0; k < 3; k++) {
        a.bx[k] = p->cam_bx[k];
        a.by[k] = p->cam_by[k];
        a.bz[k] = p->cam_bz[k];
    }
    if (p->smooth_normals && h.has_mesh) {
        if (!s->has_normals) return rtb::fail(RT_ERR_STATE, "rt_render: smooth_normals needs rt_scene_set_mesh_normals after rt_scene_set_mesh");
        a.tri_normals = s->tri_normals;
    }
    if (p->accumulate < 0) return rtb::fail(RT_ERR_INVALID, "rt_render: accumulate must be >= 0");
    if (p->accumulate > 0) {
        if (p->accumulate > 1 && s->accum_px != npx) return rtb::fail(RT_ERR_STATE, "rt_render: accumulation frame %d without a frame 1 of the same size", p->accumulate);
        if (s->accum_px != npx) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (s->accum) cudaFree(s->accum);
            if (s->linear) cudaFree(s->linear);
            s->accum = s->linear = nullptr;
            s->accum_px = 0;
            CUDA_TRY(cudaMalloc(&s->accum, npx * sizeof(float4)));
            CUDA_TRY(cudaMalloc(&s->linear, npx * sizeof(float4)));
            s->accum_px = npx;
        }
        a.linear = s->linear;
    }

    P.prelaunches = 0;
    const int tiles_x = (p->W + 15) / 16, tiles_y = (rows + 7) / 8;
    const unsigned grid = (unsigned)tiles_x * (unsigned)tiles_y;
    const bool count = (flags & RT_RENDER_COUNT_WORK) != 0;
    int variant = s->opt.variant;
    /* stochastic mode: one wavefront pass per sample (the samples of a pixel share one random stream); the
     * thread-per-pixel kernel render_stoch is the fallback and the in-library cross-check (RT_STOCH_MEGA=1) */
    const bool stoch_mega = s->opt.stoch_mega != 0;
    if (stochastic && (variant != 2 || stoch_mega)) variant = 3;
    /* tie-break rank of render_wave: (n_tris - leaf_start) and the in-leaf offset share 32 bits */
    int bits_n = 1;
    while ((1ll << bits_n) <= (long long)h.n_tris) bits_n++;
    a.rank_off_bits = std::min(32 - bits_n, 16);
    if (variant == 2 && p->push_order == 0 && h.has_mesh && (long long)s->max_leaf > (1ll << a.rank_off_bits)) variant = 1;
    const int segments = a.segments;
    if (variant == 2 && segments > WF_MAX_ROUNDS - 1) variant = stochastic ? 3 : 1;
    if (stochastic && variant == 1) variant = 3;
    s->last_was_wavefront = (variant == 2);
    if (variant != 2 && (p->camera_mode || a.tri_normals || a.linear))
        return rtb::fail(RT_ERR_UNSUPPORTED, "rt_render: the viewer-derived features (camera basis, smooth normals, accumulation) need the wavefront pipeline (variant 2, <= %d segments)", WF_MAX_ROUNDS - 1);
    if (stochastic) {
        /* start states of the random streams: once per (seed, W, H), not per launch (rt_stochastic.cuh) */
        const unsigned long long seed = p->reserved ? (unsigned long long)(unsigned int)p->reserved : 123456ull; /* optimized.cu:745 */
        const size_t frame_px = (size_t)p->W * p->H;
        if (!s->rng_states || s->rng_W != p->W || s->rng_H != p->H || s->rng_seed != seed) {
            if (s->rng_capacity < frame_px) {
                CUDA_TRY(cudaStreamSynchronize(s->stream));
                if (s->rng_states) cudaFree(s->rng_states);
                s->rng_states = nullptr;
                s->rng_capacity = 0;
                CUDA_TRY(cudaMalloc(&s->rng_states, frame_px * sizeof(rtk::XorwowState)));
                s->rng_capacity = frame_px;
            }
            rtk::xorwow_init_states<<<(unsigned)((frame_px + 255) / 256), 256, 0, s->stream>>>(seed, (unsigned)frame_px, s->rng_states);
            CUDA_TRY(cudaGetLastError());
            s->rng_W = p->W;
            s->rng_H = p->H;
            s->rng_seed = seed;
            s->jitter_valid = false;
            P.prelaunches++; /* the one-off table build is not part of the frame time */
        }
    }
    /* `./optimized 1 1`-like calls — ONE sample of ONE segment: the only random numbers that reach the image are the two of the
     * Box-Muller jitter (optimized.cu:756-758); the two drawn at the diffuse hit (:633-634) feed a bounce that is never traced, and
     * the fold (:653-660) is c = direct. Such a frame is the deterministic pipeline with jittered camera rays: no per-pixel stream state,
     * no diffuse records, no fold pass (224 B of state traffic per pixel otherwise). Same frames (tests/test_gpu_stochastic.py). */
    const bool one_shot = stochastic && variant == 2 && p->num_rays == 1 && segments == 1 && !count && s->opt.one_shot != 0;
    P.jitter = one_shot;
    if (one_shot && (!s->jitter_valid || s->jitter_sigma != p->aa_sigma || s->jitter_libm != s->opt.transcendentals)) {
        const size_t frame_px = (size_t)p->W * p->H;
        if (s->jitter_capacity < frame_px) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (s->jitter_tab) cudaFree(s->jitter_tab);
            s->jitter_tab = nullptr;
            s->jitter_capacity = 0;
            CUDA_TRY(cudaMalloc(&s->jitter_tab, frame_px * sizeof(float2)));
            s->jitter_capacity = frame_px;
        }
        rtk::jitter_table<<<(unsigned)((frame_px + 255) / 256), 256, 0, s->stream>>>(reinterpret_cast<const uint4*>(s->rng_states), (unsigned)frame_px, p->aa_sigma,
                                                                                    s->opt.transcendentals, s->jitter_tab);
        CUDA_TRY(cudaGetLastError());
        s->jitter_sigma = p->aa_sigma;
        s->jitter_libm = s->opt.transcendentals;
        s->jitter_valid = true;
        P.prelaunches++;
    }
    const bool stochastic_pipeline = stochastic && !one_shot;
    P.wide = P.anchored = P.diffuse_only = P.trav_round0 = P.dbg_times = false;
    P.tops = false;
    P.n_top = 0;
    P.npool_cap = P.spill_cap = 0;
    P.n_strips = 1;
    P.trav_smem = P.st_rng_off = P.st_total_off = P.st_rec_off = 0;
    P.task_slack = 4096;
    P.pers_grid = 0;
    P.dbg_ptr = nullptr;
    P.dbg_ints = 0;
    if (variant == 2) {
        const size_t task_slack = P.task_slack;
    /* the wide index (rt_layout.h) is the production search structure; the instrumented build counts the reference's
     * own node visits and therefore walks the two-child records, as does RT_WIDE=0 (A/B timing, cross-check) */
    /* RT_WIDE: 1 on, 0 off; unset: on for the incoherent bounce rays of the stochastic mode (measured, 6 blocks per SM:
     * stochastic 4 3 4.91 -> 4.67 ms, but the mirror 4K frame 1.23 -> 1.28 ms and no gain for the tree search of
     * coherent rays, profiles/r01_notes.md) */
    const int env_wide_v = s->opt.wide;
    /* small shards (one rank's rows of a frame split over 8 GPUs): a tree-search round is bound by the dependency depth of single rays,
     * which the wide index halves (measured, 1/8 of the 4K depth-4 frame: 0.307 -> 0.285 ms; the whole frame 1.117 -> 1.170 ms) */
    const bool small_shard = npx < 1500000;
    const bool env_wide = env_wide_v > 0 || (env_wide_v < 0 && ((stochastic_pipeline && p->indirect != 0) || small_shard));
    const bool env_wide_count = s->opt.wide_count != 0; /* timeline of the wide kernel: node_visits then counts wide nodes */
    const bool wide = env_wide && (!count || env_wide_count) && h.n_wide > 0;
    /* anchored rays (rt_bins.cuh): camera rays and shadow rays find their leaves through per-anchor bins and wf_leaves
     * tests the triangles; wf_traverse keeps the rays that start anywhere else (bounces). The instrumented build
     * counts the reference's node visits and therefore searches the tree; RT_ANCHOR=0 does so too (A/B, cross-check). */
    const int env_anchor = s->opt.anchored; /* 0 off, 1 on, -1: by mesh size */
    /* measured (tools/size_sweep.py, profiles/r01_notes.md): the bins win from 1 k to 1 M leaves (1080p: 0.25 vs 0.41 ms at
     * 1 k, 0.77 vs 1.20 ms at 241 k, 5.97 vs 6.46 ms at 1 M); at 2.5 M leaves and 4K a cell lists hundreds of leaves and
     * one task per candidate loses against the tree search (13.9 vs 12.9 ms) */
    bool anchored = (env_anchor > 0 || (env_anchor < 0 && h.n_leaves <= 1200000)) && !count && h.has_mesh && h.n_leaves > 0 && segments > 0;
    if (anchored) {
        int rc = ensure_bins(s, 0, p->cam);
        if (rc == RT_OK) rc = ensure_bins(s, 1, h.L);
        if (rc != RT_OK) return rc;
        anchored = s->bins[0].usable && s->bins[1].usable;
    }
    s->last_was_anchored = anchored;
    /* no mirror, no refractive object: the kernels without the reflection / refraction code (smaller, less instruction fetch) */
    bool diffuse_only = !(h.has_mesh && (h.mesh_mirror || h.mesh_n_in != h.mesh_n_out)) && s->opt.diffuse_kernels != 0;
    for (int k = 0; k < h.n_spheres; k++) diffuse_only = diffuse_only && !h.spheres[k].mirror && h.spheres[k].n_in == h.spheres[k].n_out;
    /* round 0 holds tree-searched queries only when a path can go on inside wf_generate: past a mirror or refractive
     * sphere, or along the indirect bounce of a pixel shaded on the spot */
    bool trav_round0 = stochastic_pipeline && p->indirect;
    for (int k = 0; k < h.n_spheres; k++) trav_round0 = trav_round0 || h.spheres[k].mirror || h.spheres[k].n_in != h.spheres[k].n_out;
    trav_round0 = trav_round0 && segments >= 2;
    /* node pool of wf_traverse: 32 lanes x (levels of the packed tree + roots of two batches + slack) */
    int npool_cap = wide ? std::min(96 * (h.wide_depth + 2), 352) : std::min(32 * (h.max_depth + 4), 256);
    if (s->opt.npool_cap > 0) npool_cap = std::max(64, std::min(s->opt.npool_cap, npool_cap)) & ~31; /* test hook: a small pool forces the spill path */
    const size_t warp_bytes = (sizeof(rtk::WfWarpSmem) + (size_t)npool_cap * sizeof(int) + 15) & ~(size_t)15;
    const size_t trav_smem = warp_bytes * (WF_THREADS / 32);
    if (trav_smem > 200 * 1024) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_render: BVH depth %d needs %zu B of shared memory per block", h.max_depth, trav_smem);
    if (s->trav_blocks_per_sm == 0 || s->trav_smem != trav_smem || s->trav_wide != wide) {
        s->trav_wide = wide;
        int nb = 0;
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<false, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaFuncSetAttribute(rtk::wf_traverse<false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trav_smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, wide ? rtk::wf_traverse<false, false, true> : rtk::wf_traverse<false, false, false>, WF_THREADS, trav_smem));
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, s->device));
        s->trav_blocks_per_sm = std::max(nb, 1);
        s->trav_smem = trav_smem;
        s->sm_count = prop.multiProcessorCount;
        if (!s->wf_counters) {
            CUDA_TRY(cudaMalloc(&s->wf_counters, RT_MAX_STRIPS * sizeof(rtk::WfCounters)));
            CUDA_TRY(cudaMallocHost(&s->h_wf_counters, RT_MAX_STRIPS * sizeof(rtk::WfCounters)));
            CUDA_TRY(cudaEventCreateWithFlags(&s->fork_ev, cudaEventDisableTiming));
            for (int k = 0; k < RT_MAX_STRIPS; k++) {
                CUDA_TRY(cudaStreamCreateWithFlags(&s->strip_stream[k], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&s->strip_done[k], cudaEventDisableTiming));
                CUDA_TRY(cudaStreamCreateWithFlags(&s->side_stream[k], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&s->side_fork[k], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&s->side_join[k], cudaEventDisableTiming));
            }
        }
    }
    if (s->wf_capacity < npx) {
        if (s->wf_queue) cudaFree(s->wf_queue);
        s->wf_queue = nullptr;
        s->wf_capacity = 0;
        CUDA_TRY(cudaMalloc(&s->wf_queue, 3 * npx * sizeof(rtk::QEntry)));
        s->wf_capacity = npx;
    }
    if (anchored) {
        const size_t need = (size_t)s->task_factor * npx + task_slack * RT_MAX_STRIPS;
        if (s->wf_tasks_cap < need) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (s->wf_tasks) cudaFree(s->wf_tasks);
            s->wf_tasks = nullptr;
            s->wf_tasks_cap = 0;
            CUDA_TRY(cudaMalloc(&s->wf_tasks, need * sizeof(int2)));
            s->wf_tasks_cap = need;
        }
    }
    const bool dbg_times = s->opt.debug_times != 0;
    const bool dbg_warps = count && s->opt.debug_warps[0];
    /* Strips: the frame is cut into bands of rows, each rendered by its own generate / traverse / shade chain
     * on its own stream. The end of a persistent traversal launch is a latency-bound tail (a few warps
     * finishing their expensive rays on an otherwise idle GPU, profiles/r01_notes.md); with strips the tail of
     * one band is covered by the bulk of the next, and only the last launch's tail is exposed. */
    const bool strips_fixed = s->opt.strips > 0;
    int n_strips = strips_fixed ? std::min(s->opt.strips, RT_MAX_STRIPS) : 2;
    {   /* host outputs: more bands, so that only the last band's copy-back is not covered by rendering */
        bool any_copy = false;
        for (int k = 0; k < 5; k++) any_copy = any_copy || copy_back[k];
        if (any_copy && !strips_fixed) n_strips = std::min(RT_MAX_STRIPS, 4);
    }
    if (rows < 64 * n_strips) n_strips = std::max(1, rows / 64);
    /* a small shard (one rank's rows of a frame split over 8 GPUs) is bound by launch latencies: one band
     * (measured: 4K depth-4 frame, 1/8 of the rows: 0.33 ms against 0.37 ms with two) */
    if (!strips_fixed && npx < 1500000 && !(flags & RT_RENDER_COUNT_WORK)) {
        bool any_copy = false;
        for (int k = 0; k < 5; k++) any_copy = any_copy || copy_back[k];
        if (!any_copy) n_strips = 1;
    }
    if (dbg_times || dbg_warps) n_strips = 1;
    const unsigned pers_grid = (unsigned)(s->sm_count * s->trav_blocks_per_sm);
    const int spill_cap = WF_SLOTS * std::max(h.max_depth + 2, (RT_WIDE - 1) * h.wide_depth + 2); /* per traversal warp: every ray slot holding a full path of pending siblings */
    {
        const size_t ints = (size_t)spill_cap * pers_grid * (WF_THREADS / 32) * n_strips;
        if (s->wf_spill_ints < ints) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (s->wf_spill) cudaFree(s->wf_spill);
            s->wf_spill = nullptr;
            s->wf_spill_ints = 0;
            CUDA_TRY(cudaMalloc(&s->wf_spill, ints * sizeof(int)));
            s->wf_spill_ints = ints;
        }
    }
    int* dbg_ptr = nullptr;
    if (dbg_warps) {
        const size_t ints = (size_t)(segments + 1) * pers_grid * (WF_THREADS / 32) * 16;
        if (s->dbg_warps_ints < ints) {
            if (s->dbg_warps) cudaFree(s->dbg_warps);
            CUDA_TRY(cudaMalloc(&s->dbg_warps, ints * sizeof(int)));
            s->dbg_warps_ints = ints;
        }
        P.dbg_ints = ints; /* cleared at the head of the frame (enqueue_frame) */
        dbg_ptr = s->dbg_warps;
    }
    /* stochastic mode: per compact pixel 32 B of stream state, 16 B of colour sum, 32 B per path segment of records */
    const size_t st_rng_off = 0, st_total_off = npx * 32, st_rec_off = st_total_off + npx * 16;
    if (stochastic_pipeline) {
        const size_t need = st_rec_off + npx * 32 * (size_t)std::max(segments, 1);
        if (s->st_buf_bytes < need) {
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            if (s->st_buf) cudaFree(s->st_buf);
            s->st_buf = nullptr;
            s->st_buf_bytes = 0;
            CUDA_TRY(cudaMalloc(&s->st_buf, need));
            s->st_buf_bytes = need;
        }
    }
        /* the table of the top levels, rebuilt when the tree changes (one kernel of one thread + a 4-byte read-back) */
        bool tops = s->opt.top_smem != 0 && h.has_mesh && h.root_ref >= 0 && !wide && !count && (long long)h.n_inner + WF_TOP_MAX < (1ll << 26);
        if (tops && s->top_generation != s->mesh_generation) {
            if (!s->top_nodes) {
                CUDA_TRY(cudaMalloc(&s->top_nodes, WF_TOP_MAX * 4 * sizeof(float4)));
                CUDA_TRY(cudaMalloc(&s->d_n_top, sizeof(int)));
            }
            rtk::build_top_table<<<1, 32, 0, s->stream>>>(reinterpret_cast<const float4*>(s->blob + h.off_nodes), h.n_inner, h.root_ref, WF_TOP_MAX, s->top_nodes, s->d_n_top);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(&s->n_top, s->d_n_top, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
            CUDA_TRY(cudaStreamSynchronize(s->stream));
            s->top_generation = s->mesh_generation;
        }
        tops = tops && s->n_top > 0;
        P.tops = tops;
        P.n_top = tops ? s->n_top : 0;
        P.wide = wide;
        P.anchored = anchored;
        P.diffuse_only = diffuse_only;
        P.trav_round0 = trav_round0;
        P.dbg_times = dbg_times;
        P.npool_cap = npool_cap;
        P.n_strips = n_strips;
        P.spill_cap = spill_cap;
        P.trav_smem = trav_smem;
        P.pers_grid = pers_grid;
        P.dbg_ptr = dbg_ptr;
        P.st_rng_off = st_rng_off;
        P.st_total_off = st_total_off;
        P.st_rec_off = st_rec_off;
    }
    P.p = *p;
    P.flags = flags;
    P.rows = rows;
    P.npx = npx;
    for (int k = 0; k < 5; k++) {
        P.user[k] = user[k];
        P.dev[k] = dev[k];
        P.copy_back[k] = copy_back[k];
        P.bytes[k] = bytes[k];
    }
    P.a = a;
    P.variant = variant;
    P.grid = grid;
    P.stochastic = stochastic_pipeline;
    P.count = count;
    P.segments = segments;
    return RT_OK;
}

/* Kernel launches of a frame with programmatic dependent launch (PDL) between consecutive kernels of one stream: the next kernel's
 * blocks are placed while the previous kernel's last blocks still run and wait in griddepcontrol.wait (first statement of every wf_*
 * kernel) until it has completed — the launch and ramp-up of the 8-30 small kernels of a frame overlap the tails of their predecessors.
 * Only a kernel whose immediate predecessor on its stream is a kernel gets the attribute; any other operation (memset, event, copy)
 * breaks the chain. Option "pdl" = 1 turns it on (default off: the graph replay already removes the launch gaps, measured). */
struct LaunchChain {
    cudaStream_t st[2 * RT_MAX_STRIPS + 2];
    bool after_kernel[2 * RT_MAX_STRIPS + 2];
    int n = 0;
    bool enabled = true;
    bool& of(cudaStream_t x) {
        for (int k = 0; k < n; k++)
            if (st[k] == x) return after_kernel[k];
        st[n] = x;
        after_kernel[n] = false;
        return after_kernel[n++];
    }
    void broke(cudaStream_t x) { of(x) = false; }
    template <typename... KArgs, typename... Args>
    cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t stream, Args&&... args) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid;
        cfg.blockDim = dim3(block);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[2];
        int na = 0;
        bool& prev = of(stream);
        if (enabled && prev) {
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na++].val.programmaticStreamSerializationAllowed = 1;
        }
        cfg.attrs = at;
        cfg.numAttrs = na;
        prev = true;
        return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    }
};

/* ---- rt_render, second half: the launches of one frame, nothing else (no allocation, no synchronisation, no host read-back):
 * the sequence is the same for every frame with the same plan, so it can be recorded once into a CUDA graph and replayed. */
int enqueue_frame(rt_scene* s, const FramePlan& P, int& launches, bool& strip_copied) {
    const rt_params* p = &P.p;
    const SceneHeader& h = s->header;
    const rtk::RenderArgs& a = P.a;
    const bool stochastic = P.stochastic, count = P.count, wide = P.wide, anchored = P.anchored, diffuse_only = P.diffuse_only, trav_round0 = P.trav_round0,
               dbg_times = P.dbg_times;
    const bool six = s->header.n_spheres == 6 && s->opt.six != 0;
#ifdef RT_TIMELINE
    if (!s->tl_buf) CUDA_TRY(cudaMalloc(&s->tl_buf, 128 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(s->tl_buf, 0xff, 64 * sizeof(unsigned long long), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->tl_buf + 64, 0, 64 * sizeof(unsigned long long), s->stream));
    s->tl_n = 0;
#define TL_MARK(name)                                                       \
    do {                                                                    \
        g.tl_min = s->tl_buf;                                               \
        g.tl_max = s->tl_buf + 64;                                          \
        g.tl_slot = s->tl_n < 64 ? s->tl_n : 63;                            \
        if (s->tl_n < 64) {                                                 \
            s->tl_name[s->tl_n] = name;                                     \
            s->tl_band[s->tl_n++] = st;                                     \
        }                                                                   \
    } while (0)
#else
#define TL_MARK(name)
#endif
    const int segments = P.segments, npool_cap = P.npool_cap, n_strips = P.n_strips, spill_cap = P.spill_cap, rows = P.rows, variant = P.variant;
    const size_t trav_smem = P.trav_smem, st_rng_off = P.st_rng_off, st_total_off = P.st_total_off, st_rec_off = P.st_rec_off, task_slack = P.task_slack;
    const unsigned pers_grid = P.pers_grid, grid = P.grid;
    int* const dbg_ptr = P.dbg_ptr;
    void* const* user = P.user;
    void* const* dev = P.dev;
    const bool* copy_back = P.copy_back;
    LaunchChain chain;
    chain.enabled = s->opt.pdl != 0 && !P.count && !P.dbg_times;
    if (variant == 2) {
        if (P.dbg_ints) CUDA_TRY(cudaMemsetAsync(s->dbg_warps, 0, P.dbg_ints * sizeof(int), s->stream));
        /* queue cursors and work statistics restart with every frame; the overflow flags live elsewhere (rt_scene::sticky) */
        CUDA_TRY(cudaMemsetAsync(s->wf_counters, 0, RT_MAX_STRIPS * sizeof(rtk::WfCounters), s->stream));
    if (n_strips > 1) CUDA_TRY(cudaEventRecord(s->fork_ev, s->stream));
    cudaEvent_t dev_ev[40];
    int n_ev = 0;
    auto mark = [&]() {
        if (dbg_times && n_ev < 40) {
            cudaEventCreate(&dev_ev[n_ev]);
            cudaEventRecord(dev_ev[n_ev++], s->stream);
        }
    };
    mark();
    int row0 = 0;
    for (int st = 0; st < n_strips; st++) {
        /* band boundaries on multiples of 8 rows (generate tiles are 8x4) */
        int row1 = (st == n_strips - 1) ? rows : (int)(((long long)rows * (st + 1) / n_strips) & ~7ll);
        if (n_strips == 2 && st == 0 && s->opt.split > 0) row1 = (int)(((long long)rows * s->opt.split / 100) & ~7ll);
        if (row1 <= row0) continue;
        const int srows = row1 - row0;
        const size_t spx = (size_t)srows * p->W, px0 = (size_t)row0 * p->W;
        cudaStream_t stream = n_strips > 1 ? s->strip_stream[st] : s->stream;
        if (n_strips > 1) CUDA_TRY(cudaStreamWaitEvent(stream, s->fork_ev, 0));
        chain.broke(stream);
        rtk::WfArgs g;
#ifdef RT_TIMELINE
        g.tl_min = g.tl_max = nullptr;
        g.tl_slot = 0;
#endif
        g.a = a;
        g.a.rows = srows;
        g.a.row0 = row0; /* the band's first compact row: image rows follow from row_begin / row_step / row_group (image_row) */
        g.a.rgb = a.rgb ? a.rgb + px0 * 3 : nullptr;
        g.a.hit_obj = a.hit_obj ? a.hit_obj + px0 : nullptr;
        g.a.hit_tri = a.hit_tri ? a.hit_tri + px0 : nullptr;
        g.a.hit_t = a.hit_t ? a.hit_t + px0 : nullptr;
        g.a.shadow = a.shadow ? a.shadow + px0 : nullptr;
        g.a.linear = a.linear ? a.linear + px0 : nullptr;
        g.qA[0] = s->wf_queue + px0;
        g.qA[1] = s->wf_queue + s->wf_capacity + px0;
        g.qS = s->wf_queue + 2 * s->wf_capacity + px0;
        g.c = s->wf_counters + st;
        g.round = 0;
        g.spill = s->wf_spill + (size_t)st * spill_cap * pers_grid * (WF_THREADS / 32);
        g.spill_cap = spill_cap;
        {
            g.run_shift = s->opt.run_shift;
            g.gss_factor = s->opt.gss;
            g.fair_share = s->opt.fair_share;
            g.top = P.tops ? s->top_nodes : nullptr;
            g.n_top = P.n_top;
        }
        g.dbg_warps = dbg_ptr;
        g.anchored = anchored ? 1 : 0;
        g.bins[0] = bins_view(s->bins[0]);
        g.bins[1] = bins_view(s->bins[1]);
        g.tasks = anchored ? s->wf_tasks + (size_t)s->task_factor * px0 + task_slack * st : nullptr;
        g.qcap = (int)spx;
        for (int k = 0; k < RT_MAX_SPHERES; k++) {
            const DevSphere& sp = h.spheres[k];
            /* the operations of Sphere::intersect / sphere_t, in their order, in float (no contraction on the host) */
            volatile float ocx = a.camx - sp.cx, ocy = a.camy - sp.cy, ocz = a.camz - sp.cz;
            volatile float xx = ocx * ocx, yy = ocy * ocy, zz = ocz * ocz;
            volatile float n2 = xx + yy;
            n2 = n2 + zz;
            volatile float cc = n2 - sp.RR;
            g.cam_sph[k] = k < h.n_spheres ? make_float4(ocx, ocy, ocz, cc) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        g.task_cap = anchored ? (int)std::min<size_t>((size_t)s->task_factor * spx + task_slack, (size_t)0x7fffffff) : 0;
        if (anchored && s->leaves_blocks_per_sm == 0) { /* one resident wave of wf_leaves: every block gets the same share of the tasks */
            int nb = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rtk::wf_leaves<false>, WF_THREADS, 0));
            s->leaves_blocks_per_sm = std::max(nb, 1);
        }
        const int env_lb = s->opt.leaves_blocks;
        const unsigned leaves_grid = (unsigned)(s->sm_count * (env_lb > 0 ? env_lb : std::max(s->leaves_blocks_per_sm, 1)));
        const dim3 gen_grid((unsigned)(((p->W + 7) / 8 + (WF_THREADS / 32) - 1) / (WF_THREADS / 32)), (unsigned)((srows + 3) / 4));
        const unsigned shade_grid = (unsigned)std::min<size_t>((spx + WF_THREADS - 1) / WF_THREADS, (size_t)s->sm_count * 16);
        g.stoch = stochastic ? 1 : 0;
        g.sample = 0;
        g.last_sample = 1;
        g.indirect = p->indirect;
        g.libm = s->opt.transcendentals;
        g.sticky = s->sticky;
        g.aa_sigma = p->aa_sigma;
        g.npx = (int)spx;
        g.rng_table = reinterpret_cast<const uint4*>(s->rng_states);
        g.jitter_tab = s->jitter_tab;
        g.rng = stochastic ? reinterpret_cast<uint4*>(s->st_buf + st_rng_off) + px0 * 2 : nullptr;
        g.total = stochastic ? reinterpret_cast<float4*>(s->st_buf + st_total_off) + px0 : nullptr;
        g.rec = stochastic ? reinterpret_cast<float4*>(s->st_buf + st_rec_off) + px0 * 2 * (size_t)std::max(segments, 1) : nullptr;
        /* deterministic mode: one pass (identical samples are traced once). Stochastic mode: one pass per sample,
         * in order, because the samples of a pixel share one random stream. */
        const int n_pass = stochastic ? p->num_rays : 1;
        for (int pass = 0; pass < n_pass; pass++) {
            g.sample = pass;
            g.last_sample = pass == n_pass - 1;
            g.round = 0;
            if (pass > 0) /* the queue counters restart with every pass; the work statistics keep adding up */
                CUDA_TRY(cudaMemsetAsync(reinterpret_cast<unsigned char*>(s->wf_counters + st) + offsetof(rtk::WfCounters, nA), 0,
                                         sizeof(rtk::WfCounters) - offsetof(rtk::WfCounters, nA), stream));
                chain.broke(stream);
            TL_MARK("generate");
            if (stochastic) {
                if (count) CUDA_TRY(chain.launch(rtk::wf_generate<true, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else if (diffuse_only) CUDA_TRY(chain.launch(rtk::wf_generate<false, true, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else CUDA_TRY(chain.launch(rtk::wf_generate<false, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
            } else if (P.jitter) {
                if (diffuse_only && anchored && six) CUDA_TRY(chain.launch(rtk::wf_generate<false, false, true, true, true, 6>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else if (diffuse_only && anchored) CUDA_TRY(chain.launch(rtk::wf_generate<false, false, true, true, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else if (diffuse_only) CUDA_TRY(chain.launch(rtk::wf_generate<false, false, true, false, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else CUDA_TRY(chain.launch(rtk::wf_generate<false, false, false, false, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
            } else {
                if (count) CUDA_TRY(chain.launch(rtk::wf_generate<true, false>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else if (diffuse_only && anchored && six) CUDA_TRY(chain.launch(rtk::wf_generate<false, false, true, true, false, 6>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else if (diffuse_only && anchored) CUDA_TRY(chain.launch(rtk::wf_generate<false, false, true, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else if (diffuse_only) CUDA_TRY(chain.launch(rtk::wf_generate<false, false, true>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                else CUDA_TRY(chain.launch(rtk::wf_generate<false, false>, dim3(gen_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
            }
            launches++;
            mark();
            /* without a mesh no query is ever posted: wf_generate runs every path to its end */
            for (int r = 0; r <= segments && segments > 0 && h.has_mesh; r++) {
                g.round = r;
                if (anchored) {
                    /* closest-hit queries of a round >= 1 start somewhere in the scene: tree search; every shadow
                     * query and the camera rays of round 0 are (ray, leaf) tasks */
                    const bool trav_now = r < segments && (r >= 1 || trav_round0);
                    /* the two kernels of the round touch disjoint entries: the tree search goes to the band's side stream
                     * and runs beside wf_leaves (its latency-bound tail is covered by the LSU-bound task kernel) */
                    const bool env_side = s->opt.side_stream != 0;
                    const bool side = trav_now && env_side && !dbg_times;
                    cudaStream_t tstream = side ? s->side_stream[st] : stream;
                    if (side) {
                        CUDA_TRY(cudaEventRecord(s->side_fork[st], stream));
                        chain.broke(stream);
                        CUDA_TRY(cudaStreamWaitEvent(tstream, s->side_fork[st], 0));
                        chain.broke(tstream);
                    }
                    if (trav_now) {
                        TL_MARK("traverse");
                        if (stochastic) {
                            if (wide) CUDA_TRY(chain.launch(rtk::wf_traverse<false, true, true>, dim3(pers_grid), WF_THREADS, trav_smem, tstream, s->header, s->blob, g, npool_cap));
                            else if (P.tops) CUDA_TRY(chain.launch(rtk::wf_traverse<false, true, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, tstream, s->header, s->blob, g, npool_cap));
                            else CUDA_TRY(chain.launch(rtk::wf_traverse<false, true, false>, dim3(pers_grid), WF_THREADS, trav_smem, tstream, s->header, s->blob, g, npool_cap));
                        } else {
                            if (wide) CUDA_TRY(chain.launch(rtk::wf_traverse<false, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, tstream, s->header, s->blob, g, npool_cap));
                            else if (P.tops) CUDA_TRY(chain.launch(rtk::wf_traverse<false, false, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, tstream, s->header, s->blob, g, npool_cap));
                            else CUDA_TRY(chain.launch(rtk::wf_traverse<false, false, false>, dim3(pers_grid), WF_THREADS, trav_smem, tstream, s->header, s->blob, g, npool_cap));
                        }
                        launches++;
                        mark();
                    }
                    TL_MARK("leaves");
                    if (stochastic) CUDA_TRY(chain.launch(rtk::wf_leaves<true>, dim3(leaves_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else CUDA_TRY(chain.launch(rtk::wf_leaves<false>, dim3(leaves_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    if (side) { /* join: wf_shade needs both */
                        CUDA_TRY(cudaEventRecord(s->side_join[st], tstream));
                        chain.broke(tstream);
                        CUDA_TRY(cudaStreamWaitEvent(stream, s->side_join[st], 0));
                        chain.broke(stream);
                    }
                } else if (stochastic) {
                    if (count) CUDA_TRY(chain.launch(rtk::wf_traverse<true, true, false>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else if (wide) CUDA_TRY(chain.launch(rtk::wf_traverse<false, true, true>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else if (P.tops) CUDA_TRY(chain.launch(rtk::wf_traverse<false, true, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else CUDA_TRY(chain.launch(rtk::wf_traverse<false, true, false>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                } else {
                    if (count && wide) CUDA_TRY(chain.launch(rtk::wf_traverse<true, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else if (count) CUDA_TRY(chain.launch(rtk::wf_traverse<true, false, false>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else if (wide) CUDA_TRY(chain.launch(rtk::wf_traverse<false, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else if (P.tops) CUDA_TRY(chain.launch(rtk::wf_traverse<false, false, false, true>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                    else CUDA_TRY(chain.launch(rtk::wf_traverse<false, false, false>, dim3(pers_grid), WF_THREADS, trav_smem, stream, s->header, s->blob, g, npool_cap));
                }
                launches++;
                mark();
                if (r == segments) break;
                TL_MARK("shade");
                if (stochastic) {
                    if (count) CUDA_TRY(chain.launch(rtk::wf_shade<true, true>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else if (diffuse_only) CUDA_TRY(chain.launch(rtk::wf_shade<false, true, true>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else CUDA_TRY(chain.launch(rtk::wf_shade<false, true>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                } else {
                    if (count) CUDA_TRY(chain.launch(rtk::wf_shade<true, false>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else if (diffuse_only && anchored && six) CUDA_TRY(chain.launch(rtk::wf_shade<false, false, true, true, 6>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else if (diffuse_only && anchored) CUDA_TRY(chain.launch(rtk::wf_shade<false, false, true, true>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else if (diffuse_only) CUDA_TRY(chain.launch(rtk::wf_shade<false, false, true>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                    else CUDA_TRY(chain.launch(rtk::wf_shade<false, false>, dim3(shade_grid), WF_THREADS, 0, stream, s->header, s->blob, g));
                }
                launches++;
                mark();
            }
            if (stochastic) {
                CUDA_TRY(chain.launch(rtk::wf_fold, dim3((unsigned)((spx + 255) / 256)), 256, 0, stream, g));
                launches++;
                mark();
            }
        }
        if (n_strips > 1) {
            /* host outputs: this band's copy-back rides on the band's stream and overlaps the other bands' kernels
             * (kernel_ms then spans the copies of all bands but the last as well) */
            const size_t elem[5] = {3, 4, 4, 4, 1};
            for (int k = 0; k < 5; k++)
                if (copy_back[k] && !P.async_copy && !(a.linear && k == 0)) { /* accumulation: the 8-bit frame exists only after accumulate_frame */
                    CUDA_TRY(cudaMemcpyAsync((unsigned char*)user[k] + px0 * elem[k], (unsigned char*)dev[k] + px0 * elem[k], spx * elem[k], cudaMemcpyDeviceToHost, stream));
                    strip_copied = true;
                }
            CUDA_TRY(cudaEventRecord(s->strip_done[st], stream));
            CUDA_TRY(cudaStreamWaitEvent(s->stream, s->strip_done[st], 0));
        }
        row0 = row1;
    }
    if (dbg_times) {
        cudaStreamSynchronize(s->stream);
        fprintf(stderr, "[times us]");
        for (int k = 1; k < n_ev; k++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, dev_ev[k - 1], dev_ev[k]);
            fprintf(stderr, " %.1f", ms * 1e3f);
        }
        fprintf(stderr, "\n");
        for (int k = 0; k < n_ev; k++) cudaEventDestroy(dev_ev[k]);
    }
        if (a.linear) { /* progressive accumulation (realtime_render.cu:1136-1140) once every band has joined */
            rtk::accumulate_frame<<<(unsigned)((P.npx + 255) / 256), 256, 0, s->stream>>>(s->accum, s->linear, (int)P.npx, p->accumulate, a.rgb, a.gamma_tab, a.gamma_mode);
            launches++;
            if (copy_back[0] && strip_copied && !P.async_copy) CUDA_TRY(cudaMemcpyAsync(user[0], dev[0], P.bytes[0], cudaMemcpyDeviceToHost, s->stream));
        }
        launches--; /* the common launches++ below counts one */
    } else if (variant == 3) {
        CUDA_TRY(cudaMemsetAsync(s->counters, 0, RT_NCOUNTERS * sizeof(unsigned long long), s->stream));
        if (count) rtk::render_stoch<true><<<grid, 128, 0, s->stream>>>(s->header, s->blob, a, s->rng_states, p->aa_sigma, p->indirect, s->opt.transcendentals);
        else rtk::render_stoch<false><<<grid, 128, 0, s->stream>>>(s->header, s->blob, a, s->rng_states, p->aa_sigma, p->indirect, s->opt.transcendentals);
    } else if (variant == 0) {
        CUDA_TRY(cudaMemsetAsync(s->counters, 0, RT_NCOUNTERS * sizeof(unsigned long long), s->stream));
        if (count) rtk::render_mega<true, false><<<grid, 128, 0, s->stream>>>(s->header, s->blob, a);
        else rtk::render_mega<false, false><<<grid, 128, 0, s->stream>>>(s->header, s->blob, a);
    } else {
        CUDA_TRY(cudaMemsetAsync(s->counters, 0, RT_NCOUNTERS * sizeof(unsigned long long), s->stream));
        if (count) rtk::render_mega<true, true><<<grid, 128, 0, s->stream>>>(s->header, s->blob, a);
        else rtk::render_mega<false, true><<<grid, 128, 0, s->stream>>>(s->header, s->blob, a);
    }
    launches++;
    CUDA_TRY(cudaGetLastError());
    for (int k = 0; k < 5; k++) /* host outputs of a frame rendered as one band (or by the one-kernel variants) */
        if (copy_back[k] && !strip_copied && !P.async_copy) CUDA_TRY(cudaMemcpyAsync(user[k], dev[k], P.bytes[k], cudaMemcpyDeviceToHost, s->stream));
    return RT_OK;
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

/* Everything a recorded frame depends on, as bytes: the plan (parameters, output pointers, kernel choices), the scene header the
 * kernels take by value, the bins, the buffers and the options. Equal keys <=> the same launches with the same arguments. */
void frame_key(rt_scene* s, const FramePlan& P, std::vector<unsigned char>& key) {
    key.clear();
    auto put = [&](const void* p, size_t n) { key.insert(key.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
    if (P.async_copy) { /* host outputs that leave on the copy stream are not part of the recorded launches: any host buffer may follow a replay */
        FramePlan Q = P;
        for (int k = 0; k < 5; k++)
            if (Q.copy_back[k]) Q.user[k] = nullptr;
        put(&Q, sizeof Q);
    } else {
        put(&P, sizeof P);
    }
    put(&s->header, sizeof s->header);
    for (int k = 0; k < 2; k++) {
        rtk::BinsView v;
        memset(&v, 0, sizeof v);
        if (P.anchored) v = bins_view(s->bins[k]);
        put(&v, sizeof v);
    }
    const void* ptrs[] = {s->top_nodes, s->blob, s->jitter_tab, s->wf_queue, s->wf_tasks, s->wf_counters, s->wf_spill, s->st_buf, s->rng_states, s->gamma_tab, s->sticky, s->stream};
    put(ptrs, sizeof ptrs);
    const size_t nums[] = {s->wf_capacity, s->wf_tasks_cap, (size_t)s->task_factor, (size_t)s->leaves_blocks_per_sm, (size_t)s->sm_count, (size_t)s->trav_blocks_per_sm};
    put(nums, sizeof nums);
    put(&s->opt, sizeof s->opt);
}

} // namespace

extern "C" {

/* The render call. plan_frame decides and sizes, enqueue_frame launches. A frame whose plan equals the previous call's is
 * recorded into a CUDA graph the second time it is seen and replayed from then on (one graph launch instead of 8-30 kernel
 * launches + stream events: what a caller that renders the same view again and again pays per frame); any change — parameters,
 * output pointers, light, mesh, options — goes back to direct launches. Option "graph" = 0 turns the replay off. */
int rt_render(rt_scene* s, const rt_params* p, uint32_t flags, uint8_t* rgb_out, int32_t* hit_obj, int32_t* hit_tri, float* hit_t,
              uint8_t* shadow, rt_stats* stats) {
    if (!s || !p) return rtb::fail(RT_ERR_INVALID, "rt_render: NULL scene or params");
    DeviceGuard g(s->device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_render: cudaSetDevice failed");
    void* user[5] = {rgb_out, hit_obj, hit_tri, hit_t, shadow};
    FramePlan P;
    memset(&P, 0, sizeof P);
    int rc = plan_frame(s, p, flags, user, P);
    if (rc != RT_OK) return rc;
    int launches = P.prelaunches;
    bool strip_copied = false;

    bool graph_ok = s->opt.graph != 0 && P.variant == 2 && !P.count && !P.dbg_times;
    for (int k = 0; k < 5 && graph_ok; k++)
        if (P.copy_back[k] && !P.async_copy && !is_pinned_host(P.user[k])) graph_ok = false; /* pageable host memory: the copy is not a pure stream operation */
    std::vector<unsigned char> key;
    if (graph_ok) frame_key(s, P, key);
    const int gs = P.async_copy ? P.scratch_set : 0;

    if (P.async_copy) {
        if (!s->copy_stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&s->copy_done[0], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->copy_done[1], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&s->frame_done, cudaEventDisableTiming));
        }
    }
    /* this frame's kernels write scratch set P.scratch_set: the copy-back of the frame that used it last must have left it */
    if (s->copy_recorded[P.scratch_set]) {
        bool uses_scratch = false;
        for (int k = 0; k < 5; k++) uses_scratch = uses_scratch || P.copy_back[k];
        if (uses_scratch) CUDA_TRY(cudaStreamWaitEvent(s->stream, s->copy_done[P.scratch_set], 0));
    }
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream)); /* kernel_ms covers everything the frame enqueues, counter resets and copies to host outputs included */
    if (graph_ok && s->graph_exec[gs] && key == s->graph_key[gs]) {
        CUDA_TRY(cudaGraphLaunch(s->graph_exec[gs], s->stream));
        launches += s->graph_launches[gs];
    } else if (graph_ok && key == s->last_key[gs]) {
        /* second identical frame in a row: record it */
        CUDA_TRY(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        int n = 0;
        rc = enqueue_frame(s, P, n, strip_copied);
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc != RT_OK || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            s->last_key[gs].clear();
            return rc != RT_OK ? rc : rtb::fail(RT_ERR_CUDA, "rt_render: cudaStreamEndCapture: %s", cudaGetErrorString(e));
        }
        if (s->graph_exec[gs]) cudaGraphExecDestroy(s->graph_exec[gs]);
        s->graph_exec[gs] = nullptr;
        const cudaError_t e2 = cudaGraphInstantiate(&s->graph_exec[gs], graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            s->graph_exec[gs] = nullptr;
            s->last_key[gs].clear();
            return rtb::fail(RT_ERR_CUDA, "rt_render: cudaGraphInstantiate: %s", cudaGetErrorString(e2));
        }
        s->graph_key[gs] = key;
        s->graph_launches[gs] = n;
        CUDA_TRY(cudaEventRecord(s->ev0, s->stream)); /* recording and instantiating are one-off costs, not frame time */
        CUDA_TRY(cudaGraphLaunch(s->graph_exec[gs], s->stream));
        launches += n;
    } else {
        rc = enqueue_frame(s, P, launches, strip_copied);
        if (rc != RT_OK) return rc;
    }
    s->last_key[gs].swap(key);
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    if (P.async_copy) { /* the frame's host outputs: behind its kernels, beside the next frame's */
        CUDA_TRY(cudaEventRecord(s->frame_done, s->stream));
        CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->frame_done, 0));
        for (int k = 0; k < 5; k++)
            if (P.copy_back[k]) CUDA_TRY(cudaMemcpyAsync(P.user[k], P.dev[k], P.bytes[k], cudaMemcpyDeviceToHost, s->copy_stream));
        CUDA_TRY(cudaEventRecord(s->copy_done[P.scratch_set], s->copy_stream));
        s->copy_recorded[P.scratch_set] = true;
        s->copies_pending = true;
    }
    s->pending = true; /* the counters are read back by rt_scene_sync, not per enqueued frame */
    s->pending_launches = launches;
    if (flags & RT_RENDER_NO_SYNC) {
        if (stats) memset(stats, 0, sizeof *stats);
        return RT_OK;
    }
    rc = rt_scene_sync(s, stats);
    if (rc == RT_ERR_AGAIN && !(flags & RT_RENDER_RETRY_)) { /* task buffer overflow: repeat with the doubled buffer (a few times at most) */
        for (int attempt = 0; attempt < 6; attempt++) {
            const int rc2 = rt_render(s, p, flags | RT_RENDER_RETRY_, rgb_out, hit_obj, hit_tri, hit_t, shadow, stats);
            if (rc2 != RT_ERR_AGAIN) return rc2;
        }
    }
    return rc;
}

/* ---- peer frame buffer: the gather of a row-sharded frame as direct NVLink copies -----------------------------------
 * One process per GPU: the destination rank allocates the frame and exports a CUDA IPC handle; the other ranks open it and
 * push their row bands straight into it (strided device-to-device copy over NVLink / NVSwitch on the scene's stream). The
 * only synchronisation left is one barrier that tells the destination all bands have landed. */
int rt_peer_alloc(int device, size_t bytes, void** ptr, uint8_t handle[64]) {
    if (!ptr || !handle || bytes == 0) return rtb::fail(RT_ERR_INVALID, "rt_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_peer_alloc: no device %d", device);
    CUDA_TRY(cudaMalloc(ptr, bytes));
    cudaIpcMemHandle_t hd;
    const cudaError_t e = cudaIpcGetMemHandle(&hd, *ptr);
    if (e != cudaSuccess) {
        cudaFree(*ptr);
        *ptr = nullptr;
        return rtb::fail(RT_ERR_CUDA, "rt_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle, &hd, 64);
    return RT_OK;
}

int rt_peer_open(int device, const uint8_t handle[64], void** ptr) {
    if (!ptr || !handle) return rtb::fail(RT_ERR_INVALID, "rt_peer_open: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_peer_open: no device %d", device);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    return RT_OK;
}

int rt_peer_close(int device, void* ptr) {
    DeviceGuard g(device);
    if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_peer_close failed");
    return RT_OK;
}

int rt_peer_free(int device, void* ptr) {
    DeviceGuard g(device);
    if (ptr) cudaFree(ptr);
    return RT_OK;
}

/* ---- completion flags in peer memory: "my band of frame k has landed" as a stream-ordered 4-byte write into the destination's memory,
 * awaited by a stream-ordered wait on the destination's stream (cuStreamWriteValue32 / cuStreamWaitValue32, bound through the runtime):
 * no kernel, no collective, no host round trip. The write is ordered behind the copy enqueued before it on the same stream. */
namespace {
struct StreamMemOps {
    CUresult (*write32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    CUresult (*wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
    bool ok = false;
};
StreamMemOps& stream_memops() {
    static StreamMemOps m;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult q1, q2;
        void *f1 = nullptr, *f2 = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &f1, cudaEnableDefault, &q1) == cudaSuccess && q1 == cudaDriverEntryPointSuccess &&
            cudaGetDriverEntryPoint("cuStreamWaitValue32", &f2, cudaEnableDefault, &q2) == cudaSuccess && q2 == cudaDriverEntryPointSuccess && f1 && f2) {
            m.write32 = reinterpret_cast<decltype(m.write32)>(f1);
            m.wait32 = reinterpret_cast<decltype(m.wait32)>(f2);
            m.ok = true;
        } else {
            cudaGetLastError();
        }
    }
    return m;
}
} // namespace

int rt_peer_signal(rt_scene* s, void* flag, uint32_t value) {
    if (!s || !flag) return rtb::fail(RT_ERR_INVALID, "rt_peer_signal: bad argument");
    DeviceGuard g(s->device);
    StreamMemOps& m = stream_memops();
    if (!m.ok) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_peer_signal: the driver exposes no stream memory operations");
    const CUresult r = m.write32((CUstream)s->stream, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) return rtb::fail(RT_ERR_CUDA, "rt_peer_signal: cuStreamWriteValue32 failed (%d)", (int)r);
    return RT_OK;
}

int rt_peer_wait(rt_scene* s, const void* flag, uint32_t value) {
    if (!s || !flag) return rtb::fail(RT_ERR_INVALID, "rt_peer_wait: bad argument");
    DeviceGuard g(s->device);
    StreamMemOps& m = stream_memops();
    if (!m.ok) return rtb::fail(RT_ERR_UNSUPPORTED, "rt_peer_wait: the driver exposes no stream memory operations");
    const CUresult r = m.wait32((CUstream)s->stream, (CUdeviceptr)(uintptr_t)flag, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return rtb::fail(RT_ERR_CUDA, "rt_peer_wait: cuStreamWaitValue32 failed (%d)", (int)r);
    return RT_OK;
}

int rt_scene_push_rows(rt_scene* s, const void* band, void* frame, int32_t W, int32_t bytes_per_pixel, int32_t row_begin, int32_t row_step, int32_t rows) {
    return rt_scene_push_row_groups(s, band, frame, W, bytes_per_pixel, row_begin, row_step, 1, rows);
}

int rt_scene_push_row_groups(rt_scene* s, const void* band, void* frame, int32_t W, int32_t bytes_per_pixel, int32_t row_begin, int32_t row_step, int32_t row_group,
                             int32_t rows) {
    if (!s || !band || !frame || W <= 0 || bytes_per_pixel <= 0 || row_begin < 0 || row_step < 1 || rows < 0 || row_group < 1 || row_step < row_group)
        return rtb::fail(RT_ERR_INVALID, "rt_scene_push_row_groups: bad argument");
    if (rows == 0) return RT_OK;
    DeviceGuard g(s->device);
    CUDA_TRY(rtb::scatter_band(frame, band, (size_t)W * bytes_per_pixel, row_begin, row_step, row_group, rows, cudaMemcpyDefault, s->stream));
    return RT_OK;
}

/* optimized.cu:748-749 evaluates `float z = -W / (2 * tan(alpha/2))` INSIDE KernelLaunch: in device code tan(float) is CUDA's
 * tanf, whose result for alpha = pi/3 is one ulp above glibc's (measured on the B200: tan 0.57735032 vs 0.57735026; with that
 * z the reference kernel's t bits, shadow flags and colours are reproduced exactly, with the host value 15 % of the t bits
 * and 0.6 % of the pixels differ). This evaluates the same expression on the device, compiled like an IEEE build of the
 * reference (no fast math, no contraction). cpu_launcher.cpp evaluates it on the host: rt_camera_z. */
__global__ void camera_z_kernel(int W, float alpha, float* out) { *out = -W / (2 * tanf(alpha / 2)); }

int rt_camera_z_device(int device, int32_t W, float alpha, float* z) {
    if (!z || W <= 0) return rtb::fail(RT_ERR_INVALID, "rt_camera_z_device: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_camera_z_device: no device %d", device);
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(float)));
    camera_z_kernel<<<1, 1>>>(W, alpha, d);
    cudaError_t e = cudaMemcpy(z, d, sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_camera_z_device: %s", cudaGetErrorString(e));
    return RT_OK;
}

/* FP32 roof of the device, measured (SURVEY.md 8d asks for an FMA micro-benchmark instead of the nominal SMs x 128 x 2 x clock):
 * every thread runs 8 independent FMA chains; 2 flop per FMA. Best of `reps` launches, CUDA-event timed. */
__global__ void __launch_bounds__(256) fma_peak_kernel(float* __restrict__ out, int iters, float b, float c) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
#pragma unroll 8
    for (int i = 0; i < iters; i++) {
        a0 = __fmaf_rn(a0, b, c);
        a1 = __fmaf_rn(a1, b, c);
        a2 = __fmaf_rn(a2, b, c);
        a3 = __fmaf_rn(a3, b, c);
        a4 = __fmaf_rn(a4, b, c);
        a5 = __fmaf_rn(a5, b, c);
        a6 = __fmaf_rn(a6, b, c);
        a7 = __fmaf_rn(a7, b, c);
    }
    const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678f) out[blockIdx.x * blockDim.x + threadIdx.x] = r; /* keeps the chains alive; practically never true */
}

int rt_selftest_fma_peak(int device, int reps, double* tflops) {
    if (!tflops || reps < 1) return rtb::fail(RT_ERR_INVALID, "rt_selftest_fma_peak: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_selftest_fma_peak: no device %d", device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.;
    cudaError_t e = cudaSuccess;
    for (int r = 0; r < reps + 1 && e == cudaSuccess; r++) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.25f);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms > 0.f) best = std::max(best, 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_selftest_fma_peak: %s", cudaGetErrorString(e));
    *tflops = best;
    return RT_OK;
}

/* Device self-test used by tests/: CUDA's single-precision logf (0) / sinf (1) / cosf (2) / tanf (3) on n arguments, compiled
 * like the rest of this translation unit (no fast math, no contraction): the functions option transcendentals = 1 and
 * rt_camera_z_device evaluate, and that the oracle restates for the CPU (oracle/rt_oracle.cpp: cuda_logf ...). */
__global__ void selftest_libm_kernel(int which, const float* __restrict__ x, int n, float* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[i];
    y[i] = which == 0 ? logf(a) : which == 1 ? sinf(a) : which == 2 ? cosf(a) : tanf(a);
}

int rt_selftest_libm(int device, int which, const float* x, int32_t n, float* y) {
    if (!x || !y || n <= 0 || which < 0 || which > 3) return rtb::fail(RT_ERR_INVALID, "rt_selftest_libm: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_selftest_libm: no device %d", device);
    float *dx = nullptr, *dy = nullptr;
    cudaError_t e = cudaMalloc(&dx, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&dy, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        selftest_libm_kernel<<<(n + 255) / 256, 256>>>(which, dx, n, dy);
        e = cudaMemcpy(y, dy, (size_t)n * 4, cudaMemcpyDeviceToHost);
    }
    if (dx) cudaFree(dx);
    if (dy) cudaFree(dy);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_selftest_libm: %s", cudaGetErrorString(e));
    return RT_OK;
}

/* Device self-test used by tests/: agreement of the reciprocal-based division with div.rn.f32.
 * out[0] mismatches (1 correction step), out[1] mismatches (2 steps), out[2] pairs tested. */
int rt_selftest_division(int device, uint64_t seed, int blocks, int per_thread, uint64_t out[3]) {
    if (!out || blocks <= 0 || per_thread <= 0) return rtb::fail(RT_ERR_INVALID, "rt_selftest_division: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_selftest_division: no device %d", device);
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 3 * sizeof(unsigned long long));
    rtk::selftest_division<<<blocks, 256>>>(seed, per_thread, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long hst[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(hst, d, sizeof hst, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_selftest_division: %s", cudaGetErrorString(e));
    out[0] = hst[0];
    out[1] = hst[1];
    out[2] = hst[2];
    return RT_OK;
}

/* Device self-test used by tests/: div3 / div3_or_zero (one refined reciprocal for three quotients) against three div.rn.f32.
 * out[0] components that differ, out[1] components tested. */
int rt_selftest_division3(int device, uint64_t seed, int blocks, int per_thread, uint64_t out[2]) {
    if (!out || blocks <= 0 || per_thread <= 0) return rtb::fail(RT_ERR_INVALID, "rt_selftest_division3: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_selftest_division3: no device %d", device);
    unsigned long long* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 2 * sizeof(unsigned long long)));
    cudaMemset(d, 0, 2 * sizeof(unsigned long long));
    rtk::selftest_division3<<<blocks, 256>>>(seed, per_thread, d);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long hst[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(hst, d, sizeof hst, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_selftest_division3: %s", cudaGetErrorString(e));
    out[0] = hst[0];
    out[1] = hst[1];
    return RT_OK;
}

/* Device self-test used by tests/: the cuRAND library's own XORWOW start states and first uniforms, the stream the
 * stochastic mode must reproduce (optimized.cu:745, 32-37). states6: n*6 words (d, v0..v4); uniforms4: n*4 floats. */
int rt_selftest_xorwow(int device, uint64_t seed, const uint32_t* subsequences, int32_t n, uint32_t* states6, float* uniforms4) {
    if (!subsequences || n <= 0 || !states6 || !uniforms4) return rtb::fail(RT_ERR_INVALID, "rt_selftest_xorwow: bad argument");
    DeviceGuard g(device);
    if (!g.ok) return rtb::fail(RT_ERR_CUDA, "rt_selftest_xorwow: no device %d", device);
    unsigned int *d_sub = nullptr, *d_st = nullptr;
    float* d_u = nullptr;
    cudaError_t e = cudaMalloc(&d_sub, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_st, (size_t)n * 24);
    if (e == cudaSuccess) e = cudaMalloc(&d_u, (size_t)n * 16);
    if (e == cudaSuccess) e = cudaMemcpy(d_sub, subsequences, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        rtk::selftest_xorwow<<<(n + 127) / 128, 128>>>((unsigned long long)seed, d_sub, n, d_st, d_u);
        e = cudaDeviceSynchronize();
    }
    if (e == cudaSuccess) e = cudaMemcpy(states6, d_st, (size_t)n * 24, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(uniforms4, d_u, (size_t)n * 16, cudaMemcpyDeviceToHost);
    if (d_sub) cudaFree(d_sub);
    if (d_st) cudaFree(d_st);
    if (d_u) cudaFree(d_u);
    if (e != cudaSuccess) return rtb::fail(RT_ERR_CUDA, "rt_selftest_xorwow: %s", cudaGetErrorString(e));
    return RT_OK;
}

} /* extern "C" */
