/*
 * rt_b200.h — C ABI of the B200-native render hot path (drop-in boundary).
 *
 * The reference (souhhcong/RaytracingGPU) has no FFI; its seams for this path
 * are (i) the process CLI `./optimized <num_rays> <num_bounce>`
 * (optimized.cu:774-785) and (ii) the kernel-launch signature
 *   KernelLaunch(char* colors, int W, int H, int num_rays, int num_bounce,
 *                TriangleIndices* indices, int indices_size,
 *                Vector* vertices, int vertices_size, float* arr_bvh)
 * (optimized.cu:670, 828-847) with the host steps around it
 * (optimized.cu:791-826, 848-862). Every entry point below replaces one of
 * those host steps; the citation on each says which.
 *
 * Conventions: plain pointers and sizes, no C++/torch types. Every function
 * returns RT_OK (0) or a negative RT_ERR_* code and never calls exit()
 * (the reference's gpuErrchk does, optimized.cu:24-30); rt_last_error()
 * gives the message of the last failure on the calling thread. The caller
 * owns host buffers, the library owns device buffers. Calls on one rt_scene
 * are not thread-safe; different scenes may be driven from different threads.
 * There is NO CPU fallback: without a CUDA device rt_scene_create fails with
 * RT_ERR_CUDA.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 3 /* 3: rt_params::row_group, rt_scene_push_row_groups */

enum {
    RT_OK = 0,
    RT_ERR_INVALID = -1,     /* bad argument */
    RT_ERR_CUDA = -2,        /* CUDA runtime error or no device */
    RT_ERR_NOMEM = -3,
    RT_ERR_IO = -4,          /* file could not be opened / written */
    RT_ERR_STATE = -5,       /* call order violated (e.g. render before scene upload) */
    RT_ERR_UNSUPPORTED = -6, /* parameter combination not implemented */
    RT_ERR_AGAIN = -7        /* rt_scene_sync after RT_RENDER_NO_SYNC frames: an internal buffer was too small and has been
                              * enlarged; the frame is incomplete, render it again (a synchronous rt_render repeats by itself) */
};

/* ---- scene description (reference: Sphere optimized.cu:117-136, Geometry :103-115) ---- */

/* One analytic sphere. `id` is the object's index in Scene::objects
 * (optimized.cu:687-725 / cpu_launcher.cpp:540-543): intersect_all scans
 * objects in ascending id and the lowest id wins exact-t ties (:549). */
typedef struct rt_sphere {
    float C[3];
    float R;
    float albedo[3];
    int32_t mirror;          /* Geometry::mirror */
    float n_in;              /* Geometry::in_refraction_index  */
    float n_out;             /* Geometry::out_refraction_index */
    int32_t id;
} rt_sphere;

/* Words per record of the reference interchange formats. */
#define RT_TRI_RECORD_WORDS 10  /* TriangleIndices: vtxi,vtxj,vtxk,uvi,uvj,uvk,ni,nj,nk,group (optimized.cu:140-147) */
#define RT_BVH_NODE_FLOATS 10   /* left,right,mn.xyz,mx.xyz,tri_start,tri_end (optimized.cu:512-534, array_bvh.cu:733-759) */

/* Per-program knobs of the reference. */
typedef struct rt_params {
    int32_t W, H;            /* optimized.cu:786-787 (hard-coded 512 there) */
    int32_t num_rays;        /* argv[1], samples per pixel (optimized.cu:785) */
    int32_t num_bounce;      /* argv[2] */
    float cam[3];            /* camera centre C, optimized.cu:747 */
    float z;                 /* -W/(2 tan(alpha/2)), computed on the host once (optimized.cu:749): rt_camera_z */
    float eps_surface;       /* P +- eps*N offset: 1e-4 optimized.cu:575, 1e-3 cpu_launcher.cpp:575 */
    float eps_tri;           /* triangle accept t > eps_tri: 0 optimized.cu:275, 1e-4 cpu_launcher.cpp:301 */
    int32_t push_order;      /* 0: push L then R (R popped first; cpu_launcher.cpp:291-292, array_bvh.cu:282-283)
                                1: push R then L (L popped first; optimized.cu:265-266). Only decides exact-t ties. */
    int32_t extra_segment;   /* 1: num_bounce+1 path segments (recursive getColor, cpu_launcher.cpp:567,710); 0: num_bounce (optimized.cu:566) */
    float aa_sigma;          /* Box-Muller jitter sigma: 0 cpu_launcher.cpp:704, 0.2 optimized.cu:753 */
    int32_t indirect;        /* 1: cosine-weighted random bounce at every diffuse hit (optimized.cu:631-649).
                                aa_sigma != 0 or indirect != 0 selects the stochastic mode: the random stream is the
                                reference's, cuRAND XORWOW curand_init(seed, global pixel index, 0) (optimized.cu:745);
                                rt_params_profile leaves both 0 (deterministic mode). */
    int32_t gamma_mode;      /* 0: trunc(min(pow((double)c, 1./2.2), 255.)) cpu_launcher.cpp:714-716
                                1: trunc(min(powf(c, (float)(1./2.2)), 255.)) optimized.cu:765-767 */
    int32_t row_begin;       /* sharding: this call renders image rows row_begin + k*row_step, k in [0,row_count) (see row_group below) */
    int32_t row_step;        /* 1 for a contiguous band, nranks for row-interleave */
    int32_t row_count;       /* 0 means "all rows from row_begin with row_step" */
    int32_t reserved;        /* stochastic mode: RNG seed, 0 = the reference's 123456 (optimized.cu:745) */
    /* ---- viewer-derived features (realtime_render.cu; the GLUT program itself is out of scope). All 0 = the launchers' behaviour. */
    int32_t camera_mode;     /* 0: fixed camera of the launchers, u_center = (j - W/2 + 0.5, H/2 - i - 0.5, z) (optimized.cu:751)
                                1: the viewer's camera basis, u_center = C + bz*z + bx*(j - W/2 + 0.5) + by*(H/2 - i - 0.5)
                                   (realtime_render.cu:1113 — the camera position IS added to the direction there; replicated) */
    float cam_bx[3], cam_by[3], cam_bz[3]; /* rt_camera_basis(yaw, pitch): Camera::rotate, realtime_render.cu:828-849 */
    int32_t smooth_normals;  /* 1: a mesh hit is shaded with the interpolated vertex normals, N = normalize(alpha Na + beta Nb + gamma Nc)
                                (get_smooth_normal, realtime_render.cu:221-245, called at :311); needs rt_scene_set_mesh_normals */
    int32_t accumulate;      /* progressive accumulation (realtime_render.cu:1136-1140): k >= 1 is the frame number — the frame's linear
                                colour is added to the scene's accumulation buffer (cleared first when k == 1) and the 8-bit frame
                                is quantise(buffer / k); 0: off. Wavefront pipeline only. */
    int32_t row_group;       /* sharding by GROUPS of consecutive rows (a power of two <= 64; 0 or 1: single rows): compact row k of this call
                                is image row row_begin + (k / row_group) * row_step + k % row_group, row_step >= row_group being the distance
                                between the starts of two groups (nranks * row_group for an interleave). Rows of one rank that are
                                neighbours in the image keep a warp's 8x4 pixel tile a tile (the average rank of eight renders its
                                share of a 4K frame 2-4 % faster; large groups unbalance the ranks) */
} rt_params;

typedef struct rt_stats {
    double kernel_ms;        /* CUDA-event time of the render kernels of this call */
    uint64_t rays;           /* intersect_all calls (primary + bounce + shadow) traced, counted on the device. The samples of a
                                pixel are identical in deterministic mode and are traced once (summed num_rays times). */
    uint64_t node_visits;    /* inner BVH nodes popped (two box tests each); 0 unless RT_RENDER_COUNT_WORK */
    uint64_t tri_tests;      /* Moller-Trumbore evaluations; 0 unless RT_RENDER_COUNT_WORK */
    int32_t launches;        /* kernels launched by this call */
    int32_t max_stack;       /* deepest traversal stack seen; 0 unless RT_RENDER_COUNT_WORK */
    uint64_t slab_fallbacks; /* box tests the certified fast path handed to the exact code; 0 unless RT_RENDER_COUNT_WORK */
    uint64_t tri_exact;      /* triangle tests that evaluated the three exact divisions; 0 unless RT_RENDER_COUNT_WORK */
} rt_stats;

/* rt_render flags */
#define RT_RENDER_COUNT_WORK 1u   /* fill node_visits / tri_tests / max_stack (slower instrumented kernel) */
#define RT_RENDER_NO_SYNC    2u   /* enqueue only; kernel_ms and rays are not filled; call rt_scene_sync later. HOST output buffers of
                                    * such a frame are filled on a copy stream behind the frame's kernels (two alternating device staging
                                    * sets: the copy of frame k runs beside the kernels of frame k + 1) and are valid after rt_scene_sync;
                                    * consecutive NO_SYNC frames must therefore be given different host buffers if both are to be read */

typedef struct rt_mesh rt_mesh;     /* host-side TriangleMeshHost replacement */
typedef struct rt_scene rt_scene;   /* device-resident Scene replacement */

/* ---- errors / device ---- */
const char* rt_last_error(void);
int rt_abi_version(void);
int rt_device_count(int* count);                    /* RT_ERR_CUDA when no driver/device */

/* ---- host mesh surface (TriangleMeshHost, optimized.cu:293-535) ---- */
int rt_mesh_create(rt_mesh** out);
void rt_mesh_destroy(rt_mesh* m);
/* readOBJ (optimized.cu:303-454): sscanf cascade, v*0.8+(0,-10,0) for 3-field vertices.
 * Unlike the reference (prints and continues, :310-313) a missing file is RT_ERR_IO. */
int rt_mesh_read_obj(rt_mesh* m, const char* path);
/* Replace the mesh by raw arrays: vertices nv*3 floats, triangles nt*3 vertex indices. */
int rt_mesh_set_triangles(rt_mesh* m, const float* vertices, int32_t nv, const int32_t* vtx_indices, int32_t nt);
/* rescale (optimized.cu:297-301): v = v*scale + offset, unfused float. */
int rt_mesh_rescale(rt_mesh* m, float scale, const float offset[3]);
/* Vertex normals (the viewer's loader keeps them, realtime_render.cu:489-493, 538-545; optimized.cu's drops them). keep != 0 before
 * rt_mesh_read_obj: the `vn` lines become the normals array and the faces' normal indices words 6-8 (ni, nj, nk) of the triangle
 * records. rt_mesh_set_normals attaches normals + nt*3 indices to a mesh given by arrays. Both before the BVH build. */
int rt_mesh_keep_normals(rt_mesh* m, int keep);
int rt_mesh_set_normals(rt_mesh* m, const float* normals, int32_t n_normals, const int32_t* normal_indices /* nt*3 */);
int rt_mesh_normal_count(const rt_mesh* m, int32_t* n_normals);
const float* rt_mesh_normals(const rt_mesh* m);            /* n_normals*3, NULL when none */
/* Merge `copies` transformed instances of the current mesh into one mesh (config 5 of BASELINE.json:
 * the reference has no instancing, so instances are baked). Instance c is v*scale[c] + offset[c]. */
int rt_mesh_instance(rt_mesh* m, int32_t copies, const float* scales, const float* offsets /* copies*3 */);
/* compute_bbox + buildBVH + bvhTreeToArray (optimized.cu:466-534, called at :809-813):
 * reorders the triangle records in place and produces the 10-float array BVH. */
int rt_mesh_build_bvh(rt_mesh* m);
/* The same builder on CUDA device `device`, a whole tree level at a time (csrc/rt_bvh_build.cuh): the identical arr_bvh
 * and triangle order (the reference builds on the host, optimized.cu:806-813, or in one device thread,
 * global_launcher.cu:298-331). build_ms (may be NULL): device time of the build without the transfers. */
int rt_mesh_build_bvh_gpu(rt_mesh* m, int device, double* build_ms);
/* After rt_mesh_build_bvh_gpu the post-build arrays STAY on that device; the host mirror (rt_mesh_tri_records, rt_mesh_arr_bvh) is
 * fetched on the first access. rt_scene_set_mesh_from (below) hands them to a scene of the same device without a host round trip. */
int rt_mesh_counts(const rt_mesh* m, int32_t* nv, int32_t* nt, int32_t* n_nodes);
const float* rt_mesh_vertices(const rt_mesh* m);          /* nv*3 */
const int32_t* rt_mesh_tri_records(const rt_mesh* m);     /* nt*RT_TRI_RECORD_WORDS, post-build order */
const float* rt_mesh_arr_bvh(const rt_mesh* m);           /* n_nodes*RT_BVH_NODE_FLOATS, NULL before build */
/* BVH shape facts used by tests: leaves, max depth, largest leaf. */
int rt_mesh_bvh_info(const rt_mesh* m, int32_t* n_leaves, int32_t* max_depth, int32_t* max_leaf);

/* ---- launcher helpers ---- */
/* z = -W / (2 * tanf(alpha/2)) in float, as optimized.cu:748-749 / cpu_launcher.cpp:666,694 evaluate it on the host. */
float rt_camera_z(int32_t W, float alpha);
/* The same expression evaluated ON THE DEVICE, as optimized.cu:748-749 does inside KernelLaunch (tan(float) there is CUDA's
 * tanf, one ulp off glibc's for alpha = pi/3): the z with which frames, t bits and shadow flags equal those of the
 * reference's own GPU program (IEEE build). rt_params_profile fills the host value; a caller reproducing optimized.cu's
 * output overrides rt_params::z with this one (the CLI does, for --profile optimized). */
int rt_camera_z_device(int device, int32_t W, float alpha, float* z);
/* Camera::rotate (realtime_render.cu:828-849): the viewer's camera basis for a yaw and a pitch, evaluated on the host in float. */
void rt_camera_basis(float yaw, float pitch, float bx[3], float by[3], float bz[3]);
/* Fill `p` with the knobs of one reference program: "cpu" (cpu_launcher.cpp), "optimized" (optimized.cu),
 * "array_bvh" (array_bvh.cu), "realtime" (realtime_render.cu: viewer camera with pitch 0.3 and a 90 degree field of view,
 * smooth normals, eps 1e-3, walls 0-5 + mesh 6 with the R = 940 floor; its light (0, 15, 40) is the caller's to set);
 * deterministic mode (aa_sigma=0, indirect=0); camera (0,0,55), alpha=pi/3. */
int rt_params_profile(rt_params* p, const char* profile, int32_t W, int32_t H, int32_t num_rays, int32_t num_bounce);
/* The six wall spheres with the ids of the given profile and the id the mesh takes
 * (cpu: walls 0-5, mesh 6, cpu_launcher.cpp:673-685; optimized: wall 0, mesh 1, walls 2-6, optimized.cu:684-726). */
int rt_default_walls(const char* profile, rt_sphere walls[6], int32_t* mesh_id);
/* PNG output, the role of stbi_write_png at optimized.cu:862. rgb is H*W*3, top row first. */
int rt_write_png(const char* path, int32_t W, int32_t H, const uint8_t* rgb);
/* Asynchronous PNG output for frame sequences: rt_png_writer_submit copies the frame and returns; a background thread
 * encodes (the scanlines in parallel bands) and writes it while the caller renders on. stbi_write_png at optimized.cu:862
 * is synchronous and single-threaded: at 4K it costs a thousand times the render. threads <= 0: all hardware threads;
 * max_pending <= 0: 4 queued frames before submit blocks. rt_png_writer_wait returns the first error of the frames written
 * so far; rt_png_writer_destroy writes what is still queued. */
typedef struct rt_png_writer rt_png_writer;
int rt_png_writer_create(rt_png_writer** out, int32_t threads, int32_t max_pending);
int rt_png_writer_submit(rt_png_writer* w, const char* path, int32_t W, int32_t H, const uint8_t* rgb);
int rt_png_writer_wait(rt_png_writer* w);
void rt_png_writer_destroy(rt_png_writer* w);
/* One step of the reference's (never launched) MoveLightSource, realtime_render.cu:1072-1090, evaluated on the host in float. */
void rt_move_light(float L[3], float angular_speed, float dt);

/* ---- device scene (Scene + the cudaMalloc/cudaMemcpy block, optimized.cu:791-826) ---- */
int rt_scene_create(rt_scene** out, int device);
void rt_scene_destroy(rt_scene* s);
/* Use an existing CUDA stream (a cudaStream_t cast to void*) instead of the scene's own. */
int rt_scene_set_stream(rt_scene* s, void* cuda_stream);
/* The stream every call on this scene enqueues on (a cudaStream_t): work of the caller that consumes a frame rendered with
 * RT_RENDER_NO_SYNC, or a band pushed with rt_scene_push_rows, must be ordered after it (event or same stream). */
int rt_scene_get_stream(rt_scene* s, void** cuda_stream);
/* Tuning and cross-check options of a scene (the reference has none: its variants are separate programs under
 * different-versions/). Every option selects among code paths that give identical results; the defaults are the production
 * choices. Options are part of the scene's state: nothing is read from the environment after rt_scene_create (which
 * presets an option from RT_<KEY> when that variable is set — a convenience for the tools/ scripts). Keys (int values):
 *   "variant" 0|1|2           0 one-thread-per-pixel kernel, exact arithmetic; 1 the same with the certified fast paths;
 *                             2 the wavefront pipeline (default)
 *   "anchored" -1|0|1         anchored-ray bins for camera / shadow rays: by mesh size (default), off (tree search), on
 *   "wide" -1|0|1             4-wide index for the tree search: default on for stochastic indirect bounces only
 *   "strips" 0..8             row bands on separate streams, 0 = chosen per call
 *   "bins_r", "task_factor", "npool_cap", "run_shift", "gss", "leaves_blocks", "side_stream", "diffuse_kernels",
 *   "stoch_mega", "wide_count", "graph", "six", "split", "fair_share", "top_smem", "debug_times", "debug_pool", "debug_bins", "debug_cost"   (see rt_device.cu: RtOptions)
 *   "transcendentals" 1|0     stochastic mode: log / cos / sin of optimized.cu:756-758, 635-636 by CUDA's single-precision
 *                             logf / cosf / sinf (1, default: what optimized.cu itself calls; frames equal those of the reference
 *                             kernel compiled without --use_fast_math bit for bit) or evaluated in double and rounded once (0)
 * Unknown key or value out of range: RT_ERR_INVALID. */
int rt_scene_set_option(rt_scene* s, const char* key, int64_t value);
int rt_scene_get_option(rt_scene* s, const char* key, int64_t* value);
int rt_scene_set_spheres(rt_scene* s, const rt_sphere* spheres, int32_t n);
/* Upload the mesh in the reference interchange formats (what optimized.cu:814-826 copies) and repack it on
 * the device into the traversal layout. nt == 0 removes the mesh. */
int rt_scene_set_mesh(rt_scene* s, const float* vertices, int32_t nv,
                      const int32_t* tri_records, int32_t nt,
                      const float* arr_bvh, int32_t n_nodes,
                      const float albedo[3], int32_t mirror, float n_in, float n_out, int32_t id);
/* The same upload with the three interchange arrays ALREADY ON THE SCENE'S DEVICE (device pointers; e.g. produced by the caller's own
 * kernels, or left there by rt_mesh_build_bvh_gpu): node relayout, leaf table and triangle repack run as kernels (csrc/rt_relayout.cuh),
 * nothing crosses the bus. No 4-wide index is built on this path (tree searches use the two-child records), and vertex normals
 * (rt_scene_set_mesh_normals) need the host upload. The arrays may be released when the call returns. */
int rt_scene_set_mesh_device(rt_scene* s, const float* d_vertices, int32_t nv, const int32_t* d_tri_records, int32_t nt,
                             const float* d_arr_bvh, int32_t n_nodes, const float albedo[3], int32_t mirror, float n_in, float n_out, int32_t id);
/* The mesh of an rt_mesh handle: rt_scene_set_mesh_device when rt_mesh_build_bvh_gpu left its arrays on this scene's device,
 * rt_scene_set_mesh with the host arrays otherwise. */
int rt_scene_set_mesh_from(rt_scene* s, rt_mesh* m, const float albedo[3], int32_t mirror, float n_in, float n_out, int32_t id);
/* Per-vertex normals for rt_params::smooth_normals (TriangleMesh::normals + TriangleIndices::ni,nj,nk, realtime_render.cu:239-241):
 * call after rt_scene_set_mesh with the mesh's normals array (rt_mesh_normals); the normal indices are words 6-8 of the triangle
 * records that call uploaded. Held beside the scene blob (not part of a broadcast: the viewer is single-GPU). n_normals == 0 removes them. */
int rt_scene_set_mesh_normals(rt_scene* s, const float* normals, int32_t n_normals);
/* Scene::L and Scene::intensity (optimized.cu:681-683). */
int rt_scene_set_light(rt_scene* s, const float L[3], float intensity);
/* Packed device scene as one contiguous blob, for a broadcast to other ranks (multi-GPU: the scene is built on
 * rank 0 only). export gives a DEVICE pointer valid until the scene changes; import adopts bytes from a
 * device buffer of that size (device-to-device copy). */
int rt_scene_blob_size(rt_scene* s, size_t* bytes);
int rt_scene_blob_export(rt_scene* s, void** device_ptr, size_t* bytes);
int rt_scene_blob_import(rt_scene* s, const void* device_ptr, size_t bytes);
/* Copy the blob into a caller-owned DEVICE buffer of at least rt_scene_blob_size bytes (e.g. the tensor a
 * broadcast collective sends from rank 0). */
int rt_scene_blob_copy_out(rt_scene* s, void* device_dst, size_t bytes);

/* The render call: replaces KernelLaunch<<<>>> + cudaDeviceSynchronize + cudaMemcpy D2H (optimized.cu:828-856).
 * rgb_out: rows*W*3 bytes, interleaved RGB, top row first, rows = rows rendered by this call. Each output
 * pointer may be a host pointer (copied back inside the call) or a device pointer on the scene's device
 * (written in place), or NULL (hit buffers only) to skip it.
 * hit_obj / hit_tri / hit_t: first-segment hit of sample 0 per pixel: object id (-1 = miss), triangle index in
 * post-BVH-build order (-1 unless the mesh was hit), t. shadow: 1 shadowed, 0 lit, 2 no diffuse hit. */
int rt_render(rt_scene* s, const rt_params* p, uint32_t flags,
              uint8_t* rgb_out, int32_t* hit_obj, int32_t* hit_tri, float* hit_t, uint8_t* shadow,
              rt_stats* stats);
int rt_scene_sync(rt_scene* s, rt_stats* stats);

/* ---- diagnostics ---- */
/* ---- peer frame buffer (multi-GPU, one process per GPU): the gather of a row-sharded frame as direct NVLink copies.
 * The destination rank allocates the frame (rt_peer_alloc) and hands the 64-byte CUDA IPC handle to the other ranks
 * (any transport); they open it (rt_peer_open) and push their row bands into it on their scene's stream
 * (rt_scene_push_rows: rows row_begin, row_begin + row_step, ... of a frame of W pixels per row). A barrier after the pushes is
 * the caller's. The reference is single-GPU; this replaces an NCCL all-gather of the bands. */
int rt_peer_alloc(int device, size_t bytes, void** ptr, uint8_t handle[64]);
int rt_peer_open(int device, const uint8_t handle[64], void** ptr);
int rt_peer_close(int device, void* ptr);
int rt_peer_free(int device, void* ptr);
/* Completion flags for the pushes: rt_peer_signal writes `value` to the 32-bit word `flag` (device memory, typically inside the destination's
 * rt_peer_alloc block) as an operation of the scene's stream, i.e. after the band pushed before it has landed; rt_peer_wait makes the scene's
 * stream wait until the word is >= value (counters that only grow: the frame number). Neither involves a kernel, a collective or the host.
 * RT_ERR_UNSUPPORTED when the driver has no stream memory operations (use a barrier then). A wait is meant for a write that arrives from
 * ANOTHER device: streams of one device may share a hardware queue, and a wait that only a later write of the same device can satisfy may hold
 * that write back for ever. */
int rt_peer_signal(rt_scene* s, void* flag, uint32_t value);
int rt_peer_wait(rt_scene* s, const void* flag, uint32_t value);
int rt_scene_push_rows(rt_scene* s, const void* band, void* frame, int32_t W, int32_t bytes_per_pixel, int32_t row_begin, int32_t row_step, int32_t rows);
/* the same for a band rendered with rt_params::row_group > 1: compact row k goes to row row_begin + (k / row_group) * row_step + k % row_group */
int rt_scene_push_row_groups(rt_scene* s, const void* band, void* frame, int32_t W, int32_t bytes_per_pixel, int32_t row_begin, int32_t row_step, int32_t row_group,
                             int32_t rows);

/* ---- multi-GPU (the reference is single-GPU, optimized.cu:774-884): pixels are independent, so a frame shards by
 * rows (row % nranks == rank through rt_params::row_begin / row_step) and an animation by frames; the only exchanges are the
 * scene broadcast before and the framebuffer gather after. NCCL is bound at run time (dlopen of libnccl.so.2).
 *   one process (or thread) per GPU:  rank 0 calls rt_comm_unique_id, the application transports the 128 bytes (MPI, a file,
 *                                     torch.distributed, ...), every rank calls rt_comm_init
 *   one process, N devices:           rt_comm_init_all (what `rt_render --gpus N` uses; peer access is enabled between the devices)
 * rt_scene_broadcast: the root's packed scene (built once: OBJ, BVH, repack) goes to every rank with ncclBroadcast on the scenes'
 *   streams and is adopted in place. Collective: every rank calls it with its own scene and communicator.
 * rt_gather_framebuffer: the row-interleaved bands (DEVICE pointers; rank r holds rows r, r + n, ... of an H-row frame, W *
 *   bytes_per_pixel bytes per row) are assembled in `frame` (device, H rows) on the root: grouped ncclSend / ncclRecv + one
 *   strided copy per source rank, on the scenes' streams. frame may be NULL on the other ranks. Collective. */
typedef struct rt_comm rt_comm;
int rt_comm_available(int* nccl_version);                       /* RT_ERR_UNSUPPORTED when libnccl.so.2 cannot be loaded */
int rt_comm_unique_id(uint8_t id[128]);
int rt_comm_init(rt_comm** out, int nranks, int rank, const uint8_t id[128], int device);
int rt_comm_init_all(rt_comm** comms /* ndev */, int ndev, const int* devices /* NULL: 0..ndev-1 */);
void rt_comm_destroy(rt_comm* c);
int rt_comm_rank(const rt_comm* c, int* rank, int* nranks);
int rt_scene_broadcast(rt_scene* s, rt_comm* c, int root, size_t* bytes /* may be NULL: size of the blob */);
int rt_gather_framebuffer(rt_scene* s, rt_comm* c, const void* band, int32_t W, int32_t H, int32_t bytes_per_pixel, void* frame, int root);
/* the same for bands of row GROUPS: rank r rendered the groups r, r + nranks, ... of row_group consecutive rows (rt_shard_rows) */
int rt_gather_framebuffer_groups(rt_scene* s, rt_comm* c, const void* band, int32_t W, int32_t H, int32_t bytes_per_pixel, int32_t row_group, void* frame,
                                 int root);
/* The row shard of one rank as rt_params fields: groups of row_group consecutive rows (a power of two <= 64; 1 = single rows) dealt out in
 * turn, group g of the frame to rank g % nranks. Sets row_begin = rank * row_group, row_step = nranks * row_group, row_group and row_count
 * (0 rows: row_count = 0 and the rank has nothing to render — rt_params::row_count == 0 would mean "all", so skip the call). Returns the
 * number of rows. No device needed. */
int rt_shard_rows(int32_t H, int32_t rank, int32_t nranks, int32_t row_group, rt_params* p);

/* Device self-test: the reciprocal-based exact division used by the fast slab test against div.rn.f32 on
 * blocks*256*per_thread pseudo-random operand pairs. out[0] = mismatches with one correction step,
 * out[1] = with two, out[2] = pairs tested. */
int rt_selftest_division(int device, uint64_t seed, int blocks, int per_thread, uint64_t out[3]);
/* the three-quotients-one-reciprocal division of the shading code (rt_math.cuh: div3) against div.rn.f32: out[0] differing
 * components, out[1] components tested */
int rt_selftest_division3(int device, uint64_t seed, int blocks, int per_thread, uint64_t out[2]);
/* Device self-test: the cuRAND library's XORWOW start states (d, v0..v4: n*6 words) and first four curand_uniform values
 * (n*4 floats) of the listed subsequences — the random stream of optimized.cu:745 that the stochastic mode reproduces. */
int rt_selftest_xorwow(int device, uint64_t seed, const uint32_t* subsequences, int32_t n, uint32_t* states6, float* uniforms4);

/* Device self-test / measurement: the FP32 FMA throughput of the device in TFLOP/s (8 independent FMA chains per thread, best of
 * `reps` event-timed launches): the measured FP32 roof bench.py reports the render path against. */
int rt_selftest_fma_peak(int device, int reps, double* tflops);
/* Device self-test: CUDA's single-precision logf (which = 0) / sinf (1) / cosf (2) / tanf (3) on n host arguments — the functions
 * option "transcendentals" = 1 and rt_camera_z_device evaluate (optimized.cu:749, 756-758, 635-636 call them in the kernel). */
int rt_selftest_libm(int device, int which, const float* x, int32_t n, float* y);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
