/*
 * host_misc.cpp — launcher-side helpers of librtb200: error channel, camera constant, the per-program knob
 * sets of the reference, the default wall spheres, the light animation step and the
 * PNG writer that takes the place of stbi_write_png (optimized.cu:862).
 */
#include "host_common.h"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <zlib.h>

namespace rtb {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
} // namespace rtb

namespace {

struct WallDef {
    float C[3], R, albedo[3];
};
/* cpu_launcher.cpp:673-678 / optimized.cu:685-722, in the order cpu_launcher adds them */
const WallDef kWalls[6] = {
    {{0, 0, -1000}, 940, {0, 1, 0}}, /* green fore wall */
    {{0, -1000, 0}, 990, {0, 0, 1}}, /* blue floor */
    {{0, 1000, 0}, 940, {1, 0, 0}},  /* red ceiling */
    {{-1000, 0, 0}, 940, {0, 1, 1}}, /* cyan left wall */
    {{1000, 0, 0}, 940, {1, 1, 0}},  /* yellow right wall */
    {{0, 0, 1000}, 940, {1, 0, 1}},  /* magenta back wall */
};

enum Profile { kCpu, kOptimized, kArrayBvh, kRealtime, kUnknown };
Profile parse_profile(const char* s) {
    if (!s) return kUnknown;
    if (!strcmp(s, "cpu")) return kCpu;
    if (!strcmp(s, "optimized")) return kOptimized;
    if (!strcmp(s, "array_bvh")) return kArrayBvh;
    if (!strcmp(s, "realtime")) return kRealtime;
    return kUnknown;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}

void png_chunk(std::vector<uint8_t>& out, const char tag[4], const uint8_t* data, size_t n) {
    put_be32(out, (uint32_t)n);
    size_t at = out.size();
    out.insert(out.end(), tag, tag + 4);
    if (n) out.insert(out.end(), data, data + n);
    uint32_t crc = (uint32_t)crc32(0L, out.data() + at, (uInt)(n + 4));
    put_be32(out, crc);
}

} // namespace

extern "C" {

const char* rt_last_error(void) { return rtb::last_error().c_str(); }
int rt_abi_version(void) { return RT_ABI_VERSION; }

float rt_camera_z(int32_t W, float alpha) {
    /* optimized.cu:748-749: float alpha = PI/3; z = -W / (2*tan(alpha/2)). Both operands are compile-time
     * constants there, so the reference's compilers fold tan() to the correctly rounded float (W=512 gives
     * -443.405029); glibc's run-time tanf is 1 ulp away for this argument. The double tan rounded to float
     * reproduces the folded value. */
    const float t = (float)std::tan((double)(alpha / 2));
    return -W / (2 * t);
}

void rt_camera_basis(float yaw, float pitch, float bx_[3], float by_[3], float bz_[3]) {
    /* Camera::rotate, realtime_render.cu:828-849, operation for operation in float (Vector * float, Vector + Vector, cross,
     * normalize = three divisions by sqrtf(norm2)); cos / sin of a float argument are the float functions in CUDA host code */
    struct V { float x, y, z; };
    auto mul = [](V a, float b) { return V{a.x * b, a.y * b, a.z * b}; };
    auto add = [](V a, V b) { return V{a.x + b.x, a.y + b.y, a.z + b.z}; };
    auto sub = [](V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; };
    auto cross = [](V a, V b) { return V{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
    auto normalize = [](V a) {
        const float n = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
        return V{a.x / n, a.y / n, a.z / n};
    };
    V bx{1, 0, 0}, by{0, 1, 0}, bz{0, 0, -1};
    const float cy = cosf(yaw), sy = sinf(yaw);
    bx = add(mul(bx, cy), mul(bz, sy));
    bz = cross(by, bx);
    const float cp = cosf(pitch), sp = sinf(pitch);
    by = sub(mul(by, cp), mul(bz, sp));
    bz = cross(bx, by);
    bx = normalize(bx);
    by = normalize(by);
    bz = normalize(bz);
    bx_[0] = bx.x; bx_[1] = bx.y; bx_[2] = bx.z;
    by_[0] = by.x; by_[1] = by.y; by_[2] = by.z;
    bz_[0] = bz.x; bz_[1] = bz.y; bz_[2] = bz.z;
}

int rt_shard_rows(int32_t H, int32_t rank, int32_t nranks, int32_t row_group, rt_params* p) {
    if (H <= 0 || nranks < 1 || rank < 0 || rank >= nranks || row_group < 1 || row_group > 64 || (row_group & (row_group - 1)) != 0) {
        rtb::fail(RT_ERR_INVALID, "rt_shard_rows: bad argument");
        return RT_ERR_INVALID;
    }
    const int G = row_group, begin = rank * G, step = nranks * G;
    int rows = 0;
    if (begin < H) {
        const int n_groups = (H - begin + step - 1) / step;
        rows = (n_groups - 1) * G + std::min(G, H - (begin + (n_groups - 1) * step));
    }
    if (p) {
        p->row_begin = begin < H ? begin : 0;
        p->row_step = step;
        p->row_group = G;
        p->row_count = rows;
    }
    return rows;
}

int rt_params_profile(rt_params* p, const char* profile, int32_t W, int32_t H, int32_t num_rays, int32_t num_bounce) {
    if (!p) return rtb::fail(RT_ERR_INVALID, "rt_params_profile: NULL params");
    const Profile pr = parse_profile(profile);
    if (pr == kUnknown) return rtb::fail(RT_ERR_INVALID, "rt_params_profile: unknown profile '%s' (cpu | optimized | array_bvh | realtime)", profile ? profile : "(null)");
    if (W <= 0 || H <= 0 || num_rays < 0 || num_bounce < 0) return rtb::fail(RT_ERR_INVALID, "rt_params_profile: bad size");
    memset(p, 0, sizeof *p);
    p->W = W;
    p->H = H;
    p->num_rays = num_rays;
    p->num_bounce = num_bounce;
    p->cam[0] = 0.f;
    p->cam[1] = 0.f;
    p->cam[2] = 55.f;                                        /* optimized.cu:747 */
    p->z = rt_camera_z(W, (float)(3.14159265358979323846 / 3)); /* :748-749 */
    p->aa_sigma = 0.f;                                       /* deterministic mode */
    p->indirect = 0;
    p->row_begin = 0;
    p->row_step = 1;
    p->row_count = 0;
    p->row_group = 1;
    switch (pr) {
    case kCpu: /* cpu_launcher.cpp:575,301,291-292,567,714 */
        p->eps_surface = 1e-3f;
        p->eps_tri = 1e-4f;
        p->push_order = 0;
        p->extra_segment = 1;
        p->gamma_mode = 0;
        break;
    case kOptimized: /* optimized.cu:575,275,265-266,566,765 */
        p->eps_surface = 1e-4f;
        p->eps_tri = 0.f;
        p->push_order = 1;
        p->extra_segment = 0;
        p->gamma_mode = 1;
        break;
    case kRealtime: /* realtime_render.cu:908,298,288-289; pov = PI/2 :1021; Camera() pitch 0.3 :809; smooth normals :311 */
        p->eps_surface = 1e-3f;
        p->eps_tri = 1e-3f;
        p->push_order = 0;
        p->extra_segment = 0;
        p->gamma_mode = 1;
        p->z = rt_camera_z(W, (float)(3.14159265358979323846 / 2));
        p->camera_mode = 1;
        rt_camera_basis(0.f, 0.3f, p->cam_bx, p->cam_by, p->cam_bz);
        p->smooth_normals = 1;
        break;
    default: /* array_bvh.cu:805,292,282-283; host gamma :1112-1118 */
        p->eps_surface = 1e-4f;
        p->eps_tri = 1e-4f;
        p->push_order = 0;
        p->extra_segment = 0;
        p->gamma_mode = 0;
        break;
    }
    return RT_OK;
}

int rt_default_walls(const char* profile, rt_sphere walls[6], int32_t* mesh_id) {
    const Profile pr = parse_profile(profile);
    if (pr == kUnknown || !walls) return rtb::fail(RT_ERR_INVALID, "rt_default_walls: bad argument");
    const int32_t mid = (pr == kOptimized) ? 1 : 6; /* optimized.cu:690-700 vs cpu_launcher.cpp:685, realtime_render.cu:1026-1050 */
    int32_t next = 0;
    for (int k = 0; k < 6; k++) {
        if (next == mid) next++;
        rt_sphere& s = walls[k];
        memcpy(s.C, kWalls[k].C, sizeof s.C);
        s.R = (pr == kRealtime && k == 1) ? 940.f : kWalls[k].R; /* the viewer's floor: realtime_render.cu:1027 */
        memcpy(s.albedo, kWalls[k].albedo, sizeof s.albedo);
        s.mirror = 0;
        s.n_in = 1.f;
        s.n_out = 1.f;
        s.id = next++;
    }
    if (mesh_id) *mesh_id = mid;
    return RT_OK;
}

void rt_move_light(float L[3], float angular_speed, float dt) {
    /* MoveLightSource, realtime_render.cu:1072-1090, about C = (0,0,0), in float on the host so that every
     * rank derives the same light position for frame f. */
    const float radius = sqrtf(powf(0.f - L[0], 2) + powf(0.f - L[2], 2));
    const float cur = atan2f(L[2], L[0]);
    const float ang = cur + angular_speed * dt;
    L[0] = 0.f + radius * cosf(ang);
    L[2] = 0.f + radius * sinf(ang);
}

/* The PNG stream of an image, encoded by `threads` threads: the scanlines are cut into bands, every band is deflated on
 * its own (raw deflate, ended with a full flush so that it ends on a byte boundary; the last band ends the stream) and the
 * pieces are concatenated behind one zlib header, with the Adler-32 of the whole assembled from the bands' checksums — any
 * inflater sees one ordinary stream (the technique of pigz). stb_image_write, which the reference calls at
 * optimized.cu:862, deflates a 4K frame on one core in ~0.3 s; this is what moves the encode off the critical path. */
static int png_encode(std::vector<uint8_t>& out, int32_t W, int32_t H, const uint8_t* rgb, int threads) {
    const size_t stride = (size_t)W * 3 + 1;
    int bands = std::max(1, std::min(threads, H / 32));
    struct Band {
        std::vector<uint8_t> z;
        uLong adler = 1;
        size_t raw_len = 0;
        int rc = Z_OK;
    };
    std::vector<Band> band((size_t)bands);
    auto work = [&](int b) {
        const int32_t y0 = (int32_t)((int64_t)H * b / bands), y1 = (int32_t)((int64_t)H * (b + 1) / bands);
        std::vector<uint8_t> raw((size_t)(y1 - y0) * stride);
        for (int32_t y = y0; y < y1; y++) { /* 8-bit RGB, filter 0 on every scanline */
            uint8_t* row = &raw[(size_t)(y - y0) * stride];
            row[0] = 0;
            memcpy(row + 1, rgb + (size_t)y * W * 3, (size_t)W * 3);
        }
        Band& o = band[(size_t)b];
        o.raw_len = raw.size();
        o.adler = adler32(1L, raw.data(), (uInt)raw.size());
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if ((o.rc = deflateInit2(&zs, 3, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY)) != Z_OK) return;
        o.z.resize(deflateBound(&zs, (uLong)raw.size()) + 64);
        zs.next_in = raw.data();
        zs.avail_in = (uInt)raw.size();
        zs.next_out = o.z.data();
        zs.avail_out = (uInt)o.z.size();
        const int rc = deflate(&zs, b == bands - 1 ? Z_FINISH : Z_FULL_FLUSH);
        o.rc = (b == bands - 1) ? (rc == Z_STREAM_END ? Z_OK : Z_BUF_ERROR) : ((rc == Z_OK && zs.avail_in == 0) ? Z_OK : Z_BUF_ERROR);
        o.z.resize(zs.total_out);
        deflateEnd(&zs);
    };
    if (bands == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int b = 1; b < bands; b++) pool.emplace_back(work, b);
        work(0);
        for (std::thread& t : pool) t.join();
    }
    std::vector<uint8_t> z;
    z.push_back(0x78); /* zlib header: deflate, 32 K window, no dictionary, check bits */
    z.push_back(0x5e);
    uLong adler = 1;
    for (int b = 0; b < bands; b++) {
        if (band[(size_t)b].rc != Z_OK) return RT_ERR_NOMEM;
        z.insert(z.end(), band[(size_t)b].z.begin(), band[(size_t)b].z.end());
        adler = b == 0 ? band[0].adler : adler32_combine(adler, band[(size_t)b].adler, (z_off_t)band[(size_t)b].raw_len);
    }
    put_be32(z, (uint32_t)adler);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    out.clear();
    out.insert(out.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)W);
    put_be32(ihdr, (uint32_t)H);
    const uint8_t tail[5] = {8, 2, 0, 0, 0}; /* bit depth 8, colour type 2 (RGB), deflate, adaptive, no interlace */
    ihdr.insert(ihdr.end(), tail, tail + 5);
    png_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    png_chunk(out, "IDAT", z.data(), z.size());
    png_chunk(out, "IEND", nullptr, 0);
    return RT_OK;
}

static int png_to_file(const char* path, int32_t W, int32_t H, const uint8_t* rgb, int threads) {
    std::vector<uint8_t> out;
    if (png_encode(out, W, H, rgb, threads) != RT_OK) return rtb::fail(RT_ERR_NOMEM, "rt_write_png: deflate failed");
    FILE* f = fopen(path, "wb");
    if (!f) return rtb::fail(RT_ERR_IO, "rt_write_png: cannot open '%s' for writing", path);
    const size_t wr = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return wr == out.size() ? RT_OK : rtb::fail(RT_ERR_IO, "rt_write_png: short write to '%s'", path);
}

static int default_png_threads() {
    if (const char* v = getenv("RT_PNG_THREADS")) return std::max(1, atoi(v));
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 32u));
}

int rt_write_png(const char* path, int32_t W, int32_t H, const uint8_t* rgb) {
    if (!path || !rgb || W <= 0 || H <= 0) return rtb::fail(RT_ERR_INVALID, "rt_write_png: bad argument");
    return png_to_file(path, W, H, rgb, default_png_threads());
}

/* ---- asynchronous PNG output: frames are copied into the writer and encoded + written by a background
 * thread (each frame by the band threads above) while the caller renders the next ones --------------------------------- */
struct rt_png_writer {
    struct Job {
        std::string path;
        int32_t W, H;
        std::vector<uint8_t> rgb;
    };
    std::mutex mu;
    std::condition_variable cv_job, cv_room;
    std::deque<Job> jobs;
    size_t max_pending = 4;
    int in_flight = 0;
    int threads = 1;
    bool stop = false;
    int first_error = RT_OK;
    std::string first_message;
    std::thread worker;
};

int rt_png_writer_create(rt_png_writer** out, int32_t threads, int32_t max_pending) {
    if (!out) return rtb::fail(RT_ERR_INVALID, "rt_png_writer_create: out is NULL");
    rt_png_writer* w = new (std::nothrow) rt_png_writer();
    if (!w) return rtb::fail(RT_ERR_NOMEM, "rt_png_writer_create: out of memory");
    w->threads = threads > 0 ? threads : default_png_threads();
    w->max_pending = (size_t)std::max(1, max_pending > 0 ? max_pending : 4);
    w->worker = std::thread([w]() {
        for (;;) {
            rt_png_writer::Job job;
            {
                std::unique_lock<std::mutex> lk(w->mu);
                w->cv_job.wait(lk, [w]() { return w->stop || !w->jobs.empty(); });
                if (w->jobs.empty()) return; /* stop requested and nothing left */
                job = std::move(w->jobs.front());
                w->jobs.pop_front();
                w->in_flight++;
            }
            w->cv_room.notify_all();
            const int rc = png_to_file(job.path.c_str(), job.W, job.H, job.rgb.data(), w->threads);
            {
                std::lock_guard<std::mutex> lk(w->mu);
                if (rc != RT_OK && w->first_error == RT_OK) {
                    w->first_error = rc;
                    w->first_message = rtb::last_error(); /* the worker's thread-local message */
                }
                w->in_flight--;
            }
            w->cv_room.notify_all();
        }
    });
    *out = w;
    return RT_OK;
}

int rt_png_writer_submit(rt_png_writer* w, const char* path, int32_t W, int32_t H, const uint8_t* rgb) {
    if (!w || !path || !rgb || W <= 0 || H <= 0) return rtb::fail(RT_ERR_INVALID, "rt_png_writer_submit: bad argument");
    rt_png_writer::Job job;
    job.path = path;
    job.W = W;
    job.H = H;
    job.rgb.assign(rgb, rgb + (size_t)W * H * 3); /* the caller's buffer is free again when this returns */
    {
        std::unique_lock<std::mutex> lk(w->mu);
        w->cv_room.wait(lk, [w]() { return w->jobs.size() < w->max_pending; }); /* back-pressure: at most max_pending frames queued */
        w->jobs.push_back(std::move(job));
    }
    w->cv_job.notify_one();
    return RT_OK;
}

int rt_png_writer_wait(rt_png_writer* w) {
    if (!w) return rtb::fail(RT_ERR_INVALID, "rt_png_writer_wait: NULL writer");
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv_room.wait(lk, [w]() { return w->jobs.empty() && w->in_flight == 0; });
    if (w->first_error != RT_OK) {
        const int rc = w->first_error;
        rtb::last_error() = w->first_message;
        w->first_error = RT_OK;
        return rc;
    }
    return RT_OK;
}

void rt_png_writer_destroy(rt_png_writer* w) {
    if (!w) return;
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->stop = true;
    }
    w->cv_job.notify_all();
    if (w->worker.joinable()) w->worker.join(); /* drains the queue first */
    delete w;
}

} /* extern "C" */
