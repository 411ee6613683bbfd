/*
 * host_misc.cpp — launcher-side helpers of librtb200: error channel, camera constant, the per-program knob
 * sets of the reference, the default wall spheres, the light animation step and the
 * PNG writer that takes the place of stbi_write_png (optimized.cu:862).
 */
#include "host_common.h"

#include <cmath>
#include <cstring>
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

enum Profile { kCpu, kOptimized, kArrayBvh, kUnknown };
Profile parse_profile(const char* s) {
    if (!s) return kUnknown;
    if (!strcmp(s, "cpu")) return kCpu;
    if (!strcmp(s, "optimized")) return kOptimized;
    if (!strcmp(s, "array_bvh")) return kArrayBvh;
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

int rt_params_profile(rt_params* p, const char* profile, int32_t W, int32_t H, int32_t num_rays, int32_t num_bounce) {
    if (!p) return rtb::fail(RT_ERR_INVALID, "rt_params_profile: NULL params");
    const Profile pr = parse_profile(profile);
    if (pr == kUnknown) return rtb::fail(RT_ERR_INVALID, "rt_params_profile: unknown profile '%s' (cpu | optimized | array_bvh)", profile ? profile : "(null)");
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
    const int32_t mid = (pr == kOptimized) ? 1 : 6; /* optimized.cu:690-700 vs cpu_launcher.cpp:685 */
    int32_t next = 0;
    for (int k = 0; k < 6; k++) {
        if (next == mid) next++;
        rt_sphere& s = walls[k];
        memcpy(s.C, kWalls[k].C, sizeof s.C);
        s.R = kWalls[k].R;
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

int rt_write_png(const char* path, int32_t W, int32_t H, const uint8_t* rgb) {
    if (!path || !rgb || W <= 0 || H <= 0) return rtb::fail(RT_ERR_INVALID, "rt_write_png: bad argument");
    /* 8-bit RGB, filter 0 on every scanline, one zlib stream */
    std::vector<uint8_t> raw((size_t)H * ((size_t)W * 3 + 1));
    for (int32_t y = 0; y < H; y++) {
        uint8_t* row = &raw[(size_t)y * ((size_t)W * 3 + 1)];
        row[0] = 0;
        memcpy(row + 1, rgb + (size_t)y * W * 3, (size_t)W * 3);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 3) != Z_OK) return rtb::fail(RT_ERR_NOMEM, "rt_write_png: deflate failed");
    std::vector<uint8_t> out;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    out.insert(out.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)W);
    put_be32(ihdr, (uint32_t)H);
    const uint8_t tail[5] = {8, 2, 0, 0, 0}; /* bit depth 8, colour type 2 (RGB), deflate, adaptive, no interlace */
    ihdr.insert(ihdr.end(), tail, tail + 5);
    png_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    png_chunk(out, "IDAT", z.data(), zlen);
    png_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(path, "wb");
    if (!f) return rtb::fail(RT_ERR_IO, "rt_write_png: cannot open '%s' for writing", path);
    const size_t wr = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return wr == out.size() ? RT_OK : rtb::fail(RT_ERR_IO, "rt_write_png: short write to '%s'", path);
}

} /* extern "C" */
