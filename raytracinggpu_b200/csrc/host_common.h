/* host_common.h — shared host-side helpers of librtb200 (error channel, small vector type). */
#pragma once
#include "../../include/rt_b200.h"

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

namespace rtb {

/* Thread-local message behind rt_last_error(). */
std::string& last_error();

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

/* rt_device.cu: the reference's BVH builder on the device (rt_bvh_build.cuh). recs10 = whole triangle records (nt x 10).
 * keep == NULL: perm[i] = index, in the current order, of the triangle that ends at position i; arr = the 10-float array BVH.
 * keep != NULL: the post-build arrays stay on the device behind an opaque handle (*keep) and perm / arr stay empty; the host mirror is
 * fetched with bvh_device_download when somebody asks for it. info = nodes, leaves, depth, largest leaf. Returns a cudaError_t value. */
int bvh_build_device(int device, const float* vertices, int nv, const int32_t* recs10, int nt, std::vector<int32_t>* perm, std::vector<float>* arr, int32_t info[4],
                     double* build_ms, void** keep);
int bvh_device_download(void* keep, int32_t* recs10_out, std::vector<float>* arr_out);
void bvh_device_free(void* keep);
/* device pointers of a kept build: vertices nv*3, records nt*10 (post-build order), arr_bvh nn*10 */
void bvh_device_arrays(void* keep, int* device, const float** vertices, const int32_t** recs10, const float** arr, int32_t* nn);

/* host_mesh.cpp: the device arrays a kept rt_mesh_build_bvh_gpu left behind (NULL: none) */
void* mesh_device_handle(struct ::rt_mesh* m);

struct Vec3 {
    float x, y, z;
    float operator[](int k) const { return k == 0 ? x : (k == 1 ? y : z); }
};

} // namespace rtb
