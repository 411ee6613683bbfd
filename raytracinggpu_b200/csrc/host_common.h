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

/* rt_device.cu: the reference's BVH builder on the device (rt_bvh_build.cuh). perm[i] = index, in the current order, of the
 * triangle that ends at position i; arr = the 10-float array BVH; info = nodes, leaves, depth, largest leaf. Returns a
 * cudaError_t value (0 = success). */
int bvh_build_device(int device, const float* vertices, int nv, const int32_t* idx3, int nt, std::vector<int32_t>* perm, std::vector<float>* arr, int32_t info[4],
                     double* build_ms);

struct Vec3 {
    float x, y, z;
    float operator[](int k) const { return k == 0 ? x : (k == 1 ? y : z); }
};

} // namespace rtb
