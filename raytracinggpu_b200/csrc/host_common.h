/* host_common.h — shared host-side helpers of librtb200 (error channel, small vector type). */
#pragma once
#include "../../include/rt_b200.h"

#include <cstdarg>
#include <cstdio>
#include <string>

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

struct Vec3 {
    float x, y, z;
    float operator[](int k) const { return k == 0 ? x : (k == 1 ? y : z); }
};

} // namespace rtb
