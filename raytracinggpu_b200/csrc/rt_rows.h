/*
 * rt_rows.h — the row sharding of rt_params (row_begin / row_step / row_group) on the host side: how many rows a rank renders and how its
 * compact band goes back into the frame. Shared by rt_device.cu (rt_scene_push_row_groups) and rt_comm.cu (rt_gather_framebuffer_groups).
 */
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>

namespace rtb {

/* rows of a frame of H rows that rank r of n renders when groups of G consecutive rows are dealt out in turn (G = 1: row % n == r) */
inline int shard_rows(int H, int r, int n, int G) {
    const int begin = r * G, step = n * G;
    if (begin >= H) return 0;
    const int n_groups = (H - begin + step - 1) / step;
    return (n_groups - 1) * G + std::min(G, H - (begin + (n_groups - 1) * step));
}

/* compact band [rows][line bytes] -> rows row_begin + (k / G) * row_step + k % G of the frame: one strided copy for the whole groups,
 * one more for a last group the frame's end cut short */
inline cudaError_t scatter_band(void* frame, const void* band, size_t line, int row_begin, int row_step, int G, int rows, cudaMemcpyKind kind, cudaStream_t stream) {
    const int full = rows / G, rest = rows % G;
    cudaError_t e = cudaSuccess;
    if (full > 0)
        e = cudaMemcpy2DAsync((unsigned char*)frame + (size_t)row_begin * line, (size_t)row_step * line, band, (size_t)G * line, (size_t)G * line, (size_t)full, kind, stream);
    if (e == cudaSuccess && rest > 0)
        e = cudaMemcpyAsync((unsigned char*)frame + ((size_t)row_begin + (size_t)full * row_step) * line, (const unsigned char*)band + (size_t)full * G * line, (size_t)rest * line, kind, stream);
    return e;
}

} // namespace rtb
