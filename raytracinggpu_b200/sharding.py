"""Host-side sharding of the render path across ranks (one process per GPU).

Pixels and frames are independent: a single frame is split by image rows, an animation by
frame index. There is no exchange during rendering; the only collective is the framebuffer gather after it.
All functions here are pure host logic (tested on CPU with a 2-rank gloo group, tests/test_multirank_gloo.py).
"""
import numpy as np


def rows_for_rank(H, rank, world, mode="interleave"):
    """(row_begin, row_step, row_count) for rt_params. 'interleave': row % world == rank (best balance, the
    mesh sits in the image centre); 'band': contiguous bands of ceil(H/world) rows."""
    if mode == "interleave":
        count = (H - rank + world - 1) // world if rank < H else 0
        return rank, world, count
    if mode == "band":
        per = (H + world - 1) // world
        begin = min(rank * per, H)
        return begin, 1, max(0, min(per, H - begin))
    raise ValueError(mode)


def padded_rows(H, world, mode="interleave"):
    """Rows every rank allocates so that an all-gather sees equal-size chunks."""
    return (H + world - 1) // world


def assemble(gathered, H, world, mode="interleave"):
    """gathered: array [world, padded_rows, W, C] as all_gather delivers it -> full frame [H, W, C]."""
    g = np.asarray(gathered)
    out = np.empty((H,) + g.shape[2:], dtype=g.dtype)
    for r in range(world):
        begin, step, count = rows_for_rank(H, r, world, mode)
        out[begin:begin + step * count:step] = g[r, :count]
    return out


def frames_for_rank(n_frames, rank, world):
    """Animation: frame f goes to rank f % world."""
    return list(range(rank, n_frames, world))


def light_positions(L0, n_frames, angular_speed, dt=0.02, move=None):
    """Light position of every frame, iterated in float exactly as MoveLightSource would step it
    (realtime_render.cu:1072-1090), so every rank derives identical positions. `move` is rt_move_light."""
    L = tuple(float(x) for x in L0)
    out = [L]
    for _ in range(1, n_frames):
        L = move(L, angular_speed, dt)
        out.append(L)
    return out
