"""Host-side sharding of the render path across ranks (one process per GPU).

Pixels and frames are independent: a single frame is split by image rows, an animation by
frame index. There is no exchange during rendering; the only collective is the framebuffer gather after it.
All functions here are pure host logic (tested on CPU with a 2-rank gloo group, tests/test_multirank_gloo.py).
"""
import numpy as np


def rows_for_rank(H, rank, world, mode="interleave", group=1):
    """(row_begin, row_step, row_count) for rt_params. 'interleave': groups of `group` consecutive rows dealt out in turn, group g of the
    frame to rank g % world (group 1: row % world == rank; rt_params.row_group = group, rt_shard_rows in the C ABI): balanced whatever the
    image, and with group >= 4 a warp's 8x4 pixel tile stays a tile; 'band': contiguous bands of ceil(H/world) rows."""
    if mode == "interleave":
        begin, step = rank * group, world * group
        if begin >= H:
            return 0, step, 0
        n_groups = (H - begin + step - 1) // step
        count = (n_groups - 1) * group + min(group, H - (begin + (n_groups - 1) * step))
        return begin, step, count
    if mode == "band":
        per = (H + world - 1) // world
        begin = min(rank * per, H)
        return begin, 1, max(0, min(per, H - begin))
    raise ValueError(mode)


def image_rows(begin, step, count, group=1):
    """Image row of every compact row of a shard: begin + (k // group) * step + k % group."""
    k = np.arange(count)
    return begin + (k // group) * step + k % group


def padded_rows(H, world, mode="interleave", group=1):
    """Rows every rank allocates so that an all-gather sees equal-size chunks."""
    return max(rows_for_rank(H, r, world, mode, group)[2] for r in range(world))


def assemble(gathered, H, world, mode="interleave", group=1):
    """gathered: array [world, padded_rows, W, C] as all_gather delivers it -> full frame [H, W, C]."""
    g = np.asarray(gathered)
    out = np.empty((H,) + g.shape[2:], dtype=g.dtype)
    for r in range(world):
        begin, step, count = rows_for_rank(H, r, world, mode, group)
        out[image_rows(begin, step, count, group if mode == "interleave" else 1)] = g[r, :count]
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
