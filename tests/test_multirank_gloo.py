"""N > 1 host logic on CPU: a world_size-2 gloo group shards one frame by rows (each rank's rows are produced
by the oracle here, standing in for its GPU), all-gathers the bands and reassembles the frame; and an
animation is split by frame index. Compared with the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raytracinggpu_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, H, W, q, group=1):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import profiles, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = scenes.torus_scene("optimized")
    p = profiles.params("optimized", W, H, 1, 1)
    p.row_begin, p.row_step, p.row_count = sharding.rows_for_rank(H, rank, world, mode, group)
    p.row_group = group
    o = scenes.run_oracle(desc, p, threads=1, want=("rgb",))
    # the product's gather helper (raytracinggpu_b200/distributed.py), on CPU tensors over gloo
    from raytracinggpu_b200 import distributed as rtd
    fg = rtd.FrameGather(H, W, world, rank, torch.device("cpu"), mode, group=group)
    assert (fg.row_begin, fg.row_step, fg.row_count) == (p.row_begin, p.row_step, p.row_count)
    q2 = profiles.params("optimized", W, H, 1, 1)
    fg.apply(q2)
    assert (q2.row_begin, q2.row_step, q2.row_count, q2.row_group) == (p.row_begin, p.row_step, p.row_count, group)
    fg.band[:o["rgb"].shape[0]] = torch.from_numpy(o["rgb"])
    frame_t = fg.gather()
    # and the plain numpy assembly of an explicit all_gather
    gathered = [torch.empty_like(fg.band) for _ in range(world)]
    dist.all_gather(gathered, fg.band)
    rays = torch.tensor([o["work"]["rays"]], dtype=torch.int64)
    dist.all_reduce(rays)
    frame = sharding.assemble(torch.stack(gathered).numpy(), H, world, mode, group)
    assert np.array_equal(frame, frame_t.numpy())
    if rank == 0:
        q.put((frame, int(rays.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,H,group", [("interleave", 45, 1), ("band", 45, 1), ("interleave", 48, 1), ("interleave", 45, 4), ("interleave", 48, 8)])
def test_two_rank_row_sharding_reassembles_the_frame(built, mode, H, group):
    from oracle import profiles, scenes
    W, world = 64, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, H, W, q, group)) for r in range(world)]
    for p in procs:
        p.start()
    frame, rays = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = scenes.run_oracle(scenes.torus_scene("optimized"), profiles.params("optimized", W, H, 1, 1), want=("rgb",))
    assert np.array_equal(frame, full["rgb"])
    assert rays == full["work"]["rays"]


def test_row_partitions_cover_every_row_once(built):
    """Single rows, row groups and contiguous bands: every image row belongs to exactly one rank; the C ABI's rt_shard_rows (what C / C++
    hosts use) gives the same shards as the Python harness."""
    import raytracinggpu_b200 as rt
    for H in (1, 7, 45, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            for mode, group in (("interleave", 1), ("interleave", 2), ("interleave", 8), ("interleave", 64), ("band", 1)):
                seen = np.zeros(H, int)
                for r in range(world):
                    b, s, c = sharding.rows_for_rank(H, r, world, mode, group)
                    assert c <= sharding.padded_rows(H, world, mode, group)
                    rows = sharding.image_rows(b, s, c, group if mode == "interleave" else 1)
                    assert c == 0 or rows.max() < H
                    seen[rows] += 1
                    if mode == "interleave":
                        p = rt.params_profile("optimized", 16, H, 1, 1)
                        assert rt.shard_rows(p, r, world, group) == c
                        assert (p.row_step, p.row_count, p.row_group) == (s, c, group) and (c == 0 or p.row_begin == b)
                        assert c == 0 or rt.shard_row_count(p) == c
                assert (seen == 1).all(), (H, world, mode, group)


def test_frames_round_robin():
    assert sharding.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((sharding.frames_for_rank(240, r, 8) for r in range(8)), [])) == list(range(240))
