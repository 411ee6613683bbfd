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


def _worker(rank, world, port, mode, H, W, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import profiles, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    desc = scenes.torus_scene("optimized")
    p = profiles.params("optimized", W, H, 1, 1)
    p.row_begin, p.row_step, p.row_count = sharding.rows_for_rank(H, rank, world, mode)
    o = scenes.run_oracle(desc, p, threads=1, want=("rgb",))
    # the product's gather helper (raytracinggpu_b200/distributed.py), on CPU tensors over gloo
    from raytracinggpu_b200 import distributed as rtd
    fg = rtd.FrameGather(H, W, world, rank, torch.device("cpu"), mode)
    assert (fg.row_begin, fg.row_step, fg.row_count) == (p.row_begin, p.row_step, p.row_count)
    fg.band[:o["rgb"].shape[0]] = torch.from_numpy(o["rgb"])
    frame_t = fg.gather()
    # and the plain numpy assembly of an explicit all_gather
    gathered = [torch.empty_like(fg.band) for _ in range(world)]
    dist.all_gather(gathered, fg.band)
    rays = torch.tensor([o["work"]["rays"]], dtype=torch.int64)
    dist.all_reduce(rays)
    frame = sharding.assemble(torch.stack(gathered).numpy(), H, world, mode)
    assert np.array_equal(frame, frame_t.numpy())
    if rank == 0:
        q.put((frame, int(rays.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,H", [("interleave", 45), ("band", 45), ("interleave", 48)])
def test_two_rank_row_sharding_reassembles_the_frame(built, mode, H):
    from oracle import profiles, scenes
    W, world = 64, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    frame, rays = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = scenes.run_oracle(scenes.torus_scene("optimized"), profiles.params("optimized", W, H, 1, 1), want=("rgb",))
    assert np.array_equal(frame, full["rgb"])
    assert rays == full["work"]["rays"]


def test_row_partitions_cover_every_row_once():
    for H in (1, 7, 45, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            for mode in ("interleave", "band"):
                seen = np.zeros(H, int)
                for r in range(world):
                    b, s, c = sharding.rows_for_rank(H, r, world, mode)
                    assert c <= sharding.padded_rows(H, world, mode)
                    seen[b:b + s * c:s] += 1
                assert (seen == 1).all(), (H, world, mode)


def test_frames_round_robin():
    assert sharding.frames_for_rank(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((sharding.frames_for_rank(240, r, 8) for r in range(8)), [])) == list(range(240))
