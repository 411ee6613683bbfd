"""The row-band push of raytracinggpu_b200.distributed.FramePush (rt_peer_alloc + rt_scene_push_rows) on one GPU: the bands of
all ranks, rendered one after the other by one process and pushed into the frame buffer, reassemble the whole frame. The
cross-process part (CUDA IPC handle, NVLink peer copies) is exercised by `bench.py --gpus N` (N > 1), which checks the pushed
frame against the NCCL all-gather's."""
import numpy as np
import pytest
import torch

import raytracinggpu_b200 as rt
from raytracinggpu_b200 import api, sharding
from oracle import profiles, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", [1, 4, 16])
def test_pushed_bands_reassemble_the_frame(built, group):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    desc = scenes.cat_scene("optimized", mirror=1) or scenes.torus_scene("optimized", mirror=1)
    sc = scenes.upload(rt.Scene(0), desc)
    W, H, world = 320, 187, 4  # ragged: the last rows are not shared evenly
    whole = sc.render(profiles.params("optimized", W, H, 1, 3), want=("rgb",))["rgb"]
    frame_ptr, handle = api.peer_alloc(0, H * W * 3)
    assert len(handle) == 64
    for r in range(world):
        p = profiles.params("optimized", W, H, 1, 3)
        p.row_begin, p.row_step, p.row_count = sharding.rows_for_rank(H, r, world, group=group)
        p.row_group = group
        if p.row_count == 0:
            continue
        band = torch.zeros((max(p.row_count, 1), W, 3), dtype=torch.uint8, device="cuda")
        sc.render_into(p, rgb=band)
        sc.push_rows(band.data_ptr(), frame_ptr, W, 3, p.row_begin, p.row_step, p.row_count, group)
    out = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
    sc.push_rows(frame_ptr, out.data_ptr(), W, 3, 0, 1, H)
    sc.sync()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), whole)
    api.peer_free(0, frame_ptr)
    sc.close()


def test_completion_flags_on_one_device(built):
    """rt_peer_signal / rt_peer_wait (stream-ordered 32-bit write / wait-until->= on the scene's stream): a flag written behind a pushed
    band is visible after the sync, and waits for values already reached do not block. (A wait that only a LATER write of the same device
    can satisfy is not tested: streams of one device may share a hardware queue, where the waiting operation would hold the write back —
    the flags are for writes that arrive from other devices, tests/mp_comm_worker.py.)"""
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    a = rt.Scene(0)
    block, _ = api.peer_alloc(0, 4096)
    try:
        zero = torch.zeros(4, dtype=torch.int32, device="cuda")
        a.push_rows(zero.data_ptr(), block, 4, 4, 0, 1, 1)
        a.sync()
        try:
            a.peer_signal(block, 5)
        except api.RtError as e:
            pytest.skip("no stream memory operations on this driver: %s" % e)
        a.peer_wait(block, 5)
        a.peer_wait(block, 3)  # already reached
        a.peer_signal(block + 4, 7)
        a.peer_wait(block + 4, 7)
        a.sync()
        out = torch.empty(4, dtype=torch.int32, device="cuda")
        a.push_rows(block, out.data_ptr(), 4, 4, 0, 1, 1)
        a.sync()
        assert out.cpu().tolist() == [5, 7, 0, 0]
    finally:
        api.peer_free(0, block)
        a.close()
