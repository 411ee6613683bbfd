"""Multi-GPU behind the C ABI (rt_comm_*, rt_scene_broadcast, rt_gather_framebuffer: NCCL bound at run time).

One device is enough for the single-rank communicator (both collectives degenerate but run through NCCL); with two or more
devices the scene is built on device 0 only, broadcast, rendered row-interleaved by one thread per device and gathered to
device 0 — against the whole frame rendered by one scene. The cross-PROCESS paths (unique id transported by torch.distributed,
CUDA IPC peer frame) have their own test below, spawned with two processes when two devices exist."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

import raytracinggpu_b200 as rt
from raytracinggpu_b200 import sharding
from oracle import profiles, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gpu(built):
    if rt.device_count() < 1:
        pytest.fail("no CUDA device: the gpu tests must run on the B200 box (there is no CPU fallback)")
    assert rt.comm_available() >= 20000, "libnccl.so.2 could not be loaded"
    return 0


def desc_mirror():
    return scenes.cat_scene("optimized", mirror=1) or scenes.torus_scene("optimized", mirror=1)


def test_single_rank_communicator(gpu):
    W, H = 320, 187
    d = desc_mirror()
    sc = scenes.upload(rt.Scene(0), d)
    p = profiles.params("optimized", W, H, 1, 3)
    whole = sc.render(p, want=("rgb",))["rgb"]
    comm = rt.Comm.init(1, 0, rt.Comm.unique_id(), 0)
    assert comm.rank() == (0, 1)
    assert comm.broadcast_scene(sc, 0) == sc.blob_size()
    band = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda:0")
    frame = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda:0")
    sc.render_into(p, rgb=band)
    comm.gather_framebuffer(sc, band.data_ptr(), W, H, 3, frame.data_ptr(), 0)
    sc.sync()
    assert np.array_equal(frame.cpu().numpy(), whole)
    comm.close()
    sc.close()


def test_one_process_n_devices(gpu):
    n = min(rt.device_count(), 4)
    if n < 2:
        pytest.skip("needs two devices (run under gpurun --gpus 2)")
    W, H = 640, 363  # ragged: 363 rows over n ranks
    d = desc_mirror()
    p0 = profiles.params("optimized", W, H, 1, 3)
    ref_scene = scenes.upload(rt.Scene(0), d)
    whole = ref_scene.render(p0, want=("rgb", "hit_tri"))
    comms = rt.Comm.init_all(n)
    scs = [rt.Scene(k) for k in range(n)]
    scenes.upload(scs[0], d)  # built on the root only
    frame = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda:0")
    errors = []

    def worker(r):
        try:
            nbytes = comms[r].broadcast_scene(scs[r], 0)
            assert nbytes == scs[0].blob_size()
            p = profiles.params("optimized", W, H, 1, 3)
            assert rt.shard_rows(p, r, n, 8) == sharding.rows_for_rank(H, r, n, group=8)[2]  # groups of 8 rows (rt_shard_rows)
            band = torch.zeros((max(p.row_count, 1), W, 3), dtype=torch.uint8, device="cuda:%d" % r)
            scs[r].render_into(p, rgb=band)
            comms[r].gather_framebuffer(scs[r], band.data_ptr(), W, H, 3, frame.data_ptr() if r == 0 else 0, 0, row_group=8)
            scs[r].sync()
        except Exception as e:  # noqa: BLE001
            errors.append((r, repr(e)))

    th = [threading.Thread(target=worker, args=(r,)) for r in range(n)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    assert np.array_equal(frame.cpu().numpy(), whole["rgb"])
    for c in comms:
        c.close()
    for s in scs + [ref_scene]:
        s.close()


def test_two_processes_nccl_and_ipc_push(gpu):
    """One process per GPU (torchrun): unique id over torch.distributed, rt_comm_init, rt_scene_broadcast,
    rt_gather_framebuffer, and the CUDA IPC peer frame (rt_peer_open + rt_scene_push_rows) of distributed.FramePush."""
    if rt.device_count() < 2:
        pytest.skip("needs two devices (run under gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(ROOT, "tests", "mp_comm_worker.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "MP_COMM_OK" in r.stdout
