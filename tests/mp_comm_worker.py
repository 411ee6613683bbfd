"""Worker of tests/test_gpu_comm.py::test_two_processes_nccl_and_ipc_push (launched by torchrun, one process per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracinggpu_b200 as rt  # noqa: E402
from raytracinggpu_b200 import distributed as rtd, sharding  # noqa: E402
from oracle import profiles, scenes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    W, H = 640, 363
    d = scenes.cat_scene("optimized", mirror=1) or scenes.torus_scene("optimized", mirror=1)
    # ---- the C ABI's own communicator: the 128-byte id travels over torch.distributed
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(rt.Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = rt.Comm.init(world, rank, bytes(uid.cpu().numpy().tobytes()), local)
    sc = rt.Scene(local)
    whole = None
    if rank == 0:
        scenes.upload(sc, d)
        whole = sc.render(profiles.params("optimized", W, H, 1, 3), want=("rgb",))["rgb"]
    comm.broadcast_scene(sc, 0)
    p = profiles.params("optimized", W, H, 1, 3)
    p.row_begin, p.row_step, p.row_count = sharding.rows_for_rank(H, rank, world)
    band = torch.zeros((max(p.row_count, 1), W, 3), dtype=torch.uint8, device=dev)
    frame = torch.zeros((H, W, 3), dtype=torch.uint8, device=dev)
    sc.render_into(p, rgb=band)
    comm.gather_framebuffer(sc, band.data_ptr(), W, H, 3, frame.data_ptr() if rank == 0 else 0, 0)
    sc.sync()
    if rank == 0:
        assert np.array_equal(frame.cpu().numpy(), whole), "rt_gather_framebuffer"
    # ---- CUDA IPC peer frame, scene on its own (non-blocking) stream: frames back to back, double-buffered; completion by flags in the
    # destination's memory (rt_peer_signal / rt_peer_wait) and by an all-reduce
    for mode in ("flags", "allreduce"):
        fp = rtd.FramePush(sc, H, W, world, rank, dev, group=8, signal=mode)
        if mode == "allreduce":
            assert fp.signal == "allreduce"
        pp = fp.apply(profiles.params("optimized", W, H, 1, 3))
        outs = []
        for k in range(5):
            sc.render_into(pp, rgb=fp.band, flags=rt.RT_RENDER_NO_SYNC)
            fp.push(release=False)
            if rank == 0:
                outs.append(fp.frame_tensor())
            fp.release()
        sc.sync()
        torch.cuda.synchronize()
        if rank == 0:
            print("FramePush", mode, "->", fp.signal)
            for o in outs:
                assert np.array_equal(o.cpu().numpy(), whole), "FramePush " + mode
        fp.close()
    comm.close()
    sc.close()
    dist.barrier()
    if rank == 0:
        print("MP_COMM_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
