"""Multi-GPU plumbing of the render path: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).

The path has no exchange during rendering. The two collectives are
  * `broadcast_scene`: the scene is loaded, BVH-built and packed on rank 0 only; its device blob (header + node
    records + triangle records, rt_layout.h) is broadcast once and adopted by every other rank;
  * `gather_frame`: every rank renders the rows `rank, rank + R, ...` of one frame; the equal-size bands are
    all-gathered and de-interleaved on the device into the full frame (NCCL has no native gather; with NVSwitch a
    flat all-gather is the cheapest exchange of R x 3 MB pieces).
  * `FramePush`: the same gather without a collective: rank `dst` allocates the frame and shares a CUDA IPC handle once;
    every rank copies its band straight into it over NVLink (one strided device-to-device copy on the render stream);
    what is left per frame is one barrier.
Frame-parallel animations need neither: rank r renders frames r, r + R, ... (sharding.frames_for_rank).
"""
import torch
import torch.distributed as dist

from . import sharding


def broadcast_scene(scene, src=0, device=None):
    """Make `scene` on every rank hold rank `src`'s packed scene. Returns the blob size in bytes."""
    rank = dist.get_rank()
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        n[0] = scene.blob_size()
    dist.broadcast(n, src)
    nbytes = int(n.item())
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if rank == src:
        scene.blob_copy_out(buf.data_ptr(), nbytes)
    dist.broadcast(buf, src)
    if rank != src:
        scene.blob_import(buf.data_ptr(), nbytes)
    return nbytes


class FrameGather:
    """Buffers for row-interleaved rendering of H x W frames on `world` ranks, reused across frames."""

    def __init__(self, H, W, world, rank, device, mode="interleave", channels=3):
        self.H, self.W, self.world, self.rank, self.mode = H, W, world, rank, mode
        self.pad = sharding.padded_rows(H, world, mode)
        self.row_begin, self.row_step, self.row_count = sharding.rows_for_rank(H, rank, world, mode)
        self.band = torch.zeros((self.pad, W, channels), dtype=torch.uint8, device=device)
        self.gathered = torch.empty((world, self.pad, W, channels), dtype=torch.uint8, device=device)
        self.frame = torch.empty((H, W, channels), dtype=torch.uint8, device=device)

    def apply(self, params):
        """Set the sharding fields of an rt_params for this rank."""
        params.row_begin, params.row_step, params.row_count = self.row_begin, self.row_step, self.row_count
        return params

    def gather(self):
        """All-gather the bands; returns the assembled [H, W, C] frame (valid on every rank)."""
        if self.world > 1:
            try:
                dist.all_gather_into_tensor(self.gathered.view(-1), self.band.view(-1))
            except (RuntimeError, NotImplementedError):  # backends without the flat form (CPU tests)
                parts = [self.gathered[r] for r in range(self.world)]
                dist.all_gather(parts, self.band)
        else:
            self.gathered[0].copy_(self.band)
        for r in range(self.world):
            b, s, c = sharding.rows_for_rank(self.H, r, self.world, self.mode)
            if c:
                self.frame[b:b + s * c:s].copy_(self.gathered[r, :c])
        return self.frame


class FramePush:
    """Row-interleaved rendering of H x W frames on `world` ranks with the bands pushed into rank `dst`'s frame buffer over
    NVLink (CUDA IPC + peer copies; NCCL backend, one process per GPU of one node). `frame` is valid on rank `dst`."""

    def __init__(self, scene, H, W, world, rank, device, dst=0, channels=3):
        from . import api
        self.scene, self.H, self.W, self.world, self.rank, self.dst, self.channels = scene, H, W, world, rank, dst, channels
        self.device_index = device.index if device.index is not None else torch.cuda.current_device()
        self.row_begin, self.row_step, self.row_count = sharding.rows_for_rank(H, rank, world, "interleave")
        self.band = torch.zeros((max(self.row_count, 1), W, channels), dtype=torch.uint8, device=device)
        handle = torch.zeros(64, dtype=torch.uint8, device=device)
        self._own = self._peer = None
        if rank == dst:
            self._own, hb = api.peer_alloc(self.device_index, H * W * channels)
            handle.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(handle, dst)
        if rank == dst:
            self.frame_ptr = self._own
        else:
            self._peer = api.peer_open(self.device_index, bytes(handle.cpu().numpy().tobytes()))
            self.frame_ptr = self._peer
        self._token = torch.zeros(1, dtype=torch.int32, device=device)

    def apply(self, params):
        params.row_begin, params.row_step, params.row_count = self.row_begin, self.row_step, self.row_count
        return params

    def push(self):
        """Enqueue this rank's copy (on the scene's stream) and the barrier that tells `dst` every band has landed."""
        self.scene.push_rows(self.band.data_ptr(), self.frame_ptr, self.W, self.channels, self.row_begin, self.row_step, self.row_count)
        if self.world > 1:
            dist.all_reduce(self._token)  # ordered after the copy on the current stream; complete when every rank's copy is

    def frame_tensor(self):
        """rank `dst` only: a copy of the assembled frame as a torch tensor (enqueued on the scene's stream)."""
        assert self.rank == self.dst
        out = torch.empty((self.H, self.W, self.channels), dtype=torch.uint8, device=self.band.device)
        self.scene.push_rows(self._own, out.data_ptr(), self.W, self.channels, 0, 1, self.H)
        return out

    def close(self):
        from . import api
        if self._peer:
            api.peer_close(self.device_index, self._peer)
            self._peer = None
        if self.world > 1:
            dist.barrier()  # nobody frees the frame while a peer still maps it
        if self._own:
            api.peer_free(self.device_index, self._own)
            self._own = None
