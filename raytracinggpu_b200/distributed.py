"""Multi-GPU plumbing of the render path: one process per GPU, `torch.distributed` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).

The path has no exchange during rendering. The two collectives are
  * `broadcast_scene`: the scene is loaded, BVH-built and packed on rank 0 only; its device blob (header + node
    records + triangle records, rt_layout.h) is broadcast once and adopted by every other rank;
  * `gather_frame`: every rank renders the rows `rank, rank + R, ...` of one frame; the equal-size bands are
    all-gathered and de-interleaved on the device into the full frame (NCCL has no native gather; with NVSwitch a
    flat all-gather is the cheapest exchange of R x 3 MB pieces).
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
