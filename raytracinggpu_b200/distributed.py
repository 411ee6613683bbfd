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


def _order_after_scene(scene):
    """Make torch's current stream wait for everything enqueued so far on the scene's stream (the render, a pushed band):
    the scene's stream is a private non-blocking stream unless the caller handed it torch's with set_stream."""
    if not torch.cuda.is_available():
        return
    sp = scene.get_stream()
    cur = torch.cuda.current_stream()
    if sp == cur.cuda_stream:
        return
    ev = torch.cuda.Event()
    ev.record(torch.cuda.ExternalStream(sp))
    cur.wait_event(ev)


def _order_scene_after_current(scene):
    """The reverse edge: later work on the scene's stream waits for what torch's current stream holds (a collective)."""
    if not torch.cuda.is_available():
        return
    sp = scene.get_stream()
    cur = torch.cuda.current_stream()
    if sp == cur.cuda_stream:
        return
    ev = torch.cuda.Event()
    ev.record(cur)
    torch.cuda.ExternalStream(sp).wait_event(ev)


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

    def __init__(self, H, W, world, rank, device, mode="interleave", channels=3, group=1):
        self.H, self.W, self.world, self.rank, self.mode = H, W, world, rank, mode
        self.group = group if mode == "interleave" else 1
        self.pad = sharding.padded_rows(H, world, mode, self.group)
        self.row_begin, self.row_step, self.row_count = sharding.rows_for_rank(H, rank, world, mode, self.group)
        self.band = torch.zeros((self.pad, W, channels), dtype=torch.uint8, device=device)
        self.gathered = torch.empty((world, self.pad, W, channels), dtype=torch.uint8, device=device)
        self.frame = torch.empty((H, W, channels), dtype=torch.uint8, device=device)

    def apply(self, params):
        """Set the sharding fields of an rt_params for this rank."""
        params.row_begin, params.row_step, params.row_count, params.row_group = self.row_begin, self.row_step, self.row_count, self.group
        return params

    def gather(self, scene=None):
        """All-gather the bands; returns the assembled [H, W, C] frame (valid on every rank). `scene`: the scene that rendered
        the band — the collective (torch's current stream) is ordered after its stream."""
        if scene is not None:
            _order_after_scene(scene)
        if self.world > 1:
            try:
                dist.all_gather_into_tensor(self.gathered.view(-1), self.band.view(-1))
            except (RuntimeError, NotImplementedError):  # backends without the flat form (CPU tests)
                parts = [self.gathered[r] for r in range(self.world)]
                dist.all_gather(parts, self.band)
        else:
            self.gathered[0].copy_(self.band)
        G = self.group
        for r in range(self.world):
            b, s, c = sharding.rows_for_rank(self.H, r, self.world, self.mode, G)
            full, rest = c // G, c % G
            if full:  # whole groups: one strided copy (the frame seen as [groups, G rows])
                dst = self.frame.as_strided((full, G) + tuple(self.frame.shape[1:]), (s * self.frame.stride(0), self.frame.stride(0)) + tuple(self.frame.stride()[1:]),
                                            self.frame.storage_offset() + b * self.frame.stride(0))
                dst.copy_(self.gathered[r, :full * G].view((full, G) + tuple(self.frame.shape[1:])))
            if rest:
                self.frame[b + full * s:b + full * s + rest].copy_(self.gathered[r, full * G:c])
        return self.frame


class FramePush:
    """Row-interleaved rendering of H x W frames on `world` ranks with the bands pushed into rank `dst`'s frame buffer over
    NVLink (CUDA IPC + peer copies; NCCL backend, one process per GPU of one node). `frame` is valid on rank `dst`.

    signal="flags" (default when the driver allows it): completion travels as stream-ordered 32-bit writes into the destination's memory
    (rt_peer_signal / rt_peer_wait): rank r writes the frame number to flag r behind its copy, the destination's stream waits for every flag;
    the destination acknowledges a frame it has consumed in an `ack` word that the peers' streams wait on before they overwrite the buffer
    of two frames ago. No kernel, no collective, no host round trip per frame. signal="allreduce": one NCCL all-reduce per frame."""

    def __init__(self, scene, H, W, world, rank, device, dst=0, channels=3, group=1, signal="flags"):
        from . import api
        self.scene, self.H, self.W, self.world, self.rank, self.dst, self.channels = scene, H, W, world, rank, dst, channels
        self.group = group
        self.device_index = device.index if device.index is not None else torch.cuda.current_device()
        self.row_begin, self.row_step, self.row_count = sharding.rows_for_rank(H, rank, world, "interleave", group)
        self.band = torch.zeros((max(self.row_count, 1), W, channels), dtype=torch.uint8, device=device)
        handle = torch.zeros(64, dtype=torch.uint8, device=device)
        self._own = self._peer = None
        # two frames alternate: the destination may still be reading frame k while the peers push frame k + 1
        self.frame_bytes = H * W * channels
        self.flags_off = (2 * self.frame_bytes + 255) & ~255  # world completion flags + one ack word behind the two frames
        self.k = 0
        if rank == dst:
            self._own, hb = api.peer_alloc(self.device_index, self.flags_off + 4 * (world + 1))
            handle.copy_(torch.frombuffer(bytearray(hb), dtype=torch.uint8))
            zeros = torch.zeros(world + 1, dtype=torch.int32, device=device)
            scene.push_rows(zeros.data_ptr(), self._own + self.flags_off, world + 1, 4, 0, 1, 1)
            scene.sync()
        if world > 1:
            dist.broadcast(handle, dst)
        if rank == dst:
            self._base = self._own
        else:
            self._peer = api.peer_open(self.device_index, bytes(handle.cpu().numpy().tobytes()))
            self._base = self._peer
        self.frame_ptr = self._base
        self._token = torch.zeros(1, dtype=torch.int32, device=device)
        # every rank must take the same path: try a harmless flag operation (value 0 to / against this rank's own flag), agree on the outcome
        self.signal = "allreduce"
        if signal == "flags" and world > 1:
            ok = 1
            try:
                scene.peer_wait(self._base + self.flags_off + 4 * rank, 0)
                if rank != dst:
                    scene.peer_signal(self._base + self.flags_off + 4 * rank, 0)
                scene.sync()
            except api.RtError:
                ok = 0
            t = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 1:
                self.signal = "flags"

    def apply(self, params):
        params.row_begin, params.row_step, params.row_count, params.row_group = self.row_begin, self.row_step, self.row_count, self.group
        return params

    def push(self, release=True):
        """Enqueue this rank's copy (on the scene's stream) and the completion signal that tells `dst` every band has landed.
        release (destination, flags mode): acknowledge the frame at once — pass False and call release() after enqueueing the work that
        reads the frame, so that the peers do not overwrite it two frames later while it is still being read."""
        self.frame_ptr = self._base + (self.k & 1) * self.frame_bytes
        self.k += 1
        k = self.k  # frame number, from 1
        flags = self._base + self.flags_off
        if self.signal == "flags" and self.rank != self.dst and k > 2:
            self.scene.peer_wait(flags + 4 * self.world, k - 2)  # the destination has consumed the frame that lived in this buffer
        self.scene.push_rows(self.band.data_ptr(), self.frame_ptr, self.W, self.channels, self.row_begin, self.row_step, self.row_count, self.group)
        if self.world == 1:
            return
        if self.signal == "flags":
            if self.rank != self.dst:
                self.scene.peer_signal(flags + 4 * self.rank, k)
            else:
                for r in range(self.world):
                    if r != self.dst:
                        self.scene.peer_wait(flags + 4 * r, k)
                if release:
                    self.release()
        else:
            _order_after_scene(self.scene)   # the collective runs on torch's current stream: behind the copy on the scene's
            dist.all_reduce(self._token)     # complete when every rank's copy is
            _order_scene_after_current(self.scene)  # what the destination enqueues next on the scene's stream sees every band

    def release(self):
        """Destination, flags mode: everything enqueued so far on the scene's stream has read the current frame; its buffer may be reused."""
        if self.signal == "flags" and self.rank == self.dst and self.world > 1:
            self.scene.peer_signal(self._base + self.flags_off + 4 * self.world, self.k)

    def frame_tensor(self):
        """rank `dst` only: a copy of the assembled frame as a torch tensor (enqueued on the scene's stream)."""
        assert self.rank == self.dst
        out = torch.empty((self.H, self.W, self.channels), dtype=torch.uint8, device=self.band.device)
        self.scene.push_rows(self.frame_ptr, out.data_ptr(), self.W, self.channels, 0, 1, self.H)
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
