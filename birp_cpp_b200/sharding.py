"""How the path shards over the GPUs of one box: every pixel and every frame is
independent, the cube is replicated, and the only exchange is the gather of finished images to
rank 0.  One process per GPU; torch.distributed is the plumbing (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def block_counts(total: int, world: int) -> List[int]:
    """contiguous blocks, the first `total % world` ranks take one extra unit"""
    return [total // world + (1 if r < total % world else 0) for r in range(world)]


def block_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    counts = block_counts(total, world)
    start = sum(counts[:rank])
    return start, start + counts[rank]


def frame_block(nframes: int, world: int, rank: int) -> Tuple[int, int]:
    """orbit sweep: frames [begin, end) rendered by `rank`"""
    return block_range(nframes, world, rank)


def row_band(height: int, world: int, rank: int) -> Tuple[int, int]:
    """single large image: rows [begin, end) rendered by `rank` (birp_render_* row_begin/row_end)"""
    return block_range(height, world, rank)


def interleaved_row_bands(height: int, band_rows: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """single large image in INTERLEAVED bands (birp_render_cube_rows / birp_render_cube_device_rows): of the bands of
    `band_rows` rows counted from the top, `rank` renders those with index = rank (mod world); the last band of the
    image may be short.  Returns the [begin, end) row ranges in image order."""
    out = []
    nbands = (height + band_rows - 1) // band_rows
    for b in range(rank, nbands, world):
        out.append((b * band_rows, min(height, (b + 1) * band_rows)))
    return out


class SharedHostResult:
    """ONE host buffer for the result of a job that several processes of a box render (rank 0 owns a POSIX
    shared-memory segment, every rank maps it): each rank's sweep / row-band call writes its share straight into it
    (birp_render_*_sweep_strided, birp_render_cube_rows), every GPU over its own PCIe link.  `register` page-locks
    the mapping with the CUDA driver (birp_host_register); on CPU-only ranks (tests) it is left out."""

    def __init__(self, shape, rank: int, world: int, register: bool = True):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.rank, self.world = rank, world
        nbytes = int(np.prod(shape)) * 4
        name = [None]
        self.shm = None
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=max(nbytes, 4))
            name[0] = self.shm.name
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        self.array = np.ndarray(shape, dtype=np.float32, buffer=self.shm.buf)
        if rank == 0:
            self.array[...] = 0
        if world > 1:
            dist.barrier()
        self._addr = self.array.ctypes.data
        self._registered = False
        if register:
            from . import capi
            capi.host_register(self.array)
            self._registered = True

    def close(self):
        import ctypes
        import torch.distributed as dist
        self.array = None
        if self.world > 1:
            dist.barrier()  # nobody unmaps while a peer may still be writing
        if self._registered:
            from . import capi
            capi.lib().birp_host_unregister(ctypes.c_void_p(self._addr))
            self._registered = False
        try:
            self.shm.close()
        except BufferError:
            pass
        if self.rank == 0:
            self.shm.unlink()


def broadcast_cube(cube_tensor, geom_tensor, src: int = 0):
    """replicate the grid and its origin/spacing from `src` to every rank (one-time)"""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(cube_tensor, src)
        dist.broadcast(geom_tensor, src)
    return cube_tensor, geom_tensor


def _gather_uneven(local, recv_views, dst: int):
    """rank r sends `local` to `dst`; `dst` receives straight into recv_views[r] (views of the output).
    Grouped point-to-point ops: ncclSend/ncclRecv inside one ncclGroup under NCCL, no padding."""
    import torch.distributed as dist
    rank = dist.get_rank()
    ops = []
    if rank == dst:
        recv_views[dst].copy_(local)
        for r, view in enumerate(recv_views):
            if r != dst and view.numel():
                ops.append(dist.P2POp(dist.irecv, view, r))
    elif local.numel():
        ops.append(dist.P2POp(dist.isend, local, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def gather_frames(local_images, counts: Sequence[int], dst: int = 0, out=None):
    """local_images: [counts[rank]][H][W] tensor.  Returns the [sum(counts)][H][W] tensor on `dst`
    (None elsewhere).  Blocks may be uneven.  `out`: optional preallocated result on `dst`."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_images
    rank = dist.get_rank()
    shape = tuple(local_images.shape[1:])
    views = None
    if rank == dst:
        if out is None:
            out = torch.empty((sum(counts),) + shape, dtype=local_images.dtype, device=local_images.device)
        offs = [sum(counts[:r]) for r in range(len(counts))]
        views = [out[offs[r]:offs[r] + counts[r]] for r in range(len(counts))]
    _gather_uneven(local_images.contiguous(), views, dst)
    return out if rank == dst else None


def gather_row_bands(local_band, bands: Sequence[int], dst: int = 0, out=None):
    """local_band: [bands[rank]][W] rows of ONE image (contiguous).  Returns the [H][W] image on `dst`."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_band
    rank = dist.get_rank()
    W = local_band.shape[-1]
    views = None
    if rank == dst:
        if out is None:
            out = torch.empty((sum(bands), W), dtype=local_band.dtype, device=local_band.device)
        offs = [sum(bands[:r]) for r in range(len(bands))]
        views = [out[offs[r]:offs[r] + bands[r]] for r in range(len(bands))]
    _gather_uneven(local_band.contiguous(), views, dst)
    return out if rank == dst else None


def frame_indices(nframes: int, world: int, rank: int, interleave: bool = True):
    """Frames of a sweep owned by `rank`.  Interleaved (frame f -> rank f % world) when the sweep divides
    evenly: neighbouring poses cost about the same, so every rank gets the same mix of cheap and
    expensive frames; contiguous blocks otherwise."""
    import numpy as np
    if interleave and world > 1 and nframes % world == 0:
        return np.arange(rank, nframes, world)
    b, e = frame_block(nframes, world, rank)
    return np.arange(b, e)


class FrameGather:
    """Gather of a sweep's images to rank `dst` in pieces, so that the transfer of piece c overlaps the
    render of piece c+1 (the NCCL point-to-point work runs on its own stream; nothing waits on it until
    finish()).  Works for both frame layouts of frame_indices()."""

    def __init__(self, nframes: int, world: int, rank: int, frame_shape, dtype, device, interleave: bool = True,
                 nchunks: int = 2, dst: int = 0):
        import torch
        self.world, self.rank, self.dst, self.nframes = world, rank, dst, nframes
        self.interleaved = bool(interleave and world > 1 and nframes % world == 0)
        self.counts = [len(frame_indices(nframes, world, r, interleave)) for r in range(world)]
        self.nchunks = max(1, min(nchunks, min(self.counts) if min(self.counts) > 0 else 1))
        self.works = []
        self.buffer = None
        if rank == dst and world > 1:
            # rank-major staging: frames of rank r at [offs[r], offs[r] + counts[r])
            self.buffer = torch.empty((nframes,) + tuple(frame_shape), dtype=dtype, device=device)
        self.offs = [sum(self.counts[:r]) for r in range(world)]

    def chunk_bounds(self, r: int, c: int):
        n = self.counts[r]
        return n * c // self.nchunks, n * (c + 1) // self.nchunks

    def post(self, local_images, c: int):
        """piece c of this rank's images (local order) has been enqueued on the current stream"""
        import torch.distributed as dist
        if self.world == 1:
            return
        ops = []
        lo, hi = self.chunk_bounds(self.rank, c)
        if self.rank == self.dst:
            o = self.offs[self.dst]
            self.buffer[o + lo:o + hi].copy_(local_images[lo:hi], non_blocking=True)
            for r in range(self.world):
                if r == self.dst:
                    continue
                rlo, rhi = self.chunk_bounds(r, c)
                if rhi > rlo:
                    ops.append(dist.P2POp(dist.irecv, self.buffer[self.offs[r] + rlo:self.offs[r] + rhi], r))
        elif hi > lo:
            ops.append(dist.P2POp(dist.isend, local_images[lo:hi], self.dst))
        if ops:
            self.works.extend(dist.batch_isend_irecv(ops))

    def finish(self):
        """wait for every posted piece; on `dst` returns the images in sweep (time) order -- a strided view
        of the rank-major staging buffer for the interleaved layout, the buffer itself for blocks"""
        for w in self.works:
            w.wait()
        self.works = []
        if self.world == 1 or self.rank != self.dst:
            return None
        if not self.interleaved:
            return self.buffer
        per = self.nframes // self.world
        shape = tuple(self.buffer.shape[1:])
        # [per][world][...] view: element (j, r) is frame j * world + r; no copy is made
        return self.buffer.view((self.world, per) + shape).transpose(0, 1)


class PeerGather:
    """Gather without a collective: rank `dst` owns the result buffer, every other rank maps it into its
    own address space (CUDA IPC) and passes that pointer to the render call, so the kernel's final pixel
    stores go straight over NVLink into the gathering rank's memory (4 bytes per ray against ~20 000
    instructions of march: the transfer hides completely under the compute).  fence() orders "every rank's
    kernel has finished" on the stream: each rank's stream stores the step number into its slot of a flag array in
    `dst`'s memory (birp_peer_signal), `dst`'s stream waits for all slots (birp_peer_wait) -- 5 us instead of the
    30-40 us of a one-element NCCL all-reduce (BIRP_FENCE=nccl keeps that form)."""

    def __init__(self, nfloats: int, rank: int, world: int, device, dst: int = 0):
        import os
        import torch
        import torch.distributed as dist

        from . import capi
        self.rank, self.world, self.dst, self.nfloats = rank, world, dst, nfloats
        self.buf = capi.DeviceBuffer(nfloats) if rank == dst else None
        self.flags = capi.DeviceBuffer(32) if rank == dst else None
        obj = [(self.buf.export_handle(), self.flags.export_handle()) if rank == dst else None]
        if world > 1:
            dist.broadcast_object_list(obj, src=dst)
            if rank != dst:
                self.buf = capi.DeviceBuffer.from_handle(obj[0][0], nfloats)
                self.flags = capi.DeviceBuffer.from_handle(obj[0][1], 32)
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._device = device
        self._step = 0
        self._nccl_fence = os.environ.get("BIRP_FENCE", "flags") == "nccl"
        if rank == dst:
            torch.as_tensor(self.flags, device=device).zero_()
            torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()

    @property
    def ptr(self) -> int:
        return self.buf.ptr

    def fence(self):
        import torch
        import torch.distributed as dist
        if self.world <= 1:
            return
        if self._nccl_fence:
            dist.all_reduce(self._flag)
            return
        from . import capi
        stream = torch.cuda.current_stream().cuda_stream
        self._step += 1
        capi.peer_signal(self.flags.ptr + 4 * self.rank, self._step, stream)
        if self.rank == self.dst:
            capi.peer_wait(self.flags.ptr, self.world, self._step, stream)

    def tensor(self, shape):
        """the gathered images as a torch tensor (on `dst` only)"""
        import torch
        if self.rank != self.dst:
            return None
        return torch.as_tensor(self.buf, device=self._device).view(shape)

    def close(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()  # nobody unmaps / frees while a peer may still be writing
        if self.buf is not None:
            self.buf.free()
            self.buf = None
        if self.flags is not None:
            self.flags.free()
            self.flags = None
