"""One process per GPU: interleaved row tiles over a replicated scene, framebuffer gathered to rank 0.

Pixels are independent (src/raytrace.cpp:228-251), so the frame shards with no data-path exchange except
the final gather: row tile t (tile_rows rows) belongs to rank t % world; every rank renders ALL samples of
its pixels (the per-pixel (jj,ii) sum order is kept, so the image is bit-identical for any world size).
torch.distributed is plumbing only: NCCL gather over NVLink on GPUs, gloo in the CPU tests of this logic.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


def n_tiles(height: int, tile_rows: int) -> int:
    return (height + tile_rows - 1) // tile_rows


def rows_owned(height: int, tile_rows: int, rank: int, world: int) -> int:
    """Number of image rows rank renders (mirrors yrt_rows_owned)."""
    rows = 0
    for t in range(rank, n_tiles(height, tile_rows), world):
        rows += min(height, (t + 1) * tile_rows) - t * tile_rows
    return rows


def global_rows(height: int, tile_rows: int, rank: int, world: int) -> np.ndarray:
    """Global row index of each packed local row of `rank`, in packed order."""
    out = []
    for t in range(rank, n_tiles(height, tile_rows), world):
        out.extend(range(t * tile_rows, min(height, (t + 1) * tile_rows)))
    return np.asarray(out, np.int64)


def gather_rows(packed: torch.Tensor, width: int, height: int, tile_rows: int, rank: int, world: int,
                group=None) -> Optional[torch.Tensor]:
    """Gather every rank's packed rows (rows_owned x width x 4 float32) on rank 0 and scatter them into the
    (height, width, 4) framebuffer.  Returns the framebuffer on rank 0, None elsewhere."""
    own = rows_owned(height, tile_rows, rank, world)
    assert packed.shape == (own, width, 4) and packed.dtype == torch.float32
    if world == 1:
        full = torch.empty((height, width, 4), dtype=torch.float32, device=packed.device)
        _unpack(packed, full, width, height, tile_rows, 0, 1)
        return full
    max_rows = max(rows_owned(height, tile_rows, r, world) for r in range(world))
    send = packed
    if own < max_rows:   # dist.gather needs equal shapes: pad the short ranks (at most one tile)
        send = torch.zeros((max_rows, width, 4), dtype=torch.float32, device=packed.device)
        send[:own] = packed
    send = send.contiguous()
    if rank == 0:
        bufs = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, gather_list=bufs, dst=0, group=group)
        full = torch.empty((height, width, 4), dtype=torch.float32, device=packed.device)
        for r in range(world):
            n = rows_owned(height, tile_rows, r, world)
            if n:
                _unpack(bufs[r][:n], full, width, height, tile_rows, r, world)
        return full
    dist.gather(send, gather_list=None, dst=0, group=group)
    return None


def _unpack(packed: torch.Tensor, full: torch.Tensor, width, height, tile_rows, rank, world) -> None:
    if packed.is_cuda:
        st = torch.cuda.current_stream(packed.device).cuda_stream
        check(_lib.load().yrt_unpack_rows(packed.contiguous().data_ptr(), full.data_ptr(), width, height, tile_rows, rank, world, st))
    else:
        idx = torch.from_numpy(global_rows(height, tile_rows, rank, world))
        full[idx] = packed


def render_sharded(scene, width: int, height: int, samples: int, amb=0.1, tile_rows: int = 1, group=None, want_stats=False):
    """Render this rank's tiles with `scene` (a render.Scene bound to this process's GPU) and gather on rank 0.
    Returns (framebuffer on rank 0 | None, Stats | None)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    own = rows_owned(height, tile_rows, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    packed = torch.empty((max(own, 1), width, 4), dtype=torch.float32, device=dev)[:own]
    st = torch.cuda.current_stream(dev).cuda_stream
    stats = scene.render_rows_into(packed.data_ptr() if own else torch.empty(4, device=dev).data_ptr(), width, height, samples, amb,
                                   tile_rows, rank, world, st, want_stats)
    full = gather_rows(packed, width, height, tile_rows, rank, world, group)
    return full, stats


class SharedFrame:
    """The full framebuffer in rank 0's HBM, mapped into every other rank over NVLink (CUDA IPC).  Every rank's
    resolve kernel stores its rows at their final position, so the per-frame exchange is the stores themselves plus one
    barrier — no gather copy, no unpack, no collective: the barrier is a one-thread kernel per rank that adds its arrival
    to a counter in the same peer memory and waits for the others (yrt_frame_barrier).  Frames alternate between two
    buffers, which is what makes ONE barrier per frame enough (see include/yrt_b200.h).  barrier="nccl" keeps round 1's
    protocol (one buffer, two 1-element all-reduces per frame) for systems without peer atomics."""

    def __init__(self, width: int, height: int, group=None, barrier: str = "peer"):
        import ctypes as C
        self.width, self.height, self.group = width, height, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.barrier = barrier
        self._frames = 0
        lib = _lib.load()
        self._ptr = C.c_void_p()
        self._owner = self.rank == 0
        handle = C.create_string_buffer(64)
        err = None
        if self._owner:
            try:
                check(lib.yrt_frame_alloc(width, height, C.byref(self._ptr)))
                if self.world > 1:
                    check(lib.yrt_frame_export(self._ptr, handle))
            except Exception as e:          # the other ranks are waiting in the broadcast: tell them before raising
                err = e
        if self.world > 1:
            box = [(handle.raw if err is None else None, None if err is None else str(err)) if self._owner else None]
            dist.broadcast_object_list(box, src=0, group=group)
            raw, msg = box[0]
            if raw is None:
                raise err if err is not None else RuntimeError(f"rank 0 could not share the frame: {msg}")
            if not self._owner:
                check(lib.yrt_frame_import(raw, C.byref(self._ptr)))
        elif err is not None:
            raise err
        self._token = torch.zeros(1, device=torch.device("cuda", torch.cuda.current_device()))

    @property
    def ptr(self) -> int:
        return self._ptr.value

    def _buffer(self, frame: int) -> int:
        return self.ptr + (frame % 2 if self.barrier == "peer" else 0) * self.width * self.height * 16

    def tensor(self) -> torch.Tensor:
        """Rank 0: the last rendered frame as a (height, width, 4) float32 CUDA tensor view (no copy); valid until the
        next-but-one render call (read it on the current stream, or finish reading before the next render call)."""
        assert self._owner
        iface = {"shape": (self.height, self.width, 4), "typestr": "<f4", "data": (self._buffer(max(self._frames - 1, 0)), False), "version": 3, "strides": None}
        holder = type("_Frame", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=torch.device("cuda", torch.cuda.current_device()))

    def render(self, scene, samples: int, amb=0.1, tile_rows: int = 1, want_stats: bool = False):
        """Render this rank's rows into the current buffer of the shared frame, then the frame's barrier — all enqueued on the
        current stream, no host synchronisation.  Frame k's buffer is written again in frame k + 2, after barrier k + 1, which
        rank 0 joins only after the reads of frame k it enqueued before calling render(k + 1)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        st = torch.cuda.current_stream(dev).cuda_stream
        if self.barrier != "peer":
            if self.world > 1:
                dist.all_reduce(self._token, group=self.group)
            stats = scene.render_rows_into_frame(self.ptr, self.width, self.height, samples, amb, tile_rows, self.rank, self.world, st, want_stats)
            if self.world > 1:
                dist.all_reduce(self._token, group=self.group)
            self._frames += 1
            return stats
        stats = scene.render_rows_into_frame(self._buffer(self._frames), self.width, self.height, samples, amb, tile_rows, self.rank, self.world, st, want_stats)
        self._frames += 1
        if self.world > 1:
            check(_lib.load().yrt_frame_barrier(self.ptr, self.width, self.height, self.world, self._frames, st))
        return stats

    def close(self):
        lib = _lib.load()
        if self._ptr:
            if self._owner:
                lib.yrt_frame_free(self._ptr)
            else:
                lib.yrt_frame_release(self._ptr)
            self._ptr = None


class SharedHostFrame:
    """The full framebuffer in HOST memory shared by all ranks of the node (POSIX shared memory, page-locked by every rank):
    each rank renders its interleaved rows and copies them itself into their final positions — pitched device->host copies
    per rank over that GPU's own PCIe link, all ranks at once — so the frame reaches the host N times faster than through
    rank 0's single link, and the GPUs exchange nothing.

    Frames alternate between TWO buffers and every frame ends with one barrier (a counter in the same shared memory,
    yrt_host_barrier: microseconds, no collective).  That orders everything: render(k) returns on every rank after all rows
    of frame k have landed; the consumer (rank 0) reads frame k and then calls render(k+1); a rank can only write into frame
    k's buffer again in render(k+2), i.e. after the barrier of frame k+1, which rank 0 joins after it is done with frame k."""

    def __init__(self, width: int, height: int, group=None):
        from multiprocessing import shared_memory
        self.width, self.height, self.group = width, height, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._fbytes = width * height * 16
        nbytes = 2 * self._fbytes + 64
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            self._shm.buf[2 * self._fbytes: 2 * self._fbytes + 64] = bytes(64)      # the barrier counter starts at zero
            name[0] = self._shm.name
        if self.world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
            if self.rank != 0:
                self._shm = shared_memory.SharedMemory(name=name[0])
                try:      # Python < 3.13 registers attached segments for unlinking at process exit; only the owner (rank 0) unlinks
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self._shm._name, "shared_memory")
                except Exception:
                    pass
        self._frames_np = [np.ndarray((height, width, 4), np.float32, buffer=self._shm.buf, offset=k * self._fbytes) for k in range(2)]
        self._ctr = np.ndarray((8,), np.int64, buffer=self._shm.buf, offset=2 * self._fbytes)
        self._base = self._frames_np[0].ctypes.data
        self._pinned = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self._base, 2 * self._fbytes, 1)   # cudaHostRegisterPortable
            self._pinned = int(rc) == 0
        self._frames = 0
        self.array = self._frames_np[0]
        if self.world > 1:
            dist.barrier(group=group)     # every rank has mapped (and page-locked) the memory before the first frame

    def render(self, scene, samples: int, amb=0.1, tile_rows: int = 1, want_stats: bool = False):
        """Render this rank's rows into the current buffer; returns after every rank's rows of this frame have landed.
        self.array is the frame (valid until the next-but-one render call)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        st = torch.cuda.current_stream(dev).cuda_stream
        buf = self._frames_np[self._frames % 2]
        stats = scene.render_rows_to_host(buf.ctypes.data, self.width, self.height, samples, amb, tile_rows, self.rank, self.world, st, want_stats)
        torch.cuda.current_stream(dev).synchronize()
        self._frames += 1
        if self.world > 1:
            check(_lib.load().yrt_host_barrier(self._ctr.ctypes.data, self.world, self._frames))
        self.array = buf
        return stats

    def close(self):
        try:
            if self._pinned:
                torch.cuda.cudart().cudaHostUnregister(self._base)
            self.array = None
            self._frames_np = None
            self._ctr = None
            self._shm.close()
            if self.rank == 0:
                self._shm.unlink()
        except Exception:
            pass
