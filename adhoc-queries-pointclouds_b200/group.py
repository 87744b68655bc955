"""Python twin of the multi-GPU part of the C ABI (include/pcq.h: pcq_group_*, pcq_dataset_*, pcq_result_*).

Harness-side only (tests, bench.py): every call goes straight to libpcq.so.  A `Group` is either ONE process driving
n GPUs (`Group.local(n)`) or one process per GPU (`Group.from_torch_distributed()`: the NCCL id made by rank 0 travels
through torch.distributed's store / broadcast, the communicator itself is the library's own).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import binding as B
from .binding import check, lib
from .searcher import Context, DeviceFile, Searcher


def shard_plan(points_per_file: Sequence[int], world: int, mode: int = B.SHARD_RANGES) -> np.ndarray:
    """pcq_shard_plan (host only): -> array of SHARD_DTYPE ordered by (file, rank)."""
    n = len(points_per_file)
    ppf = (C.c_uint64 * max(n, 1))(*[int(v) for v in points_per_file])
    n_out = C.c_uint64()
    check(lib.pcq_shard_plan(ppf, n, int(world), int(mode), None, 0, C.byref(n_out)))
    out = np.zeros(int(n_out.value), dtype=B.SHARD_DTYPE)
    check(lib.pcq_shard_plan(ppf, n, int(world), int(mode), C.c_void_p(out.ctypes.data), out.shape[0], C.byref(n_out)))
    return out


class _MemberContext(Context):
    """A context owned by a group (never destroyed from Python)."""

    def __init__(self, handle, device):  # noqa: D401 - no pcq_ctx_create
        self.handle = C.c_void_p(handle)
        self.device = device

    def close(self):
        self.handle = None


class Result:
    def __init__(self, handle, kind: int):
        self.handle = C.c_void_p(handle)
        self.kind = kind

    def counts(self) -> list:
        p = C.POINTER(C.c_uint64)()
        n = C.c_uint32()
        check(lib.pcq_result_counts(self.handle, C.byref(p), C.byref(n)))
        return [int(p[i]) for i in range(n.value)]

    def point_count(self) -> int:
        return sum(self.counts())

    def points(self, lane: int = 0) -> Optional[np.ndarray]:
        if self.kind == B.COLLECT_COUNT:
            return None
        p, n = C.c_void_p(), C.c_uint64()
        check(lib.pcq_result_points(self.handle, int(lane), C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=B.POINT_DTYPE)
        raw = (C.c_uint8 * (31 * n.value)).from_address(p.value)
        return np.frombuffer(raw, dtype=B.POINT_DTYPE).copy()

    def release(self):
        if self.handle:
            lib.pcq_result_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Dataset:
    def __init__(self, group: "Group", handle, keepalive=None):
        self.group = group
        self.handle = C.c_void_p(handle)
        self._keepalive = keepalive

    def release(self):
        if self.handle:
            lib.pcq_dataset_release(self.handle)
            self.handle = None
        self._keepalive = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Group:
    def __init__(self, handle):
        self.handle = C.c_void_p(handle)
        self.world = int(lib.pcq_group_world(self.handle))
        self.n_local = int(lib.pcq_group_local_count(self.handle))
        self.local_ranks = [int(lib.pcq_group_local_rank(self.handle, i)) for i in range(self.n_local)]

    # ---- construction --------------------------------------------------------------------------------------------
    @classmethod
    def local(cls, n_gpus: int, devices: Optional[Sequence[int]] = None) -> "Group":
        """One process, n GPUs (pcq_group_create)."""
        h = C.c_void_p()
        devs = (C.c_int * n_gpus)(*[int(d) for d in devices]) if devices is not None else None
        check(lib.pcq_group_create(devs, int(n_gpus), C.byref(h)))
        g = cls(h.value)
        g._devices = [int(d) for d in devices] if devices is not None else list(range(n_gpus))
        return g

    @classmethod
    def from_torch_distributed(cls, device: int) -> "Group":
        """One process per GPU under torchrun: rank 0 makes the NCCL id, torch.distributed only carries its bytes."""
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        ident = [None]
        if rank == 0 and world > 1:
            buf = (C.c_uint8 * B.GROUP_ID_BYTES)()
            check(lib.pcq_group_unique_id(buf))
            ident = [bytes(buf)]
        if world > 1:
            dist.broadcast_object_list(ident, src=0)
        h = C.c_void_p()
        idb = (C.c_uint8 * B.GROUP_ID_BYTES).from_buffer_copy(ident[0]) if world > 1 else None
        check(lib.pcq_group_create_rank(int(device), rank, world, idb, C.byref(h)))
        g = cls(h.value)
        g._devices = [int(device)]
        return g

    def ctx(self, local_index: int = 0) -> Context:
        h = lib.pcq_group_ctx(self.handle, int(local_index))
        return _MemberContext(h, self._device_of(local_index))

    def _device_of(self, local_index: int) -> int:
        return getattr(self, "_devices", list(range(self.n_local)))[local_index]

    @property
    def launch_count(self) -> int:
        return int(lib.pcq_group_launch_count(self.handle))

    def synchronize(self):
        check(lib.pcq_group_synchronize(self.handle))

    @property
    def last_stats(self) -> dict:
        st = B.GroupStats()
        check(lib.pcq_group_last_stats(self.handle, C.byref(st)))
        return {n: getattr(st, n) for n, _ in B.GroupStats._fields_}

    # ---- datasets ------------------------------------------------------------------------------------------------
    def stage_host_files(self, images: Sequence[tuple], shard_mode: int = B.SHARD_RANGES) -> Dataset:
        n = len(images)
        addrs = [B.buffer_address(b) for b, _ in images]
        ptrs = (C.c_void_p * n)(*[a for a, _ in addrs])
        sizes = (C.c_size_t * n)(*[s for _, s in addrs])
        exts = (C.c_char_p * n)(*[e.encode() for _, e in images])
        h = C.c_void_p()
        check(lib.pcq_group_stage_host_files(self.handle, ptrs, sizes, exts, n, int(shard_mode), C.byref(h)))
        return Dataset(self, h.value)

    def wrap_files(self, points_per_file: Sequence[int], pieces: Sequence[tuple]) -> Dataset:
        """pieces: (DeviceFile made on self.ctx(local member), local member, file index)."""
        nf, n = len(points_per_file), len(pieces)
        ppf = (C.c_uint64 * max(nf, 1))(*[int(v) for v in points_per_file])
        fh = (C.c_void_p * max(n, 1))(*[p[0].handle for p in pieces])
        lm = (C.c_uint32 * max(n, 1))(*[int(p[1]) for p in pieces])
        fi = (C.c_uint32 * max(n, 1))(*[int(p[2]) for p in pieces])
        h = C.c_void_p()
        check(lib.pcq_group_wrap_files(self.handle, ppf, nf, fh, lm, fi, n, C.byref(h)))
        return Dataset(self, h.value, keepalive=[p[0] for p in pieces])

    # ---- searches ------------------------------------------------------------------------------------------------
    @staticmethod
    def _grid_args(kind, grid):
        if kind != B.COLLECT_GRID:
            return None, None, 0.0
        gmin, gmax, cell = grid
        return B.d3(gmin), B.d3(gmax), float(cell)

    def search(self, dataset: Dataset, searchers: Sequence[Searcher], kind: int, per_file: bool, grid=None) -> list:
        """pcq_group_search: one Result per searcher."""
        nq = len(searchers)
        qs = (B.Query * nq)(*[s._query() for s in searchers])
        out = (C.c_void_p * nq)()
        a, b, cell = self._grid_args(kind, grid)
        check(lib.pcq_group_search(self.handle, dataset.handle, qs, nq, int(kind), a, b, cell, int(bool(per_file)), out))
        return [Result(out[i], kind) for i in range(nq)]

    def search_host_files(self, images: Sequence[tuple], searchers: Sequence[Searcher], kind: int, per_file: bool, grid=None,
                          shard_mode: int = B.SHARD_RANGES) -> list:
        """pcq_group_search_host_files: one Result per searcher; `images` = (buffer or (address, nbytes), ext)."""
        n, nq = len(images), len(searchers)
        addrs = [b if isinstance(b, tuple) else B.buffer_address(b) for b, _ in images]
        ptrs = (C.c_void_p * n)(*[a for a, _ in addrs])
        sizes = (C.c_size_t * n)(*[s for _, s in addrs])
        exts = (C.c_char_p * n)(*[e.encode() for _, e in images])
        qs = (B.Query * nq)(*[s._query() for s in searchers])
        out = (C.c_void_p * nq)()
        a, b, cell = self._grid_args(kind, grid)
        check(lib.pcq_group_search_host_files(self.handle, ptrs, sizes, exts, n, qs, nq, int(kind), a, b, cell,
                                              int(bool(per_file)), int(shard_mode), out))
        return [Result(out[i], kind) for i in range(nq)]

    def close(self):
        if self.handle:
            lib.pcq_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
