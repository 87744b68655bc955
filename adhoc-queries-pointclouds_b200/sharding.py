"""Multi-GPU plumbing for the scan path: one process per GPU, torch.distributed for the exchange.

The reference has exactly one parallelism strategy — one rayon task per file (main.rs:153-161) — and
no communication.  Here files AND point ranges inside files shard across ranks:

  * count / select queries need no collective: every rank scans its ranges, the host sums the counts
    (or concatenates the per-range record streams in range order);
  * the max-density query has one real exchange step: each rank builds a local cell table over its
    ranges, exports ONE candidate per locally occupied cell partitioned by owner = mix64(key) % world,
    the parts travel with one all-to-all over NCCL (NVLink/NVSwitch), and each owner folds what it
    receives into a merge table.  Ties are broken by the GLOBAL scan index (file order, then point
    index), so the result equals the sequential fold of grid_sampling.rs:97-102 over the whole dataset.

`exchange_parts` is backend-agnostic (gloo on CPU tensors in the tests, NCCL on device tensors).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

TILE = 512  # kTilePts: ranges are cut at tile boundaries so every rank's kernels see whole tiles


@dataclass(frozen=True)
class PointRange:
    file: int         # index into the dataset's file list (scan order)
    first_point: int  # first record of the range inside the file
    n_points: int
    scan_base: int    # global scan index of `first_point` (sum of the sizes of all earlier files + first_point)


def plan_point_ranges(points_per_file: Sequence[int], world: int) -> List[List[PointRange]]:
    """Cut the dataset's scan order into `world` contiguous, tile-aligned, near-equal pieces."""
    total = int(sum(points_per_file))
    starts = np.concatenate([[0], np.cumsum(points_per_file)]).astype(np.int64)
    # ideal cut positions in global scan index space, snapped to tile boundaries inside their file
    cuts = [0]
    for r in range(1, world):
        g = total * r // world
        f = int(np.searchsorted(starts, g, side="right") - 1)
        f = min(max(f, 0), len(points_per_file) - 1)
        local = (g - int(starts[f])) // TILE * TILE
        cuts.append(max(cuts[-1], int(starts[f]) + local))
    cuts.append(total)
    plan: List[List[PointRange]] = []
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        ranges = []
        for f, n in enumerate(points_per_file):
            a, b = max(lo, int(starts[f])), min(hi, int(starts[f]) + int(n))
            if b > a:
                ranges.append(PointRange(f, a - int(starts[f]), b - a, a))
        plan.append(ranges)
    return plan


def plan_files(n_files: int, world: int) -> List[List[int]]:
    """Whole-file sharding for count / select queries: contiguous blocks of files per rank."""
    return [list(range(n_files * r // world, n_files * (r + 1) // world)) for r in range(world)]


def gather_selected(ranges: Sequence[PointRange], selected: Sequence[np.ndarray], n_files: int, per_file: bool, group=None):
    """Select (BufferCollector) over sharded point ranges: every rank passes its ranges and the scan-ordered records
    its GPU selected from each; the per-range streams are concatenated on the host in (file, first point) order, which
    is exactly the order one BufferCollector would have seen them in.  No data-path collective — only this gather.

    per_file=False -> one array (run_search_sequential: one collector over all files, main.rs:122-144)
    per_file=True  -> one array per file (run_search_parallel: one collector per file, main.rs:146-183)."""
    import torch.distributed as dist

    mine = [(r.file, r.first_point, np.ascontiguousarray(p).tobytes()) for r, p in zip(ranges, selected)]
    world = dist.get_world_size(group)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    pieces = sorted((item for part in gathered for item in part), key=lambda t: (t[0], t[1]))
    from .binding import POINT_DTYPE

    def cat(items):
        raw = b"".join(b for _, _, b in items)
        return np.frombuffer(raw, dtype=POINT_DTYPE).copy()

    if not per_file:
        return cat(pieces)
    return [cat([it for it in pieces if it[0] == f]) for f in range(n_files)]


def mix64(x: np.ndarray) -> np.ndarray:
    """The owner hash of kernels.cu (murmur3 finaliser) on uint64 arrays."""
    x = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xFF51AFD7ED558CCD)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xC4CEB9FE1A85EC53)
        x ^= x >> np.uint64(33)
    return x


def owner_of(keys: np.ndarray, world: int) -> np.ndarray:
    return (mix64(keys) % np.uint64(world)).astype(np.int64)


def exchange_parts(send, send_counts: Sequence[int], item_bytes: int, group=None):
    """All-to-all of variable-sized parts.  `send` is a flat uint8 tensor holding world parts back to back
    (part p = send_counts[p] items of item_bytes).  -> (recv tensor, recv_counts).  Two collectives: the counts
    (all_to_all_single of int64) and the payload (all_to_all_single with split sizes)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = send.device
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=dev)
    rc = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(v) for v in rc.tolist()]
    recv = torch.empty(sum(recv_counts) * item_bytes, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv, send[: sum(send_counts) * item_bytes],
                           output_split_sizes=[c * item_bytes for c in recv_counts],
                           input_split_sizes=[int(c) * item_bytes for c in send_counts], group=group)
    return recv, recv_counts


class _DevicePtr:
    """zero-copy view of library-owned device memory as a torch tensor (__cuda_array_interface__)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_bytes_as_tensor(ptr: int, nbytes: int, device):
    import torch

    if nbytes == 0 or not ptr:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)


def density_exchange(local_collector, make_collector, group=None):
    """The exchange step of a sharded max-density query.  `local_collector` holds this rank's cells;
    returns a collector holding the final winners of the cells this rank owns."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    ptr, counts = local_collector.export_candidates(world)
    send = device_bytes_as_tensor(ptr, sum(counts) * 64, dev)
    torch.cuda.current_stream().synchronize()
    recv, recv_counts = exchange_parts(send, counts, 64, group)
    torch.cuda.current_stream().synchronize()
    merged = make_collector()
    merged.import_candidates(recv.data_ptr() if recv.numel() else 0, sum(recv_counts))
    merged._keepalive = recv
    return merged
