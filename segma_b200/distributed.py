"""Multi-GPU sharding of the path: one process per GPU, files partitioned statically, no collective on the
data path, one exchange step at the end -- an all-gather of the int32 interval tables.

The reference is single-process, single-device (inference.py:442-458); this module is new work.  The
atomic work item is a file (the LSTM couples the windows of a forward call, so a file's batches stay on
one rank); files are assigned longest-processing-time-first so ranks finish together.
"""
from __future__ import annotations

import heapq
import os

import torch
import torch.distributed as dist


def assign_files(sizes: list[int], world_size: int) -> list[list[int]]:
    """Longest-processing-time-first assignment of file indices to ranks; deterministic on every rank.
    Each rank's list is returned in ascending file order."""
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out: list[list[int]] = [[] for _ in range(world_size)]
    for idx in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        load, r = heapq.heappop(heap)
        out[r].append(idx)
        heapq.heappush(heap, (load + max(int(sizes[idx]), 1), r))
    return [sorted(v) for v in out]


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun contract).
    Returns (rank, world_size, local_rank); a single process needs no initialisation."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


class UnitQueue:
    """Work units handed out on demand instead of up front: every rank walks the same longest-first order and claims
    the next unit with an atomic add on the process group's key-value store (one ~0.1 ms round trip per unit, no
    collective).  Boards under the same power cap differ by a few percent in speed; with static shares a run ends when
    the slowest board finishes, with claims the faster boards simply take more units.  ``name`` must be the same on
    every rank and unique per walk (``infer_corpus`` numbers its calls)."""

    def __init__(self, name: str, n_units: int, store=None):
        if store is None:
            from torch.distributed.distributed_c10d import _get_default_store

            store = _get_default_store()
        self.store, self.key, self.n_units = store, f"segma/unit_queue/{name}", int(n_units)

    def claim(self) -> int | None:
        """Index (into the common order) of the next unclaimed unit, or None when all are taken."""
        i = int(self.store.add(self.key, 1)) - 1
        return i if i < self.n_units else None


def all_gather_tables(table: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather variable-length int32 ``(n_i, 4)`` interval tables: one count exchange, then one padded
    ``all_gather_into_tensor``; rows come back ordered by rank, then in each rank's own order.
    ``table[:, 0]`` must already hold *global* file indices."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return table
    world = dist.get_world_size(group)
    dev = table.device
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n, group=group)
    cmax = int(counts.max().item())
    padded = torch.zeros((max(cmax, 1), 4), dtype=torch.int32, device=dev)
    padded[: table.shape[0]] = table
    gathered = torch.empty((world * max(cmax, 1), 4), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    parts = [gathered[r * max(cmax, 1): r * max(cmax, 1) + int(counts[r])] for r in range(world)]
    return torch.cat(parts, dim=0)


def gather_file_tables(local_table: torch.Tensor, group=None) -> torch.Tensor:
    """``all_gather_tables`` followed by a stable sort on the global file index, so every rank ends up
    with the same table ordered by file, then label, then time."""
    full = all_gather_tables(local_table, group)
    if full.shape[0] == 0:
        return full
    order = torch.sort(full[:, 0].to(torch.int64), stable=True).indices
    return full[order]


def merge_split_files(table: torch.Tensor, merge_fn=None) -> torch.Tensor:
    """Rows of files whose window batches were decoded in several pieces (on several ranks) -> the table of the whole
    files.  Sorted by (file, label, start); a run of active frames that crosses a piece boundary arrives as two intervals
    that touch (end == next start) and is fused by the gap-0 merge of the interval post-processing kernel -- inside one
    piece two intervals of a label never touch (at least one inactive frame lies between them), so nothing else merges."""
    if table.shape[0] == 0:
        return table
    order = torch.sort(table[:, 2], stable=True).indices
    order = order[torch.sort(table[order, 1], stable=True).indices]
    order = order[torch.sort(table[order, 0], stable=True).indices]
    table = table[order].contiguous()
    if merge_fn is None:
        from . import ops

        merge_fn = lambda t: ops.postprocess_intervals(t, 0, 0)  # noqa: E731
    return merge_fn(table)


def gather_corpus_tables(file_indices, tables, counts, device=None, gather: bool = True, group=None,
                         sample_offsets=None, merge_split: bool = False, merge_fn=None) -> torch.Tensor:
    """The single exchange step at the end of a sharded corpus run.

    ``tables[k]`` is the worst-case-sized int32 ``(cap_k, 4)`` table of this rank's k-th work unit (a file, or a range
    of a file's window batches; global file index ``file_indices[k]``, first sample ``sample_offsets[k]``) and
    ``counts[k]`` the 1-element tensor with its valid row count, both still where the decode kernel left them.  Reads all
    counts back at once (the only host synchronisation of the run), compacts, stamps the global file index into column
    0, shifts the pieces of split files to their place on the file timeline and all-gathers; with ``merge_split`` the
    pieces are then fused (``merge_split_files``).  Every rank returns the same table ordered by file, label, time."""
    if device is None:
        device = tables[0].device if tables else torch.device("cpu")
    parts = []
    if tables:
        host_counts = torch.cat([c.reshape(1) for c in counts]).cpu().tolist()
        offs = sample_offsets if sample_offsets is not None else [0] * len(tables)
        for i, t, c, off in zip(file_indices, tables, host_counts, offs):
            part = t[:c].clone()
            part[:, 0] = i
            if off:
                part[:, 2:4] += int(off)
            parts.append(part)
    local = torch.cat(parts) if parts else torch.empty((0, 4), dtype=torch.int32, device=device)
    full = gather_file_tables(local, group) if gather else local
    return merge_split_files(full, merge_fn) if merge_split else full
