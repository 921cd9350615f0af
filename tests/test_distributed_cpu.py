"""Multi-rank host logic on CPU: file assignment and the final interval-table all-gather over gloo
(world_size 2), the N > 1 path of SURVEY.md section 8e."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from segma_b200.distributed import all_gather_tables, assign_files, gather_corpus_tables, gather_file_tables


def test_assign_files_is_balanced_and_complete():
    sizes = [100, 90, 80, 10, 10, 10, 5, 5, 1, 0]
    parts = assign_files(sizes, 3)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= max(sizes)
    assert assign_files(sizes, 3) == parts  # deterministic
    assert assign_files([5], 4) == [[0], [], [], []]
    thousand = assign_files([57_600_000] * 1000, 8)
    assert [len(p) for p in thousand] == [125] * 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    files = assign_files([300, 200, 100, 50, 10], world)[rank]
    rows = []
    for f in files:
        n = f + 1 + rank  # different, non-empty counts per file
        rows += [(f, k % 4, 320 * k, 320 * (k + 1)) for k in range(n)]
    table = torch.tensor(rows, dtype=torch.int32).reshape(-1, 4)
    if rank == 1:
        empty = all_gather_tables(torch.zeros((0, 4), dtype=torch.int32))
        assert empty.shape[1] == 4
    else:
        all_gather_tables(torch.zeros((0, 4), dtype=torch.int32))
    full = gather_file_tables(table)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    # end-of-run exchange of a sharded corpus: per-file worst-case tables + device-side counts, one gather
    tables, counts = [], []
    for f in files:
        n = f + 1 + rank
        t = torch.full((n + 7, 4), -1, dtype=torch.int32)  # capacity > count: rows beyond the count are junk
        t[:n] = torch.tensor([(0, k % 4, 320 * k, 320 * (k + 1)) for k in range(n)], dtype=torch.int32)
        tables.append(t)
        counts.append(torch.tensor([n], dtype=torch.int32))
    corpus = gather_corpus_tables(files, tables, counts)
    assert torch.equal(corpus, full)
    # a file cut into batch ranges across the ranks: pieces carry their sample offset, touching runs fuse after the gather
    whole = _split_file_truth()
    pieces = _cut(whole, [0, 6400, 12800, 32000])
    mine = [k for k in range(len(pieces)) if k % world == rank]
    tables, counts = [], []
    for k in mine:
        t = torch.full((len(pieces[k][1]) + 3, 4), -7, dtype=torch.int32)
        t[:len(pieces[k][1])] = torch.tensor(pieces[k][1], dtype=torch.int32).reshape(-1, 4)
        tables.append(t)
        counts.append(torch.tensor([len(pieces[k][1])], dtype=torch.int32))
    merged = gather_corpus_tables([3] * len(mine), tables, counts, sample_offsets=[pieces[k][0] for k in mine],
                                  merge_split=True, merge_fn=_merge_touching)
    assert merged.tolist() == [[3, lab, s, e] for lab, s, e in whole], (merged.tolist(), whole)
    dist.destroy_process_group()


def _split_file_truth():
    """(label, start, end) of one file, sorted by label then time; several runs cross the cuts at 6400 / 12800."""
    return [(0, 0, 640), (0, 3200, 9600), (0, 12160, 12800), (1, 6400, 6720), (1, 9600, 20000), (2, 320, 31680), (3, 12800, 13120)]


def _cut(whole, edges):
    """Decode the file piece by piece: every interval is clipped to the piece and given piece-relative samples."""
    out = []
    for lo, hi in zip(edges, edges[1:]):
        rows = [(0, lab, max(s, lo) - lo, min(e, hi) - lo) for lab, s, e in whole if s < hi and e > lo]
        out.append((lo, rows))
    return out


def _merge_touching(table):
    rows = []
    for f, lab, s, e in table.tolist():
        if rows and rows[-1][0] == f and rows[-1][1] == lab and rows[-1][3] >= s:
            rows[-1][3] = max(rows[-1][3], e)
        else:
            rows.append([f, lab, s, e])
    return torch.tensor(rows, dtype=torch.int32).reshape(-1, 4)


def test_interval_table_all_gather_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b)
    assert list(a[:, 0]) == sorted(a[:, 0])  # ordered by global file index
    assert set(a[:, 0]) == {0, 1, 2, 3, 4}
    # every file's rows stay in their rank-local (label, time) order
    for f in range(5):
        rows = a[a[:, 0] == f]
        assert list(rows[:, 2]) == [320 * k for k in range(len(rows))]


def test_single_process_gather_is_identity():
    t = torch.arange(12, dtype=torch.int32).reshape(3, 4)
    assert torch.equal(all_gather_tables(t), t)


def _queue_worker(rank, world, port, out_dir):
    import time

    from segma_b200.distributed import UnitQueue

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = []
    for walk, n_units in enumerate([37, 0, 5]):  # three walks over one store: the keys must not collide
        queue = UnitQueue(f"test{walk}", n_units)
        mine = []
        while (i := queue.claim()) is not None:
            mine.append(i)
            time.sleep(0.002 if rank == 0 else 0.02)  # rank 1 is the slow board
        got.append(mine)
        dist.barrier()
    torch.save(got, os.path.join(out_dir, f"claims{rank}.pt"))
    dist.destroy_process_group()


def test_unit_queue_hands_every_unit_out_once_world2(tmp_path):
    """`distributed.UnitQueue` (the on-demand form of the file / window-batch partition): two ranks claiming from one
    counter cover every unit exactly once, and the slower rank ends up with fewer."""
    mp.spawn(_queue_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f"claims{r}.pt") for r in range(2))
    for walk, n_units in enumerate([37, 0, 5]):
        assert sorted(a[walk] + b[walk]) == list(range(n_units))
        assert a[walk] == sorted(a[walk]) and b[walk] == sorted(b[walk])  # longest first on every rank
    assert len(a[0]) > len(b[0])
