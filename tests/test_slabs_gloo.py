"""CPU: the multi-rank host logic (dctz_b200/slabs.py) with world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dctz_b200 import slabs
from tests import reflib


def test_partition_covers_field_with_whole_blocks():
    for n in (1, 63, 64, 65, 64 * 7, 64 * 7 + 13, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 4, 8):
            parts = slabs.partition(n, world)
            assert len(parts) == world
            pos = 0
            for r, (start, count) in enumerate(parts):
                assert start == pos or count == 0
                if count:
                    assert start % 64 == 0
                    if start + count != n:
                        assert count % 64 == 0  # only the slab that ends the field may hold the tail block
                pos += count
            assert pos == n


def test_scaling_factor_matches_oracle():
    for v in (0.0031, 0.1, 0.99, 1.0, 1.5, 9.99, 10.0, 10.5, 123.0, 1e5, 99999.9, 1e-5):
        x = np.array([v, -v / 3, v / 7])
        assert slabs.scaling_factor(v) == reflib.oracle_stat(x)["sf"]
        xf = x.astype(np.float32)
        assert slabs.scaling_factor(float(np.float32(v)), single=True) == reflib.oracle_stat(xf)["sf"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, qfile):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        field = rng.standard_normal(n) * 37.0 + 3.0
        start, count = slabs.partition(n, world)[rank]
        mine = field[start:start + count]
        # stand-in for k_stats on this rank's slab
        local = torch.tensor([np.abs(mine).max(), np.abs(mine).min(), mine.sum()], dtype=torch.float64)
        gathered = slabs.all_gather_stats(local, world)
        triples = gathered.reshape(world, 3).tolist()
        mx, mn, s = slabs.merge_stats(triples)
        s -= field[0]  # util.c:21-25 never adds element 0
        want = reflib.oracle_stat(field)
        assert mx == want["max"] and mn == want["min"], (mx, want)
        assert abs(s - want["sum"]) <= 1e-9 * np.abs(field).sum()
        assert slabs.scaling_factor(mx) == want["sf"]
        # QT table exchange: positions 1..63 max over ranks, position 0 from the last rank
        q = torch.full((64,), float(rank + 1), dtype=torch.float64)
        q[0] = 100.0 + rank
        q = slabs.all_reduce_qtable(q, rank, world)
        assert q[0].item() == 100.0 + world - 1 and torch.all(q[1:] == float(world))
        # a field of two blocks: ranks beyond the data hold nothing, entry 0 must come from the last rank WITH data
        src = slabs.last_rank_with_data(100, world)
        assert src == 1 and slabs.partition(100, world)[src][1] == 36
        q = torch.full((64,), float(rank + 1), dtype=torch.float64)
        q[0] = 100.0 + rank
        q = slabs.all_reduce_qtable(q, rank, world, src_last=src)
        assert q[0].item() == 101.0
        # outlier segments are concatenated in rank order: exchange counts only to place them (host side)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([count // 7], dtype=torch.int64))
        offs = np.concatenate([[0], np.cumsum([int(c) for c in counts])])
        assert offs[rank] == sum(p[1] // 7 for p in slabs.partition(n, world)[:rank])
        if rank == 0:
            open(qfile, "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stats_exchange_over_gloo(world, tmp_path):
    qfile = str(tmp_path / "ok")
    mp.spawn(_worker, args=(world, _free_port(), 64 * 1000 + 17, qfile), nprocs=world, join=True)
    assert open(qfile).read() == "ok"
