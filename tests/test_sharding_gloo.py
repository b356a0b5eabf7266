"""CPU, world_size 2 over gloo: the multi-GPU path's host logic (contiguous range sharding + final gather).
The per-rank "solver" here is the CPU oracle (test infrastructure) so the test runs without GPUs; on GPUs each rank calls
LexLSEBatch on the same slice."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_cover_the_batch():
    from lexls_b200.shard import shard_range, shard_ranges
    for batch in (1, 7, 16, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            rs = shard_ranges(batch, world)
            assert rs[0][0] == 0 and rs[-1][1] == batch
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
            assert shard_range(batch, world, world - 1) == rs[-1]


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from lexls_b200 import problems
    from lexls_b200.shard import gather_to_rank0, shard_range
    w = problems.workload("tiny_lse", 37)          # ragged: 19 + 18
    data, _ = problems.generate(w)
    lo, hi = shard_range(w.batch, world, rank)
    loc = oracle.lse_solve_batch(w.nVar, w.ObjDim, [d[lo:hi] for d in data])
    x = gather_to_rank0(loc["x"], w.batch)
    ranks = gather_to_rank0(loc["ranks"], w.batch)
    if rank == 0:
        full = oracle.lse_solve_batch(w.nVar, w.ObjDim, data)
        ok = np.array_equal(x, full["x"]) and np.array_equal(ranks, full["ranks"])
        open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(os.path.join(str(tmp_path), "ok")).read() == "1"


def test_tapered_ranges_cover_the_batch_without_empty_chunks():
    from lexls_b200.shard import tapered_ranges
    for batch in (1, 2, 3, 15, 63, 64, 65, 1000, 16384, 65536):
        for chunks in (1, 2, 3, 4, 5, 8):
            r = tapered_ranges(batch, chunks)
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(hi > lo for lo, hi in r) and all(r[i][1] == r[i + 1][0] for i in range(len(r) - 1))
            if chunks >= 3 and batch >= 8192:
                assert (r[-1][1] - r[-1][0]) * 3 < (r[0][1] - r[0][0])  # the last chunk is the small one
