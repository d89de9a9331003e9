"""CPU, world_size 2, gloo: the batch-sharding plumbing used by bench.py --workload c4 under torchrun
(contiguous instance shards per rank, no data-path collective, results gathered in rank order, timing =
max over ranks, units = sum over ranks).  The per-instance "solver" here is a stand-in closed form so the test
needs no GPU; on the GPU box the same helpers wrap solve_batched."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import __graft_entry__ as graft
    pkg = graft.load_package()
    lo, hi = pkg.shard_range(B, world, rank)
    rng = np.random.default_rng(0)
    C = rng.normal(size=(B, n))                       # every rank generates the same batch deterministically
    local = np.sort(C[lo:hi], axis=1)                 # stand-in for solve_batched on the shard
    full = pkg.gather_shards(dist, local, B, world, rank)
    ms, units = pkg.reduce_time_units(dist, 10.0 + rank, hi - lo)
    # the NCCL unique id of a partitioned-engine communicator travels from rank 0 to every rank the same way
    uid = pkg.broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    if rank == 0:
        np.save(Path(out_dir) / "full.npy", full)
        np.save(Path(out_dir) / "tu.npy", np.array([ms, units]))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_shard_gather_world2(tmp_path):
    B, n, world = 37, 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, n, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    rng = np.random.default_rng(0)
    assert np.array_equal(full, np.sort(rng.normal(size=(B, n)), axis=1))
    ms, units = np.load(tmp_path / "tu.npy")
    assert ms == 11.0 and units == B


def test_row_blocks_cover_all_rows_and_balance_nonzeros():
    """The row blocks of the partitioned solve (one per rank): contiguous, complete, balanced by nonzeros."""
    import __graft_entry__ as graft
    pkg = graft.load_package()
    lp = pkg.synth_lp("powerlaw", 5000, 9000, 120000)
    rp = lp["rowPtr"]
    for P in (1, 2, 3, 8):
        b = pkg.row_blocks_by_nnz(rp, P)
        assert b[0] == 0 and b[-1] == 5000 and all(b[i] <= b[i + 1] for i in range(P))
        nz = [int(rp[b[i + 1]] - rp[b[i]]) for i in range(P)]
        assert sum(nz) == int(rp[-1])
        assert max(nz) - min(nz) <= 2 * int(np.max(np.diff(rp)))     # within two longest rows of the ideal split


def test_shard_range_partitions_exactly():
    import __graft_entry__ as graft
    pkg = graft.load_package()
    for B in (1, 7, 32, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [pkg.shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
