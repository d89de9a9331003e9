"""GPU (-m gpu): the row-block partitioned solve (x-block ownership: reduce-scatter of the partial A^T y, x-update on
n/P entries, all-gather of x_hat) against the single-GPU engine on the same LP: same status, same iteration count (same
restart decisions), iterates to rounding (the cross-rank sum changes the summation order of A^T y only).
* `local` tests run on ONE GPU: P logical ranks of the same partitioned engine code with host-synchronised exchanges
  (csrc/collective.cu) -- the partitioned path is parity-checked on every box;
* NCCL tests need >= 2 GPUs (gpurun --gpus 2) and skip otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("kind,m,n,nnz", [("uniform", 3000, 9000, 90000), ("powerlaw", 20000, 50000, 1000000)])
def test_partitioned_matches_single_gpu(pkg, engine, kind, m, n, nnz, exchange, monkeypatch):
    """exchange = p2p: our fused reduce-scatter + x-update + all-gather kernel over NVLink peer memory;
    nccl: ncclReduceScatter + x-update kernel + ncclAllGather."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    monkeypatch.setenv("HPRLP_EXCHANGE", exchange)
    lp = pkg.synth_lp(kind, m, n, nnz, with_solution=True)
    for prm in (dict(stop_tol=1e-6), dict(max_iter=300, stop_tol=1e-30)):
        p = pkg.Parameters.default(use_presolve=False, **prm)
        model = engine.create_model(lp)
        one = engine.solve(model, p, main=True)
        two = engine.solve_partitioned(model, p, n_gpus=2)
        engine.free_model(model)
        assert two["info"]["peer_exchange"] == (1 if exchange == "p2p" else 0)
        assert one["status"] == two["status"] and one["iter"] == two["iter"], (prm, one["iter"], two["iter"])
        assert abs(one["primal_obj"] - two["primal_obj"]) <= 1e-9 * (1 + abs(one["primal_obj"]))
        for k in "xyz":
            assert np.max(np.abs(one[k] - two[k])) <= 1e-8 * max(1.0, np.max(np.abs(one[k]))), (prm, k)
    assert abs(two["primal_obj"] - lp["obj_star"]) / (1 + abs(lp["obj_star"])) < 1e-5 or prm.get("max_iter")


@pytest.mark.parametrize("ranks", [2, 3, 4])
@pytest.mark.parametrize("kind,m,n,nnz", [("uniform", 3000, 9000, 90000), ("powerlaw", 20000, 50000, 1000000)])
def test_partitioned_local_ranks_match_single_gpu(pkg, engine, kind, m, n, nnz, ranks):
    """P logical ranks on one GPU: row blocks, x-block ownership, every exchange -- against the plain engine."""
    lp = pkg.synth_lp(kind, m, n, nnz, with_solution=True)
    for prm in (dict(stop_tol=1e-6), dict(max_iter=300, stop_tol=1e-30)):
        p = pkg.Parameters.default(use_presolve=False, **prm)
        model = engine.create_model(lp)
        one = engine.solve(model, p, main=True)
        par = engine.solve_partitioned(model, p, n_gpus=ranks, local=True)
        engine.free_model(model)
        assert par["info"]["reserved0"] == ranks
        assert one["status"] == par["status"] and one["iter"] == par["iter"], (prm, one["iter"], par["iter"])
        assert abs(one["primal_obj"] - par["primal_obj"]) <= 1e-9 * (1 + abs(one["primal_obj"]))
        for k in "xyz":
            assert par[k].shape == one[k].shape
            assert np.max(np.abs(one[k] - par[k])) <= 1e-8 * max(1.0, np.max(np.abs(one[k]))), (prm, k)


def test_partitioned_local_ragged_blocks(pkg, engine):
    """n not a multiple of the block size, more ranks than convenient, empty rows/columns: the padded exchange blocks."""
    rng = np.random.default_rng(5)
    m, n = 37, 131
    dense = (rng.random((m, n)) < 0.08) * rng.normal(size=(m, n))
    dense[5, :] = 0.0; dense[:, 7] = 0.0; dense[0, 0] = 1.0
    rp = np.zeros(m + 1, np.int32); cols = []; vals = []
    for i in range(m):
        nzc = np.nonzero(dense[i])[0]
        cols += list(nzc); vals += list(dense[i, nzc]); rp[i + 1] = len(cols)
    xs = rng.random(n); ax = dense @ xs
    lp = dict(m=m, n=n, rowPtr=rp, colIndex=np.array(cols, np.int32), values=np.array(vals), AL=ax - 0.5, AU=ax + 0.5,
              l=np.zeros(n), u=np.full(n, 2.0), c=rng.normal(size=n))
    p = pkg.Parameters.default(use_presolve=False, stop_tol=1e-7, max_iter=20000)
    model = engine.create_model(lp)
    one = engine.solve(model, p, main=True)
    for ranks in (2, 5):
        par = engine.solve_partitioned(model, p, n_gpus=ranks, local=True)
        assert one["status"] == par["status"] and one["iter"] == par["iter"], (ranks, one["iter"], par["iter"])
        for k in "xyz":
            assert np.max(np.abs(one[k] - par[k])) <= 1e-8 * max(1.0, np.max(np.abs(one[k]))), (ranks, k)
    engine.free_model(model)


def test_partitioned_single_gpu_falls_through(pkg, engine):
    """n_gpus = 1 (or one visible device) is the ordinary engine."""
    lp = pkg.synth_lp("uniform", 300, 900, 3600)
    p = pkg.Parameters.default(use_presolve=False, stop_tol=1e-6)
    model = engine.create_model(lp)
    one = engine.solve(model, p, main=True)
    same = engine.solve_partitioned(model, p, n_gpus=1)
    engine.free_model(model)
    assert one["status"] == same["status"] and one["iter"] == same["iter"]
    for k in "xyz":
        assert np.array_equal(one[k], same[k])


def test_device_generator_matches_host_generator(pkg, engine):
    """The on-device shard generator (csrc/synth_device.cu) reproduces tools/synth_lp.c bit for bit, any row block."""
    m, n, K = 600, 900, 40          # K/n large enough that duplicate column draws (re-draw path) occur
    lp = pkg.synth_lp("uniform", m, n, m * K)
    for row0, rows in ((0, m), (137, 200), (m - 1, 1)):
        col, val = engine.synth_rows(n, K, row0, rows)
        assert np.array_equal(col, lp["colIndex"][row0 * K:(row0 + rows) * K])
        assert np.array_equal(val, lp["values"][row0 * K:(row0 + rows) * K])


@pytest.mark.parametrize("gpus", [1, 2])
def test_device_generated_partitioned_solve_matches_host_lp(pkg, engine, gpus):
    if _ngpu() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    m, n, K = 4000, 12000, 25
    lp = pkg.synth_lp("uniform", m, n, m * K, with_solution=True)
    p = pkg.Parameters.default(use_presolve=False, stop_tol=1e-6)
    model = engine.create_model(lp)
    host = engine.solve(model, p, main=True)
    engine.free_model(model)
    dev = engine.solve_partitioned_synth(m, n, K, p, n_gpus=gpus)
    assert abs(dev["obj_star"] - lp["obj_star"]) <= 1e-10 * (1 + abs(lp["obj_star"]))    # same LP up to rounding of c
    assert dev["status"] == host["status"] == "OPTIMAL" and dev["iter"] == host["iter"]
    assert abs(dev["primal_obj"] - host["primal_obj"]) <= 1e-8 * (1 + abs(host["primal_obj"]))
    for k in "xyz":
        assert np.max(np.abs(dev[k] - host[k])) <= 1e-6 * max(1.0, np.max(np.abs(host[k]))), k


@pytest.mark.parametrize("transport", ["local", "nccl"])
def test_a_failing_rank_does_not_hang_its_peers(pkg, engine, transport, monkeypatch):
    """if one rank throws, the others used to block forever in the next collective.  Now the failing rank
    aborts every endpoint (ncclCommAbort / poisoned host barrier): the call returns "ERROR" promptly and the library stays
    usable.  HPRLP_TEST_FAIL_RANK makes rank 1 fail before its first collective while rank 0 is already waiting in one."""
    import time
    if transport == "nccl" and _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    lp = pkg.synth_lp("uniform", 3000, 9000, 90000)
    p = pkg.Parameters.default(use_presolve=False, stop_tol=1e-6)
    model = engine.create_model(lp)
    monkeypatch.setenv("HPRLP_TEST_FAIL_RANK", "1")
    t0 = time.perf_counter()
    bad = engine.solve_partitioned(model, p, n_gpus=2, local=(transport == "local"))
    assert bad["status"] == "ERROR" and bad["x"] is None
    assert time.perf_counter() - t0 < 60
    monkeypatch.delenv("HPRLP_TEST_FAIL_RANK")
    ok = engine.solve_partitioned(model, p, n_gpus=2, local=(transport == "local"))
    engine.free_model(model)
    assert ok["status"] == "OPTIMAL"
