"""GPU (-m gpu, needs >= 2 GPUs): the row-block partitioned multi-GPU solve (NCCL all-reduce of A^T y over NVLink)
against the single-GPU engine on the same LP: same status, same iteration count (same restart decisions), iterates
to rounding (the cross-GPU sum changes the summation order of A^T y only)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("kind,m,n,nnz", [("uniform", 3000, 9000, 90000), ("powerlaw", 20000, 50000, 1000000)])
def test_partitioned_matches_single_gpu(pkg, engine, kind, m, n, nnz):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    lp = pkg.synth_lp(kind, m, n, nnz, with_solution=True)
    for prm in (dict(stop_tol=1e-6), dict(max_iter=300, stop_tol=1e-30)):
        p = pkg.Parameters.default(use_presolve=False, **prm)
        model = engine.create_model(lp)
        one = engine.solve(model, p, main=True)
        two = engine.solve_partitioned(model, p, n_gpus=2)
        engine.free_model(model)
        assert one["status"] == two["status"] and one["iter"] == two["iter"], (prm, one["iter"], two["iter"])
        assert abs(one["primal_obj"] - two["primal_obj"]) <= 1e-9 * (1 + abs(one["primal_obj"]))
        for k in "xyz":
            assert np.max(np.abs(one[k] - two[k])) <= 1e-8 * max(1.0, np.max(np.abs(one[k]))), (prm, k)
    assert abs(two["primal_obj"] - lp["obj_star"]) / (1 + abs(lp["obj_star"])) < 1e-5 or prm.get("max_iter")


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
