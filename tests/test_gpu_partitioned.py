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
