"""GPU (-m gpu): BASELINE.json's full-size config 2 (uniform LP m=1e5, n=1e6, nnz=1e7) -- too large for the CPU oracle
in a test, so parity is taken against the reference's own CUDA build on the same arrays, plus size-independent
properties: the constructed optimum, feasibility of the returned point in the ORIGINAL space, run-to-run
bitwise reproducibility."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(pkg):
    return pkg.synth_lp("uniform", 100_000, 1_000_000, 10_000_000, with_solution=True)


def test_c2_time_to_1e4_matches_reference(pkg, engine, reference, c2):
    p = pkg.Parameters.default(stop_tol=1e-4, use_presolve=False)
    outs = {}
    for tag, lib in (("new", engine), ("ref", reference)):
        model = lib.create_model(c2)
        outs[tag] = lib.solve(model, p)
        lib.free_model(model)
    a, b = outs["new"], outs["ref"]
    assert a["status"] == b["status"] == "OPTIMAL"
    assert a["iter"] == b["iter"]                                              # same restart / stopping decisions
    assert abs(a["primal_obj"] - b["primal_obj"]) / (1 + abs(b["primal_obj"])) <= 1e-6   # north_star: <= 1e-6 at tol 1e-4
    assert abs(a["residuals"] - b["residuals"]) <= 1e-9
    assert abs(a["primal_obj"] - c2["obj_star"]) / (1 + abs(c2["obj_star"])) < 5e-4      # constructed optimum
    # returned point: original-space feasibility and dual consistency at the KKT tolerance
    A = sp.csr_matrix((c2["values"], c2["colIndex"], c2["rowPtr"]), shape=(c2["m"], c2["n"]))
    ax = A @ a["x"]
    nb = 1 + np.linalg.norm(np.maximum(np.abs(np.where(np.isfinite(c2["AL"]), c2["AL"], 0)), np.abs(np.where(np.isfinite(c2["AU"]), c2["AU"], 0))))
    assert np.linalg.norm(np.maximum(c2["AL"] - ax, 0) + np.maximum(ax - c2["AU"], 0)) / nb < 1e-4
    assert np.linalg.norm(c2["c"] - A.T @ a["y"] - a["z"]) / (1 + np.linalg.norm(c2["c"])) < 1e-4
    assert np.all(a["x"] >= c2["l"] - 1e-9) and np.all(a["x"] <= c2["u"] + 1e-9)


def test_c2_iterates_match_reference_and_are_reproducible(pkg, engine, reference, c2):
    p = pkg.Parameters.default(max_iter=200, stop_tol=1e-30, use_presolve=False)
    model = engine.create_model(c2)
    a1 = engine.solve(model, p)
    a2 = engine.solve(model, p)
    engine.free_model(model)
    model = reference.create_model(c2)
    b = reference.solve(model, p)
    reference.free_model(model)
    for k in "xyz":
        assert np.array_equal(a1[k], a2[k]), k                                # bitwise reproducible
        assert np.max(np.abs(a1[k] - b[k])) <= 1e-10 * max(1.0, np.max(np.abs(b[k]))), k
