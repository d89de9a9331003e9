"""CPU: pins of the oracle itself (oracle/hpr_oracle.c).  The reference has no tests and no CPU path; the
pins are the toy LP known answer of its examples, the constructed optimum of the synthetic LPs, algebraic
invariants of the scaling, and golden records of the reference's own CUDA build (tests/golden/ref_*.json,
recorded on a B200 by tests/golden/make_ref_golden.py)."""
import json
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

GOLD = Path(__file__).resolve().parent / "golden"


def test_toy_lp_known_answer(pkg, oracle):
    # reference examples/cpp/example_direct_lp.cpp:14 -- x = (2.8, 3.6), objective -26.4
    for tol, xtol in ((1e-4, 2e-3), (1e-8, 1e-6)):
        r = oracle.solve(pkg.TOY_LP, pkg.Parameters.default(stop_tol=tol))
        assert r["status"] == "OPTIMAL"
        assert np.allclose(r["x"], [2.8, 3.6], atol=xtol)
        assert abs(r["primal_obj"] + 26.4) < 30 * tol
    # 180 iterations at 1e-4, 530 at 1e-8 for the reference algorithm
    assert oracle.solve(pkg.TOY_LP, pkg.Parameters.default(stop_tol=1e-4))["iter"] == 180
    assert oracle.solve(pkg.TOY_LP, pkg.Parameters.default(stop_tol=1e-8))["iter"] == 530


@pytest.mark.parametrize("kind,m,n,nnz", [("uniform", 400, 1500, 400 * 15), ("powerlaw", 3000, 6000, 60000)])
def test_synthetic_lp_reaches_constructed_optimum(pkg, oracle, kind, m, n, nnz):
    lp = pkg.synth_lp(kind, m, n, nnz, with_solution=True)
    A = sp.csr_matrix((lp["values"], lp["colIndex"], lp["rowPtr"]), shape=(m, n))
    # the constructed pair satisfies primal feasibility and stationarity exactly (up to rounding)
    ax = A @ lp["xs"]
    assert np.all(ax >= lp["AL"] - 1e-9) and np.all(ax <= lp["AU"] + 1e-9)
    assert np.allclose(A.T @ lp["ys"] + lp["zs"], lp["c"], atol=1e-12)
    r = oracle.solve(lp, pkg.Parameters.default(stop_tol=1e-6))
    assert r["status"] == "OPTIMAL"
    assert abs(r["primal_obj"] - lp["obj_star"]) / (1 + abs(lp["obj_star"])) < 1e-5


@pytest.mark.parametrize("kind", ["banded", "blocked"])
def test_structured_twins_of_the_synthetic_lp(pkg, oracle, kind):
    """The structured generators behind bench.py's c3band / c3block workloads: sorted distinct columns inside the window,
    dense 8x8 blocks for 'blocked' (the 8 rows of a group share their columns, which come in aligned runs of 8), and an LP
    the oracle solves to the constructed optimum like the uniform ones."""
    m, n, per_row = 640, 6000, 24
    lp = pkg.synth_lp(kind, m, n, m * per_row, with_solution=True)
    rp, col = lp["rowPtr"], lp["colIndex"]
    assert rp[-1] == m * per_row and col.min() >= 0 and col.max() < n
    for i in range(m):
        c = col[rp[i]:rp[i + 1]]
        assert np.all(np.diff(c) > 0)
        assert c[-1] - c[0] < 4096
    if kind == "blocked":
        for g in range(0, m, 8):
            first = col[rp[g]:rp[g + 1]]
            assert np.all(first % 8 == np.arange(per_row) % 8) and np.all(first[::8] % 8 == 0)   # aligned runs of 8
            for i in range(g + 1, g + 8):
                assert np.array_equal(col[rp[i]:rp[i + 1]], first)
        lens = np.bincount(col, minlength=n)
        assert np.all(lens % 8 == 0)            # column counts are multiples of 8: the uneven rows of the transpose
    r = oracle.solve(lp, pkg.Parameters.default(stop_tol=1e-6))
    assert r["status"] == "OPTIMAL"
    assert abs(r["primal_obj"] - lp["obj_star"]) / (1 + abs(lp["obj_star"])) < 1e-5


def test_iter_limit_and_stale_bar_quirk(pkg, oracle):
    lp = pkg.synth_lp("uniform", 120, 400, 120 * 10)
    r = oracle.solve(lp, pkg.Parameters.default(max_iter=40, stop_tol=1e-12))
    assert r["status"] == "ITER_LIMIT" and r["iter"] == 40
    # max_iter not a multiple of step(): bars are those of the last check iteration (40), quirk #2
    r45 = oracle.solve(lp, pkg.Parameters.default(max_iter=45, stop_tol=1e-12))
    assert r45["iter"] == 45 and np.array_equal(r45["x"], r["x"])


def test_scaling_invariants(pkg, oracle):
    lp = pkg.synth_lp("powerlaw", 500, 800, 9000)
    m, n = lp["m"], lp["n"]
    A = sp.csr_matrix((lp["values"], lp["colIndex"], lp["rowPtr"]), shape=(m, n))
    for flags in ((True, True, True, True), (False, True, True, False), (True, False, False, True)):
        p = pkg.Parameters.default(use_CR_scaling=flags[0], use_Ruiz_scaling=flags[1], use_Pock_Chambolle_scaling=flags[2],
                                   use_bc_scaling=flags[3])
        s = oracle.scale(lp, p)
        As = sp.csr_matrix((s["A_val"], lp["colIndex"], lp["rowPtr"]), shape=(m, n))
        want = sp.diags(1.0 / s["row_norm"]) @ A @ sp.diags(1.0 / s["col_norm"])
        assert abs(As - want).max() < 1e-12 * max(1.0, abs(want).max())
        ATs = sp.csr_matrix((s["AT_val"], s["AT_col"], s["AT_rowPtr"]), shape=(n, m))
        assert (ATs != As.T.tocsr()).nnz == 0          # both copies stay bit-identical
        bs, cs = s["scalars"][0], s["scalars"][1]
        fin = np.isfinite(lp["AU"])
        assert np.allclose(s["AU"][fin], lp["AU"][fin] / s["row_norm"][fin] / bs, rtol=1e-12)
        assert np.allclose(s["c"], lp["c"] / s["col_norm"] / cs, rtol=1e-12)
        if flags[1] or flags[2]:
            assert abs(As).max() <= 1.0 + 1e-9


def _golden_files():
    return sorted(GOLD.glob("ref_*.json"))


@pytest.mark.parametrize("path", _golden_files() or [None])
def test_oracle_matches_reference_golden(pkg, oracle, path):
    """Golden records of the reference's own CUDA build (status, objective, iterates at max_iter=k)."""
    if path is None:
        pytest.skip("no golden records yet (tests/golden/make_ref_golden.py runs on the GPU box)")
    rec = json.loads(path.read_text())
    lp = pkg.synth_lp(**rec["lp"]) if rec["lp"] != "toy" else pkg.TOY_LP
    z0 = np.array(rec["power_z0"]) if rec.get("power_z0") is not None else None
    for run in rec["runs"]:
        p = pkg.Parameters.default(**run["param"])
        r = oracle.solve(lp, p, power_z0=z0)
        assert r["status"] == run["status"] and r["iter"] == run["iter"], (path.name, run["param"], r["iter"], run["iter"])
        assert abs(r["primal_obj"] - run["primal_obj"]) / (1 + abs(run["primal_obj"])) < 1e-9
        for k in "xyz":   # 1e-10: the iterate tolerance BASELINE.json's north_star states (measured: <= 5e-12)
            ref = np.array(run[k])
            assert np.max(np.abs(r[k] - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), (path.name, run["param"], k)
