"""Small end-to-end run of every device path, meant to be executed under compute-sanitizer:
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py   (where the tool is available; this pool has it closed)
(single LP incl. cut rows and check passes, graph replay, uneven rows, batched, the row partition on logical ranks)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
eng = pkg.load_engine()


def solve(lp, **prm):
    p = pkg.Parameters.default(use_presolve=False, **prm)
    model = eng.create_model(lp)
    r = eng.solve(model, p, main=True)
    eng.free_model(model)
    return r


def main():
    r = solve(pkg.TOY_LP, stop_tol=1e-8)
    assert r["status"] == "OPTIMAL" and np.allclose(r["x"], [2.8, 3.6], atol=1e-6)
    lp = pkg.synth_lp("powerlaw", 4000, 9000, 120000, with_solution=True)      # rows cut by item boundaries, look-back
    r = solve(lp, max_iter=400, stop_tol=1e-30)                                   # > 300 iterations: graph replay
    assert r["status"] == "ITER_LIMIT" and np.all(np.isfinite(r["x"]))
    lp = pkg.synth_lp("blocked", 2400, 6000, 2400 * 50)                           # 4 lanes per row on the transpose
    r = solve(lp, max_iter=120, stop_tol=1e-30)
    assert np.all(np.isfinite(r["y"]))
    base = pkg.synth_lp("uniform", 400, 1500, 6000)
    vs = [pkg.synth_vectors(base, pkg.SEED, pkg.SEED + k) for k in range(40)]
    st = lambda key: np.stack([v[key] for v in vs])
    model = eng.create_model(base)
    p = pkg.Parameters.default(use_presolve=False, max_iter=120, stop_tol=1e-30)
    rb = eng.solve_batched(model, st("c"), st("AL"), st("AU"), st("l"), st("u"), None, p)
    assert np.all(np.isfinite(rb["x"]))
    lp = pkg.synth_lp("uniform", 3000, 9000, 90000)
    model2 = eng.create_model(lp)
    p = pkg.Parameters.default(use_presolve=False, max_iter=120, stop_tol=1e-30)
    one = eng.solve(model2, p, main=True)
    par = eng.solve_partitioned(model2, p, n_gpus=3, local=True)
    assert np.max(np.abs(one["x"] - par["x"])) <= 1e-8 * max(1.0, np.max(np.abs(one["x"])))
    eng.free_model(model); eng.free_model(model2)
    print("sanitize_smoke: ok")


if __name__ == "__main__":
    main()
