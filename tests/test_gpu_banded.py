"""GPU (-m gpu): column-banded passes over A (Engine::build_bands -- the layout used when the gathered n-vector exceeds
the L2, e.g. BASELINE configs[4]) against the plain single-pass kernels on the same LP.  HPRLP_BAND_COLS forces bands on
small problems.  The banded pass adds the same products band by band, so only the summation order of a row changes:
same status, same iteration count, iterates to rounding."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def band_env():
    old = os.environ.get("HPRLP_BAND_COLS")
    yield
    if old is None:
        os.environ.pop("HPRLP_BAND_COLS", None)
    else:
        os.environ["HPRLP_BAND_COLS"] = old


@pytest.mark.parametrize("kind,m,n,nnz,band", [
    ("uniform", 3000, 9000, 90000, 1000),          # 9 bands, ~3 entries per row and band
    ("powerlaw", 20000, 50000, 1000000, 7001),     # 8 ragged bands, long rows spanning many items inside a band
    ("uniform", 300, 900, 3600, 64),               # 15 bands, most (row, band) cells empty
    ("uniform", 2000, 70000, 100000, 1000),        # 70 bands requested -> clamped to 64
])
def test_banded_matches_plain(pkg, engine, band_env, kind, m, n, nnz, band):
    lp = pkg.synth_lp(kind, m, n, nnz, with_solution=True)
    out = {}
    for tag, cols in (("plain", None), ("banded", band)):
        if cols is None:
            os.environ.pop("HPRLP_BAND_COLS", None)
        else:
            os.environ["HPRLP_BAND_COLS"] = str(cols)
        res = []
        for prm in (dict(stop_tol=1e-6), dict(max_iter=300, stop_tol=1e-30)):
            p = pkg.Parameters.default(use_presolve=False, **prm)
            model = engine.create_model(lp)
            r = engine.solve_ex(model, p)
            engine.free_model(model)
            assert (r["info"]["bands_A"] > 1) == (cols is not None), r["info"]["bands_A"]   # the banded path really ran
            res.append(r)
        out[tag] = res
    for a, b in zip(out["plain"], out["banded"]):
        assert a["status"] == b["status"] and a["iter"] == b["iter"], (a["iter"], b["iter"])
        assert abs(a["primal_obj"] - b["primal_obj"]) <= 1e-9 * (1 + abs(a["primal_obj"]))
        for k in "xyz":
            assert np.max(np.abs(a[k] - b[k])) <= 1e-8 * max(1.0, np.max(np.abs(a[k]))), k
    assert abs(out["banded"][0]["primal_obj"] - lp["obj_star"]) / (1 + abs(lp["obj_star"])) < 1e-4
