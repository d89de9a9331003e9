"""Records golden vectors from the UNMODIFIED reference (its own CUDA build, oracle/_ref/libhprlp_ref.so)
on a GPU box:   gpurun -- python tests/golden/make_ref_golden.py   -> gpurun_out/golden/ref_*.json
(copy the files into tests/golden/).  The CPU test tests/test_oracle.py::test_oracle_matches_reference_golden
then pins the oracle against them without a GPU.  The power-iteration start vector is the cuRAND
XORWOW(seed 1) normal sequence both libraries draw; it is stored because the CPU oracle cannot regenerate it."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

CASES = {
    "uniform_small": dict(kind="uniform", m=60, n=150, nnz=60 * 8),
    "powerlaw_small": dict(kind="powerlaw", m=80, n=200, nnz=1600),
    "toy": "toy",
}


def main():
    pkg = graft.load_package()
    ref = pkg.load_reference()
    eng = pkg.load_engine()
    out_dir = ROOT / "gpurun_out" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    for name, spec in CASES.items():
        lp = pkg.TOY_LP if spec == "toy" else pkg.synth_lp(**spec)
        z0 = eng.power_start(lp["m"])
        runs = []
        for cusparse in (False, True):
            for prm in (dict(stop_tol=1e-4), dict(stop_tol=1e-8), dict(max_iter=10, stop_tol=1e-30), dict(max_iter=100, stop_tol=1e-30),
                        dict(max_iter=160, stop_tol=1e-30), dict(max_iter=500, stop_tol=1e-30), dict(max_iter=1000, stop_tol=1e-30)):
                prm = dict(prm, use_presolve=False, CUSPARSE_spmv=cusparse)
                model = ref.create_model(lp)
                r = ref.solve(model, pkg.Parameters.default(**prm))
                ref.free_model(model)
                runs.append(dict(param=prm, status=r["status"], iter=r["iter"], primal_obj=r["primal_obj"], residuals=r["residuals"],
                                 x=r["x"].tolist(), y=r["y"].tolist(), z=r["z"].tolist(), tol=1e-9))
        rec = dict(lp=spec, power_z0=z0.tolist(), runs=runs,
                   source="oracle/_ref/libhprlp_ref.so (reference v0.1.2 built by oracle/build_ref.sh) on NVIDIA B200")
        (out_dir / f"ref_{name}.json").write_text(json.dumps(rec))
        print(name, [(r["status"], r["iter"]) for r in runs])


if __name__ == "__main__":
    main()
