"""Runs the reference build (oracle/_ref) on a bench workload for a few hundred iterations; meant to be run under
`ncu --metrics gpu__time_duration.sum` to list the reference's own kernels and their durations next to ours."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
from bench import WORKLOADS
pkg = graft.load_package()
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 400
which = sys.argv[3] if len(sys.argv) > 3 else "ref"
spec = WORKLOADS[w]
lp = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
lib = pkg.load_reference() if which == "ref" else pkg.load_engine()
p = pkg.Parameters.default(stop_tol=0.0, max_iter=iters, use_presolve=False)
model = lib.create_model(lp)
r = lib.solve(model, p)
print(r["status"], r["iter"], r["time"])
