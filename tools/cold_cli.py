"""Cold-process comparison of the two CLIs: configs[1] (or any synthetic LP) written to an MPS file,
then `build/solve_mps_file` (this repo) and `oracle/_ref/solve_mps_file` (the reference's own build) each started as a
fresh process: process wall time = CUDA context + MPS parse + (presolve) + solve.  Prints one JSON line.
  python tools/cold_cli.py [--m 100000 --n 1000000 --nnz 10000000] [--presolve true|false] [--runs 2]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import __graft_entry__ as graft  # noqa: E402
from make_big_mps import write_mps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=100_000); ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--nnz", type=int, default=10_000_000); ap.add_argument("--kind", default="uniform")
ap.add_argument("--presolve", default="false"); ap.add_argument("--runs", type=int, default=2)
a = ap.parse_args()
pkg = graft.load_package()
lp = pkg.synth_lp(a.kind, a.m, a.n, a.nnz, with_solution=True)
tmp = Path(tempfile.mkdtemp()) / "lp.mps"
t0 = time.perf_counter()
write_mps(lp, tmp)
t_write = time.perf_counter() - t0


def run(binary):
    t0 = time.perf_counter()
    pr = subprocess.run([str(binary), "-i", str(tmp), "--presolve", a.presolve], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    wall = time.perf_counter() - t0
    out = dict(wall_s=wall, rc=pr.returncode)
    for ln in pr.stdout.splitlines():
        s = ln.strip()
        if s.startswith("Iterations:"): out["iters"] = int(s.split(":")[1])
        elif s.startswith("Time:"): out["solver_time_s"] = float(s.split(":")[1].split()[0])
        elif s.startswith("Status:"): out["status"] = s.split(":")[1].strip()
        elif s.startswith("Primal Objective:"): out["primal_obj"] = float(s.split(":")[1])
    return out


res = {}
for name, binary in (("reference", ROOT / "oracle" / "_ref" / "solve_mps_file"), ("ours", ROOT / "build" / "solve_mps_file")):
    if binary.exists():
        res[name] = [run(binary) for _ in range(a.runs)]
best = {k: min(v, key=lambda r: r["wall_s"]) for k, v in res.items()}
print(json.dumps(dict(what="cold CLI process wall: CUDA context + MPS parse + solve to KKT<1e-4", m=a.m, n=a.n, nnz=a.nnz, kind=a.kind,
                      presolve=a.presolve, mps_bytes=tmp.stat().st_size, mps_write_s=t_write, host_cores=os.cpu_count(),
                      constructed_optimum=lp["obj_star"], runs=res, best=best,
                      speedup=(best["reference"]["wall_s"] / best["ours"]["wall_s"]) if len(best) == 2 else None)))
os.remove(tmp)
