"""BASELINE.json config 5: synthetic uniform LP with nnz = m*K generated on the GPUs and solved row-partitioned (NCCL).
python tools/run_c5.py --gpus 8 --m 20000000 --n 50000000 --k 300 --tol 1e-4"""
import argparse, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8); ap.add_argument("--m", type=int, default=20_000_000)
ap.add_argument("--n", type=int, default=50_000_000); ap.add_argument("--k", type=int, default=300)
ap.add_argument("--tol", type=float, default=1e-4); ap.add_argument("--max-iter", type=int, default=2**31 - 1)
ap.add_argument("--time-limit", type=float, default=600.0)
a = ap.parse_args()
pkg = graft.load_package(); eng = pkg.load_engine()
p = pkg.Parameters.default(use_presolve=False, stop_tol=a.tol, max_iter=a.max_iter, time_limit=a.time_limit)
t0 = time.perf_counter()
r = eng.solve_partitioned_synth(a.m, a.n, a.k, p, n_gpus=a.gpus, want_solution=False, quiet=False)
wall = time.perf_counter() - t0
i = r["info"]
nnz = a.m * a.k
print(json.dumps(dict(config="configs[4]: synthetic uniform LP, row-block partitioned over NVLink (peer-memory exchange)", m=a.m, n=a.n, nnz=nnz, gpus=a.gpus,
                      status=r["status"], iters=r["iter"], primal_obj=r["primal_obj"], obj_star=r["obj_star"], residuals=r["residuals"],
                      wall_s=wall, generate_and_transpose_s=i["setup_seconds"], scaling_s=i["scaling_seconds"], power_s=i["power_seconds"],
                      power_iters=i["power_iters"], solver_time_s=r["time"], loop_ms=i["loop_device_ms"],
                      loop_ms_per_iter=i["loop_device_ms"] / max(r["iter"], 1), iters_per_s=1e3 * r["iter"] / max(i["loop_device_ms"], 1e-9),
                      algorithmic_GBps_per_gpu=((24 * nnz + 68 * a.n + 52 * a.m) / a.gpus + 20 * a.n) / (i["loop_device_ms"] / max(r["iter"], 1) * 1e-3) / 1e9)))
