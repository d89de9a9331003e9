"""Where the wall time of one solve() call goes (setup / scaling / power iteration / loop), engine vs reference."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
from bench import WORKLOADS
pkg = graft.load_package(); eng = pkg.load_engine()
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
for w in (sys.argv[1:] or ["c2"]):
    spec = WORKLOADS[w]; lp = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
    for rep in range(3):
        p = pkg.Parameters.default(stop_tol=1e-4, use_presolve=False)
        model = eng.create_model(lp)
        os.dup2(devnull, 1)
        t0 = time.perf_counter(); r = eng.solve_ex(model, p, quiet=True); wall = time.perf_counter() - t0
        os.dup2(saved, 1)
        eng.free_model(model)
        i = r["info"]
        print(json.dumps(dict(w=w, rep=rep, wall=round(wall, 4), solver_time=round(r["time"], 4), setup=round(i["setup_seconds"], 4),
                              scaling=round(i["scaling_seconds"], 4), power=round(i["power_seconds"], 4), power_iters=i["power_iters"],
                              loop_ms=round(i["loop_device_ms"], 2), iters=r["iter"], launches=i["kernel_launches"])), flush=True)
