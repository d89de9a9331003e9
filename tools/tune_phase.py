"""Times the two fused kernels (CUDA events, back to back) for every library variant in lib/variants/ (+ lib/libhprlp.so)."""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
sys.path.insert(0, str(ROOT))
from bench import WORKLOADS, algorithmic_bytes  # noqa: E402

import os
wl = sys.argv[1:] or ["c2"]
libs = [ROOT / "lib" / "libhprlp.so"] + sorted((ROOT / "lib" / "variants").glob("*.so"))
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
rows = []
for w in wl:
    spec = WORKLOADS[w]
    lp = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
    ab = algorithmic_bytes(lp["m"], lp["n"], int(lp["values"].shape[0]))
    for path in libs:
        os.dup2(devnull, 1)
        try:
            lib = pkg.HprLib(path, extended=True)
            param = pkg.Parameters.default(stop_tol=0.0, use_presolve=False)
            model = lib.create_model(lp)
            h = lib.lib.hprlp_b200_engine_create(model, C.byref(param))
            lib.lib.hprlp_b200_engine_run(h, 1000)          # past the check-every-10 regime
            ms100 = lib.lib.hprlp_b200_engine_run(h, 500) / 5.0   # sustained (power-capped) rate, 1 check per 100 iterations
            tx = lib.lib.hprlp_b200_engine_time_phase(h, 0, 50)
            ty = lib.lib.hprlp_b200_engine_time_phase(h, 1, 50)
            lib.lib.hprlp_b200_engine_destroy(h)
            lib.free_model(model)
        finally:
            os.dup2(saved, 1)
        rows.append(dict(workload=w, lib=path.name, x_us=tx * 1e3, y_us=ty * 1e3, x_gbs=ab["x"] / tx / 1e6, y_gbs=ab["y"] / ty / 1e6,
                         iters_per_s=100 / ms100 * 1e3))
        print(json.dumps(rows[-1]), flush=True)
