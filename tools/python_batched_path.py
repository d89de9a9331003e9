"""the Python solve_batched path.  The reference's binding receives C, l, u as (n, B) and AL, AU as (m, B)
NumPy arrays and re-packs all five element by element into column-major std::vectors before every call
(bindings/python/src/hprlp_pybind.cpp:343-356, 413-455).  This script times, on configs[3] (n=2e5, m=5e4, B=256):
  repack_numpy_s   the same re-packing done by NumPy (np.asfortranarray of the five arrays: a vectorised copy, i.e. a LOWER
                   bound of the binding's scalar strided loops)
  repack_scalar_s  a C restatement of the binding's double loop with the same strides (compiled here with gcc -O2), per call
  call_colmajor_s  solve_batched on the re-packed arrays (the library call the binding then makes)
  call_rowmajor_s  hprlp_b200_solve_batched_layout(layout=1) on the ORIGINAL C-ordered arrays: no host re-packing at all
and checks that both calls return the same bits.  python tools/python_batched_path.py [--iters 300]"""
import argparse, ctypes, json, os, subprocess, sys, tempfile, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
from bench import WORKLOADS, make_batch, Quiet

ap = argparse.ArgumentParser(); ap.add_argument("--iters", type=int, default=300); ap.add_argument("--workload", default="c4")
a = ap.parse_args()
pkg = graft.load_package(); eng = pkg.load_engine()
spec = WORKLOADS[a.workload]; B = spec["batch"]
base, d = make_batch(pkg, spec, 0, B)                      # d[k]: (B, n) C-ordered = column-major ABI layout
nB = {k: np.ascontiguousarray(v.T) for k, v in d.items()}  # what a Python user holds: (n, B) / (m, B), C-ordered
t0 = time.perf_counter(); packed = {k: np.asfortranarray(v) for k, v in nB.items()}; repack_numpy = time.perf_counter() - t0
# the binding's loop: out[j*rows + i] = *(base + i*stride0 + j*stride1)
src = r'''
#include <stddef.h>
void repack(const char *base, long s0, long s1, int rows, int cols, double *out) {
    for (int j = 0; j < cols; ++j) for (int i = 0; i < rows; ++i) out[(size_t)j * rows + i] = *(const double *)(base + i * s0 + j * s1);
}'''
tmp = Path(tempfile.mkdtemp()); (tmp / "r.c").write_text(src)
subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", str(tmp / "r.so"), str(tmp / "r.c")], check=True)
R = ctypes.CDLL(str(tmp / "r.so"))
R.repack.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
t0 = time.perf_counter()
for k, v in nB.items():
    out = np.empty(v.size)
    R.repack(v.ctypes.data, v.strides[0], v.strides[1], v.shape[0], v.shape[1], out.ctypes.data)
repack_scalar = time.perf_counter() - t0
p = pkg.Parameters.default(stop_tol=0.0, max_iter=a.iters, use_presolve=False)
model = eng.create_model(base)
with Quiet():
    eng.solve_batched(model, d["C"], d["AL"], d["AU"], d["l"], d["u"], None, p)          # warm
    r0 = eng.solve_batched(model, *[packed[k].T for k in ("C", "AL", "AU", "l", "u")], None, p)
    r1 = eng.solve_batched_nB(model, nB["C"], nB["AL"], nB["AU"], nB["l"], nB["u"], None, p)
eng.free_model(model)
same = bool(np.array_equal(r0["x"], r1["x"].T) and np.array_equal(r0["y"], r1["y"].T) and np.array_equal(r0["iter"], r1["iter"]))
print(json.dumps(dict(workload=spec["name"], n=base["n"], m=base["m"], B=B, iters=a.iters, host_cores=os.cpu_count(),
                      repack_numpy_s=repack_numpy, repack_scalar_s=repack_scalar, call_colmajor_s=r0["call_seconds"],
                      call_rowmajor_s=r1["call_seconds"], reference_style_total_s=repack_scalar + r0["call_seconds"],
                      speedup_of_the_python_call=(repack_scalar + r0["call_seconds"]) / r1["call_seconds"], same_bits=same)))
