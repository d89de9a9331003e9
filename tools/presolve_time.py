"""Host-side PSLP presolve time of a bench workload (CPU only): ours (in-process bridge, csrc/presolve.cpp) and, when the
reference build is present, the reference's (forked worker + pipes, src/pslp_integration.cpp).
python tools/presolve_time.py [c2|c3|small]"""
import ctypes as C, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
from bench import WORKLOADS
pkg = graft.load_package()
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
spec = WORKLOADS[w]
lp = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
out = dict(workload=w, m=lp["m"], n=lp["n"], nnz=int(lp["values"].shape[0]), host_cores=os.cpu_count())
eng = pkg.load_engine()
p = pkg.Parameters.default(use_presolve=True)
model = eng.create_model(lp)
red, h = pkg.LPInfoCpu(), C.c_void_p()
f = eng.lib.hprlp_b200_presolve
f.restype = C.c_int
f.argtypes = [C.POINTER(pkg.LPInfoCpu), C.POINTER(pkg.Parameters), C.POINTER(pkg.LPInfoCpu), C.POINTER(C.c_void_p)]
t0 = time.perf_counter(); ok = f(model, C.byref(p), C.byref(red), C.byref(h)); out["ours_s"] = time.perf_counter() - t0
out["ours_ok"] = int(ok)
if ok:
    out["reduced"] = dict(m=red.m, n=red.n, nnz=red.A.contents.numElements)
    g = eng.lib.hprlp_b200_presolve_free
    g.argtypes = [C.c_void_p, C.POINTER(pkg.LPInfoCpu)]
    g(h, C.byref(red))
eng.free_model(model)
if pkg.REF_LIB_PATH.exists():
    ref = pkg.load_reference()
    model = ref.create_model(lp)
    red, h = pkg.LPInfoCpu(), C.c_void_p()
    f = getattr(ref.lib, "_Z26run_embedded_pslp_presolvePK11LP_info_cpuPK16HPRLP_parametersPS_PPv")
    f.restype = C.c_bool
    f.argtypes = [C.POINTER(pkg.LPInfoCpu), C.POINTER(pkg.Parameters), C.POINTER(pkg.LPInfoCpu), C.POINTER(C.c_void_p)]
    t0 = time.perf_counter(); ok = f(model, C.byref(p), C.byref(red), C.byref(h)); out["reference_s"] = time.perf_counter() - t0
    out["reference_ok"] = int(bool(ok))
    ref.free_model(model)
print(json.dumps(out))
