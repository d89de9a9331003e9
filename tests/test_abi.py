"""CPU: the C-ABI library loads, exports every symbol include/*.h declares, and the public structs
have the reference's byte layout (measured with offsetof on the reference headers)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

EXPECTED = {
    "HPRLP_parameters": (40, dict(max_iter=0, stop_tol=8, time_limit=16, device_number=24, check_iter=28, CUSPARSE_spmv=32,
                                  autotune_verbose=33, use_CR_scaling=34, use_Ruiz_scaling=35,
                                  use_Pock_Chambolle_scaling=36, use_bc_scaling=37, use_presolve=38)),
    "HPRLP_results": (160, dict(residuals=0, primal_obj=8, gap=16, time4=24, time6=32, time8=40, time=48, iter4=56, iter6=60,
                                iter8=64, iter=68, status=72, x=136, y=144, z=152)),
    "HPRLP_batched_results": (112, dict(m=0, n=4, batch_size=8, x=16, y=24, z=32, primal_obj=40, residuals=48, gap=56, iter=64,
                                        status=72, time=80, setup_time=88, solve_time=96, power_time=104)),
    "LP_info_cpu": (64, dict(m=0, n=4, A=8, AL=16, AU=24, c=32, l=40, u=48, obj_constant=56)),
    "sparseMatrix": (40, dict(row=0, col=4, numElements=8, colIndex=16, rowPtr=24, value=32)),
}


def test_struct_layout_compiled_header(tmp_path):
    src = ['#include <cstdio>', '#include <cstddef>', '#include "HPRLP.h"', '#include "hprlp_b200.h"', 'int main(){']
    for name, (size, fields) in EXPECTED.items():
        src.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            src.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    src.append("HPRLP_parameters p; printf(\"defaults %d %g %g %d %d %d %d %d %d %d %d %d\\n\", p.max_iter, p.stop_tol, p.time_limit,"
               " p.device_number, p.check_iter, p.CUSPARSE_spmv, p.autotune_verbose, p.use_CR_scaling, p.use_Ruiz_scaling,"
               " p.use_Pock_Chambolle_scaling, p.use_bc_scaling, p.use_presolve);")
    src.append("return 0;}")
    cpp = tmp_path / "probe.cpp"
    cpp.write_text("\n".join(src))
    exe = tmp_path / "probe"
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wno-invalid-offsetof", f"-I{ROOT/'include'}", str(cpp), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    got = {ln.split()[0]: ln.split()[1:] for ln in out}
    for name, (size, fields) in EXPECTED.items():
        assert int(got[name][0]) == size, name
        for f, off in fields.items():
            assert int(got[f"{name}.{f}"][0]) == off, f"{name}.{f}"
    assert got["defaults"] == ["2147483647", "0.0001", "3600", "0", "150", "0", "0", "1", "1", "1", "1", "1"]


def test_ctypes_mirror_matches(pkg):
    for cls, name in ((pkg.Parameters, "HPRLP_parameters"), (pkg.Results, "HPRLP_results"),
                      (pkg.BatchedResults, "HPRLP_batched_results"), (pkg.LPInfoCpu, "LP_info_cpu"),
                      (pkg.SparseMatrix, "sparseMatrix")):
        size, fields = EXPECTED[name]
        assert C.sizeof(cls) == size
        for f, off in fields.items():
            assert getattr(cls, f).offset == off, f"{name}.{f}"


def _declared_functions():
    names = []
    for h in ("HPRLP.h", "batched_solver.h", "hprlp_b200.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//.*", "", text)
        for mobj in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", text):
            nm = mobj.group(1)
            if nm not in ("defined",):
                names.append(nm)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(pkg):
    lib = C.CDLL(str(pkg.LIB_PATH))
    names = _declared_functions()
    assert set(pkg.REFERENCE_SYMBOLS) <= set(names)
    assert set(pkg.EXTENDED_SYMBOLS) <= set(names)
    for nm in names:
        assert hasattr(lib, nm), f"libhprlp.so does not export {nm}"


def test_static_library_and_cli_exist(pkg):
    assert (ROOT / "lib" / "libhprlp.a").exists()
    assert (ROOT / "build" / "solve_mps_file").exists()
    nm = subprocess.run(["nm", "-g", "--defined-only", str(ROOT / "lib" / "libhprlp.a")], capture_output=True, text=True).stdout
    for sym in pkg.REFERENCE_SYMBOLS:
        assert re.search(rf"\bT {sym}\b", nm), sym


def test_no_vendor_sparse_blas_on_link_line(pkg):
    """north_star: no cuSPARSE / cuBLAS / cuSOLVER anywhere in the engine (cuRAND only for the start vector)."""
    out = subprocess.run(["readelf", "-d", str(pkg.LIB_PATH)], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not [n for n in needed if re.search(r"cusparse|cublas|cusolver", n)], needed
    assert any("curand" in n for n in needed)


def test_error_behaviour_without_gpu(engine, pkg):
    """NULL / bad input handling is host-side and must match the reference (src/HPRLP.cu:329-337,493-498)."""
    bad = dict(pkg.TOY_LP)
    assert not engine.lib.create_model_from_arrays(0, 2, 4, None, None, None, None, None, None, None, None, False)
    assert not engine.lib.create_model_from_arrays(2, 2, 4, None, None, None, None, None, None, None, None, False)
    assert not engine.lib.create_model_from_mps(None)
    assert not engine.lib.create_model_from_mps(b"/nonexistent/file.mps")
    res = engine.lib.solve(None, None)
    assert res.status == b"ERROR" and not res.x and not res.y and not res.z
    engine.lib.free_model(None)
    br = engine.lib.solve_batched(None, 3, None, None, None, None, None, None, None)
    assert br.batch_size == 3 and C.string_at(br.status, 5) == b"ERROR"
    engine.lib.free_batched_results(C.byref(br))
    assert not br.status
    del bad


def test_cli_argument_handling_matches_reference_cli():
    """build/solve_mps_file: the reference CLI's diagnostics and exit codes for the paths that need no GPU
    (reference src/solve_mps_file.cpp:34-122)."""
    import subprocess
    exe = str(ROOT / "build" / "solve_mps_file")
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True, timeout=60)
    r = run("--help")
    assert r.returncode == 0 and "Usage:" in r.stdout and "--presolve <true/false>" in r.stdout
    r = run()
    assert r.returncode == 1 and "Error: Input file is required. Use -i or --input option." in r.stderr
    r = run("-i")
    assert r.returncode == 1 and "Missing value for option: -i" in r.stderr
    r = run("--tol")
    assert r.returncode == 1 and "Missing value for option: --tol" in r.stderr
    r = run("--bogus", "1")
    assert r.returncode == 1 and "Unknown option: --bogus" in r.stderr
    r = run("-i", "/no/such/file.mps")
    assert r.returncode == 1 and "Input file does not exist: /no/such/file.mps" in r.stderr


def test_error_paths_same_as_reference_build(engine, reference, pkg):
    """The host-only error paths side by side with the reference's own build: same NULL / non-NULL results, same status
    strings, same batched error shape."""
    outcomes = []
    for lib in (engine, reference):
        L = lib.lib
        o = []
        o.append(bool(L.create_model_from_arrays(0, 2, 4, None, None, None, None, None, None, None, None, False)))
        o.append(bool(L.create_model_from_arrays(2, 2, 4, None, None, None, None, None, None, None, None, False)))
        o.append(bool(L.create_model_from_arrays(2, 0, 4, None, None, None, None, None, None, None, None, True)))
        o.append(bool(L.create_model_from_mps(None)))
        o.append(bool(L.create_model_from_mps(b"/nonexistent/file.mps")))
        r = L.solve(None, None)
        o.append((r.status, bool(r.x), bool(r.y), bool(r.z), r.iter))
        for B in (3, 0, -2):
            br = L.solve_batched(None, B, None, None, None, None, None, None, None)
            o.append((br.batch_size, C.string_at(br.status, 5) if br.status else None, bool(br.x), bool(br.iter)))
            L.free_batched_results(C.byref(br))
            o.append(bool(br.status))
        outcomes.append(o)
    assert outcomes[0] == outcomes[1], outcomes
