"""CPU: the reference's own consumers build UNCHANGED against include/ + lib/ (drop-in boundary).
The sources are compiled from where they lie under /root/reference (nothing is copied into the repo); the GPU box has no
reference tree, so these tests run in the build container only."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("HPRLP_REFERENCE_DIR", "/root/reference"))
NVCC = "/usr/local/cuda/bin/nvcc"

pytestmark = pytest.mark.skipif(not (REF / "examples").is_dir(), reason="reference tree not present")


def _sh(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, **kw)
    assert r.returncode == 0, (" ".join(map(str, cmd)), r.stdout[-1500:], r.stderr[-3000:])
    return r


@pytest.mark.parametrize("src", ["example_direct_lp.c", "example_mps_file.c", "example_batched_lp.c"])
def test_reference_c_examples_compile_and_link(tmp_path, src):
    """examples/c/Makefile:45-54: nvcc -x cu, -lhprlp -lcublas -lcusolver -lcusparse -lstdc++."""
    _sh([NVCC, "-w", "-O2", "-arch=sm_100", "-x", "cu", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include",
         str(REF / "examples" / "c" / src), "-o", str(tmp_path / "a.out"), f"-L{ROOT / 'lib'}", "-L/usr/local/cuda/lib64",
         "-lhprlp", "-lcublas", "-lcusolver", "-lcusparse", "-lstdc++", "-Xlinker", "-rpath", "-Xlinker", str(ROOT / "lib")])


@pytest.mark.parametrize("src", ["example_direct_lp.cpp", "example_mps_file.cpp"])
def test_reference_cpp_examples_compile_and_link(tmp_path, src):
    """examples/cpp/Makefile: nvcc --std=c++17, -lhprlp -lcublas -lcusolver -lcusparse (example_direct_lp.cpp takes INFINITY
    from the library header)."""
    _sh([NVCC, "-w", "-O2", "--std=c++17", "-arch=sm_100", f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include",
         str(REF / "examples" / "cpp" / src), "-o", str(tmp_path / "a.out"), f"-L{ROOT / 'lib'}", "-L/usr/local/cuda/lib64",
         "-lhprlp", "-lcublas", "-lcusolver", "-lcusparse", "-Xlinker", "-rpath", "-Xlinker", str(ROOT / "lib")])


def test_reference_pybind_module_builds_and_runs_host_calls(tmp_path):
    """bindings/python/src/hprlp_pybind.cpp (includes HPRLP.h, structs.h, mps_reader.h) against this library: build,
    import, and drive the host-only calls (MPS parse, model from arrays, parameter struct)."""
    pybind11 = pytest.importorskip("pybind11")
    import sysconfig
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = tmp_path / f"_hprlp_core{ext}"
    _sh(["g++", "-O1", "-shared", "-fPIC", "-std=c++17", f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}",
         f"-I{ROOT / 'include'}", "-I/usr/local/cuda/include", str(REF / "bindings" / "python" / "src" / "hprlp_pybind.cpp"),
         f"-L{ROOT / 'lib'}", "-lhprlp", f"-Wl,-rpath,{ROOT / 'lib'}", "-o", str(out)])
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {str(tmp_path)!r})
        import _hprlp_core as core
        p = core.Parameters()
        assert (p.max_iter, p.stop_tol, p.time_limit, p.check_iter, p.use_presolve) == (2**31 - 1, 1e-4, 3600.0, 150, True)
        m = core.create_model_from_mps({str(ROOT / 'tests' / 'golden' / 'model.mps')!r})
        assert m.is_valid() and (m.m, m.n) == (2, 2)
        core.free_model(m)
        rp = np.array([0, 2, 4], np.int32); ci = np.array([0, 1, 0, 1], np.int32); v = np.array([1., 2., 3., 1.])
        m2 = core.create_model_from_arrays(2, 2, 4, rp, ci, v, np.array([-np.inf, -np.inf]), np.array([10., 12.]),
                                           np.zeros(2), np.full(2, np.inf), np.array([-3., -5.]), False)
        assert m2.is_valid() and (m2.m, m2.n) == (2, 2)
        core.free_model(m2)
        print("BINDING_OK")
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "BINDING_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])

    # the reference's pure-Python package on top of that module (a temporary copy next to the freshly built core, as its
    # setup.py lays it out): Model.from_arrays / Model.from_mps / Parameters through this library
    import shutil
    pkg_dir = tmp_path / "site" / "hprlp"
    shutil.copytree(REF / "bindings" / "python" / "hprlp", pkg_dir)
    shutil.copy(out, pkg_dir / out.name)
    code2 = textwrap.dedent(f"""
        import sys, numpy as np, scipy.sparse as sp
        sys.path.insert(0, {str(tmp_path / 'site')!r})
        import hprlp
        A = sp.csr_matrix(np.array([[1., 2.], [3., 1.]]))
        m = hprlp.Model.from_arrays(A, np.array([-np.inf, -np.inf]), np.array([10., 12.]), np.zeros(2), np.full(2, np.inf),
                                    np.array([-3., -5.]))
        assert m.is_valid() and (m.m, m.n) == (2, 2)
        m2 = hprlp.Model.from_mps({str(ROOT / 'tests' / 'golden' / 'model.mps')!r})
        assert (m2.m, m2.n) == (2, 2)
        p = hprlp.Parameters()
        p.stop_tol = 1e-6
        print("PACKAGE_OK")
    """)
    r = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=300)
    assert "PACKAGE_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_reference_matlab_mex_gateway_compiles_against_these_headers():
    """bindings/matlab/src/hprlp_mex.cpp includes HPRLP.h, mps_reader.h and preprocess.h from the library's include
    directory (bindings/matlab/install.sh:85-99).  No MATLAB in the image: tests/stubs/{mex,matrix}.h declare the MEX
    API it uses, and the gateway is compiled (not linked) against include/ to prove the headers offer every type and
    function it needs."""
    src = REF / "bindings" / "matlab" / "src" / "hprlp_mex.cpp"
    if not src.exists():
        pytest.skip("MATLAB binding source not present")
    _sh(["g++", "-std=c++17", "-fsyntax-only", f"-I{ROOT / 'tests' / 'stubs'}", f"-I{ROOT / 'include'}",
         "-I/usr/local/cuda/include", str(src)])


def test_julia_wrapper_symbols_and_struct_sizes():
    """bindings/julia/package/src/wrapper.jl ccalls the seven symbols by name and mirrors the structs by size: the names must
    be exported unmangled and the sizes must be the ones the wrapper hard-codes through its field lists (40 / 160 / 112 /
    64 / 40 bytes)."""
    wrapper = REF / "bindings" / "julia" / "package" / "src" / "wrapper.jl"
    if not wrapper.exists():
        pytest.skip("Julia wrapper not present")
    import re
    text = wrapper.read_text()
    called = set(re.findall(r"ccall\(\s*\(\s*:(\w+)", text))
    assert called, "no ccall found in the Julia wrapper"
    out = subprocess.run(["nm", "-D", "--defined-only", str(ROOT / "lib" / "libhprlp.so")], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    libc = {"free", "malloc"}
    missing = sorted(s for s in called if s not in exported and s not in libc)
    assert not missing, missing
