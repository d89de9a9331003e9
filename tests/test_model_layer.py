"""CPU: host model layer (create_model_from_arrays / create_model_from_mps / CSC conversion /
transpose order).  When the reference's own build is present (oracle/_ref) the LP_info_cpu arrays
are compared bit-exactly against it ("MPS parsing, CSR/CSC construction ... bit-exact")."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

GOLD = __import__("pathlib").Path(__file__).resolve().parent / "golden"
ROOT = GOLD.parent.parent


def _rand_csr(rng, m, n, density, empty_rows=()):
    A = sp.random(m, n, density=density, random_state=rng, format="csr", dtype=np.float64)
    A.data[:] = rng.uniform(-2, 2, A.nnz)
    A = A.tolil()
    for r in empty_rows:
        A.rows[r] = []
        A.data[r] = []
    A = A.tocsr()
    A.sort_indices()
    return A


def test_toy_mps_parses_to_known_model(engine):
    model = engine.create_model_from_mps(GOLD / "model.mps")
    assert model
    a = engine.model_arrays(model)
    engine.free_model(model)
    assert (a["m"], a["n"], a["nnz"]) == (2, 2, 4)
    assert a["rowPtr"].tolist() == [0, 2, 4] and a["colIndex"].tolist() == [0, 1, 0, 1]
    assert a["values"].tolist() == [1.0, 2.0, 3.0, 1.0]
    assert np.all(np.isneginf(a["AL"])) and a["AU"].tolist() == [10.0, 12.0]
    assert a["l"].tolist() == [0.0, 0.0] and np.all(np.isposinf(a["u"])) and a["c"].tolist() == [-3.0, -5.0]
    assert a["obj_constant"] == 0.0


def test_tricky_mps_semantics(engine):
    model = engine.create_model_from_mps(GOLD / "tricky.mps")
    assert model
    a = engine.model_arrays(model)
    engine.free_model(model)
    assert (a["m"], a["n"]) == (5, 5)
    dense = sp.csr_matrix((a["values"], a["colIndex"], a["rowPtr"]), shape=(5, 5)).toarray()
    want = np.array([[2.0, 1.0, 0, 0, 0], [-1.0, 0, 1.0, 2.5, 0], [0, 3.0, 1.0, 0, 0], [0, -4.5, 0, 1.0, 0],
                     [0, 0, 0.5, -1.0, 7.0]])
    assert np.array_equal(dense, want)
    assert a["c"].tolist() == [1.5, -2.25, 0.0, 0.125, 0.0]
    assert a["obj_constant"] == 3.5               # RHS on the objective row: c0 = -(-3.5)
    # e1: rhs 4, range +2 -> [4,6]; l1: rhs 6, |range| 3 -> [3,6]; g1: rhs 1, |range| 1.5 -> [1,2.5];
    # e2: rhs -2, range -0.5 -> [-2.5,-2]; g2: rhs from the rim set "rhs2" ignored -> [0, inf)
    assert a["AL"].tolist() == [4.0, 3.0, 1.0, -2.5, 0.0]
    assert a["AU"][:4].tolist() == [6.0, 6.0, 2.5, -2.0] and np.isposinf(a["AU"][4])
    # x1: UP -1 with no lower -> l=-inf; x2: MI -> (-inf, inf); x3: integer-marked default [0,1];
    # x4: FX 2.5; x5: FR, the rim bound set "bnd2" ignored
    assert np.isneginf(a["l"][0]) and a["u"][0] == -1.0
    assert np.isneginf(a["l"][1]) and np.isposinf(a["u"][1])
    assert (a["l"][2], a["u"][2]) == (0.0, 1.0)
    assert (a["l"][3], a["u"][3]) == (2.5, 2.5)
    assert np.isneginf(a["l"][4]) and np.isposinf(a["u"][4])


@pytest.mark.parametrize("name", ["model.mps", "tricky.mps", "dup.mps"])
def test_mps_bit_exact_vs_reference_build(engine, reference, name):
    mine = engine.create_model_from_mps(GOLD / name)
    ref = reference.create_model_from_mps(GOLD / name)
    assert mine and ref
    a, b = engine.model_arrays(mine), reference.model_arrays(ref)
    engine.free_model(mine); reference.free_model(ref)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True), k


def test_gz_mps(engine, tmp_path):
    import gzip
    p = tmp_path / "model.mps.gz"
    p.write_bytes(gzip.compress((GOLD / "model.mps").read_bytes()))
    model = engine.create_model_from_mps(p)
    assert model
    assert engine.model_arrays(model)["values"].tolist() == [1.0, 2.0, 3.0, 1.0]
    engine.free_model(model)


@pytest.mark.parametrize("seed,m,n,dens", [(0, 7, 5, 0.5), (1, 40, 90, 0.1), (2, 300, 120, 0.03)])
def test_arrays_roundtrip_and_csc(engine, pkg, seed, m, n, dens):
    rng = np.random.default_rng(seed)
    A = _rand_csr(rng, m, n, dens, empty_rows=(0, m - 1))
    lp = dict(m=m, n=n, rowPtr=A.indptr, colIndex=A.indices, values=A.data, AL=rng.normal(size=m), AU=rng.normal(size=m) + 3,
              l=np.zeros(n), u=np.full(n, np.inf), c=rng.normal(size=n))
    model = engine.create_model(lp)
    a = engine.model_arrays(model)
    engine.free_model(model)
    assert np.array_equal(a["rowPtr"], A.indptr) and np.array_equal(a["colIndex"], A.indices) and np.array_equal(a["values"], A.data)
    for k in ("AL", "AU", "l", "u", "c"):
        assert np.array_equal(a[k], lp[k])
    # CSC input (is_csc=true) must produce the same CSR, entries ordered by column within each row
    Acsc = A.tocsc(); Acsc.sort_indices()
    lpc = dict(lp, rowPtr=Acsc.indptr, colIndex=Acsc.indices, values=Acsc.data)
    model = engine.create_model(lpc, is_csc=True)
    b = engine.model_arrays(model)
    engine.free_model(model)
    assert np.array_equal(b["rowPtr"], A.indptr) and np.array_equal(b["colIndex"], A.indices) and np.array_equal(b["values"], A.data)


@pytest.mark.parametrize("seed", [3, 4])
def test_arrays_bit_exact_vs_reference_build(engine, reference, seed):
    rng = np.random.default_rng(seed)
    m, n = 60, 45
    A = _rand_csr(rng, m, n, 0.08, empty_rows=(5,))
    Acsc = A.tocsc(); Acsc.sort_indices()
    base = dict(m=m, n=n, AL=rng.normal(size=m), AU=rng.normal(size=m) + 3, l=np.zeros(n), u=np.full(n, np.inf), c=rng.normal(size=n))
    for is_csc, M in ((False, A), (True, Acsc)):
        lp = dict(base, rowPtr=M.indptr, colIndex=M.indices, values=M.data)
        mine, ref = engine.create_model(lp, is_csc=is_csc), reference.create_model(lp, is_csc=is_csc)
        a, b = engine.model_arrays(mine), reference.model_arrays(ref)
        engine.free_model(mine); reference.free_model(ref)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (is_csc, k)


def test_transpose_order_matches_reference_counting_sort(oracle):
    """A^T entry order = stable counting sort by column (reference CSR_transpose_host, src/utils.cu:203-232):
    within a transposed row, entries are ordered by original row index."""
    rng = np.random.default_rng(7)
    A = _rand_csr(rng, 50, 30, 0.2, empty_rows=(3,))
    trp, tci, tv = oracle.transpose(50, 30, A.indptr, A.indices, A.data)
    AT = A.T.tocsr(); AT.sort_indices()
    assert np.array_equal(trp, AT.indptr) and np.array_equal(tci, AT.indices) and np.array_equal(tv, AT.data)


# ------------------------------------------------------------------------------------------------
# PSLP presolve hand-off (host): the reduced problem must be bit-identical to the reference's
# ("presolve output ... bit-exact"); the reference runs PSLP in a forked worker, we run it in-process.
# ------------------------------------------------------------------------------------------------
def _presolve_lp(pkg, seed):
    """LP with presolve opportunities: singleton rows/columns, duplicate rows, fixed variables, empty column."""
    rng = np.random.default_rng(seed)
    base = pkg.synth_lp("uniform", 60, 140, 60 * 6, seed=pkg.SEED + seed)
    A = sp.csr_matrix((base["values"], base["colIndex"], base["rowPtr"]), shape=(60, 140)).tolil()
    A[3, :] = 0; A[3, 7] = 2.0                      # singleton row
    A[10, :] = A[11, :] * 2.0                      # parallel rows
    A[:, 20] = 0                                   # empty column
    A[:, 21] = 0; A[5, 21] = 1.5                   # singleton column
    A = A.tocsr(); A.sort_indices(); A.eliminate_zeros()
    lp = dict(m=60, n=140, rowPtr=A.indptr.astype(np.int32), colIndex=A.indices.astype(np.int32), values=A.data.copy())
    v = pkg.synth_vectors(lp, pkg.SEED + seed, pkg.SEED + seed)      # bounds/cost rebuilt from a feasible primal-dual pair
    lp.update({k: v[k] for k in ("AL", "AU", "l", "u", "c")})
    lp["l"][30] = lp["u"][30] = v["xs"][30]         # fixed variable
    del rng
    return lp


@pytest.mark.parametrize("seed", [0, 1])
def test_presolve_reduced_model_bit_exact_vs_reference(pkg, engine, reference, seed):
    import ctypes as C
    lp = _presolve_lp(pkg, seed) if seed else pkg.TOY_LP
    p = pkg.Parameters.default()
    outs = {}
    # ours: C hook around the in-process PSLP bridge
    model = engine.create_model(lp)
    red, h = pkg.LPInfoCpu(), C.c_void_p()
    engine.lib.hprlp_b200_presolve.argtypes = [C.POINTER(pkg.LPInfoCpu), C.POINTER(pkg.Parameters), C.POINTER(pkg.LPInfoCpu), C.POINTER(C.c_void_p)]
    ok = engine.lib.hprlp_b200_presolve(model, C.byref(p), C.byref(red), C.byref(h))
    if not ok:
        engine.free_model(model)
        pytest.skip("PSLP not linked into this build of libhprlp.so")
    outs["new"] = engine.model_arrays(C.pointer(red))
    engine.lib.hprlp_b200_presolve_free.argtypes = [C.c_void_p, C.POINTER(pkg.LPInfoCpu)]
    engine.lib.hprlp_b200_presolve_free(h, C.byref(red))
    engine.free_model(model)
    # reference: its exported C++ bridge (forked PSLP worker), src/pslp_integration.cpp:628
    f = getattr(reference.lib, "_Z26run_embedded_pslp_presolvePK11LP_info_cpuPK16HPRLP_parametersPS_PPv")
    f.restype = C.c_bool
    f.argtypes = [C.POINTER(pkg.LPInfoCpu), C.POINTER(pkg.Parameters), C.POINTER(pkg.LPInfoCpu), C.POINTER(C.c_void_p)]
    model = reference.create_model(lp)
    red2, h2 = pkg.LPInfoCpu(), C.c_void_p()
    assert f(model, C.byref(p), C.byref(red2), C.byref(h2))
    outs["ref"] = reference.model_arrays(C.pointer(red2))
    g = getattr(reference.lib, "_Z28free_embedded_pslp_presolverPv")
    g.argtypes = [C.c_void_p]
    g(h2)
    reference.free_model(model)
    a, b = outs["new"], outs["ref"]
    assert (a["m"], a["n"], a["nnz"]) == (b["m"], b["n"], b["nnz"])
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True), k


def _write_random_mps(path, seed=7, m=1500, n=12000, dups=True):
    """A few hundred KB of free-format MPS with the spellings and layouts the reader must agree on with the reference:
    two entries per card, comments and blank lines inside sections, integer markers, duplicate cards, '+'/exponent/'D'
    number spellings, objective entries in the middle of a column, RANGES, all bound types."""
    rng = np.random.default_rng(seed)
    fmt = [lambda v: f"{v:.17g}", lambda v: f"{v:+.6e}", lambda v: f"{v:.3f}", lambda v: f"{int(v)}" if v == int(v) else f"{v:.9g}",
           lambda v: f"{v:.4f}D0"]
    with open(path, "w") as f:
        f.write("* generated\nNAME rnd\nROWS\n N obj\n N rim\n")
        types = rng.choice(list("ELG"), m)
        for i in range(m):
            f.write(f" {types[i]}  R{i}\n")
        f.write("COLUMNS\n")
        for j in range(n):
            if j % 997 == 0:
                f.write("    MARKER  'MARKER'  'INTORG'\n" if (j // 997) % 2 == 0 else "    MARKER  'MARKER'  'INTEND'\n")
            if j % 501 == 0:
                f.write("* a comment inside COLUMNS\n\n")
            k = int(rng.integers(1, 9))
            rows = rng.choice(m, k, replace=False)
            vals = np.round(rng.uniform(-5, 5, k), int(rng.integers(0, 6)))
            cards = [(f"R{r}", v) for r, v in zip(rows, vals)]
            if rng.random() < 0.5:
                cards.insert(int(rng.integers(0, len(cards) + 1)), ("obj", float(np.round(rng.uniform(-2, 2), 3))))
            if dups and rng.random() < 0.05:
                cards.append(cards[0])                      # duplicate (row, col) card
            if rng.random() < 0.02:
                cards.append(("rim", 1.0))
            i = 0
            while i < len(cards):
                fm = fmt[int(rng.integers(0, len(fmt)))]
                if i + 1 < len(cards) and rng.random() < 0.5:
                    f.write(f"    C{j}  {cards[i][0]}  {fm(cards[i][1])}  {cards[i + 1][0]}  {fm(cards[i + 1][1])}\n")
                    i += 2
                else:
                    f.write(f"    C{j}\t{cards[i][0]}   {fm(cards[i][1])}\n")
                    i += 1
        f.write("RHS\n    rhs  obj  -1.25\n")
        for i in range(0, m, 2):
            f.write(f"    rhs  R{i}  {rng.uniform(-3, 3):.6g}  R{min(i + 1, m - 1)}  {rng.uniform(-3, 3):.6g}\n")
        f.write("RANGES\n")
        for i in range(0, m, 7):
            f.write(f"    rng  R{i}  {rng.uniform(-2, 2):.5g}\n")
        f.write("BOUNDS\n")
        for j in range(0, n, 3):
            bt = ["UP", "LO", "FX", "FR", "MI", "PL", "BV"][int(rng.integers(0, 7))]
            f.write(f" {bt} bnd  C{j}" + ("" if bt in ("FR", "MI", "PL", "BV") else f"  {rng.uniform(-4, 4):.6g}") + "\n")
        f.write("ENDATA\n")


@pytest.mark.parametrize("threads,dups", [(1, True), (7, True), (5, False)])
def test_parallel_mps_reader_bit_exact_vs_reference(pkg, reference, tmp_path, threads, dups):
    """The multi-threaded COLUMNS parse (chunk boundaries fall inside column runs) against the reference's sequential
    reader, byte for byte; run in a subprocess so OMP_NUM_THREADS takes effect."""
    import subprocess, sys, textwrap
    path = tmp_path / "rnd.mps"
    _write_random_mps(path, dups=dups)      # dups=False: the direct sorted-list-is-CSR path of the reader
    assert path.stat().st_size > 8 * 65536          # enough bytes for several parse chunks
    code = textwrap.dedent(f"""
        import sys, numpy as np
        sys.path.insert(0, {str(ROOT)!r})
        import __graft_entry__ as graft
        pkg = graft.load_package()
        out = []
        for lib in (pkg.load_engine(), pkg.load_reference()):
            mdl = lib.create_model_from_mps({str(path)!r})
            assert mdl
            out.append(lib.model_arrays(mdl)); lib.free_model(mdl)
        a, b = out
        bad = [k for k in a if not np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True)]
        print("MISMATCH" if bad else "IDENTICAL", bad, a["m"], a["n"], a["nnz"])
    """)
    env = dict(**__import__("os").environ, OMP_NUM_THREADS=str(threads))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert "IDENTICAL" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


_EDGE_MPS = {
    # CRLF line endings, tab separators, blank and comment lines between cards
    "crlf_tabs": "NAME\tt\r\nROWS\r\n N\tobj\r\n L\tr1\r\n G\tr2\r\n\r\n* comment\r\nCOLUMNS\r\n\tx\tobj\t1\tr1\t2\r\n\tx\tr2\t3\r\n\ty\tr1\t-1.5\r\n"
                 "RHS\r\n\trhs\tr1\t4\tr2\t1\r\nBOUNDS\r\n UP\tb\tx\t10\r\nENDATA\r\n",
    # no ENDATA, no trailing newline, RHS/RANGES/BOUNDS absent
    "no_endata": "NAME t\nROWS\n N obj\n E r1\nCOLUMNS\n x obj 1 r1 1\n y r1 2",
    # unknown row in COLUMNS / RHS / RANGES, unknown column in BOUNDS, short lines, unknown bound type, rim RHS/RANGES/BOUNDS sets
    "unknowns": "NAME t\nROWS\n N obj\n L r1\n G r2\nCOLUMNS\n x obj 1 nosuch 5\n x r1 1\n y r2 2 r1 1\n z\n z r1\n"
                "RHS\n rhs r1 3 nosuch 1\n rhs2 r2 9\n rhs\nRANGES\n rng nosuch 1\n rng r1 2\n rng2 r2 5\n"
                "BOUNDS\n UP b nosuch 1\n XX b x 1\n LO b y -2\n UP b2 y 5\n UP b\nENDATA\n",
    # duplicate row name (re-bound), N row after constraints (rim), second objective entry overriding the first
    "rebinding": "NAME t\nROWS\n N obj\n L r1\n G r1\n E r2\n N late\nCOLUMNS\n x obj 1 r1 1\n x obj 7\n x late 3\n y r2 1 r1 4\nRHS\n rhs r1 2 r2 1\nENDATA\n",
    # number spellings: sign, exponent, leading dot, Fortran D (atof stops at D), inf, trailing junk
    "numbers": "NAME t\nROWS\n N obj\n L r1\n G r2\n E r3\nCOLUMNS\n a obj +1.5 r1 1e0\n a r2 .5 r3 -2.5E-1\n b r1 1.0D3 r2 3abc\n b r3 1e400\n c r1 -0 r2 0x10\n"
               "RHS\n rhs r1 +4 r2 -1e-3\n rhs r3 1e30\nBOUNDS\n UP b a 1e30\n LO b b -inf\n UP b c Infinity\nENDATA\n",
    # negative upper bound without a lower bound, MI then UP, FX, BV, PL, integer markers with default [0,1]
    "bounds": "NAME t\nROWS\n N obj\n L r1\nCOLUMNS\n a r1 1 obj 1\n MARKER 'MARKER' 'INTORG'\n b r1 1\n c r1 1\n MARKER 'MARKER' 'INTEND'\n d r1 1\n e r1 1\n f r1 1\n"
              "RHS\n rhs r1 5\nBOUNDS\n UP bnd a -3\n MI bnd d\n UP bnd d 2\n FX bnd e 1.5\n BV bnd f\n PL bnd c\n LI bnd b 0\n UI bnd b 7\nENDATA\n",
    # RANGES on E rows with both signs, on L and G rows, on the objective (error), OBJSENSE MAX (ignored)
    "ranges": "NAME t\nOBJSENSE\n MAX\nROWS\n N obj\n E e1\n E e2\n L l1\n G g1\nCOLUMNS\n x obj 1 e1 1\n x e2 1 l1 1\n x g1 1\n"
              "RHS\n rhs e1 1 e2 1\n rhs l1 1 g1 1\nRANGES\n rng e1 2 e2 -2\n rng l1 -3 g1 -4\n rng obj 1\nENDATA\n",
}


@pytest.mark.parametrize("name", sorted(_EDGE_MPS))
def test_mps_edge_cases_bit_exact_vs_reference(engine, reference, tmp_path, name):
    """Reader corner cases (tokenising, error paths that must not change the model, number spellings, bound and range
    rules): the arrays must equal the reference reader's byte for byte."""
    path = tmp_path / f"{name}.mps"
    path.write_bytes(_EDGE_MPS[name].encode())
    mine = engine.create_model_from_mps(path)
    ref = reference.create_model_from_mps(path)
    assert bool(mine) == bool(ref)
    if not mine:
        return
    a, b = engine.model_arrays(mine), reference.model_arrays(ref)
    engine.free_model(mine); reference.free_model(ref)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k]), equal_nan=True), (name, k, a[k], b[k])


def test_csc_input_multithreaded_conversion_bit_exact(engine, reference, pkg):
    """CSC -> CSR of create_model_from_arrays(is_csc=true) on a matrix large enough for the multi-threaded conversion
    (csrc/host_utils.cpp): same arrays as the reference's single-threaded counting sort, and as the CSR input path."""
    lp = pkg.synth_lp("powerlaw", 40000, 90000, 3_000_000)
    A = sp.csr_matrix((lp["values"], lp["colIndex"], lp["rowPtr"]), shape=(lp["m"], lp["n"]))
    Ac = A.tocsc(); Ac.sort_indices()
    lpc = dict(lp, rowPtr=Ac.indptr.astype(np.int32), colIndex=Ac.indices.astype(np.int32), values=Ac.data)
    got = {}
    for tag, lib, d, csc in (("ours_csc", engine, lpc, True), ("ref_csc", reference, lpc, True), ("ours_csr", engine, lp, False)):
        mdl = lib.create_model(d, is_csc=csc)
        got[tag] = lib.model_arrays(mdl)
        lib.free_model(mdl)
    for other in ("ref_csc", "ours_csr"):
        for k in got["ours_csc"]:
            assert np.array_equal(np.asarray(got["ours_csc"][k]), np.asarray(got[other][k])), (other, k)


@pytest.mark.parametrize("fault", ["segv", "abort"])
def test_presolve_crash_falls_back_to_original_model(pkg, engine, monkeypatch, fault):
    """The reference runs PSLP in a forked worker, so a presolver crash only loses the presolve
    (src/pslp_integration.cpp:677-691).  Here PSLP runs in-process inside a signal guard: a synchronous fault on the
    presolving thread must come back as "presolve failed" (the caller then solves the original model), not kill the
    process.  HPRLP_TEST_PRESOLVE_FAULT raises the fault inside the guarded region."""
    lp = pkg.synth_lp("uniform", 60, 150, 600)
    p = pkg.Parameters.default()
    model = engine.create_model(lp)
    engine.lib.hprlp_b200_presolve.argtypes = [C.POINTER(pkg.LPInfoCpu), C.POINTER(pkg.Parameters), C.POINTER(pkg.LPInfoCpu), C.POINTER(C.c_void_p)]
    engine.lib.hprlp_b200_presolve_free.argtypes = [C.c_void_p, C.POINTER(pkg.LPInfoCpu)]
    red, h = pkg.LPInfoCpu(), C.c_void_p()
    if not engine.lib.hprlp_b200_presolve(model, C.byref(p), C.byref(red), C.byref(h)):
        engine.free_model(model)
        pytest.skip("PSLP not linked into this build of libhprlp.so")
    engine.lib.hprlp_b200_presolve_free(h, C.byref(red))
    monkeypatch.setenv("HPRLP_TEST_PRESOLVE_FAULT", fault)
    red, h = pkg.LPInfoCpu(), C.c_void_p()
    assert engine.lib.hprlp_b200_presolve(model, C.byref(p), C.byref(red), C.byref(h)) == 0     # survived, reported failure
    monkeypatch.delenv("HPRLP_TEST_PRESOLVE_FAULT")
    red, h = pkg.LPInfoCpu(), C.c_void_p()
    assert engine.lib.hprlp_b200_presolve(model, C.byref(p), C.byref(red), C.byref(h)) == 1     # and the bridge still works
    engine.lib.hprlp_b200_presolve_free(h, C.byref(red))
    engine.free_model(model)
