"""GPU (-m gpu): behaviour at the edges of the C ABI -- nothing is thrown across it, cached device memory can be handed
back, and the in-kernel look-back of rows cut by item boundaries is reproducible under concurrency."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_bytes():
    import torch
    free, _ = torch.cuda.mem_get_info(0)
    return free


def test_cuda_failure_becomes_error_status_not_an_exception(pkg, engine):
    """A CUDA failure inside solve() (here: a device that does not exist) must come back as status "ERROR" with null
    vectors, as the reference's own error results do (src/HPRLP.cu:66-79) -- never as a C++ exception through ctypes."""
    lp = pkg.synth_lp("uniform", 300, 900, 3600)
    model = engine.create_model(lp)
    bad = pkg.Parameters.default(use_presolve=False, device_number=63)
    r = engine.solve(model, bad, main=True)
    assert r["status"] == "ERROR" and r["x"] is None and r["y"] is None and r["z"] is None
    assert engine.lib.hprlp_b200_engine_create(model, C.byref(bad)) is None
    rb = engine.solve_batched(model, np.zeros((2, 900)), np.zeros((2, 300)), np.zeros((2, 300)), np.zeros((2, 900)),
                              np.ones((2, 900)), None, bad)
    assert rb["status"] == ["ERROR", "ERROR"]
    rp = engine.solve_partitioned(model, bad, n_gpus=2, local=True)
    assert rp["status"] == "ERROR"
    # the library is still usable afterwards
    ok = engine.solve(model, pkg.Parameters.default(use_presolve=False, stop_tol=1e-6), main=True)
    assert ok["status"] == "OPTIMAL"
    engine.free_model(model)


def test_release_cached_memory_returns_the_arena(pkg, engine):
    """Finished solves keep their arena in the engines' private pool; hprlp_b200_release_cached_memory hands it back."""
    import torch
    torch.cuda.init()
    lp = pkg.synth_lp("uniform", 100_000, 400_000, 8_000_000)     # arena of a few hundred MB
    p = pkg.Parameters.default(use_presolve=False, max_iter=20, stop_tol=1e-30)
    model = engine.create_model(lp)
    engine.solve(model, p, main=True)                             # warm: context, modules, cuRAND
    engine.release_cached_memory()
    torch.cuda.synchronize()
    before = _free_bytes()
    engine.solve(model, p, main=True)
    cached = before - _free_bytes()
    engine.release_cached_memory()
    after = _free_bytes()
    engine.free_model(model)
    assert cached > 150 << 20, f"the arena was expected to stay cached ({cached} bytes)"
    assert before - after < 32 << 20, f"cached memory not returned: {before - after} bytes still held"


def test_lookback_bitwise_reproducible_under_concurrency(pkg, engine):
    """Rows longer than many 256-nonzero items force the in-kernel look-back (publish / consume of partial sums between
    CTAs, chunk order from an atomic ticket).  Two engines solving concurrently on one GPU (two host threads, two
    streams: CTAs of both kernels interleave on the SMs) must give bit-identical results to a solo run."""
    lp = pkg.synth_lp("powerlaw", 3000, 40000, 3_000_000)         # mean row length 1000, longest rows ~ 40000 nonzeros
    p = pkg.Parameters.default(use_presolve=False, max_iter=400, stop_tol=1e-30)
    model = engine.create_model(lp)
    solo = engine.solve(model, p, main=True)
    outs = [None, None]

    def work(i):
        outs[i] = engine.solve(model, p, main=True)

    for _ in range(2):
        ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for o in outs:
            assert o["status"] == solo["status"] and o["iter"] == solo["iter"]
            for k in "xyz":
                assert np.array_equal(o[k], solo[k]), k
    engine.free_model(model)


def test_vectors_beyond_the_texture_limit_use_plain_gathers(pkg, engine, monkeypatch):
    """Gathered vectors longer than cudaDevAttrMaxTexture1DLinearWidth cannot be bound as linear textures; the engine then
    dispatches the ld.global.nc variants of the same kernels.  HPRLP_TEX_LIMIT=0 forces them: same bits."""
    lp = pkg.synth_lp("powerlaw", 4000, 12000, 200000)
    p = pkg.Parameters.default(use_presolve=False, max_iter=300, stop_tol=1e-30)
    model = engine.create_model(lp)
    a = engine.solve(model, p, main=True)
    monkeypatch.setenv("HPRLP_TEX_LIMIT", "0")
    b = engine.solve(model, p, main=True)
    engine.free_model(model)
    assert a["status"] == b["status"] and a["iter"] == b["iter"]
    for k in "xyz":
        assert np.array_equal(a[k], b[k]), k


def test_time_limit_is_honoured_between_checks(pkg, engine):
    """The reference looks at the clock every iteration (src/HPRLP.cu:184-198); here the run to the next residual check is
    capped by what the remaining time allows, so a solve stops close to time_limit even when checks are 150 iterations apart."""
    import time
    lp = pkg.synth_lp("uniform", 100_000, 400_000, 8_000_000)
    model = engine.create_model(lp)
    engine.solve(model, pkg.Parameters.default(use_presolve=False, max_iter=20, stop_tol=1e-30), main=True)   # warm
    p = pkg.Parameters.default(use_presolve=False, stop_tol=1e-30, time_limit=0.25, check_iter=100000)
    t0 = time.perf_counter()
    r = engine.solve(model, p, main=True)
    wall = time.perf_counter() - t0
    engine.free_model(model)
    assert r["status"] == "TIME_LIMIT"
    assert 0.25 <= r["time"] < 0.6, r["time"]          # not tens of thousands of iterations later
    assert wall < 3.0


def test_ticket_counter_is_rezeroed_before_it_can_wrap(pkg, engine):
    """The chunk ticket of csr_stream_kernel is a 32-bit counter that only ever grows by n_items per launch; the host
    re-zeroes it between launches long before 2^32.  HPRLP_TICKET_WRAP makes that happen every few launches (direct
    launches and the CUDA-graph path of small LPs): same bits as without.  (The variable is read once per process, so
    the forced run happens in a child process.)"""
    import json, os, subprocess, sys, textwrap
    code = textwrap.dedent('''
        import json, sys
        import numpy as np
        sys.path.insert(0, %r)
        import __graft_entry__ as g
        pkg = g.load_package(); eng = pkg.load_engine()
        out = {}
        for name, (kind, m, n, nnz, it) in dict(direct=("powerlaw", 20000, 50000, 3000000, 300), graph=("uniform", 300, 900, 3600, 700)).items():
            lp = pkg.synth_lp(kind, m, n, nnz)
            p = pkg.Parameters.default(use_presolve=False, max_iter=it, stop_tol=1e-30)
            model = eng.create_model(lp)
            r = eng.solve(model, p, main=True)
            eng.free_model(model)
            out[name] = [float(np.sum(r["x"])), float(np.sum(r["y"])), float(np.sum(np.abs(r["z"]))), r["iter"]]
        print("RESULT " + json.dumps(out))
    ''') % str(pkg.ROOT)
    def run(env_extra):
        env = dict(os.environ, **env_extra)
        pr = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert pr.returncode == 0, pr.stderr[-2000:]
        return json.loads([ln for ln in pr.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])
    assert run({"HPRLP_TICKET_WRAP": "3000"}) == run({})


def test_warmup_then_solve(pkg, engine):
    """hprlp_b200_warmup starts the context / cuRAND warm-up on a background thread; the next solve joins it."""
    engine.lib.hprlp_b200_warmup.argtypes = [C.c_int]
    engine.lib.hprlp_b200_warmup.restype = None
    engine.lib.hprlp_b200_warmup(0)
    engine.lib.hprlp_b200_warmup(0)          # idempotent
    lp = pkg.synth_lp("uniform", 300, 900, 3600)
    model = engine.create_model(lp)
    r = engine.solve(model, pkg.Parameters.default(use_presolve=False, stop_tol=1e-6), main=True)
    engine.free_model(model)
    assert r["status"] == "OPTIMAL"
