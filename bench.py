#!/usr/bin/env python
"""bench.py -- HPR iterations/s of the B200-native engine on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|small]

One "step" = ITERS_PER_STEP (100) consecutive HPR iterations of the full driver (fused x/y phase
kernels, the check iterations and residual passes the reference's schedule puts among them, restarts
and sigma updates) on one synthetic LP that is already resident in HBM.
  value        iterations/s over exactly K timed steps (CUDA events on the engine stream, max over ranks)
  e2e          iterations/s through the reference-facing C ABI: solve(model, param) with HOST arrays --
               H2D of the model, scaling, power iteration, the loop to KKT < 1e-4, D2H of x,y,z all timed
  roofline     the slower of the two fused SpMV+prox kernels, timed alone with CUDA events
  cpu_baseline the CPU oracle (OpenMP port) on a bounded sample of the same workload
--impl reference runs the UNMODIFIED reference (its own CUDA build, oracle/_ref/libhprlp_ref.so -- the
reference has no CPU path, BASELINE.json) through the same solve() call on the same LP.
N > 1: the single-instance path does not shard, so every rank runs an independent replica of the same
workload ("replicas only") and `value` is the sum; no collective on the data path.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

ITERS_PER_STEP = 100
PREROLL = 1000

WORKLOADS = {
    # BASELINE.json configs[1]: synthetic uniform-density LP m=1e5 n=1e6 nnz=1e7
    "c2": dict(kind="uniform", m=100_000, n=1_000_000, nnz=10_000_000,
               name="configs[1]: synthetic uniform LP m=1e5 n=1e6 nnz=1e7, tol 1e-4"),
    # BASELINE.json configs[2]: synthetic power-law LP m=2e6 n=5e6 nnz=1e8
    "c3": dict(kind="powerlaw", m=2_000_000, n=5_000_000, nnz=100_000_000,
               name="configs[2]: synthetic power-law LP m=2e6 n=5e6 nnz=1e8, tol 1e-4"),
    "small": dict(kind="uniform", m=5_000, n=20_000, nnz=200_000, name="debug: uniform m=5e3 n=2e4 nnz=2e5"),
    # BASELINE.json configs[3]: solve_batched shared-A m=5e4 n=2e5 nnz=2e6, batch 256, batch-sharded over the GPUs
    "c4": dict(kind="uniform", m=50_000, n=200_000, nnz=2_000_000, batch=256, iters=300,
               name="configs[3]: solve_batched shared-A m=5e4 n=2e5 nnz=2e6, batch 256, batch-sharded"),
    "c4small": dict(kind="uniform", m=2_000, n=8_000, nnz=80_000, batch=64, iters=300, name="debug: batched m=2e3 n=8e3 B=64"),
}


def algorithmic_bytes(m, n, nnz):
    """per normal iteration 24 nnz + 68 n + 52 m; x-phase 12 nnz + 60 n + 8 m; y-phase 12 nnz + 44 m + 8 n."""
    return dict(iter=24 * nnz + 68 * n + 52 * m, x=12 * nnz + 60 * n + 8 * m, y=12 * nnz + 44 * m + 8 * n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def reduce_max_sum(dist, local, ms, units):
    if dist is None:
        return ms, units
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    u = torch.tensor([units], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


def cpu_baseline(pkg, lp, iters=10):
    import ctypes as C
    O = pkg.load_oracle()
    f = O.lib.oracle_time_iterations
    f.restype = C.c_double
    thr = C.c_int(0)
    ip, dp = pkg._ip, pkg._dp
    secs = f(lp["m"], lp["n"], ip(lp["rowPtr"]), ip(lp["colIndex"]), dp(lp["values"]), dp(lp["AL"]), dp(lp["AU"]), dp(lp["l"]),
             dp(lp["u"]), dp(lp["c"]), iters, C.byref(thr))
    return dict(value=iters / secs, unit="HPR iterations/s", cores=int(thr.value), kind="port",
                sample=f"{iters} HPR iterations (fused x+y phase, unscaled data) of the same LP by oracle/hpr_oracle.c, OpenMP over rows")


def run_engine_arm(args, pkg, spec, lp, rank, world, local, dist):
    eng = pkg.load_engine()          # no fallback: raises if lib/libhprlp.so is missing
    m, n, nnz = lp["m"], lp["n"], int(lp["values"].shape[0])
    param = pkg.Parameters.default(stop_tol=0.0, use_presolve=False, device_number=local)   # tol 0: the loop never stops
    model = eng.create_model(lp)
    h = eng.lib.hprlp_b200_engine_create(model, __import__("ctypes").byref(param))
    if not h:
        raise RuntimeError("engine_create failed")
    info = pkg.B200Info()
    # pre-roll (untimed): the first PREROLL iterations carry a residual check every 10 iterations, afterwards every
    # 100 (reference schedule, src/utils.cu:100-102).  The timed window starts in the steady regime, the same
    # window the reference arm's `value` is taken from (iterations 1000..3000).
    eng.lib.hprlp_b200_engine_run(h, PREROLL)
    for _ in range(max(args.warmup, 3)):
        eng.lib.hprlp_b200_engine_run(h, ITERS_PER_STEP)
    eng.lib.hprlp_b200_engine_info(h, __import__("ctypes").byref(info))
    launches0 = info.kernel_launches
    barrier(dist, local)
    eng.lib.hprlp_b200_profiler_start()      # no-op unless run under `ncu --profile-from-start off`
    with ClockSampler(local) as clk:
        t_wall = time.perf_counter()
        ms = 0.0
        for _ in range(args.steps):
            ms += eng.lib.hprlp_b200_engine_run(h, ITERS_PER_STEP)     # CUDA events on the engine stream
        wall_ms = (time.perf_counter() - t_wall) * 1e3
    eng.lib.hprlp_b200_profiler_stop()
    barrier(dist, local)
    eng.lib.hprlp_b200_engine_info(h, __import__("ctypes").byref(info))
    launches = int(info.kernel_launches - launches0)
    iters = args.steps * ITERS_PER_STEP
    ms_max, iters_sum = reduce_max_sum(dist, local, max(ms, 0.0), iters)
    # roofline: each fused kernel alone, back to back, CUDA events (inputs > L2 for c2/c3: no flush needed)
    tx = eng.lib.hprlp_b200_engine_time_phase(h, 0, 50)
    ty = eng.lib.hprlp_b200_engine_time_phase(h, 1, 50)
    eng.lib.hprlp_b200_engine_destroy(h)
    eng.free_model(model)
    if rank != 0:
        return None
    ab = algorithmic_bytes(m, n, nnz)
    peaks = {}
    pk_file = ROOT / "MEASURED_PEAKS.json"
    peak_src = "fallback"
    peak = 6650.0
    if pk_file.exists():
        peaks = json.loads(pk_file.read_text())
        if "hbm_gbs" in peaks:
            peak, peak_src = float(peaks["hbm_gbs"]), "measured"
    dom = "y" if ty >= tx else "x"
    dur_ms = ty if dom == "y" else tx
    achieved = ab[dom] / (dur_ms * 1e-3) / 1e9
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    # second, tighter bound (profiles/r1_v4_experiments.md): on a uniformly random matrix every nonzero is one L1->crossbar
    # request (a 32-byte sector) and an SM issues one request per clock; measured ceiling of the 12 B/nnz stream + gather +
    # FMA on this launch shape (tools/gather_bench.cu, SPMV_TEX) = 262e9 nnz/s with the gathered vector resident in L2.
    port = None
    gc = ROOT / "profiles" / "r1_gather_ceiling.json"
    if gc.exists():
        try:
            rows = [r for r in json.loads(gc.read_text())["results"] if r["mode"] == "SPMV_TEX"]
            vec = m if dom == "x" else n     # x-phase gathers y (m entries), y-phase gathers x_hat (n entries)
            best = min(rows, key=lambda r: abs(r["V"] - vec))
            port = dict(bound="l1-to-crossbar request port (1 request/clk/SM)", achieved_gnnz_per_s=nnz / (dur_ms * 1e-3) / 1e9,
                        ceiling_gnnz_per_s=best["gnnz_per_s"], frac=nnz / (dur_ms * 1e-3) / 1e9 / best["gnnz_per_s"],
                        source="profiles/r1_gather_ceiling.json SPMV_TEX, gathered vector of %d doubles" % best["V"])
        except Exception:
            port = None
    roof = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, frac_of_nominal_8000_gbs=achieved / 8000.0,
                traffic=traffic, request_port=port,
                peak_source=peak_src, kernel=f"csr_stream_kernel<{'YPhaseOp' if dom == 'y' else 'XPhaseOp'}<false>> ({dom}-phase)",
                algorithmic_bytes_per_launch=ab[dom], launch_ms=dur_ms,
                x_phase=dict(ms=tx, gbs=ab["x"] / (tx * 1e-3) / 1e9), y_phase=dict(ms=ty, gbs=ab["y"] / (ty * 1e-3) / 1e9),
                iteration_gbs=ab["iter"] / ((tx + ty) * 1e-3) / 1e9)
    out = dict(metric="HPR iterations/s (time-to-1e-4 KKT in e2e); fused SpMV+prox HBM GB/s in roofline",
               value=iters_sum / (ms_max * 1e-3), unit="HPR iterations/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
               ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
               data="synthetic",
               config=dict(workload=spec["name"], m=m, n=n, nnz=nnz, iters_per_step=ITERS_PER_STEP,
                           window="timed steps start at HPR iteration %d (steady regime: residual check every 100 iterations)" % (PREROLL + ITERS_PER_STEP * max(args.warmup, 3)),
                           l2="per-iteration working set %.0f MB > 126 MB L2 (no flush)" % (ab["iter"] / 1e6),
                           parallelism="replicas only" if world > 1 else "1 GPU",
                           lanes_A=info.lanes_A, lanes_AT=info.lanes_AT),
               gpu_launches=launches, wall_ms_per_step=wall_ms / args.steps, clocks=clk.summary(), roofline=roof)
    return out


def make_batch(pkg, spec, lo, hi):
    """Shared matrix + instances lo..hi-1 (instance k draws its primal-dual pair with seed+k)."""
    base = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
    vs = [pkg.synth_vectors(base, pkg.SEED, pkg.SEED + k) for k in range(lo, hi)]
    st = lambda key: np.stack([v[key] for v in vs])
    return base, dict(C=st("c"), AL=st("AL"), AU=st("AU"), l=st("l"), u=st("u"))


def run_batched(args, pkg, spec, lib, rank, world, local, dist, impl):
    """One step = one solve_batched call of `iters` iterations on this rank's contiguous instance shard
    (A replicated, no data-path collective). value = instance-iterations/s over all ranks, max-over-ranks time."""
    B = spec["batch"]
    lo, hi = pkg.shard_range(B, world, rank)
    base, d = make_batch(pkg, spec, lo, hi)
    param = pkg.Parameters.default(stop_tol=0.0, max_iter=spec["iters"], use_presolve=False, device_number=local)
    model = lib.create_model(base)
    steps = max(1, min(args.steps, 3))
    warm = 1
    times, walls = [], []
    barrier(dist, local)
    with ClockSampler(local) as clk:
        for i in range(warm + steps):
            t0 = time.perf_counter()
            r = lib.solve_batched(model, d["C"], d["AL"], d["AU"], d["l"], d["u"], None, param)
            w = time.perf_counter() - t0
            if i >= warm:
                times.append(r["solve_time"]); walls.append(w)
    barrier(dist, local)
    lib.free_model(model)
    n_inst_iters = (hi - lo) * spec["iters"]
    t_solve, units = reduce_max_sum(dist, local, 1e3 * min(times), n_inst_iters)
    t_wall, _ = reduce_max_sum(dist, local, 1e3 * min(walls), 0)
    if rank != 0:
        return None
    m, n, nnz = base["m"], base["n"], int(base["values"].shape[0])
    bytes_iter = 24 * nnz + 4 * (n + m) + B * (64 * n + 48 * m)    # SURVEY.md 8(d), whole batch, A streamed once
    out = dict(metric="HPR instance-iterations/s, solve_batched shared-A", value=units / (t_solve * 1e-3), unit="instance-iterations/s",
               n_gpus=world, steps=steps, warmup=warm, ms_per_step=t_solve, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", impl=impl,
               config=dict(workload=spec["name"], m=m, n=n, nnz=nnz, batch=B, iters_per_step=spec["iters"],
                           parallelism="batch-sharded dp%d, A replicated, no collective" % world,
                           l2="batched vectors %.0f MB >> 126 MB L2" % (B * (64 * n + 48 * m) / 1e6)),
               e2e=dict(value=units / (t_wall * 1e-3), unit="instance-iterations/s",
                        h2d_bytes_per_step=8 * (hi - lo) * (3 * n + 2 * m) + 24 * nnz, d2h_bytes_per_step=8 * (hi - lo) * (2 * n + m)),
               roofline=dict(bound="hbm", achieved=bytes_iter * spec["iters"] / (t_solve * 1e-3) / 1e9 / world, peak=6549.1, unit="GB/s",
                             frac=bytes_iter * spec["iters"] / (t_solve * 1e-3) / 1e9 / world / 6549.1, traffic=None,
                             note="whole loop (fused SpMM+prox x/y kernels + checks); per GPU; algorithmic bytes 24nnz+4(n+m)+B(64n+48m)"),
               clocks=clk.summary(), gpu_launches=None)
    return out


def steady_state_rate(pkg, lib, lp, local, k1=1000, k2=None):
    """Loop iterations/s of a library that can only be driven through solve(): difference of the library's own
    results.time between max_iter=k2 and max_iter=k1 (setup, power iteration and the reference's autotune cancel).
    The span k2-k1 is sized so the difference is seconds, not tenths (the reference's power iteration + autotune vary by
    ~0.1 s between runs: a 2000-iteration span on the nnz=1e7 LP once read 10.5k it/s against 5.1k in another run)."""
    if k2 is None:
        k2 = k1 + (2000 if int(lp["values"].shape[0]) >= 50_000_000 else 20000)
    ts = []
    for k in (k1, k2):
        param = pkg.Parameters.default(stop_tol=0.0, max_iter=k, use_presolve=False, device_number=local)
        model = lib.create_model(lp)
        r = lib.solve(model, param)
        lib.free_model(model)
        ts.append(r["time"])
    return (k2 - k1) / max(ts[1] - ts[0], 1e-9), ts


def run_e2e(pkg, lib, lp, local, tol=1e-4):
    """solve() through the C ABI with host arrays: H2D + setup + scaling + power iteration + loop + D2H timed."""
    import contextlib, io
    param = pkg.Parameters.default(stop_tol=tol, use_presolve=False, device_number=local)
    model = lib.create_model(lp)
    t0 = time.perf_counter()
    r = lib.solve(model, param)
    wall = time.perf_counter() - t0
    lib.free_model(model)
    nnz = int(lp["values"].shape[0])
    h2d = 12 * nnz + 4 * (lp["m"] + 1) + 8 * (2 * lp["m"] + 3 * lp["n"])   # A only: A^T is built on the device
    d2h = 8 * (2 * lp["n"] + lp["m"])
    return dict(value=r["iter"] / wall, unit="HPR iterations/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                time_to_tol_s=wall, solver_time_s=r["time"], iters=r["iter"], status=r["status"], primal_obj=r["primal_obj"],
                residuals=r["residuals"], tol=tol)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))   # configs[2]: the LP BASELINE.json's target is quoted on
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    pkg = graft.load_package()
    rank, world, local, dist = dist_setup(args.gpus)
    spec = WORKLOADS[args.workload]

    if "batch" in spec:
        lib = pkg.load_engine() if args.impl == "ours" else (pkg.load_reference() if pkg.REF_LIB_PATH.exists() else None)
        if args.impl == "reference" and (rank != 0 or lib is None):
            if rank == 0:
                print(json.dumps(dict(impl="reference", unavailable="oracle/_ref/libhprlp_ref.so was not built")))
            return 0
        if args.impl == "reference":
            world, dist = 1, None          # the reference is single-GPU: rank 0 runs the whole batch
        sys.stdout.flush()
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
        try:
            out = run_batched(args, pkg, spec, lib, rank, world, local, dist, args.impl)
        finally:
            os.dup2(saved, 1); os.close(devnull); os.close(saved)
        if rank == 0:
            print(json.dumps(out))
        if dist is not None:
            dist.destroy_process_group()
        return 0

    if args.impl == "reference":
        if rank != 0:
            return 0
        lp = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
        if not pkg.REF_LIB_PATH.exists():
            print(json.dumps(dict(impl="reference", unavailable="oracle/_ref/libhprlp_ref.so was not built (needs /root/reference at build time)")))
            return 0
        ref = pkg.load_reference()
        sys.stdout.flush()
        # the reference prints its log with std::cout: keep stdout clean for the JSON line
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
        try:
            with ClockSampler(local) as clk:
                runs = [run_e2e(pkg, ref, lp, local) for _ in range(max(1, min(args.steps, 3)))]
                rate, rate_ts = steady_state_rate(pkg, ref, lp, local)
        finally:
            os.dup2(saved, 1); os.close(devnull); os.close(saved)
        best = max(runs, key=lambda r: r["value"])
        out = dict(impl="reference", metric="HPR iterations/s (time-to-1e-4 KKT in e2e); fused SpMV+prox HBM GB/s in roofline",
                   value=rate, unit="HPR iterations/s", n_gpus=1, steps=len(runs), warmup=0,
                   ms_per_step=1e3 * ITERS_PER_STEP / rate, steady_state=dict(
                       how="(k2-k1) iterations / (results.time[max_iter=k2] - results.time[max_iter=k1]), k1=1000, k2-k1=2000 (nnz>=5e7) or 20000", times=rate_ts), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                   data="synthetic", config=dict(workload=spec["name"], m=lp["m"], n=lp["n"], nnz=int(lp["values"].shape[0]),
                                                 note="value = steady-state loop rate from two max_iter runs; e2e = one solve() call to KKT<1e-4 through the reference's own C API"),
                   cpu_baseline=dict(value=rate, unit="HPR iterations/s", cores=1, kind="reference",
                                     sample="the reference has no CPU path (BASELINE.json): this arm is its own CUDA build "
                                            "(oracle/_ref/libhprlp_ref.so, autotuned fused/cuSPARSE backend) on the same B200, driven by one "
                                            "host thread; value = the line's steady-state loop rate, e2e = whole solve() calls to KKT<1e-4"),
                   e2e=dict(best, h2d_bytes_per_step=0, d2h_bytes_per_step=0), clocks=clk.summary(), all_runs=runs)
        print(json.dumps(out))
        return 0

    lp = pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])
    # keep the engine's own log off stdout (single JSON line contract)
    sys.stdout.flush()
    devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
    try:
        out = run_engine_arm(args, pkg, spec, lp, rank, world, local, dist)
        e2e = None
        if not args.no_e2e:
            # same protocol as the reference arm: up to 3 whole solve() calls, the best one is reported (the first call of
            # a process also loads cuRAND's kernels for the power-iteration start vector, ~0.6 s), all are listed
            runs = [run_e2e(pkg, pkg.load_engine(), lp, local) for _ in range(max(1, min(args.steps, 3)))]
            e = max(runs, key=lambda r: r["value"])
            v, _ = reduce_max_sum(dist, local, 0.0, e["value"])
            e2e = dict(e, value=_ if world > 1 else e["value"], all_runs_time_to_tol_s=[r["time_to_tol_s"] for r in runs])
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu:
            cpu = cpu_baseline(pkg, lp, iters=10 if spec["nnz"] >= 5_000_000 else 200)
    finally:
        os.dup2(saved, 1); os.close(devnull); os.close(saved)
    if rank == 0:
        if e2e is not None:
            out["e2e"] = e2e
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
