#!/usr/bin/env python
"""bench.py -- HPR iterations/s of the B200-native engine on BASELINE.json's configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c4|small|c3band|c3block]

One "step" = ITERS_PER_STEP (100) consecutive HPR iterations of the full driver (fused x/y phase kernels, the check
iterations and residual passes the reference's schedule puts among them, restarts and sigma updates) on one synthetic
LP that is already resident in HBM.
  value        HPR iterations/s over exactly K timed steps (CUDA events on the engine stream, max over ranks)
  e2e          the same metric through the reference-facing C ABI: solve(model, param) with HOST arrays -- H2D of the
               model, device transposition, scaling, power iteration, the loop to KKT < 1e-4, D2H of x,y,z all timed
  roofline     the slower of the two fused SpMV+prox kernels, timed alone with CUDA events
  cpu_baseline the CPU oracle (OpenMP port) on a bounded sample of the same workload (N = 1 only)
  batched      configs[3] (solve_batched shared-A, B = 256), batch-sharded over the N GPUs: instance-iterations/s
  parity       (N > 1) asserted in the run: a small LP solved row-partitioned over the N GPUs vs on one GPU, and a
               batch solved sharded vs unsharded
--impl reference runs the UNMODIFIED reference (its own CUDA build, oracle/_ref/libhprlp_ref.so -- the reference has
no CPU path, BASELINE.json) through the same solve() call on the same LP; it is single-GPU, so rank 0 alone runs it.

N > 1 (torchrun, one process per GPU): the SAME LP (configs[2], nnz = 1e8) is solved ROW-PARTITIONED over the N GPUs --
GPU p owns a block of rows of A and an x-block; per iteration one reduce-scatter of the partial A^T y and one all-gather
of the x_hat blocks (NCCL over NVLink).  `value` is the iteration rate of that one solve ("scaling": "strong": total
work fixed), `partitioned.speedup_vs_1gpu` is measured against the 1-GPU engine in the same run, `batched` carries
configs[3] sharded over the N GPUs, and at N = 8 `c5` carries configs[4] (nnz = 6e9, does not fit one GPU).
"""
from __future__ import annotations

import argparse
import ctypes
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
    # BASELINE.json configs[0]: the bundled toy LP through the CLI with default parameters (presolve on)
    "c1": dict(kind="mps", name="configs[0]: bundled data/model.mps (m=n=2), CLI defaults incl. presolve, tol 1e-4"),
    # BASELINE.json configs[1]: synthetic uniform-density LP m=1e5 n=1e6 nnz=1e7
    "c2": dict(kind="uniform", m=100_000, n=1_000_000, nnz=10_000_000,
               name="configs[1]: synthetic uniform LP m=1e5 n=1e6 nnz=1e7, tol 1e-4"),
    # BASELINE.json configs[2]: synthetic power-law LP m=2e6 n=5e6 nnz=1e8
    "c3": dict(kind="powerlaw", m=2_000_000, n=5_000_000, nnz=100_000_000,
               name="configs[2]: synthetic power-law LP m=2e6 n=5e6 nnz=1e8, tol 1e-4"),
    # the same kernels on a STRUCTURED matrix of configs[2]'s size: columns of row i lie in a window around i*n/m, so
    # gathers coalesce into few sectors (shows what the kernels reach when the input has locality)
    "c3band": dict(kind="banded", m=2_000_000, n=5_000_000, nnz=100_000_000,
                   name="structured twin of configs[2]: banded LP m=2e6 n=5e6 nnz=1e8 (columns within a 4096-wide window)"),
    # ... and on a BLOCK-structured one (dense 8x8 blocks): a warp's gathers coalesce into whole 64-byte segments in both passes
    "c3block": dict(kind="blocked", m=2_000_000, n=5_000_000, nnz=100_000_000,
                    name="block-structured twin of configs[2]: m=2e6 n=5e6 nnz=1e8, dense 8x8 blocks within a 4096-wide window"),
    "small": dict(kind="uniform", m=5_000, n=20_000, nnz=200_000, name="debug: uniform m=5e3 n=2e4 nnz=2e5"),
    # BASELINE.json configs[3]: solve_batched shared-A m=5e4 n=2e5 nnz=2e6, batch 256, batch-sharded over the GPUs
    "c4": dict(kind="uniform", m=50_000, n=200_000, nnz=2_000_000, batch=256, iters=300,
               name="configs[3]: solve_batched shared-A m=5e4 n=2e5 nnz=2e6, batch 256, batch-sharded"),
    "c4small": dict(kind="uniform", m=2_000, n=8_000, nnz=80_000, batch=64, iters=300, name="debug: batched m=2e3 n=8e3 B=64"),
}
C5 = dict(m=20_000_000, n=50_000_000, K=300, name="configs[4]: synthetic uniform LP nnz=6e9 (m=2e7 n=5e7, 300 per row), row-block partitioned")
PARITY_LP = dict(kind="powerlaw", m=20_000, n=50_000, nnz=1_000_000)


def algorithmic_bytes(m, n, nnz):
    """per normal iteration 24 nnz + 68 n + 52 m; x-phase 12 nnz + 60 n + 8 m; y-phase 12 nnz + 44 m + 8 n."""
    return dict(iter=24 * nnz + 68 * n + 52 * m, x=12 * nnz + 60 * n + 8 * m, y=12 * nnz + 44 * m + 8 * n)


def measured_peak():
    pk_file = ROOT / "MEASURED_PEAKS.json"
    if pk_file.exists():
        try:
            peaks = json.loads(pk_file.read_text())
            if "hbm_gbs" in peaks:
                return float(peaks["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


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


class Quiet:
    """The libraries print their logs to stdout (std::cout / printf): keep stdout clean for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.devnull = os.open(os.devnull, os.O_WRONLY)
        self.saved = os.dup(1)
        os.dup2(self.devnull, 1)
        return self

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.devnull)
        os.close(self.saved)


def dist_setup(n_gpus):
    with Quiet():     # NCCL announces its version on stdout when the first communicator comes up
        return _dist_setup(n_gpus)


def _dist_setup(n_gpus):
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
        dist.barrier(device_ids=[local])
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


def fresh_uid(eng, dist, rank, local):
    """A new NCCL unique id made by rank 0 and broadcast with torch.distributed: one per communicator of the engine."""
    return graft.load_package().broadcast_unique_id(dist, eng.nccl_unique_id, rank)


def cpu_baseline(pkg, lp, iters=10):
    O = pkg.load_oracle()
    f = O.lib.oracle_time_iterations
    f.restype = ctypes.c_double
    thr = ctypes.c_int(0)
    ip, dp = pkg._ip, pkg._dp
    secs = f(lp["m"], lp["n"], ip(lp["rowPtr"]), ip(lp["colIndex"]), dp(lp["values"]), dp(lp["AL"]), dp(lp["AU"]), dp(lp["l"]),
             dp(lp["u"]), dp(lp["c"]), iters, ctypes.byref(thr))
    return dict(value=iters / secs, unit="HPR iterations/s", cores=int(thr.value), kind="port",
                sample=f"{iters} HPR iterations (fused x+y phase, unscaled data) of the same LP by oracle/hpr_oracle.c, OpenMP over rows")


def make_lp(pkg, spec):
    return pkg.synth_lp(spec["kind"], spec["m"], spec["n"], spec["nnz"])


# ----------------------------------------------------------------------------------------------------------------
# resident engine, CUDA-event timed: the `value` of both the 1-GPU and the row-partitioned arm
# ----------------------------------------------------------------------------------------------------------------
def time_resident_engine(eng, pkg, h, steps, warmup, dist, local, profile=False):
    info = pkg.B200Info()
    eng.lib.hprlp_b200_engine_run(h, PREROLL)
    for _ in range(warmup):
        eng.lib.hprlp_b200_engine_run(h, ITERS_PER_STEP)
    eng.lib.hprlp_b200_engine_info(h, ctypes.byref(info))
    launches0 = info.kernel_launches
    barrier(dist, local)
    if profile:
        eng.lib.hprlp_b200_profiler_start()      # no-op unless run under `ncu --profile-from-start off`
    with ClockSampler(local) as clk:
        t_wall = time.perf_counter()
        ms = 0.0
        for _ in range(steps):
            ms += eng.lib.hprlp_b200_engine_run(h, ITERS_PER_STEP)     # CUDA events on the engine stream
        wall_ms = (time.perf_counter() - t_wall) * 1e3
    if profile:
        eng.lib.hprlp_b200_profiler_stop()
    barrier(dist, local)
    eng.lib.hprlp_b200_engine_info(h, ctypes.byref(info))
    return dict(ms=max(ms, 0.0), wall_ms=wall_ms, launches=int(info.kernel_launches - launches0), clocks=clk.summary(), info=info)


def run_engine_arm(args, pkg, spec, lp, local):
    """N = 1: the whole LP on one GPU."""
    eng = pkg.load_engine()          # no fallback: raises if lib/libhprlp.so is missing
    m, n, nnz = lp["m"], lp["n"], int(lp["values"].shape[0])
    param = pkg.Parameters.default(stop_tol=0.0, use_presolve=False, device_number=local)   # tol 0: the loop never stops
    model = eng.create_model(lp)
    h = eng.lib.hprlp_b200_engine_create(model, ctypes.byref(param))
    if not h:
        raise RuntimeError("engine_create failed")
    warm = max(args.warmup, 3)
    # pre-roll (untimed): the first PREROLL iterations carry a residual check every 10 iterations, afterwards every
    # 100 (reference schedule, src/utils.cu:100-102).  The timed window starts in the steady regime, the same
    # window the reference arm's `value` is taken from (iterations 1000..3000).
    t = time_resident_engine(eng, pkg, h, args.steps, warm, None, local, profile=True)
    iters = args.steps * ITERS_PER_STEP
    # roofline: each fused kernel alone, back to back, CUDA events (inputs > L2 for c2/c3: no flush needed)
    tx = eng.lib.hprlp_b200_engine_time_phase(h, 0, 50)
    ty = eng.lib.hprlp_b200_engine_time_phase(h, 1, 50)
    eng.lib.hprlp_b200_engine_destroy(h)
    eng.free_model(model)
    ab = algorithmic_bytes(m, n, nnz)
    peak, peak_src = measured_peak()
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
    if gc.exists() and spec["kind"] not in ("banded", "blocked"):
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
    info = t["info"]
    return dict(metric="HPR iterations/s (time-to-1e-4 KKT in e2e); fused SpMV+prox HBM GB/s in roofline",
                value=iters / (t["ms"] * 1e-3), unit="HPR iterations/s", n_gpus=1, steps=args.steps, warmup=warm,
                ms_per_step=t["ms"] / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic",
                config=dict(workload=spec["name"], m=m, n=n, nnz=nnz, iters_per_step=ITERS_PER_STEP,
                            window="timed steps start at HPR iteration %d (steady regime: residual check every 100 iterations)" % (PREROLL + ITERS_PER_STEP * warm),
                            l2="per-iteration working set %.0f MB > 126 MB L2 (no flush)" % (ab["iter"] / 1e6),
                            parallelism="1 GPU", lanes_A=info.lanes_A, lanes_AT=info.lanes_AT),
                gpu_launches=t["launches"], wall_ms_per_step=t["wall_ms"] / args.steps, clocks=t["clocks"], roofline=roof)


def run_partitioned_arm(args, pkg, spec, lp, rank, world, local, dist, comm):
    """N > 1: ONE LP row-partitioned over the N GPUs, one process per GPU (this process = row block `rank`)."""
    eng = pkg.load_engine()
    m, n, nnz = lp["m"], lp["n"], int(lp["values"].shape[0])
    param = pkg.Parameters.default(stop_tol=0.0, use_presolve=False, device_number=local)
    model = eng.create_model(lp)
    h = eng.lib.hprlp_b200_engine_create_rank(model, ctypes.byref(param), comm)
    if not h:
        raise RuntimeError("engine_create_rank failed")
    warm = max(args.warmup, 3)
    t = time_resident_engine(eng, pkg, h, args.steps, warm, dist, local)
    iters = args.steps * ITERS_PER_STEP
    ms_max, _ = reduce_max_sum(dist, local, t["ms"], 0)
    # per-GPU pieces of one iteration, each timed alone (CUDA events, max over ranks)
    parts = {}
    for name, which in (("partial_ATy_pass", 2), ("exchange_reduce_scatter_all_gather", 4), ("x_update_block", 3), ("y_phase", 1)):
        barrier(dist, local)
        v = eng.lib.hprlp_b200_engine_time_phase(h, which, 30)
        parts[name + "_ms"], _ = reduce_max_sum(dist, local, v, 0)
    eng.lib.hprlp_b200_engine_destroy(h)
    eng.free_model(model)
    # the 1-GPU engine on the same LP in the same run (rank 0; the others wait): the denominator of speedup_vs_1gpu
    one_gpu = None
    if rank == 0 and not args.no_single_ref:
        h1 = eng.lib.hprlp_b200_engine_create(model_or_new(eng, lp), ctypes.byref(param))
        t1 = time_resident_engine(eng, pkg, h1, max(3, args.steps // 2), 3, None, local)
        one_gpu = (max(3, args.steps // 2) * ITERS_PER_STEP) / (t1["ms"] * 1e-3)
        eng.lib.hprlp_b200_engine_destroy(h1)
        eng.release_cached_memory()
    barrier(dist, local)
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    ms_iter = ms_max / iters
    # algorithmic bytes per GPU per iteration: its share of both matrix passes and of the y-side vectors, the x-side
    # vectors of ITS x-block (68 n / P: counted once over the node), plus what the row partition itself requires on
    # every GPU: the full x_hat gathered by the y-phase (8 n) and the partial w = A_p^T y_p (row pointers of the n-row
    # transpose 4 n, write 8 n).
    per_gpu = (24 * nnz + 52 * m + 68 * n) / world + 20 * n
    roof = dict(bound="hbm", achieved=per_gpu / (ms_iter * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=per_gpu / (ms_iter * 1e-3) / 1e9 / peak,
                traffic=None, peak_source=peak_src, kernel="whole partitioned iteration per GPU (A_p^T pass + exchange + x-update + fused y-phase)",
                algorithmic_bytes_per_gpu_per_iteration=per_gpu, nvlink_bytes_per_gpu_per_iteration=2 * 8 * n * (world - 1) / world,
                **parts)
    info = t["info"]
    out = dict(metric="HPR iterations/s (time-to-1e-4 KKT in e2e); fused SpMV+prox HBM GB/s in roofline",
               value=iters / (ms_max * 1e-3), unit="HPR iterations/s", n_gpus=world, steps=args.steps, warmup=warm,
               ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=spec["name"], m=m, n=n, nnz=nnz, iters_per_step=ITERS_PER_STEP,
                           window="timed steps start at HPR iteration %d (steady regime)" % (PREROLL + ITERS_PER_STEP * warm),
                           l2="per-GPU per-iteration working set %.0f MB > 126 MB L2 (no flush)" % (per_gpu / 1e6),
                           parallelism="row-block partitioned over %d GPUs (one process per GPU): reduce-scatter of partial A^T y + all-gather of x_hat "
                                       "blocks per iteration, NCCL over NVLink; x-block ownership" % world,
                           exchange="fused reduce-scatter + x-update + all-gather kernel over NVLink peer memory (P2P loads/stores)"
                                    if info.peer_exchange else "ncclReduceScatter + x-update kernel + ncclAllGather",
                           lanes_A=info.lanes_A, lanes_AT=info.lanes_AT),
               gpu_launches=t["launches"], wall_ms_per_step=t["wall_ms"] / args.steps, clocks=t["clocks"], roofline=roof,
               partitioned=dict(peer_exchange=bool(info.peer_exchange), ms_per_iteration=ms_iter, iters_per_s=1e3 / ms_iter, one_gpu_iters_per_s=one_gpu,
                                speedup_vs_1gpu=(1e3 / ms_iter) / one_gpu if one_gpu else None, **parts))
    return out


def model_or_new(eng, lp):
    model_or_new.keep = eng.create_model(lp)     # kept alive until the process ends (bench only)
    return model_or_new.keep


def run_e2e(pkg, lib, lp, local, tol=1e-4):
    """solve() through the C ABI with host arrays: H2D + setup + scaling + power iteration + loop + D2H timed."""
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


def run_e2e_partitioned(pkg, eng, lp, rank, world, local, dist, comm, tol=1e-4):
    """The partitioned solve through the C ABI with HOST arrays on every rank (each uploads its row block).  The
    communicator handle exists already (like torch.distributed's process group, it is created once per job)."""
    param = pkg.Parameters.default(stop_tol=tol, use_presolve=False, device_number=local)
    model = eng.create_model(lp)
    barrier(dist, local)
    t0 = time.perf_counter()
    r = eng.solve_partitioned_rank(model, param, comm)
    wall = time.perf_counter() - t0
    eng.free_model(model)
    wall_max, _ = reduce_max_sum(dist, local, wall, 0)
    nnz = int(lp["values"].shape[0])
    return dict(value=r["iter"] / wall_max, unit="HPR iterations/s",
                h2d_bytes_per_step=(12 * nnz + 16 * lp["m"]) // world + 24 * lp["n"] + 4 * (lp["m"] // world + 1),
                d2h_bytes_per_step=8 * (2 * lp["n"] + lp["m"]),
                time_to_tol_s=wall_max, solver_time_s=r["time"], iters=r["iter"], status=r["status"], primal_obj=r["primal_obj"],
                residuals=r["residuals"], tol=tol, power_iters=r["info"]["power_iters"], power_seconds=r["info"]["power_seconds"])


# ----------------------------------------------------------------------------------------------------------------
# solve_batched (configs[3])
# ----------------------------------------------------------------------------------------------------------------
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
            r = lib.solve_batched(model, d["C"], d["AL"], d["AU"], d["l"], d["u"], None, param)
            if i >= warm:      # e2e = the C-ABI call itself (host arrays in, host arrays out), not the test harness's numpy copies
                times.append(r["solve_time"]); walls.append(r["call_seconds"])
    barrier(dist, local)
    lib.free_model(model)
    n_inst_iters = (hi - lo) * spec["iters"]
    t_solve, units = reduce_max_sum(dist, local, 1e3 * min(times), n_inst_iters)
    t_wall, _ = reduce_max_sum(dist, local, 1e3 * min(walls), 0)
    if rank != 0:
        return None
    m, n, nnz = base["m"], base["n"], int(base["values"].shape[0])
    peak, peak_src = measured_peak()
    bytes_iter = 24 * nnz + 4 * (n + m) + B * (64 * n + 48 * m)    # SURVEY.md 8(d), whole batch, A streamed once
    per_gpu_gbs = bytes_iter * spec["iters"] / (t_solve * 1e-3) / 1e9 / world
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("c4", {}).get("iteration")
        except Exception:
            traffic = None
    out = dict(metric="HPR instance-iterations/s, solve_batched shared-A", value=units / (t_solve * 1e-3), unit="instance-iterations/s",
               n_gpus=world, steps=steps, warmup=warm, ms_per_step=t_solve, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", impl=impl,
               config=dict(workload=spec["name"], m=m, n=n, nnz=nnz, batch=B, iters_per_step=spec["iters"],
                           parallelism="batch-sharded dp%d, A replicated, no collective" % world,
                           l2="batched vectors %.0f MB >> 126 MB L2" % (B * (64 * n + 48 * m) / 1e6)),
               e2e=dict(value=units / (t_wall * 1e-3), unit="instance-iterations/s",
                        h2d_bytes_per_step=8 * (hi - lo) * (3 * n + 2 * m) + 24 * nnz, d2h_bytes_per_step=8 * (hi - lo) * (2 * n + m)),
               roofline=dict(bound="hbm", achieved=per_gpu_gbs, peak=peak, unit="GB/s", frac=per_gpu_gbs / peak, traffic=traffic,
                             peak_source=peak_src,
                             note="whole loop (fused SpMM+prox x/y kernels + checks); per GPU; algorithmic bytes 24nnz+4(n+m)+B(64n+48m)"),
               clocks=clk.summary(), gpu_launches=None)
    return out


def batched_summary(b):
    if b is None:
        return None
    return dict(workload=b["config"]["workload"], instance_iterations_per_s=b["value"], e2e_instance_iterations_per_s=b["e2e"]["value"],
                solve_ms=b["ms_per_step"], n_gpus=b["n_gpus"], roofline_frac_per_gpu=b["roofline"]["frac"], batch=b["config"]["batch"],
                iters=b["config"]["iters_per_step"], parallelism=b["config"]["parallelism"])


# ----------------------------------------------------------------------------------------------------------------
# in-run parity of the multi-GPU paths (N > 1)
# ----------------------------------------------------------------------------------------------------------------
def run_parity(pkg, eng, rank, world, local, dist, comm):
    out = {}
    lp = pkg.synth_lp(PARITY_LP["kind"], PARITY_LP["m"], PARITY_LP["n"], PARITY_LP["nnz"])
    ok_all = True
    cases = []
    for prm in (dict(stop_tol=1e-6), dict(max_iter=300, stop_tol=1e-30)):
        p = pkg.Parameters.default(use_presolve=False, device_number=local, **prm)
        model = eng.create_model(lp)
        one = eng.solve(model, p, main=True)
        par = eng.solve_partitioned_rank(model, p, comm)
        eng.free_model(model)
        err = max(float(np.max(np.abs(one[k] - par[k])) / max(1.0, float(np.max(np.abs(one[k]))))) for k in "xyz")
        ok = one["status"] == par["status"] and one["iter"] == par["iter"] and err <= 1e-8
        cases.append(dict(params=prm, status=[one["status"], par["status"]], iters=[one["iter"], par["iter"]], max_rel_err_xyz=err, ok=bool(ok)))
        ok_all = ok_all and ok
    out["partitioned_vs_single_gpu"] = dict(lp=PARITY_LP, tolerance=1e-8, cases=cases, ok=bool(ok_all))
    # sharded vs unsharded batch
    spec = WORKLOADS["c4small"]
    B = spec["batch"]
    lo, hi = pkg.shard_range(B, world, rank)
    base, d = make_batch(pkg, spec, 0, B)
    p = pkg.Parameters.default(stop_tol=1e-6, max_iter=2000, use_presolve=False, device_number=local)
    model = eng.create_model(base)
    mine = eng.solve_batched(model, d["C"][lo:hi], d["AL"][lo:hi], d["AU"][lo:hi], d["l"][lo:hi], d["u"][lo:hi], None, p)
    xs = pkg.gather_shards(dist, mine["x"], B, world, rank)
    its = pkg.gather_shards(dist, mine["iter"].astype(np.float64).reshape(-1, 1), B, world, rank)
    bok, berr = True, 0.0
    if rank == 0:
        full = eng.solve_batched(model, d["C"], d["AL"], d["AU"], d["l"], d["u"], None, p)
        berr = float(np.max(np.abs(full["x"] - xs)) / max(1.0, float(np.max(np.abs(full["x"])))))
        bok = bool(np.array_equal(full["iter"], its.reshape(-1).astype(full["iter"].dtype)) and berr <= 1e-9)
    eng.free_model(model)
    out["batch_sharded_vs_unsharded"] = dict(workload=spec["name"], tolerance=1e-9, max_rel_err_x=berr, iteration_arrays_equal=bok, ok=bool(bok))
    out["ok"] = bool(ok_all and bok)
    flag, _ = reduce_max_sum(dist, local, 0.0 if out["ok"] else 1.0, 0)
    out["ok_all_ranks"] = flag == 0.0
    return out


def run_c5(pkg, eng, rank, world, local, dist, comm):
    """configs[4]: nnz = 6e9, generated shard by shard on the GPUs, row-partitioned, solved to KKT < 1e-4."""
    p = pkg.Parameters.default(use_presolve=False, stop_tol=1e-4, time_limit=300.0, device_number=local)
    barrier(dist, local)
    t0 = time.perf_counter()
    r = eng.solve_partitioned_synth_rank(C5["m"], C5["n"], C5["K"], p, comm, want_solution=False)
    wall = time.perf_counter() - t0
    wall, _ = reduce_max_sum(dist, local, wall, 0)
    if rank != 0:
        return None
    i = r["info"]
    nnz = C5["m"] * C5["K"]
    ms_it = i["loop_device_ms"] / max(r["iter"], 1)
    per_gpu = (24 * nnz + 52 * C5["m"] + 68 * C5["n"]) / world + 20 * C5["n"]
    peak, _ = measured_peak()
    return dict(workload=C5["name"], m=C5["m"], n=C5["n"], nnz=nnz, n_gpus=world, status=r["status"], iters=r["iter"],
                primal_obj=r["primal_obj"], constructed_optimum=r["obj_star"],
                rel_obj_error=abs(r["primal_obj"] - r["obj_star"]) / (1 + abs(r["obj_star"])), residuals=r["residuals"],
                wall_s=wall, generate_and_transpose_s=i["setup_seconds"], scaling_s=i["scaling_seconds"], power_s=i["power_seconds"],
                power_iters=i["power_iters"], time_to_1e4_s=r["time"], ms_per_iteration=ms_it, iters_per_s=1e3 / ms_it,
                roofline_frac_per_gpu=per_gpu / (ms_it * 1e-3) / 1e9 / peak, bands_A=i["bands_A"])


# ----------------------------------------------------------------------------------------------------------------
# configs[0]: the CLI on the bundled toy LP, cold process, default parameters
# ----------------------------------------------------------------------------------------------------------------
def run_cli(binary, mps, extra=()):
    t0 = time.perf_counter()
    pr = subprocess.run([str(binary), "-i", str(mps), *extra], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    wall = time.perf_counter() - t0
    it, tm, status, obj = None, None, None, None
    for ln in pr.stdout.splitlines():
        s = ln.strip()
        if s.startswith("Iterations:"):
            it = int(s.split(":")[1])
        elif s.startswith("Time:"):
            tm = float(s.split(":")[1].split()[0])
        elif s.startswith("Status:"):
            status = s.split(":")[1].strip()
        elif s.startswith("Primal Objective:"):
            obj = float(s.split(":")[1])
    return dict(wall_s=wall, iters=it, solver_time_s=tm, status=status, primal_obj=obj, returncode=pr.returncode)


def run_c1(args, impl):
    spec = WORKLOADS["c1"]
    mps = ROOT / "tests" / "golden" / "model.mps"
    binary = ROOT / "build" / "solve_mps_file" if impl == "ours" else ROOT / "oracle" / "_ref" / "solve_mps_file"
    if not binary.exists():
        return dict(impl=impl, unavailable=f"{binary.relative_to(ROOT)} was not built")
    runs = [run_cli(binary, mps) for _ in range(max(2, min(args.steps, 5)))]
    best = min(runs, key=lambda r: r["wall_s"])
    out = dict(metric="HPR iterations/s (time-to-1e-4 KKT in e2e); fused SpMV+prox HBM GB/s in roofline",
               value=(best["iters"] or 0) / max(best["solver_time_s"] or 1e-2, 1e-2), unit="HPR iterations/s", n_gpus=1, steps=len(runs), warmup=0,
               ms_per_step=1e3 * best["wall_s"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="bundled",
               config=dict(workload=spec["name"], note="each step = one cold CLI process (CUDA context, MPS parse, PSLP presolve, solve); "
                           "value = iterations / the CLI's printed solver time (2 decimals, floored at 0.01 s); e2e = iterations / process wall"),
               e2e=dict(value=(best["iters"] or 0) / best["wall_s"], unit="HPR iterations/s", h2d_bytes_per_step=0 if impl == "reference" else 4 * 12 + 6 * 16,
                        d2h_bytes_per_step=0 if impl == "reference" else 48, time_to_tol_s=best["wall_s"], iters=best["iters"], status=best["status"],
                        primal_obj=best["primal_obj"]),
               all_runs=runs, gpu_launches=None, roofline=None)
    if impl == "reference":
        out["impl"] = "reference"
        out["cpu_baseline"] = dict(value=out["value"], unit="HPR iterations/s", cores=1, kind="reference", sample="the reference CLI (its own CUDA build) on data/model.mps")
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


def run_reference_arm(args, pkg, spec, local):
    lp = make_lp(pkg, spec)
    if not pkg.REF_LIB_PATH.exists():
        return dict(impl="reference", unavailable="oracle/_ref/libhprlp_ref.so was not built (needs /root/reference at build time)")
    ref = pkg.load_reference()
    with Quiet():
        with ClockSampler(local) as clk:
            runs = [run_e2e(pkg, ref, lp, local) for _ in range(max(1, min(args.steps, 3)))]
            rate, rate_ts = steady_state_rate(pkg, ref, lp, local)
    best = max(runs, key=lambda r: r["value"])
    return dict(impl="reference", metric="HPR iterations/s (time-to-1e-4 KKT in e2e); fused SpMV+prox HBM GB/s in roofline",
                value=rate, unit="HPR iterations/s", n_gpus=1, steps=len(runs), warmup=0,
                ms_per_step=1e3 * ITERS_PER_STEP / rate, steady_state=dict(
                    how="(k2-k1) iterations / (results.time[max_iter=k2] - results.time[max_iter=k1]), k1=1000, k2-k1=2000 (nnz>=5e7) or 20000",
                    times=rate_ts,
                    window_note="the reference can only be driven through solve(): its value is a difference of two host-clock results.time "
                                "(iterations 1000..3000); our arm's value is CUDA events over iterations 1300..2300 of a resident engine -- same "
                                "steady regime (one residual check per 100 iterations), different clocks"),
                higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
                data="synthetic", config=dict(workload=spec["name"], m=lp["m"], n=lp["n"], nnz=int(lp["values"].shape[0]),
                                              note="value = steady-state loop rate from two max_iter runs; e2e = one solve() call to KKT<1e-4 through the reference's own C API; "
                                                   "the reference is single-GPU: at --gpus N > 1 this line is still its 1-GPU run"),
                cpu_baseline=dict(value=rate, unit="HPR iterations/s", cores=1, kind="reference",
                                  sample="the reference has no CPU path (BASELINE.json): this arm is its own CUDA build "
                                         "(oracle/_ref/libhprlp_ref.so, autotuned fused/cuSPARSE backend) on the same B200, driven by one "
                                         "host thread; value = the line's steady-state loop rate, e2e = whole solve() calls to KKT<1e-4"),
                e2e=dict(best, h2d_bytes_per_step=0, d2h_bytes_per_step=0), clocks=clk.summary(), all_runs=runs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))   # configs[2]: the LP BASELINE.json's target is quoted on
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the configs[3] solve_batched leg of the default line")
    ap.add_argument("--no-c5", action="store_true", help="skip configs[4] (nnz = 6e9) at --gpus 8")
    ap.add_argument("--no-single-ref", action="store_true", help="N > 1: skip the 1-GPU run that speedup_vs_1gpu is measured against")
    args = ap.parse_args()

    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun exports OMP_NUM_THREADS=1; the host side (LP generator, oracle, model copies) is OpenMP/threads code
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // int(os.environ["WORLD_SIZE"])))
    pkg = graft.load_package()
    rank, world, local, dist = dist_setup(args.gpus)
    spec = WORKLOADS[args.workload]

    if spec["kind"] == "mps":
        if rank == 0:
            print(json.dumps(run_c1(args, args.impl)))
        return 0

    if "batch" in spec:
        lib = pkg.load_engine() if args.impl == "ours" else (pkg.load_reference() if pkg.REF_LIB_PATH.exists() else None)
        if args.impl == "reference" and (rank != 0 or lib is None):
            if rank == 0:
                print(json.dumps(dict(impl="reference", unavailable="oracle/_ref/libhprlp_ref.so was not built")))
            return 0
        if args.impl == "reference":
            world, dist = 1, None          # the reference is single-GPU: rank 0 runs the whole batch
        with Quiet():
            out = run_batched(args, pkg, spec, lib, rank, world, local, dist, args.impl)
        if rank == 0:
            print(json.dumps(out))
        if dist is not None:
            dist.destroy_process_group()
        return 0

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference_arm(args, pkg, spec, local)))
        return 0

    lp = make_lp(pkg, spec)
    eng = pkg.load_engine()
    with Quiet():
        if world == 1:
            out = run_engine_arm(args, pkg, spec, lp, local)
            if not args.no_e2e:
                # same protocol as the reference arm: up to 3 whole solve() calls, the best one is reported (the first call of
                # a process also loads cuRAND's kernels for the power-iteration start vector, ~0.6 s), all are listed
                runs = [run_e2e(pkg, eng, lp, local) for _ in range(max(1, min(args.steps, 3)))]
                e = max(runs, key=lambda r: r["value"])
                out["e2e"] = dict(e, all_runs_time_to_tol_s=[r["time_to_tol_s"] for r in runs])
            eng.release_cached_memory()
            if not args.no_batched and args.workload == "c3":
                out["batched"] = batched_summary(run_batched(args, pkg, WORKLOADS["c4"], eng, 0, 1, local, None, "ours"))
            if not args.no_cpu:
                out["cpu_baseline"] = cpu_baseline(pkg, lp, iters=10 if spec["nnz"] >= 5_000_000 else 200)
        else:
            comm = eng.comm_create(fresh_uid(eng, dist, rank, local), rank, world, local)   # one communicator for the whole run
            parity = run_parity(pkg, eng, rank, world, local, dist, comm)
            out = run_partitioned_arm(args, pkg, spec, lp, rank, world, local, dist, comm)
            e2e = None
            if not args.no_e2e:
                runs = [run_e2e_partitioned(pkg, eng, lp, rank, world, local, dist, comm) for _ in range(2)]
                e2e = dict(max(runs, key=lambda r: r["value"]), all_runs_time_to_tol_s=[r["time_to_tol_s"] for r in runs])
            eng.release_cached_memory()
            batched = None
            if not args.no_batched and args.workload == "c3":
                batched = batched_summary(run_batched(args, pkg, WORKLOADS["c4"], eng, rank, world, local, dist, "ours"))
            c5 = None
            if world == 8 and not args.no_c5 and args.workload == "c3":
                eng.release_cached_memory()
                c5 = run_c5(pkg, eng, rank, world, local, dist, comm)
            if rank == 0:
                out["parity"] = parity
                if e2e is not None:
                    out["e2e"] = e2e
                if batched is not None:
                    out["batched"] = batched
                if c5 is not None:
                    out["c5"] = c5
            eng.comm_destroy(comm)
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        barrier(dist, local)
        dist.destroy_process_group()
    if world > 1 and rank == 0 and not out.get("parity", {}).get("ok_all_ranks", True):
        print("[bench] in-run parity FAILED: " + json.dumps(out["parity"]), file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
