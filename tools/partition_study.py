"""Break-even study of the row-block partitioned solve: loop time per HPR iteration on 1..N GPUs for
growing synthetic LPs.  Usage: python tools/partition_study.py [--gpus 1,2,4,8] [--sizes small,c2,c3] [--iters 200]"""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

SIZES = {
    "s1e6": dict(kind="uniform", m=20_000, n=100_000, nnz=1_000_000),
    "c2": dict(kind="uniform", m=100_000, n=1_000_000, nnz=10_000_000),
    "u3e7": dict(kind="uniform", m=300_000, n=3_000_000, nnz=30_000_000),
    "c3": dict(kind="powerlaw", m=2_000_000, n=5_000_000, nnz=100_000_000),
    "u4e8": dict(kind="uniform", m=4_000_000, n=10_000_000, nnz=400_000_000),
}

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="1,2")
ap.add_argument("--sizes", default="s1e6,c2,u3e7,c3")
ap.add_argument("--iters", type=int, default=200)
args = ap.parse_args()
pkg = graft.load_package()
eng = pkg.load_engine()
import torch
avail = torch.cuda.device_count()
devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1)
for name in args.sizes.split(","):
    lp = pkg.synth_lp(**SIZES[name])
    nnz = int(lp["values"].shape[0])
    base = None
    gpu_list = [int(x) for x in args.gpus.split(",")]
    if 1 not in gpu_list:
        gpu_list = [1] + gpu_list          # the 1-GPU engine is always measured: it is the denominator of speedup_vs_1gpu
    for g in gpu_list:
        if g > avail:
            continue
        p = pkg.Parameters.default(use_presolve=False, stop_tol=0.0, max_iter=args.iters)
        model = eng.create_model(lp)
        os.dup2(devnull, 1)
        try:
            # warm-up: CUDA module load, NCCL lazy connection setup, graph instantiation
            eng.solve_partitioned(model, pkg.Parameters.default(use_presolve=False, stop_tol=0.0, max_iter=20), n_gpus=g)
            r = eng.solve_partitioned(model, p, n_gpus=g)
        finally:
            os.dup2(saved, 1)
        eng.free_model(model)
        ms_it = r["info"]["loop_device_ms"] / max(r["iter"], 1)
        if g == 1:
            base = ms_it
        print(json.dumps(dict(size=name, m=lp["m"], n=lp["n"], nnz=nnz, gpus=g, iters=r["iter"], loop_ms_per_iter=ms_it,
                              iters_per_s=1e3 / ms_it, speedup_vs_1gpu=base / ms_it, power_s=r["info"]["power_seconds"],
                              setup_s=r["info"]["setup_seconds"], scaling_s=r["info"]["scaling_seconds"],
                              exchange=("peer-memory push (fused)" if r["info"]["peer_exchange"] else "nccl reduce-scatter + all-gather") if g > 1 else None,
                              nvlink_bytes_per_gpu_per_iter=2 * 8 * lp["n"] * (g - 1) // g if g > 1 else 0)), flush=True)
