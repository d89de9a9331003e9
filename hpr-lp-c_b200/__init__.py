"""hprlp_b200 -- ctypes mirror of the HPR-LP C ABI (include/HPRLP.h, include/batched_solver.h,
include/hprlp_b200.h) for tests and bench.py.

The product is the C-ABI shared library ``lib/libhprlp.so`` (CUDA, sm_100a); this module only
declares the struct layouts and function signatures, exactly as the reference's Julia wrapper does
(reference bindings/julia/package/src/wrapper.jl:92-164).  It also knows how to load, side by
side, the reference's own CUDA build (``oracle/_ref/libhprlp_ref.so``, same seven symbols), the CPU
oracle (``oracle/liboracle.so``) and the synthetic LP generator (``tools/libsynth.so``) -- those three
are test/bench infrastructure, never used by the product path.

There is no CPU fallback: loading or calling the engine without the CUDA library / a GPU raises.
The directory is named ``hpr-lp-c_b200``; import it with ``__graft_entry__.load_package()``.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ["HPRLP_LIB"]) if os.environ.get("HPRLP_LIB") else ROOT / "lib" / "libhprlp.so"   # HPRLP_LIB: tuning variants (tools/build_variants.sh)
REF_LIB_PATH = ROOT / "oracle" / "_ref" / "libhprlp_ref.so"
ORACLE_LIB_PATH = ROOT / "oracle" / "liboracle.so"
SYNTH_LIB_PATH = ROOT / "tools" / "libsynth.so"

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


# ---------------------------------------------------------------------------------------------
# struct mirrors (include/structs.h)
# ---------------------------------------------------------------------------------------------
class SparseMatrix(C.Structure):
    _fields_ = [("row", C.c_int), ("col", C.c_int), ("numElements", C.c_int),
                ("colIndex", c_int_p), ("rowPtr", c_int_p), ("value", c_double_p)]


class Parameters(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("stop_tol", C.c_double), ("time_limit", C.c_double),
                ("device_number", C.c_int), ("check_iter", C.c_int),
                ("CUSPARSE_spmv", C.c_bool), ("autotune_verbose", C.c_bool),
                ("use_CR_scaling", C.c_bool), ("use_Ruiz_scaling", C.c_bool),
                ("use_Pock_Chambolle_scaling", C.c_bool), ("use_bc_scaling", C.c_bool),
                ("use_presolve", C.c_bool)]

    @classmethod
    def default(cls, **kw):
        p = cls(2**31 - 1, 1e-4, 3600.0, 0, 150, False, False, True, True, True, True, True)
        for k, v in kw.items():
            setattr(p, k, v)
        return p


class Results(C.Structure):
    _fields_ = [("residuals", C.c_double), ("primal_obj", C.c_double), ("gap", C.c_double),
                ("time4", C.c_double), ("time6", C.c_double), ("time8", C.c_double), ("time", C.c_double),
                ("iter4", C.c_int), ("iter6", C.c_int), ("iter8", C.c_int), ("iter", C.c_int),
                ("status", C.c_char * 64),
                ("x", c_double_p), ("y", c_double_p), ("z", c_double_p)]


class BatchedResults(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("batch_size", C.c_int),
                ("x", c_double_p), ("y", c_double_p), ("z", c_double_p),
                ("primal_obj", c_double_p), ("residuals", c_double_p), ("gap", c_double_p),
                ("iter", c_int_p), ("status", C.POINTER(C.c_char)),
                ("time", C.c_double), ("setup_time", C.c_double), ("solve_time", C.c_double),
                ("power_time", C.c_double)]


class LPInfoCpu(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("A", C.POINTER(SparseMatrix)),
                ("AL", c_double_p), ("AU", c_double_p), ("c", c_double_p),
                ("l", c_double_p), ("u", c_double_p), ("obj_constant", C.c_double)]


class B200Info(C.Structure):
    _fields_ = [("lambda_max", C.c_double), ("sigma", C.c_double),
                ("setup_seconds", C.c_double), ("scaling_seconds", C.c_double), ("power_seconds", C.c_double),
                ("loop_device_ms", C.c_double), ("restarts", C.c_int), ("power_iters", C.c_int),
                ("kernel_launches", C.c_longlong),
                ("b_scale", C.c_double), ("c_scale", C.c_double), ("norm_b", C.c_double), ("norm_c", C.c_double),
                ("norm_b_org", C.c_double), ("norm_c_org", C.c_double),
                ("lanes_A", C.c_int), ("lanes_AT", C.c_int), ("items_A", C.c_int), ("items_AT", C.c_int),
                ("bands_A", C.c_int), ("reserved0", C.c_int), ("peer_exchange", C.c_int), ("reserved1", C.c_int)]


REFERENCE_SYMBOLS = ["create_model_from_arrays", "create_model_from_mps", "solve", "free_model",
                     "HPRLP_main_solve", "solve_batched", "free_batched_results"]
EXTENDED_SYMBOLS = ["hprlp_b200_solve_ex", "hprlp_b200_power_start", "hprlp_b200_engine_create",
                    "hprlp_b200_engine_run", "hprlp_b200_engine_time_phase", "hprlp_b200_engine_residuals",
                    "hprlp_b200_engine_info", "hprlp_b200_engine_destroy", "hprlp_b200_scale_only",
                    "hprlp_b200_solve_batched_multi", "hprlp_b200_solve_partitioned", "hprlp_b200_presolve", "hprlp_b200_presolve_free", "hprlp_b200_solve_partitioned_synth", "hprlp_b200_synth_rows", "hprlp_b200_profiler_start", "hprlp_b200_profiler_stop", "hprlp_b200_version",
                    "hprlp_b200_solve_partitioned_local", "hprlp_b200_nccl_unique_id", "hprlp_b200_solve_partitioned_rank",
                    "hprlp_b200_engine_create_rank", "hprlp_b200_nccl_exchange_ms", "hprlp_b200_solve_partitioned_synth_rank",
                    "hprlp_b200_comm_create", "hprlp_b200_comm_destroy", "hprlp_b200_warmup", "hprlp_b200_solve_batched_layout",
                    "hprlp_b200_release_cached_memory"]


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class HprLib:
    """One loaded libhprlp (ours or the reference build): the seven reference symbols."""

    def __init__(self, path, extended=False):
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (run __graft_entry__.build())")
        self.path = path
        self.lib = C.CDLL(str(path), mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
        L = self.lib
        L.create_model_from_arrays.restype = C.POINTER(LPInfoCpu)
        L.create_model_from_arrays.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p,
                                               c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_bool]
        L.create_model_from_mps.restype = C.POINTER(LPInfoCpu)
        L.create_model_from_mps.argtypes = [C.c_char_p]
        L.solve.restype = Results
        L.solve.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters)]
        L.HPRLP_main_solve.restype = Results
        L.HPRLP_main_solve.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters)]
        L.free_model.restype = None
        L.free_model.argtypes = [C.POINTER(LPInfoCpu)]
        L.solve_batched.restype = BatchedResults
        L.solve_batched.argtypes = [C.POINTER(LPInfoCpu), C.c_int, c_double_p, c_double_p, c_double_p,
                                    c_double_p, c_double_p, c_double_p, C.POINTER(Parameters)]
        L.free_batched_results.restype = None
        L.free_batched_results.argtypes = [C.POINTER(BatchedResults)]
        self.extended = extended
        if extended:
            L.hprlp_b200_solve_ex.restype = Results
            L.hprlp_b200_solve_ex.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters), c_double_p, C.c_int, c_int_p,
                                              c_double_p, c_double_p, c_double_p, C.c_int, C.POINTER(B200Info)]
            L.hprlp_b200_power_start.restype = C.c_int
            L.hprlp_b200_power_start.argtypes = [C.c_int, C.c_int, c_double_p]
            L.hprlp_b200_engine_create.restype = C.c_void_p
            L.hprlp_b200_engine_create.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters)]
            L.hprlp_b200_engine_run.restype = C.c_double
            L.hprlp_b200_engine_run.argtypes = [C.c_void_p, C.c_int]
            L.hprlp_b200_engine_time_phase.restype = C.c_double
            L.hprlp_b200_engine_time_phase.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.hprlp_b200_engine_residuals.restype = C.c_int
            L.hprlp_b200_engine_residuals.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
            L.hprlp_b200_engine_info.restype = None
            L.hprlp_b200_engine_info.argtypes = [C.c_void_p, C.POINTER(B200Info)]
            L.hprlp_b200_engine_destroy.restype = None
            L.hprlp_b200_engine_destroy.argtypes = [C.c_void_p]
            L.hprlp_b200_scale_only.restype = C.c_int
            L.hprlp_b200_scale_only.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters)] + [c_double_p, c_int_p, c_int_p] + \
                [c_double_p] * 9
            L.hprlp_b200_version.restype = C.c_char_p
            L.hprlp_b200_solve_partitioned_synth.restype = Results
            L.hprlp_b200_solve_partitioned_synth.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(Parameters), C.c_int,
                                                             C.c_int, C.c_int, c_double_p, C.POINTER(B200Info)]
            L.hprlp_b200_synth_rows.restype = C.c_int
            L.hprlp_b200_synth_rows.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_longlong, C.c_int, c_int_p, c_double_p]
            L.hprlp_b200_solve_partitioned.restype = Results
            L.hprlp_b200_solve_partitioned.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters), C.c_int, C.c_int, C.POINTER(B200Info)]
            L.hprlp_b200_solve_batched_multi.restype = BatchedResults
            L.hprlp_b200_solve_batched_multi.argtypes = L.solve_batched.argtypes + [C.c_int]
            L.hprlp_b200_solve_batched_layout.restype = BatchedResults
            L.hprlp_b200_solve_batched_layout.argtypes = L.solve_batched.argtypes + [C.c_int]
            L.hprlp_b200_solve_partitioned_local.restype = Results
            L.hprlp_b200_solve_partitioned_local.argtypes = L.hprlp_b200_solve_partitioned.argtypes
            L.hprlp_b200_nccl_unique_id.restype = C.c_int
            L.hprlp_b200_nccl_unique_id.argtypes = [C.c_char_p]
            L.hprlp_b200_comm_create.restype = C.c_void_p
            L.hprlp_b200_comm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
            L.hprlp_b200_comm_destroy.restype = None
            L.hprlp_b200_comm_destroy.argtypes = [C.c_void_p]
            L.hprlp_b200_solve_partitioned_rank.restype = Results
            L.hprlp_b200_solve_partitioned_rank.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters), C.c_void_p,
                                                            C.c_int, C.POINTER(B200Info)]
            L.hprlp_b200_engine_create_rank.restype = C.c_void_p
            L.hprlp_b200_engine_create_rank.argtypes = [C.POINTER(LPInfoCpu), C.POINTER(Parameters), C.c_void_p]
            L.hprlp_b200_nccl_exchange_ms.restype = C.c_int
            L.hprlp_b200_nccl_exchange_ms.argtypes = [C.c_int, C.c_longlong, C.c_int, c_double_p]
            L.hprlp_b200_solve_partitioned_synth_rank.restype = Results
            L.hprlp_b200_solve_partitioned_synth_rank.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_ulonglong, C.POINTER(Parameters),
                                                                  C.c_void_p, C.c_int, C.c_int, c_double_p, C.POINTER(B200Info)]
            L.hprlp_b200_release_cached_memory.restype = None
            L.hprlp_b200_release_cached_memory.argtypes = []

    # -- model layer -----------------------------------------------------------------------------
    def create_model(self, lp, is_csc=False):
        """lp: dict with m,n,rowPtr,colIndex,values,AL,AU,l,u,c (numpy)."""
        keep = {k: (_i32(lp[k]) if k in ("rowPtr", "colIndex") else _f64(lp[k]))
                for k in ("rowPtr", "colIndex", "values", "AL", "AU", "l", "u", "c")}
        nnz = int(keep["values"].shape[0])
        model = self.lib.create_model_from_arrays(int(lp["m"]), int(lp["n"]), nnz, _ip(keep["rowPtr"]), _ip(keep["colIndex"]),
                                                  _dp(keep["values"]), _dp(keep["AL"]), _dp(keep["AU"]), _dp(keep["l"]),
                                                  _dp(keep["u"]), _dp(keep["c"]), bool(is_csc))
        return model

    def create_model_from_mps(self, path):
        return self.lib.create_model_from_mps(str(path).encode())

    def free_model(self, model):
        self.lib.free_model(model)

    @staticmethod
    def model_arrays(model):
        """Copy the arrays of an LP_info_cpu* back to numpy (for bit-exact model comparisons)."""
        mm = model.contents
        A = mm.A.contents
        m, n, nnz = mm.m, mm.n, A.numElements
        return dict(m=m, n=n, nnz=nnz,
                    rowPtr=np.ctypeslib.as_array(A.rowPtr, (m + 1,)).copy(),
                    colIndex=np.ctypeslib.as_array(A.colIndex, (nnz,)).copy(),
                    values=np.ctypeslib.as_array(A.value, (nnz,)).copy(),
                    AL=np.ctypeslib.as_array(mm.AL, (m,)).copy(), AU=np.ctypeslib.as_array(mm.AU, (m,)).copy(),
                    c=np.ctypeslib.as_array(mm.c, (n,)).copy(), l=np.ctypeslib.as_array(mm.l, (n,)).copy(),
                    u=np.ctypeslib.as_array(mm.u, (n,)).copy(), obj_constant=mm.obj_constant)

    # -- solve -------------------------------------------------------------------------------------
    def _take(self, res, m, n):
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        out = dict(status=res.status.decode(), iter=res.iter, residuals=res.residuals, primal_obj=res.primal_obj,
                   gap=res.gap, time=res.time, time4=res.time4, iter4=res.iter4)
        for name, ln in (("x", n), ("y", m), ("z", n)):
            p = getattr(res, name)
            out[name] = np.ctypeslib.as_array(p, (ln,)).copy() if p else None
            if p:
                libc.free(C.cast(p, C.c_void_p))
        return out

    def solve(self, model, param=None, main=False):
        mm = model.contents
        fn = self.lib.HPRLP_main_solve if main else self.lib.solve
        res = fn(model, C.byref(param) if param is not None else None)
        return self._take(res, mm.m, mm.n)

    def solve_ex(self, model, param, power_z0=None, trace_iters=(), quiet=True):
        assert self.extended
        mm = model.contents
        m, n = mm.m, mm.n
        ti = _i32(list(trace_iters))
        nt = int(ti.shape[0])
        tx = np.zeros((max(nt, 1), n)); ty = np.zeros((max(nt, 1), m)); tz = np.zeros((max(nt, 1), n))
        z0 = _f64(power_z0) if power_z0 is not None else None
        info = B200Info()
        res = self.lib.hprlp_b200_solve_ex(model, C.byref(param), _dp(z0), nt, _ip(ti), _dp(tx), _dp(ty), _dp(tz),
                                           1 if quiet else 0, C.byref(info))
        out = self._take(res, m, n)
        out["info"] = {f[0]: getattr(info, f[0]) for f in B200Info._fields_}
        out["trace"] = {int(k): (tx[i].copy(), ty[i].copy(), tz[i].copy()) for i, k in enumerate(ti)}
        return out

    def solve_partitioned(self, model, param, n_gpus, quiet=True, local=False):
        """Row-partitioned solve driven from this process: n_gpus GPUs over NCCL, or (local=True) n_gpus logical ranks
        on ONE GPU with host-synchronised exchanges (the parity-test path of 1-GPU boxes)."""
        mm = model.contents
        info = B200Info()
        fn = self.lib.hprlp_b200_solve_partitioned_local if local else self.lib.hprlp_b200_solve_partitioned
        res = fn(model, C.byref(param), int(n_gpus), 1 if quiet else 0, C.byref(info))
        out = self._take(res, mm.m, mm.n)
        out["info"] = {f[0]: getattr(info, f[0]) for f in B200Info._fields_}
        return out

    def nccl_unique_id(self):
        buf = C.create_string_buffer(128)
        if self.lib.hprlp_b200_nccl_unique_id(buf) != 0:
            raise RuntimeError("hprlp_b200_nccl_unique_id failed")
        return buf.raw

    def comm_create(self, uid, rank, nranks, device):
        """This rank's endpoint of a process-per-GPU partitioned solve (NCCL communicator); reuse it across solves."""
        h = self.lib.hprlp_b200_comm_create(uid, int(rank), int(nranks), int(device))
        if not h:
            raise RuntimeError("hprlp_b200_comm_create failed")
        return h

    def comm_destroy(self, comm):
        self.lib.hprlp_b200_comm_destroy(comm)

    def solve_partitioned_rank(self, model, param, comm, quiet=True):
        """One process per GPU: this process owns one row block; every rank gets the full solution."""
        mm = model.contents
        info = B200Info()
        res = self.lib.hprlp_b200_solve_partitioned_rank(model, C.byref(param), comm, 1 if quiet else 0, C.byref(info))
        out = self._take(res, mm.m, mm.n)
        out["info"] = {f[0]: getattr(info, f[0]) for f in B200Info._fields_}
        return out

    def solve_partitioned_synth_rank(self, m, n, K, param, comm, seed=None, want_solution=False, quiet=True):
        info = B200Info()
        obj = C.c_double(0.0)
        res = self.lib.hprlp_b200_solve_partitioned_synth_rank(int(m), int(n), int(K), SEED if seed is None else seed, C.byref(param),
                                                               comm, 1 if quiet else 0, 1 if want_solution else 0, C.byref(obj),
                                                               C.byref(info))
        out = self._take(res, int(m), int(n))
        out["info"] = {f[0]: getattr(info, f[0]) for f in B200Info._fields_}
        out["obj_star"] = obj.value
        return out

    def nccl_exchange_ms(self, n_gpus, count, reps=20):
        out = np.zeros(2)
        if self.lib.hprlp_b200_nccl_exchange_ms(int(n_gpus), int(count), int(reps), _dp(out)) != 0:
            raise RuntimeError("hprlp_b200_nccl_exchange_ms failed")
        return dict(reduce_scatter_all_gather_ms=float(out[0]), all_reduce_ms=float(out[1]))

    def release_cached_memory(self):
        self.lib.hprlp_b200_release_cached_memory()

    def solve_partitioned_synth(self, m, n, K, param, n_gpus, seed=None, want_solution=True, quiet=True):
        info = B200Info()
        obj = C.c_double(0.0)
        res = self.lib.hprlp_b200_solve_partitioned_synth(int(m), int(n), int(K), SEED if seed is None else seed, C.byref(param),
                                                          int(n_gpus), 1 if quiet else 0, 1 if want_solution else 0, C.byref(obj),
                                                          C.byref(info))
        out = self._take(res, int(m), int(n))
        out["info"] = {f[0]: getattr(info, f[0]) for f in B200Info._fields_}
        out["obj_star"] = obj.value
        return out

    def synth_rows(self, n, K, row0, rows, seed=None):
        col = np.zeros(rows * K, np.int32); val = np.zeros(rows * K)
        rc = self.lib.hprlp_b200_synth_rows(int(n), int(K), SEED if seed is None else seed, int(row0), int(rows), _ip(col), _dp(val))
        if rc != 0:
            raise RuntimeError("hprlp_b200_synth_rows failed")
        return col, val

    def power_start(self, m, device=0):
        out = np.zeros(m)
        rc = self.lib.hprlp_b200_power_start(int(m), int(device), _dp(out))
        if rc != 0:
            raise RuntimeError("hprlp_b200_power_start failed")
        return out

    def scale_only(self, model, param):
        mm = model.contents
        m, n, nnz = mm.m, mm.n, mm.A.contents.numElements
        o = dict(A_val=np.zeros(nnz), AT_rowPtr=np.zeros(n + 1, np.int32), AT_col=np.zeros(nnz, np.int32),
                 AT_val=np.zeros(nnz), AL=np.zeros(m), AU=np.zeros(m), l=np.zeros(n), u=np.zeros(n), c=np.zeros(n),
                 row_norm=np.zeros(m), col_norm=np.zeros(n), scalars=np.zeros(6))
        rc = self.lib.hprlp_b200_scale_only(model, C.byref(param), _dp(o["A_val"]), _ip(o["AT_rowPtr"]), _ip(o["AT_col"]),
                                            _dp(o["AT_val"]), _dp(o["AL"]), _dp(o["AU"]), _dp(o["l"]), _dp(o["u"]), _dp(o["c"]),
                                            _dp(o["row_norm"]), _dp(o["col_norm"]), _dp(o["scalars"]))
        if rc != 0:
            raise RuntimeError("hprlp_b200_scale_only failed")
        return o

    def solve_batched_nB(self, model, C_, AL, AU, l, u, obj_constants=None, param=None):
        """The reference's Python calling convention: C, l, u of shape (n, B) and AL, AU of shape (m, B), any memory order.
        C-ordered arrays (numpy's default) go through hprlp_b200_solve_batched_layout(layout=1) without any host copy;
        Fortran-ordered ones are already the ABI's column-major layout.  Returns x, z as (n, B) and y as (m, B) views."""
        arrs = [np.asarray(a, dtype=np.float64) for a in (C_, AL, AU, l, u)]
        B = arrs[0].shape[1]
        if all(a.flags.c_contiguous for a in arrs):
            layout = 1
        else:
            arrs = [np.asfortranarray(a) for a in arrs]
            layout = 0
        out = self.solve_batched(model, *[a if layout == 1 else a.T for a in arrs], obj_constants, param, _layout=layout, _B=B)
        for k in ("x", "y", "z"):
            if k in out:
                out[k] = out[k].T
        return out

    def solve_batched(self, model, C_, AL, AU, l, u, obj_constants=None, param=None, n_gpus=None, _layout=None, _B=None):
        """Dense inputs column-major: arrays of shape (B, n) / (B, m) in C order == n x B column-major."""
        mm = model.contents
        m, n = mm.m, mm.n
        C_, AL, AU, l, u = (_f64(a) for a in (C_, AL, AU, l, u))
        B = C_.shape[0] if _B is None else _B
        oc = _f64(obj_constants) if obj_constants is not None else None
        args = (model, B, _dp(C_), _dp(AL), _dp(AU), _dp(l), _dp(u), _dp(oc), C.byref(param) if param is not None else None)
        import time as _time
        t0 = _time.perf_counter()
        if _layout is not None:
            res = self.lib.hprlp_b200_solve_batched_layout(*args, int(_layout))
        else:
            res = self.lib.solve_batched(*args) if n_gpus is None else self.lib.hprlp_b200_solve_batched_multi(*args, int(n_gpus))
        call_seconds = _time.perf_counter() - t0      # the C-ABI call alone: host arrays in, malloc'ed host arrays out
        out = dict(m=res.m, n=res.n, batch_size=res.batch_size, time=res.time, setup_time=res.setup_time,
                   solve_time=res.solve_time, power_time=res.power_time, call_seconds=call_seconds)
        if res.status:
            raw = C.string_at(res.status, 64 * B)
            out["status"] = [raw[64 * k:64 * k + 64].split(b"\0")[0].decode() for k in range(B)]
        if res.x:
            out["x"] = np.ctypeslib.as_array(res.x, (B, n)).copy()
            out["y"] = np.ctypeslib.as_array(res.y, (B, m)).copy()
            out["z"] = np.ctypeslib.as_array(res.z, (B, n)).copy()
            out["primal_obj"] = np.ctypeslib.as_array(res.primal_obj, (B,)).copy()
            out["residuals"] = np.ctypeslib.as_array(res.residuals, (B,)).copy()
            out["gap"] = np.ctypeslib.as_array(res.gap, (B,)).copy()
            out["iter"] = np.ctypeslib.as_array(res.iter, (B,)).copy()
        self.lib.free_batched_results(C.byref(res))
        return out


_cache = {}


def load_engine():
    """The product library. Raises if it is not built -- there is no fallback."""
    if "eng" not in _cache:
        _cache["eng"] = HprLib(LIB_PATH, extended=True)
    return _cache["eng"]


def load_reference():
    """The reference's own CUDA build (oracle/_ref), test/bench infrastructure."""
    if "ref" not in _cache:
        _cache["ref"] = HprLib(REF_LIB_PATH, extended=False)
    return _cache["ref"]


# ---------------------------------------------------------------------------------------------
# CPU oracle (oracle/hpr_oracle.c) -- test infrastructure only
# ---------------------------------------------------------------------------------------------
class OracleParams(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("stop_tol", C.c_double), ("time_limit", C.c_double), ("check_iter", C.c_int),
                ("use_cr", C.c_int), ("use_ruiz", C.c_int), ("use_pc", C.c_int), ("use_bc", C.c_int)]


class OracleInfo(C.Structure):
    _fields_ = [(k, C.c_double) for k in "residuals primal_obj dual_obj gap err_rp err_rd".split()] + \
               [("iter", C.c_int), ("status", C.c_char * 64), ("lambda_max", C.c_double), ("sigma", C.c_double),
                ("restarts", C.c_int)] + \
               [(k, C.c_double) for k in "b_scale c_scale norm_b norm_c norm_b_org norm_c_org".split()] + \
               [("power_iters", C.c_int), ("solve_seconds", C.c_double)]


class Oracle:
    def __init__(self):
        if not ORACLE_LIB_PATH.exists():
            raise FileNotFoundError(f"{ORACLE_LIB_PATH} not built (make -C oracle liboracle.so)")
        self.lib = C.CDLL(str(ORACLE_LIB_PATH))
        self.lib.oracle_power.restype = C.c_double

    @staticmethod
    def params_from(param: Parameters):
        return OracleParams(param.max_iter, param.stop_tol, param.time_limit, param.check_iter,
                            int(param.use_CR_scaling), int(param.use_Ruiz_scaling),
                            int(param.use_Pock_Chambolle_scaling), int(param.use_bc_scaling))

    def solve(self, lp, param: Parameters, power_z0=None, trace_iters=(), obj_constant=0.0):
        m, n = int(lp["m"]), int(lp["n"])
        rp, ci, v = _i32(lp["rowPtr"]), _i32(lp["colIndex"]), _f64(lp["values"])
        AL, AU, l, u, c = (_f64(lp[k]) for k in ("AL", "AU", "l", "u", "c"))
        x, y, z = np.zeros(n), np.zeros(m), np.zeros(n)
        ti = _i32(list(trace_iters)); nt = int(ti.shape[0])
        tx = np.zeros((max(nt, 1), n)); ty = np.zeros((max(nt, 1), m)); tz = np.zeros((max(nt, 1), n))
        z0 = _f64(power_z0) if power_z0 is not None else None
        info = OracleInfo()
        op = self.params_from(param)
        self.lib.oracle_solve(m, n, _ip(rp), _ip(ci), _dp(v), _dp(AL), _dp(AU), _dp(l), _dp(u), _dp(c),
                              C.c_double(obj_constant), C.byref(op), _dp(z0), _dp(x), _dp(y), _dp(z), C.byref(info),
                              nt, _ip(ti), _dp(tx), _dp(ty), _dp(tz))
        out = dict(status=info.status.decode(), iter=info.iter, residuals=info.residuals, primal_obj=info.primal_obj,
                   dual_obj=info.dual_obj, gap=info.gap, x=x, y=y, z=z,
                   info={f[0]: getattr(info, f[0]) for f in OracleInfo._fields_ if f[0] != "status"},
                   trace={int(k): (tx[i].copy(), ty[i].copy(), tz[i].copy()) for i, k in enumerate(ti)})
        return out

    def transpose(self, rows, cols, rp, ci, v):
        rp, ci, v = _i32(rp), _i32(ci), _f64(v)
        nnz = int(v.shape[0])
        trp = np.zeros(cols + 1, np.int32); tci = np.zeros(nnz, np.int32); tv = np.zeros(nnz)
        self.lib.oracle_transpose(int(rows), int(cols), nnz, _ip(rp), _ip(ci), _dp(v), _ip(trp), _ip(tci), _dp(tv))
        return trp, tci, tv

    def scale(self, lp, param: Parameters):
        m, n = int(lp["m"]), int(lp["n"])
        rp, ci = _i32(lp["rowPtr"]), _i32(lp["colIndex"])
        v = _f64(lp["values"]).copy()
        AL, AU, l, u, c = (_f64(lp[k]).copy() for k in ("AL", "AU", "l", "u", "c"))
        nnz = int(v.shape[0])
        o = dict(AT_rowPtr=np.zeros(n + 1, np.int32), AT_col=np.zeros(nnz, np.int32), AT_val=np.zeros(nnz),
                 row_norm=np.zeros(m), col_norm=np.zeros(n), scalars=np.zeros(6))
        op = self.params_from(param)
        self.lib.oracle_scale(m, n, _ip(rp), _ip(ci), _dp(v), _dp(AL), _dp(AU), _dp(l), _dp(u), _dp(c), C.byref(op),
                              _ip(o["AT_rowPtr"]), _ip(o["AT_col"]), _dp(o["AT_val"]), _dp(o["row_norm"]), _dp(o["col_norm"]),
                              _dp(o["scalars"]))
        o.update(A_val=v, AL=AL, AU=AU, l=l, u=u, c=c)
        return o


def load_oracle():
    if "oracle" not in _cache:
        _cache["oracle"] = Oracle()
    return _cache["oracle"]


# ---------------------------------------------------------------------------------------------
# synthetic LPs (tools/synth_lp.c)
# ---------------------------------------------------------------------------------------------
SEED = 20251018


def synth_lp(kind, m, n, nnz, seed=SEED, vec_seed=None, with_solution=False):
    """kind: 'uniform' | 'powerlaw' | 'banded' (uniform row lengths, columns of row i within a 4096-wide window around
    i*n/m: gathers with cache locality) | 'blocked' (dense 8x8 blocks inside that window: gathers that coalesce into whole
    64-byte segments in both passes). Returns the dict create_model() takes
    (+ x*,y*,z*,obj*)."""
    if not SYNTH_LIB_PATH.exists():
        raise FileNotFoundError(f"{SYNTH_LIB_PATH} not built")
    L = C.CDLL(str(SYNTH_LIB_PATH))
    L.synth_lp_rowptr.restype = C.c_longlong
    L.synth_lp_rowptr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_uint64, c_int_p]
    L.synth_lp_matrix_rows.restype = None
    L.synth_lp_matrix_rows.argtypes = [C.c_int, C.c_uint64, c_int_p, C.c_int, C.c_int, c_int_p, c_double_p]
    L.synth_lp_vectors.restype = C.c_double
    L.synth_lp_vectors.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, C.c_uint64, C.c_uint64] + [c_double_p] * 8
    k = {"uniform": 0, "powerlaw": 1, "banded": 0, "blocked": 0}[kind]
    rp = np.zeros(m + 1, np.int32)
    tot = L.synth_lp_rowptr(k, m, n, int(nnz), seed, _ip(rp))
    col = np.zeros(tot, np.int32); val = np.zeros(tot)
    if kind == "banded":
        L.synth_lp_matrix_rows_banded.restype = None
        L.synth_lp_matrix_rows_banded.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, c_int_p, C.c_int, C.c_int, c_int_p, c_double_p]
        L.synth_lp_matrix_rows_banded(m, n, 4096, seed, _ip(rp), 0, m, _ip(col), _dp(val))
    elif kind == "blocked":
        L.synth_lp_matrix_rows_blocked.restype = None
        L.synth_lp_matrix_rows_blocked.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, c_int_p, C.c_int, C.c_int, c_int_p, c_double_p]
        L.synth_lp_matrix_rows_blocked(m, n, 4096, int(os.environ.get("SYNTH_BLOCK_RUN", "8")), seed, _ip(rp), 0, m, _ip(col), _dp(val))
    else:
        L.synth_lp_matrix_rows(n, seed, _ip(rp), 0, m, _ip(col), _dp(val))
    lp = dict(m=m, n=n, rowPtr=rp, colIndex=col, values=val)
    lp.update(synth_vectors(lp, seed, seed if vec_seed is None else vec_seed, _lib=L))
    if not with_solution:
        for kx in ("xs", "ys", "zs"):
            lp.pop(kx)
    return lp


def synth_vectors(lp, seed=SEED, vec_seed=SEED, _lib=None):
    L = _lib or C.CDLL(str(SYNTH_LIB_PATH))
    if _lib is None:
        L.synth_lp_vectors.restype = C.c_double
        L.synth_lp_vectors.argtypes = [C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, C.c_uint64, C.c_uint64] + [c_double_p] * 8
    m, n = lp["m"], lp["n"]
    AL, AU, l, u, c = np.zeros(m), np.zeros(m), np.zeros(n), np.zeros(n), np.zeros(n)
    xs, ys, zs = np.zeros(n), np.zeros(m), np.zeros(n)
    obj = L.synth_lp_vectors(m, n, _ip(lp["rowPtr"]), _ip(lp["colIndex"]), _dp(lp["values"]), seed, vec_seed,
                             _dp(AL), _dp(AU), _dp(l), _dp(u), _dp(c), _dp(xs), _dp(ys), _dp(zs))
    return dict(AL=AL, AU=AU, l=l, u=u, c=c, xs=xs, ys=ys, zs=zs, obj_star=obj)


TOY_LP = dict(m=2, n=2, rowPtr=np.array([0, 2, 4], np.int32), colIndex=np.array([0, 1, 0, 1], np.int32),
              values=np.array([1.0, 2.0, 3.0, 1.0]), AL=np.array([-np.inf, -np.inf]), AU=np.array([10.0, 12.0]),
              l=np.zeros(2), u=np.full(2, np.inf), c=np.array([-3.0, -5.0]))
"""The toy LP of every reference example (examples/cpp/example_direct_lp.cpp:14): x*=(2.8,3.6), obj=-26.4."""


# ---------------------------------------------------------------------------------------------
# multi-process batch sharding helpers (bench.py under torchrun; tests/test_multi_cpu.py with gloo)
# ---------------------------------------------------------------------------------------------
def shard_range(B, world, rank):
    """Contiguous instance shard of rank `rank` -- the same split hprlp_b200_solve_batched_multi uses."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(dist, local, B, world, rank):
    """All-gather per-rank result rows (shape (shard, k)) into the (B, k) array in instance order."""
    import torch
    k = int(local.shape[1])
    sizes = [shard_range(B, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros((width, k), dtype=torch.float64, device=dev)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([outs[r][: hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)], axis=0)


def broadcast_unique_id(dist, make_id, rank):
    """The 128-byte NCCL unique id of a new engine communicator: made by rank 0 (make_id() -> bytes), broadcast with
    torch.distributed (one process per GPU; bench.py under torchrun).  Works on the nccl and gloo backends."""
    import torch
    dev = f"cuda:{torch.cuda.current_device()}" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def row_blocks_by_nnz(rowPtr, P):
    """Row blocks of the partitioned solve, balanced by nonzeros -- the split csrc/partitioned.cu uses
    (block p = rows [b[p], b[p+1]))."""
    rp = np.asarray(rowPtr, dtype=np.int64)
    m, nnz = rp.shape[0] - 1, int(rp[-1])
    b = [0] * (P + 1)
    b[P] = m
    for p in range(1, P):
        t = nnz * p // P
        b[p] = min(max(int(np.searchsorted(rp, t, side="left")), b[p - 1]), m)
    return b


def reduce_time_units(dist, ms, units):
    """Timing contract of bench.py: max over ranks of the device time, sum over ranks of the processed units."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(ms)], dtype=torch.float64, device=dev)
    u = torch.tensor([float(units)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
