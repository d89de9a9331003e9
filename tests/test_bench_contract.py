"""CPU: bench.py's own logic (argument handling, the JSON line and its contract keys, the reference-arm line) with the
engine replaced by a stand-in -- the numbers are meaningless, the shape of the line is what the driver depends on.
The real thing runs on the B200 box (`python bench.py`, see profiles/r1_bench_*.json)."""
import io
import json
import sys
from contextlib import redirect_stdout
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class _FakeCLib:
    def __init__(self):
        self.launches = 0

    def hprlp_b200_engine_create(self, model, param):
        return 1

    def hprlp_b200_engine_run(self, h, iters):
        self.launches += 2 * iters
        return 0.125 * iters          # ms

    def hprlp_b200_engine_info(self, h, info_ref):
        info = info_ref._obj
        info.kernel_launches = self.launches
        info.lanes_A, info.lanes_AT = 8, 1
        return 0

    def hprlp_b200_engine_time_phase(self, h, which, reps):
        return 0.05 + 0.01 * which    # ms

    def hprlp_b200_engine_destroy(self, h):
        return None

    def hprlp_b200_profiler_start(self):
        return None

    def hprlp_b200_profiler_stop(self):
        return None


class _FakeLib:
    def __init__(self):
        self.lib = _FakeCLib()

    def create_model(self, lp, is_csc=False):
        return object()

    def free_model(self, model):
        return None

    def solve(self, model, param, main=False):
        return dict(iter=520, time=0.25, status="OPTIMAL", primal_obj=-1.0, residuals=9e-5)

    def release_cached_memory(self):
        return None


REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "gpu_launches", "clocks", "roofline", "e2e", "cpu_baseline"}


def _run_bench(monkeypatch, argv):
    import __graft_entry__ as graft
    import bench
    pkg = graft.load_package()
    fake = _FakeLib()
    monkeypatch.setattr(pkg, "load_engine", lambda: fake)
    monkeypatch.setattr(pkg, "load_reference", lambda: fake)
    monkeypatch.setattr(sys, "argv", ["bench.py", *argv])
    buf = io.StringIO()
    with redirect_stdout(buf):
        rc = bench.main()
    assert rc == 0
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1, lines          # ONE JSON line
    return json.loads(lines[0])


def test_engine_arm_line_has_the_contract_keys(monkeypatch):
    out = _run_bench(monkeypatch, ["--workload", "small", "--steps", "2", "--warmup", "3"])
    assert REQUIRED <= set(out), sorted(REQUIRED - set(out))
    assert out["unit"] == "HPR iterations/s" and out["dtype"] == "f64" and out["data"] == "synthetic"
    assert out["higher_is_better"] is True and out["vs_baseline"] is None and out["n_gpus"] == 1 and out["steps"] == 2
    assert "workload" in out["config"] and "model" not in out["config"]
    r = out["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = out["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0
    c = out["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["cores"] >= 1
    assert out["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(out["clocks"])
    assert abs(out["value"] - 200 / (0.125 * 200 * 1e-3)) < 1e-6     # iterations / device seconds of the timed steps


def test_reference_arm_line(monkeypatch):
    import __graft_entry__ as graft
    pkg = graft.load_package()
    if not pkg.REF_LIB_PATH.exists():
        pytest.skip("oracle/_ref not built")
    out = _run_bench(monkeypatch, ["--workload", "small", "--impl", "reference", "--steps", "2"])
    assert out["impl"] == "reference" and out["unit"] == "HPR iterations/s"
    assert out["cpu_baseline"]["kind"] == "reference" and out["cpu_baseline"]["value"] == out["value"]
    assert out["e2e"]["h2d_bytes_per_step"] == 0 and out["e2e"]["d2h_bytes_per_step"] == 0
