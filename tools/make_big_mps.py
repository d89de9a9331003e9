"""Writes a large free-format MPS file from the synthetic uniform generator (reader throughput test):
python tools/make_big_mps.py OUT.mps [m n nnz]"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft
pkg = graft.load_package()
out = sys.argv[1]
m, n, nnz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (20000, 100000, 2000000)
lp = pkg.synth_lp("uniform", m, n, nnz)
rp, ci, v = lp["rowPtr"], lp["colIndex"], lp["values"]
rows = np.repeat(np.arange(m), np.diff(rp))
order = np.argsort(ci, kind="stable")
cs, rs, vs = ci[order], rows[order], v[order]
start = np.searchsorted(cs, np.arange(n + 1))
AL, AU, c, l, u = lp["AL"], lp["AU"], lp["c"], lp["l"], lp["u"]
with open(out, "w") as f:
    f.write("NAME big\nROWS\n N obj\n")
    for i in range(m):
        t = ("E" if AL[i] == AU[i] else "G") if np.isfinite(AL[i]) and np.isfinite(AU[i]) else ("L" if np.isfinite(AU[i]) else "G")
        f.write(f" {t} r{i}\n")
    f.write("COLUMNS\n")
    lines = []
    for j in range(n):
        if c[j] != 0:
            lines.append(f" x{j} obj {c[j]:.17g}\n")
        lines.extend(f" x{j} r{rs[k]} {vs[k]:.17g}\n" for k in range(start[j], start[j + 1]))
    f.write("".join(lines))
    f.write("RHS\n")
    for i in range(m):
        f.write(f" rhs r{i} {(AU[i] if np.isfinite(AU[i]) else AL[i]):.17g}\n")
    f.write("RANGES\n")
    for i in range(m):
        if np.isfinite(AL[i]) and np.isfinite(AU[i]) and AL[i] != AU[i]:
            f.write(f" rng r{i} {AU[i] - AL[i]:.17g}\n")
    f.write("BOUNDS\n")
    for j in range(n):
        if np.isfinite(u[j]):
            f.write(f" UP bnd x{j} {u[j]:.17g}\n")
    f.write("ENDATA\n")
