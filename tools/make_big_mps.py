"""Writes a large free-format MPS file from the synthetic generator (reader throughput test, cold-CLI comparison):
python tools/make_big_mps.py OUT.mps [m n nnz [kind]]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def write_mps(lp, out):
    """min c'x, AL <= Ax <= AU, l <= x <= u of the synthetic generator (rows: equality, <=, or ranged; l = 0, u in {1, inf})."""
    m, n = lp["m"], lp["n"]
    rp, ci, v = lp["rowPtr"], lp["colIndex"], lp["values"]
    rows = np.repeat(np.arange(m), np.diff(rp))
    order = np.argsort(ci, kind="stable")
    cs, rs, vs = ci[order], rows[order], v[order]
    start = np.searchsorted(cs, np.arange(n + 1))
    AL, AU, c, l, u = lp["AL"], lp["AU"], lp["c"], lp["l"], lp["u"]
    with open(out, "w") as f:
        f.write("NAME big\nROWS\n N obj\n")
        fin_l, fin_u = np.isfinite(AL), np.isfinite(AU)
        # E: equality; L: AU finite (a RANGES entry below turns it into [AL, AU]); G: only AL finite
        types = np.where(fin_l & fin_u & (AL == AU), "E", np.where(fin_u, "L", "G"))
        f.write("".join(f" {types[i]} r{i}\n" for i in range(m)))
        f.write("COLUMNS\n")
        for j0 in range(0, n, 50000):
            lines = []
            for j in range(j0, min(n, j0 + 50000)):
                if c[j] != 0:
                    lines.append(f" x{j} obj {c[j]:.17g}\n")
                lines.extend(f" x{j} r{rs[k]} {vs[k]:.17g}\n" for k in range(start[j], start[j + 1]))
            f.write("".join(lines))
        f.write("RHS\n")
        f.write("".join(f" rhs r{i} {(AU[i] if fin_u[i] else AL[i]):.17g}\n" for i in range(m)))
        f.write("RANGES\n")
        f.write("".join(f" rng r{i} {AU[i] - AL[i]:.17g}\n" for i in range(m) if fin_l[i] and fin_u[i] and AL[i] != AU[i]))
        f.write("BOUNDS\n")
        f.write("".join(f" UP bnd x{j} {u[j]:.17g}\n" for j in range(n) if np.isfinite(u[j])))
        f.write("ENDATA\n")


if __name__ == "__main__":
    import __graft_entry__ as graft
    pkg = graft.load_package()
    m, n, nnz = (int(a) for a in sys.argv[2:5]) if len(sys.argv) >= 5 else (20000, 100000, 2000000)
    kind = sys.argv[5] if len(sys.argv) >= 6 else "uniform"
    write_mps(pkg.synth_lp(kind, m, n, nnz), sys.argv[1])
