"""Micro-benchmark (not a test): host-pointer set_coeffs + solve on the 216^3 bench system with pageable arrays,
for tuning the staged copies.  usage: perf_e2e.py [n]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import numpy as np  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import decompose  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
reg = decompose.local_box_region(n, 0, 1)
ctx = ldub200.Context(0)
A = ldub200.lduMatrix(ctx, reg["nCells"], reg["lower"], reg["upper"])
ctl = dict(solver="PCG", preconditioner="DIC", tolerance=0.0, relTol=0.0, maxIter=49)
solver = ldub200.lduMatrix.solver.New("p", A, ctl)
diag, up, src = reg["diag"].copy(), reg["upperCoef"].copy(), reg["source"].copy()
for rep in range(4):
    psi = np.zeros(reg["nCells"])
    t0 = time.perf_counter()
    A.set_coeffs(diag, up)
    t1 = time.perf_counter()
    perf = solver.solve(psi, src)
    t2 = time.perf_counter()
print(f"set_coeffs {1e3*(t1-t0):.1f} ms ({(diag.nbytes+up.nbytes)/(t1-t0)/1e9:.1f} GB/s)  solve {1e3*(t2-t1):.1f} ms  "
      f"-> {50/(t2-t0):.1f} it/s")
