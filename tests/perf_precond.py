"""Micro-benchmark (not a test): repeated DIC applications on an nx*ny*nz box, for ncu.
usage: perf_precond.py nx ny nz [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import numpy as np  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
s = meshes.laplacian_system(a[0], a[1], a[2])
ctx = ldub200.Context(0)
A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
A.set_coeffs(s["diag"], s["upperCoef"])
P = ldub200.lduMatrix.preconditioner.New(A, "DIC")
for _ in range(a[3] if len(a) > 3 else 3):
    w = P.precondition(s["source"])
print("ok", float(np.abs(w).sum()))
