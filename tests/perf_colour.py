"""Micro-benchmark (not a test): Gauss-Seidel smoothing and GAMG on the same box in lexicographic and in
colour-ordered numbering (ldu_colour_order).  usage: perf_colour.py n [sweeps]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes, renumber  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
base = meshes.laplacian_system(n, n, n)
t0 = time.perf_counter()
col = renumber.colour_order(base)
t_col = time.perf_counter() - t0
stream = torch.cuda.Stream()
ctx = ldub200.Context(0, stream.cuda_stream)
for label, s in (("lexicographic", base), (f"colour-ordered ({col['nColours']} colours)", col)):
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
    A.set_coeffs(s["diag"], s["upperCoef"])
    A.set_face_weights(s["faceWeights"])
    d_psi = ldub200.DeviceField(ctx, s["nCells"])
    d_src = ldub200.DeviceField(ctx, s["nCells"], s["source"])
    out = {}
    import os
    only = os.environ.get("PERF_ONLY")
    for name, ctl in (("smoothSolver GaussSeidel", dict(solver="smoothSolver", smoother="GaussSeidel", nSweeps=-sweeps)),
                      ("PCG+DIC", dict(solver="PCG", preconditioner="DIC", tolerance=0, relTol=0, maxIter=sweeps - 1)),
                      ("GAMG GaussSeidel", dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair",
                                                nCellsInCoarsestLevel=10, mergeLevels=1, cacheAgglomeration=True,
                                                nPreSweeps=0, nPostSweeps=2, nFinestSweeps=2, tolerance=0, relTol=0,
                                                maxIter=10))):
        if only and only not in name:
            continue
        solver = ldub200.lduMatrix.solver.New("p", A, ctl)
        for rep in range(3):
            d_psi.zero()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            perf = solver.solve_device(d_psi, d_src)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        units = sweeps if "GAMG" not in name else perf.nIterations
        out[name] = (dt / units * 1e3, perf.finalResidual)
    print(f"{n}^3 {label}: " + "; ".join(f"{k} {v[0]:.3f} ms per {'cycle' if 'GAMG' in k else 'sweep/iteration'} (res {v[1]:.2e})"
                                        for k, v in out.items()))
    A.destroy()
print(f"colouring + renumbering on the host: {t_col:.2f} s")
