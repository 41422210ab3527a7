"""Micro-benchmark (not a test): GAMG V-cycles per second (GaussSeidel smoother, faceAreaPair
agglomeration, the motorBike / pitzDaily fvSolution settings) on an nx*ny*nz box.
usage: perf_gamg.py nx ny nz [cycles]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402

a = [int(x) for x in sys.argv[1:]]
nx, ny, nz = a[:3]
cycles = a[3] if len(a) > 3 else 10
s = meshes.laplacian_system(nx, ny, nz)
stream = torch.cuda.Stream()
ctx = ldub200.Context(0, stream.cuda_stream)
A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
A.set_coeffs(s["diag"], s["upperCoef"])
A.set_face_weights(s["faceWeights"])
d_psi = ldub200.DeviceField(ctx, s["nCells"])
d_src = ldub200.DeviceField(ctx, s["nCells"], s["source"])
ctl = dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair", nCellsInCoarsestLevel=10,
           mergeLevels=1, cacheAgglomeration=True, nPreSweeps=0, nPostSweeps=2, nFinestSweeps=2,
           tolerance=0, relTol=0, maxIter=cycles)
solver = ldub200.lduMatrix.solver.New("p", A, ctl)
t_first = None
for rep in range(3):
    d_psi.zero()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    perf = solver.solve_device(d_psi, d_src)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if t_first is None:
        t_first = dt
print(f"GAMG {nx}x{ny}x{nz}: cells {s['nCells']}  {perf.nIterations} cycles in {dt*1e3:.1f} ms -> "
      f"{perf.nIterations/dt:.1f} V-cycles/s (first solve incl. agglomeration {t_first*1e3:.0f} ms); "
      f"residual {perf.initialResidual:.3g} -> {perf.finalResidual:.3g}")
