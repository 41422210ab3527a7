"""Profiling helper (not a test): ONE warm GAMG solve of a few V-cycles, to be run under
`ncu --metrics gpu__time_duration.sum`; argv: nx ny nz smoother cycles"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402

nx, ny, nz = (int(x) for x in sys.argv[1:4])
smoother = sys.argv[4]
cycles = int(sys.argv[5])
s = meshes.laplacian_system(nx, ny, nz)
ctx = ldub200.Context(0)
A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
A.set_coeffs(s["diag"], s["upperCoef"])
A.set_face_weights(s["faceWeights"])
d_psi = ldub200.DeviceField(ctx, s["nCells"])
d_src = ldub200.DeviceField(ctx, s["nCells"], s["source"])
ctl = dict(solver="GAMG", smoother=smoother, agglomerator="faceAreaPair", nCellsInCoarsestLevel=10, mergeLevels=1,
           cacheAgglomeration=True, nPreSweeps=0, nPostSweeps=2, nFinestSweeps=2, tolerance=1e-14, relTol=0, maxIter=cycles)
d_psi.zero()
print(ldub200.lduMatrix.solver.New("p", A, ctl).solve_device(d_psi, d_src))
