"""Debug helper (not a test): one GAMG solve on a small box with the library's own agglomeration against the oracle
(hierarchy sizes, iteration count, psi) -- a few seconds, for a last look after touching the host side of gamg.cu."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402
from oracle import oracle as O  # noqa: E402

s = meshes.laplacian_system(14, 12, 10, variable=True)
ctl = dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair", nCellsInCoarsestLevel=6,
           mergeLevels=1, tolerance=1e-9, relTol=0, referenceOrderSums=True)
w = O.World([s])
psi_o, perf_o = w.solve(ctl, s["psi0"], s["source"])
ctx = ldub200.Context(0)
A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
if "faceWeights" in s:
    A.set_face_weights(s["faceWeights"])
psi = s["psi0"].copy()
perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])
print("iterations", perf.nIterations, perf_o["nIterations"], "psi identical", bool(np.array_equal(psi, psi_o[0])),
      "final", perf.finalResidual, perf_o["finalResidual"])
A.destroy()
ctx.close()
