# measured deviation of the fast (parallel-tree) mode from the oracle
import sys
sys.path.insert(0,'tests'); sys.path.insert(0,'.'); sys.path.insert(0,'openfoam-2.2.x_b200')
import numpy as np, cases, ldub200
from oracle import oracle as O
ctx = ldub200.Context(0)
for name, ctl in cases.SOLVES + cases.GAMG_SOLVES:
    s = cases.system(name)
    psi_o, po = O.World([s]).solve(ctl, s["psi0"], s["source"])
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"]); A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
    if s.get("faceWeights") is not None: A.set_face_weights(s["faceWeights"])
    psi = s["psi0"].copy()
    try:
        p = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])
        rel = abs(p.finalResidual-po["finalResidual"])/max(po["finalResidual"],1e-300)
        print(f"{name:20s} {ctl['solver']:12s} {str(ctl.get('preconditioner', ctl.get('smoother')))[:14]:14s} it {p.nIterations:4d}/{po['nIterations']:4d} final {po['finalResidual']:.3e} reldiff {rel:.2e} psidiff {np.abs(psi-psi_o[0]).max()/ (np.abs(psi_o[0]).max()+1e-300):.2e}")
    except Exception as e:
        print(name, ctl['solver'], "ERROR", e)
    A.destroy()
