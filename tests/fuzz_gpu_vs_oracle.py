"""Differential fuzzing of the CUDA path against the CPU restatement (needs a GPU; not collected by
pytest: `python tests/fuzz_gpu_vs_oracle.py [nCases] [seed]`).  Same generators as
tests/fuzz_oracle_vs_ref.py (random LDU graphs, coefficients, dictionaries, initial guesses, a random
cyclic pair every fourth system), one region per case; with referenceOrderSums every solve must be
bit-identical, operators / preconditioners / smoothers always."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "openfoam-2.2.x_b200"), str(ROOT / "tests")]

import cases  # noqa: E402
import fuzz_oracle_vs_ref as F  # noqa: E402
import ldub200  # noqa: E402
from oracle import oracle as O  # noqa: E402


def matrix(ctx, s):
    its = s.get("interfaces") or []
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in its]
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"], ifs)
    A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"], [it["bouCoeffs"] for it in its],
                 [it["intCoeffs"] for it in its])
    A.set_face_weights(s["faceWeights"])
    return A


def solve_case(ctx, seed):
    rng = np.random.default_rng(seed)
    s = F.random_system(rng)
    ctl = F.random_controls(rng, s, False)
    psi0 = rng.standard_normal(s["nCells"]) if rng.random() < 0.5 else np.zeros(s["nCells"])
    if s["nCells"] >= 6 and rng.random() < 0.25:
        F.add_random_cyclic(rng, s, 0)
    try:
        po, perf_o = O.World([s]).solve(ctl, psi0.copy(), s["source"])
    except AssertionError:
        po = None
    A = matrix(ctx, s)
    try:
        psi = psi0.copy()
        perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, s["source"])
    except ldub200.LduError as e:
        A.destroy()
        return None if po is None else f"seed {seed}: library refused ({e}) what the oracle solved: {ctl}"
    A.destroy()
    if po is None:
        return f"seed {seed}: library solved what the oracle (and the reference) refuse: {ctl}"
    if (perf.nIterations, perf.initialResidual, perf.finalResidual, perf.converged, perf.singular) != \
            (perf_o["nIterations"], perf_o["initialResidual"], perf_o["finalResidual"], perf_o["converged"],
             perf_o["singular"]):
        return f"seed {seed}: performance {perf} vs {perf_o}  n={s['nCells']} ctl={ctl}"
    if not np.array_equal(psi, po[0], equal_nan=True):
        return f"seed {seed}: psi differs (max {np.abs(psi - po[0]).max():.3e}) n={s['nCells']} ctl={ctl}"
    return None


def operator_case(ctx, seed):
    rng = np.random.default_rng(10_000_000 + seed)
    s = F.random_system(rng)
    if s["nCells"] >= 6 and rng.random() < 0.5:
        F.add_random_cyclic(rng, s, 0)
    w = O.World([s])
    A = matrix(ctx, s)
    x = rng.standard_normal(s["nCells"])
    src = s["source"]
    bad = None
    checks = [("Amul", A.Amul(x), w.amul(x)[0]), ("Tmul", A.Tmul(x), w.tmul(x)[0]), ("sumA", A.sumA(), w.sumA()[0]),
              ("residual", A.residual(x, src), w.residual(x, src)[0])]
    if not s.get("interfaces"):
        checks += [("H", A.H(x), w.H(x)[0]), ("H1", A.H1(), w.H1()[0])]
        if s["nFaces"]:
            checks.append(("faceH", A.faceH(x), w.faceH(x)[0]))
    for pre in cases.PRECONDITIONERS:
        if cases.selectable(s, pre):
            P = ldub200.lduMatrix.preconditioner.New(A, pre)
            checks.append(("precondition " + pre, P.precondition(src), w.precondition(pre, src)[0]))
            if pre == "DILU":
                checks.append(("preconditionT DILU", P.preconditionT(src), w.precondition(pre, src, True)[0]))
    nsw = int(rng.integers(1, 4))
    for sm in cases.SMOOTHERS:
        if cases.selectable(s, sm):
            psi = x.copy()
            ldub200.lduMatrix.smoother.New("p", A, sm).smooth(psi, src, nsw)
            checks.append((f"smoother {sm} x{nsw}", psi, w.smooth(sm, x, src, nsw)[0]))
    for name, a, b in checks:
        if not np.array_equal(a, b):
            bad = f"operator seed {seed}: {name} differs (max {np.abs(a - b).max():.3e}) n={s['nCells']}"
            break
    A.destroy()
    return bad


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    ctx = ldub200.Context(0)
    bad = 0
    for k in range(n):
        for msg in (solve_case(ctx, seed0 + k), operator_case(ctx, seed0 + k) if k % 4 == 0 else None):
            if msg:
                bad += 1
                print(msg, flush=True)
    print(f"{n} cases from seed {seed0}: {bad} differences")
