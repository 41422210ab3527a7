#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference by oracle/build_ref.py).  Run in the build container only:
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own for this path, so
these fixtures are outputs of the reference's own code on the systems of
tests/cases.py: they pin the CPU restatement wherever oracle/_ref is absent."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
sys.path.insert(0, str(HERE.parent))

import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    assert O.ref_available(), "build oracle/_ref first (python oracle/build_ref.py)"
    rng = np.random.default_rng(2024)
    for name in ("cavity20x20", "box9x7x5_dirichlet", "asym10"):
        s = cases.system(name)
        x = rng.standard_normal(s["nCells"])
        out = dict(x=x)
        out["amul"] = O.ref_run(s, "amul", psi=x)[0]
        out["tmul"] = O.ref_run(s, "tmul", psi=x)[0]
        out["sumA"] = O.ref_run(s, "suma")[0]
        out["residual"] = O.ref_run(s, "residual", psi=x)[0]
        out["H"] = O.ref_run(s, "H", psi=x)[0]
        out["H1"] = O.ref_run(s, "H1")[0]
        out["faceH"] = O.ref_run(s, "faceH", psi=x)[0]
        for pre in cases.PRECONDITIONERS:
            if cases.selectable(s, pre):
                out[f"pre_{pre}"] = O.ref_run(s, "precondition", pre)[0]
        if cases.selectable(s, "DILU"):
            out["preT_DILU"] = O.ref_run(s, "preconditionT", "DILU")[0]
        for sm in cases.SMOOTHERS:
            if cases.selectable(s, sm):
                out[f"smooth_{sm}"] = O.ref_run(s, "smooth", O.dict_text(dict(smoother=sm)), 2, psi=x)[0]
        np.savez_compressed(HERE / f"ops_{name}.npz", **out)
        print("ops", name, len(out))
    solves = {}
    for i, (name, ctl) in enumerate(cases.SOLVES + cases.GAMG_SOLVES):
        s = cases.system(name)
        psi, perf = O.ref_solve(s, cases.ref_controls(ctl))
        solves[f"psi_{i}"] = psi
        solves[f"perf_{i}"] = np.array([perf["initialResidual"], perf["finalResidual"],
                                        perf["nIterations"], perf["converged"], perf["singular"]], dtype=np.float64)
        print("solve", i, name, ctl["solver"], perf["nIterations"])
    np.savez_compressed(HERE / "solves.npz", **solves)
    agg = {}
    for name, merge, weights in [("cavity20x20", 1, False), ("box12_var", 2, True), ("asym10", 1, False)]:
        s = cases.system(name)
        ctl = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=10, mergeLevels=merge,
                   agglomerator="faceAreaPair" if weights else "algebraicPair")
        for lev, L in enumerate(O.ref_agglom(s, cases.ref_controls(ctl))):
            agg[f"{name}_m{merge}_l{lev}"] = L["restrict"].astype(np.int32)
    np.savez_compressed(HERE / "agglomeration.npz", **agg)
    print("agglomeration", len(agg))


if __name__ == "__main__":
    main()
