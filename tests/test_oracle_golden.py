"""The CPU restatement against the committed outputs of the reference itself
(tests/golden/*.npz, written by tests/golden/make_golden.py from oracle/_ref).
Bit-for-bit.  Runs everywhere (no GPU, no /root/reference, no oracle/_ref)."""
from pathlib import Path

import numpy as np
import pytest

import cases
from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["cavity20x20", "box9x7x5_dirichlet", "asym10"])
def test_operators_preconditioners_smoothers(name):
    g = np.load(GOLD / f"ops_{name}.npz")
    s = cases.system(name)
    w = O.World([s])
    x = g["x"]
    assert np.array_equal(w.amul(x)[0], g["amul"])
    assert np.array_equal(w.tmul(x)[0], g["tmul"])
    assert np.array_equal(w.sumA()[0], g["sumA"])
    assert np.array_equal(w.H(x)[0], g["H"])
    assert np.array_equal(w.H1()[0], g["H1"])
    assert np.array_equal(w.faceH(x)[0], g["faceH"])
    assert np.array_equal(w.residual(x, s["source"])[0], g["residual"])
    checked = 0
    for key in g.files:
        if key.startswith("pre_"):
            assert np.array_equal(w.precondition(key[4:], s["source"])[0], g[key]), key
            checked += 1
        elif key.startswith("preT_"):
            assert np.array_equal(w.precondition(key[5:], s["source"], True)[0], g[key]), key
            checked += 1
        elif key.startswith("smooth_"):
            assert np.array_equal(w.smooth(key[7:], x, s["source"], 2)[0], g[key]), key
            checked += 1
    assert checked >= 8


@pytest.mark.parametrize("case", range(len(cases.SOLVES) + len(cases.GAMG_SOLVES)))
def test_solves(case):
    g = np.load(GOLD / "solves.npz")
    name, ctl = (cases.SOLVES + cases.GAMG_SOLVES)[case]
    s = cases.system(name)
    psi, perf = O.World([s]).solve(ctl, s["psi0"], s["source"])
    ref = g[f"perf_{case}"]
    assert perf["initialResidual"] == ref[0]
    assert perf["finalResidual"] == ref[1]
    assert perf["nIterations"] == int(ref[2])
    assert perf["converged"] == bool(ref[3]) and perf["singular"] == bool(ref[4])
    assert np.array_equal(psi[0], g[f"psi_{case}"])


def test_agglomeration():
    g = np.load(GOLD / "agglomeration.npz")
    for name, merge, weights in [("cavity20x20", 1, False), ("box12_var", 2, True), ("asym10", 1, False)]:
        s = cases.system(name)
        ctl = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=10, mergeLevels=merge,
                   agglomerator="faceAreaPair" if weights else "algebraicPair")
        levels = O.World([s]).gamg_levels(ctl)
        keys = [k for k in g.files if k.startswith(f"{name}_m{merge}_l")]
        assert len(levels) == len(keys) > 0
        for lev, L in enumerate(levels):
            assert np.array_equal(L["restrict"], g[f"{name}_m{merge}_l{lev}"])


def test_do_while_off_by_one():
    """PCG runs maxIter+1 iterations, GAMG maxIter (PCG.C:174-178 vs GAMGSolverSolve.C:109-113)."""
    s = cases.system("box12_var")
    _, p = O.World([s]).solve(dict(solver="PCG", preconditioner="DIC", tolerance=0, relTol=0, maxIter=10),
                              s["psi0"], s["source"])
    assert p["nIterations"] == 11 and not p["converged"]
    _, p = O.World([s]).solve(dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair",
                                   nCellsInCoarsestLevel=10, mergeLevels=1, tolerance=0, relTol=0, maxIter=3),
                              s["psi0"], s["source"])
    assert p["nIterations"] == 3


# --- multi-region: outputs of the reference run as one process per region (coupled through
# --- oracle/pstream_shm), stored gathered into the global cell order
@pytest.mark.parametrize("case", range(len(cases.MULTI_REGION_SOLVES)))
def test_multi_region_solves(case):
    from ldub200 import decompose
    g = np.load(GOLD / "multi_region.npz")
    name, R, part, ctl = cases.MULTI_REGION_SOLVES[case]
    s, regs = cases.regions(name, R, part)
    psi, perf = O.World(regs).solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    ref = g[f"perf_{case}"]
    assert perf["initialResidual"] == ref[0]
    assert perf["finalResidual"] == ref[1]
    assert perf["nIterations"] == int(ref[2])
    assert perf["converged"] == bool(ref[3]) and perf["singular"] == bool(ref[4])
    assert np.array_equal(decompose.gather_field(regs, psi, s["nCells"]), g[f"psi_{case}"])


@pytest.mark.parametrize("name,R,part", [("asym4x35x13", 4, "slab"), ("box12_var", 3, "random")])
def test_multi_region_operators_and_smoothers(name, R, part):
    from ldub200 import decompose
    g = np.load(GOLD / "multi_region.npz")
    s, regs = cases.regions(name, R, part)
    key = f"{name}_{R}_{part}"
    x = g[key + "_x"]
    xs = [x[r["cells"]] for r in regs]
    src = [r["source"] for r in regs]
    w = O.World(regs)

    def gathered(fields):
        return decompose.gather_field(regs, fields, s["nCells"])

    assert np.array_equal(gathered(w.amul(xs)), g[key + "_amul"])
    assert np.array_equal(gathered(w.tmul(xs)), g[key + "_tmul"])
    assert np.array_equal(gathered(w.sumA()), g[key + "_suma"])
    assert np.array_equal(gathered(w.residual(xs, src)), g[key + "_residual"])
    checked = 0
    for k in g.files:
        if k.startswith(key + "_smooth_"):
            sm = k[len(key) + 8:]
            assert np.array_equal(gathered(w.smooth(sm, xs, src, 2)), g[k]), sm
            checked += 1
    assert checked >= 5


# --- cyclic (periodic) patches
@pytest.mark.parametrize("name,axis", cases.CYCLIC_SYSTEMS[:3])
def test_cyclic_operators_and_smoothers(name, axis):
    g = np.load(GOLD / "cyclic.npz")
    s = cases.cyclic_system(name, axis)
    key = f"{name}_{axis}"
    x = g[key + "_x"]
    w = O.World([s])
    assert np.array_equal(w.amul(x)[0], g[key + "_amul"])
    assert np.array_equal(w.tmul(x)[0], g[key + "_tmul"])
    assert np.array_equal(w.sumA()[0], g[key + "_suma"])
    assert np.array_equal(w.residual(x, s["source"])[0], g[key + "_residual"])
    checked = 0
    for k in g.files:
        if k.startswith(key + "_smooth_"):
            sm = k[len(key) + 8:]
            assert np.array_equal(w.smooth(sm, x, s["source"], 2)[0], g[k]), sm
            checked += 1
    assert checked >= 5


@pytest.mark.parametrize("case", range(len(cases.CYCLIC_SOLVES)))
def test_cyclic_solves(case):
    g = np.load(GOLD / "cyclic.npz")
    name, axis, ctl = cases.CYCLIC_SOLVES[case]
    s = cases.cyclic_system(name, axis)
    psi, perf = O.World([s]).solve(ctl, s["psi0"], s["source"])
    ref = g[f"perf_{case}"]
    assert perf["initialResidual"] == ref[0]
    assert perf["finalResidual"] == ref[1]
    assert perf["nIterations"] == int(ref[2])
    assert np.array_equal(psi[0], g[f"psi_{case}"])


# --- edge cases of the solver front end (diagonal / faceless matrices, maxIter 0/1, initial guesses,
# --- zero sources, tolerance corner cases)
@pytest.mark.parametrize("case", range(len(cases.EDGE_SOLVES)))
def test_edge_cases(case):
    g = np.load(GOLD / "edge_cases.npz")
    s, ctl, psi0, source = cases.edge_case(case)
    psi, perf = O.World([s]).solve(ctl, psi0.copy(), source)
    ref = g[f"perf_{case}"]
    assert perf["initialResidual"] == ref[0] and perf["finalResidual"] == ref[1]
    assert perf["nIterations"] == int(ref[2])
    assert perf["converged"] == bool(ref[3]) and perf["singular"] == bool(ref[4])
    assert np.array_equal(psi[0], g[f"psi_{case}"])


@pytest.mark.parametrize("case", range(len(cases.SINGULAR_SOLVES)))
def test_singular_matrix(case):
    """the reference's answer (oracle/_ref run recorded in test_oracle_vs_ref.py::test_singular_matrix):
    singular, not converged, 0 iterations, residuals 1, psi untouched"""
    s, ctl = cases.singular_case(case)
    psi, perf = O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])
    assert perf["singular"] and not perf["converged"] and perf["nIterations"] == 0
    assert perf["initialResidual"] == 1.0 and perf["finalResidual"] == 1.0
    assert np.array_equal(psi[0], s["psi0"])


@pytest.mark.parametrize("case", range(len(cases.CACHE_SOLVES)))
def test_cached_agglomeration_across_solves(case):
    g = np.load(GOLD / "cache_solves.npz")
    name, ctl = cases.CACHE_SOLVES[case]
    s = cases.system(name)
    w = O.World([s])
    _, perf1 = w.solve(ctl, s["psi0"].copy(), s["source"])
    w.set_coeffs(0, *O.second_coeffs(s))
    psi2, perf2 = w.solve(ctl, s["psi0"].copy(), s["source"])
    ref = g[f"perf_{case}"]
    assert (perf1["nIterations"], perf1["finalResidual"]) == (int(ref[0]), ref[1])
    assert (perf2["nIterations"], perf2["finalResidual"]) == (int(ref[2]), ref[3])
    assert np.array_equal(psi2[0], g[f"psi_{case}"])


@pytest.mark.parametrize("case", range(len(cases.GAMG_OPTION_SOLVES)))
def test_gamg_options(case):
    g = np.load(GOLD / "gamg_options.npz")
    name, ctl = cases.GAMG_OPTION_SOLVES[case]
    s = cases.system(name)
    psi, perf = O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])
    ref = g[f"perf_{case}"]
    assert perf["initialResidual"] == ref[0] and perf["finalResidual"] == ref[1]
    assert perf["nIterations"] == int(ref[2]) and perf["converged"] == bool(ref[3])
    assert np.array_equal(cases.digest(psi[0]), g[f"sha_psi_{case}"])
