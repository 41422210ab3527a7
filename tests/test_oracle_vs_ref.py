"""Pin the CPU restatement (oracle/ldu_oracle.c) against the UNMODIFIED reference
compiled from /root/reference (oracle/_ref, built by oracle/build_ref.py).
Bit-for-bit: same psi, same residuals, same iteration counts.
Skipped where oracle/_ref is absent (it travels to the GPU box prebuilt)."""
import numpy as np
import pytest

import cases
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("name", list(cases.SYSTEMS))
def test_operators(name):
    s = cases.system(name)
    w = O.World([s])
    x = np.random.default_rng(5).standard_normal(s["nCells"])
    assert np.array_equal(w.amul(x)[0], O.ref_run(s, "amul", psi=x)[0])
    assert np.array_equal(w.tmul(x)[0], O.ref_run(s, "tmul", psi=x)[0])
    assert np.array_equal(w.sumA()[0], O.ref_run(s, "suma")[0])
    assert np.array_equal(w.H(x)[0], O.ref_run(s, "H", psi=x)[0])
    assert np.array_equal(w.H1()[0], O.ref_run(s, "H1")[0])
    assert np.array_equal(w.faceH(x)[0], O.ref_run(s, "faceH", psi=x)[0])
    assert np.array_equal(w.residual(x, s["source"])[0], O.ref_run(s, "residual", psi=x)[0])


@pytest.mark.parametrize("name", ["cavity20x20", "box12_var", "asym10"])
@pytest.mark.parametrize("pre", cases.PRECONDITIONERS)
def test_preconditioners(name, pre):
    s = cases.system(name)
    if not cases.selectable(s, pre):
        pytest.skip("not in the reference's table for this matrix type")
    w = O.World([s])
    assert np.array_equal(w.precondition(pre, s["source"])[0], O.ref_run(s, "precondition", pre)[0])
    if pre == "DILU":
        assert np.array_equal(w.precondition(pre, s["source"], True)[0],
                              O.ref_run(s, "preconditionT", pre)[0])


@pytest.mark.parametrize("name", ["cavity20x20", "box9x7x5_dirichlet", "asym10"])
@pytest.mark.parametrize("sm", cases.SMOOTHERS)
def test_smoothers(name, sm):
    s = cases.system(name)
    if not cases.selectable(s, sm):
        pytest.skip("not in the reference's table for this matrix type")
    w = O.World([s])
    psi0 = np.random.default_rng(2).standard_normal(s["nCells"])
    want, _ = O.ref_run(s, "smooth", O.dict_text(dict(smoother=sm)), 3, psi=psi0)
    assert np.array_equal(w.smooth(sm, psi0, s["source"], 3)[0], want)


@pytest.mark.parametrize("case", range(len(cases.SOLVES) + len(cases.GAMG_SOLVES)))
def test_solves(case):
    name, ctl = (cases.SOLVES + cases.GAMG_SOLVES)[case]
    s = cases.system(name)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
    psi_r, perf_r = O.ref_solve(s, cases.ref_controls(ctl))
    assert perf_o["nIterations"] == perf_r["nIterations"]
    assert perf_o["initialResidual"] == perf_r["initialResidual"]
    assert perf_o["finalResidual"] == perf_r["finalResidual"]
    assert perf_o["converged"] == perf_r["converged"]
    assert np.array_equal(psi_o[0], psi_r)


@pytest.mark.parametrize("name,merge,weights", [("cavity20x20", 1, False), ("box12_var", 1, True),
                                                ("box12_var", 2, True), ("box9x7x5_dirichlet", 3, True),
                                                ("asym10", 1, False)])
def test_agglomeration(name, merge, weights):
    s = cases.system(name)
    ctl = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=10, mergeLevels=merge,
               agglomerator="faceAreaPair" if weights else "algebraicPair")
    mine = O.World([s]).gamg_levels(ctl)
    ref = O.ref_agglom(s, cases.ref_controls(ctl))
    assert len(mine) == len(ref) and len(mine) > 0
    for a, b in zip(mine, ref):
        assert a["nCoarse"] == b["nCoarse"]
        assert np.array_equal(a["restrict"], b["restrict"])


# --- cyclic (periodic) patches: two interfaces of the one region pointing at each other; the driver
# --- gives the reference cyclicLduInterface(Field) objects of type "cyclic" (oracle/ref_driver.C)
@pytest.mark.parametrize("name,axis", cases.CYCLIC_SYSTEMS)
def test_cyclic_operators_and_smoothers(name, axis):
    s = cases.cyclic_system(name, axis)
    w = O.World([s])
    x = np.random.default_rng(5).standard_normal(s["nCells"])
    assert np.array_equal(w.amul(x)[0], O.ref_run(s, "amul", psi=x)[0])
    assert np.array_equal(w.tmul(x)[0], O.ref_run(s, "tmul", psi=x)[0])
    assert np.array_equal(w.sumA()[0], O.ref_run(s, "suma")[0])
    assert np.array_equal(w.residual(x, s["source"])[0], O.ref_run(s, "residual", psi=x)[0])
    for sm in cases.SMOOTHERS:
        if cases.selectable(s, sm):
            want, _ = O.ref_run(s, "smooth", O.dict_text(dict(smoother=sm)), 3, psi=x)
            assert np.array_equal(w.smooth(sm, x, s["source"], 3)[0], want), sm


@pytest.mark.parametrize("case", range(len(cases.CYCLIC_SOLVES)))
def test_cyclic_solves(case):
    name, axis, ctl = cases.CYCLIC_SOLVES[case]
    s = cases.cyclic_system(name, axis)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
    psi_r, perf_r = O.ref_solve(s, cases.ref_controls(ctl))
    assert perf_o["nIterations"] == perf_r["nIterations"]
    assert perf_o["initialResidual"] == perf_r["initialResidual"]
    assert perf_o["finalResidual"] == perf_r["finalResidual"]
    assert np.array_equal(psi_o[0], psi_r)


# --- edge cases of the solver front end: diagonal / faceless matrices, maxIter 0 and 1, converged or
# --- random initial guesses, zero sources, relTol-only and huge tolerances
@pytest.mark.parametrize("case", range(len(cases.EDGE_SOLVES)))
def test_edge_cases(case):
    s, ctl, psi0, source = cases.edge_case(case)
    psi_o, perf_o = O.World([s]).solve(ctl, psi0.copy(), source)
    psi_r, so = O.ref_run(s, "solve", O.dict_text(cases.ref_controls(ctl)), psi=psi0, source=source)
    perf_r = O.parse_perf(so)
    for key in ("initialResidual", "finalResidual", "nIterations", "converged", "singular"):
        assert perf_o[key] == perf_r[key], (key, perf_o, perf_r)
    assert np.array_equal(psi_o[0], psi_r)
    if cases.EDGE_SOLVES[case][0] == "diagonal6":
        assert perf_r["solverName"] == "diagonal"


def test_gamg_without_coarse_levels_is_an_error():
    """GAMGSolver.C:108-126: the reference stops with "No coarse levels created"; so does the oracle"""
    s = cases.system("line50")
    ctl = dict(solver="GAMG", smoother="GaussSeidel", agglomerator="algebraicPair", nCellsInCoarsestLevel=100,
               mergeLevels=1, tolerance=1e-8, relTol=0)
    with pytest.raises(AssertionError):
        O.World([s]).solve(ctl, s["psi0"], s["source"])
    with pytest.raises(RuntimeError, match="No coarse levels created"):
        O.ref_run(s, "solve", O.dict_text(ctl))


@pytest.mark.parametrize("name,ctl", [
    ("box12_var", dict(solver="ICCG", preconditioner="diagonal", tolerance=1e-8, relTol=0)),
    ("box12_var", dict(solver="ICCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("asym10", dict(solver="BICCG", preconditioner="diagonal", tolerance=1e-8, relTol=0)),
])
def test_iccg_biccg_are_aliases(name, ctl):
    """selected at run time, ICCG / BICCG pass the dictionary on to PCG / PBiCG (ICCG.C:67-86): the
    dictionary's own preconditioner is used, and the performance line says <preconditioner>PCG"""
    s = cases.system(name)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
    psi_r, perf_r = O.ref_solve(s, ctl)
    assert perf_r["solverName"] == ctl["preconditioner"] + ("PCG" if ctl["solver"] == "ICCG" else "PBiCG")
    assert perf_o["nIterations"] == perf_r["nIterations"] and perf_o["finalResidual"] == perf_r["finalResidual"]
    assert np.array_equal(psi_o[0], psi_r)
    with pytest.raises(RuntimeError, match="keyword preconditioner is undefined"):
        O.ref_solve(s, {k: v for k, v in ctl.items() if k != "preconditioner"})


@pytest.mark.parametrize("case", range(len(cases.SINGULAR_SOLVES)))
def test_singular_matrix(case):
    s, ctl = cases.singular_case(case)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])
    psi_r, perf_r = O.ref_solve(s, ctl)
    assert perf_r["singular"] and not perf_r["converged"] and perf_r["nIterations"] == 0
    for key in ("initialResidual", "finalResidual", "nIterations", "converged", "singular"):
        assert perf_o[key] == perf_r[key], key
    assert np.array_equal(psi_o[0], psi_r)


@pytest.mark.parametrize("case", range(len(cases.CACHE_SOLVES)))
def test_cached_agglomeration_across_solves(case):
    """two solves on the same mesh with changed coefficients in between (driver op solve2)"""
    name, ctl = cases.CACHE_SOLVES[case]
    s = cases.system(name)
    w = O.World([s])
    _, perf1 = w.solve(ctl, s["psi0"].copy(), s["source"])
    w.set_coeffs(0, *O.second_coeffs(s))
    psi2, perf2 = w.solve(ctl, s["psi0"].copy(), s["source"])
    psi_r, so = O.ref_run(s, "solve2", O.dict_text(cases.ref_controls(ctl)))
    ref1, ref2 = O.parse_perfs(so)
    assert perf1["nIterations"] == ref1["nIterations"] and perf1["finalResidual"] == ref1["finalResidual"]
    assert perf2["nIterations"] == ref2["nIterations"] and perf2["finalResidual"] == ref2["finalResidual"]
    assert np.array_equal(psi2[0], psi_r)


def test_pbicg_with_gamg_preconditioner_is_an_error():
    """GAMGPreconditioner implements precondition() only; PBiCG needs preconditionT() and the reference
    stops with "Not implemented" (lduMatrix.H:492-505).  Oracle and library refuse the combination."""
    s = cases.system("asym10")
    ctl = dict(solver="PBiCG", tolerance=1e-9, relTol=0,
               preconditioner=dict(preconditioner="GAMG", smoother="GaussSeidel", agglomerator="algebraicPair",
                                   nCellsInCoarsestLevel=10, mergeLevels=1, cacheAgglomeration=False,
                                   tolerance=1e-5, relTol=0, nVcycles=2))
    with pytest.raises(AssertionError):
        O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])
    with pytest.raises(RuntimeError, match="Not implemented"):
        O.ref_solve(s, ctl)


@pytest.mark.parametrize("case", range(len(cases.GAMG_OPTION_SOLVES)))
def test_gamg_options(case):
    name, ctl = cases.GAMG_OPTION_SOLVES[case]
    s = cases.system(name)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])
    psi_r, perf_r = O.ref_solve(s, cases.ref_controls(ctl))
    for key in ("initialResidual", "finalResidual", "nIterations", "converged"):
        assert perf_o[key] == perf_r[key], key
    assert np.array_equal(psi_o[0], psi_r)
