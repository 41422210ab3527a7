"""`smoother multiColourGaussSeidel` (north star: "a multi-colour Gauss-Seidel smoother").

Not a smoother of the reference: its Gauss-Seidel visited colour by colour of a greedy colouring -- every level of a
GAMG hierarchy with its own colouring -- all cells of a colour in parallel.  Pinned two ways:

  * CPU: the oracle's restatement equals the REFERENCE's lexicographic GaussSeidel run on the mesh renumbered by
    colour (ldub200.renumber.colour_order, what renumberMesh would hand over): same iterates up to the rounding of
    the row sums (the renumbered mesh lists a row's faces in another order), on structured, asymmetric and
    scrambled systems.
  * GPU: the CUDA smoother is bit-identical to the oracle's restatement: as a smoother, inside smoothSolver, on
    every level of GAMG (reference-order sums), on coupled regions.
and the price in GAMG iterations against the reference's lexicographic smoother is measured and stated."""
import numpy as np
import pytest

import cases
from ldub200 import renumber
from oracle import oracle as O

NAMES = ["cavity20x20", "box12_var", "asym10", "scrambled9", "box7x41x3", "line50", "single"]


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", NAMES)
def test_oracle_is_the_references_gauss_seidel_on_the_colour_ordered_mesh(name):
    s = cases.system(name)
    p = renumber.colour_order(s)
    x0 = np.sin(0.3 * np.arange(s["nCells"]))
    mine = O.World([s]).smooth("multiColourGaussSeidel", x0, s["source"], 3)[0]
    ref, _ = O.ref_run(p, "smooth", O.dict_text(dict(smoother="GaussSeidel")), 3, psi=x0[np.argsort(p["perm"])],
                       source=p["source"])
    back = ref[p["perm"]]          # value of old cell c sits at new index perm[c]
    scale = np.abs(back).max() + 1e-300
    assert np.abs(mine - back).max() <= 1e-13 * scale


@pytest.mark.parametrize("name,n_regions", [("box12_var", 1), ("box12_var", 3), ("asym10", 2)])
def test_gamg_iteration_price_is_small(name, n_regions):
    """GAMG with the multi-colour smoother against GAMG with the reference's lexicographic one (oracle, CPU):
    the same hierarchy, a few per cent more or fewer cycles"""
    s, regs = cases.regions(name, n_regions, "slab") if n_regions > 1 else (cases.system(name), None)
    regs = regs or [s]
    ctl = dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair", nCellsInCoarsestLevel=10,
               mergeLevels=1, tolerance=1e-8, relTol=0)
    w = O.World(regs)
    _, lex = w.solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    _, mc = w.solve(dict(ctl, smoother="multiColourGaussSeidel"), [r["psi0"] for r in regs],
                    [r["source"] for r in regs])
    assert mc["converged"] and lex["converged"]
    assert abs(mc["nIterations"] - lex["nIterations"]) <= max(2, lex["nIterations"] // 3), (mc, lex)


# --------------------------------------------------------------------------------------- GPU
def _matrix(ctx, s):
    import ldub200
    its = s.get("interfaces") or []
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in its]
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"], ifs)
    A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"], [it["bouCoeffs"] for it in its],
                 [it["intCoeffs"] for it in its])
    if s.get("faceWeights") is not None:
        A.set_face_weights(s["faceWeights"])
    return A


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.SYSTEMS))
def test_gpu_smoother_bit_exact(ctx, name):
    import ldub200
    s = cases.system(name)
    x0 = np.cos(0.2 * np.arange(s["nCells"]))
    want = O.World([s]).smooth("multiColourGaussSeidel", x0, s["source"], 3)[0]
    A = _matrix(ctx, s)
    psi = x0.copy()
    ldub200.lduMatrix.smoother.New("p", A, "multiColourGaussSeidel").smooth(psi, s["source"], 3)
    assert np.array_equal(psi, want)
    A.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("name,ctl", [
    ("box12_var", dict(solver="smoothSolver", smoother="multiColourGaussSeidel", nSweeps=2, tolerance=1e-6, relTol=0)),
    ("asym10", dict(solver="smoothSolver", smoother="multiColourGaussSeidel", nSweeps=1, tolerance=1e-7, relTol=0)),
    ("box12_var", dict(solver="GAMG", smoother="multiColourGaussSeidel", agglomerator="faceAreaPair",
                       nCellsInCoarsestLevel=10, mergeLevels=1, tolerance=1e-8, relTol=0)),
    ("scrambled17", dict(solver="GAMG", smoother="multiColourGaussSeidel", agglomerator="algebraicPair",
                         nCellsInCoarsestLevel=10, mergeLevels=2, tolerance=1e-8, relTol=0, nPreSweeps=1)),
    ("box40x30x20", dict(solver="PCG", tolerance=1e-9, relTol=0,
                         preconditioner=dict(preconditioner="GAMG", smoother="multiColourGaussSeidel",
                                             agglomerator="faceAreaPair", nCellsInCoarsestLevel=10, mergeLevels=1,
                                             tolerance=1e-5, relTol=0, nVcycles=2))),
])
def test_gpu_solves_bit_exact(ctx, name, ctl):
    import ldub200
    s = cases.system(name)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
    A = _matrix(ctx, s)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"] and perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(psi, psi_o[0])
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])     # default sums
    assert perf.nIterations == perf_o["nIterations"]
    A.destroy()


@pytest.mark.gpu
def test_gpu_cyclic_region_bit_exact(ctx):
    """interfaces enter through bPrime exactly as in GaussSeidel (a cyclic pair: one region, two interfaces)"""
    import ldub200
    s = cases.cyclic_system("box12_var", 0)
    x0 = np.cos(0.2 * np.arange(s["nCells"]))
    want = O.World([s]).smooth("multiColourGaussSeidel", x0, s["source"], 2)[0]
    A = _matrix(ctx, s)
    psi = x0.copy()
    ldub200.lduMatrix.smoother.New("p", A, "multiColourGaussSeidel").smooth(psi, s["source"], 2)
    assert np.array_equal(psi, want)
    A.destroy()
