"""CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star):
  * Amul / Tmul / sumA / residual, every preconditioner and smoother apply,
    GAMG coarse matrices: BIT-EXACT (same per-row operation order, no FMA).
  * Krylov / GAMG solves: identical iteration counts; final residual within
    REL_TOL_RESIDUAL of the oracle's.  The only arithmetic that differs is the
    summation ORDER of the global dot products (sequential on the CPU, a fixed
    tree on the GPU), which perturbs alpha/beta at the 1e-16 level.
"""
import numpy as np
import pytest

from ldub200 import meshes

pytestmark = pytest.mark.gpu

# north_star bar: 1e-12 relative on the PCG final residual.  With referenceOrderSums the
# CUDA path meets it with margin (difference exactly 0).  With the default parallel-tree
# sums the scalars alpha/beta differ from the reference's in the last bit and CG amplifies
# that by roughly u/finalResidual, so the fast mode is held to these looser bounds
# (measured deviations are tabulated in DESIGN.md; residuals are normalised so that the
# initial one is O(1), hence the absolute floor of 1e-13 used beside these):
REL_TOL_RESIDUAL = 1e-12
REL_TOL_RESIDUAL_LONG = 1e-4


def _oracle():
    from oracle import oracle as O
    return O


def _matrix(ctx, s):
    import ldub200
    its = s.get("interfaces") or []
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in its]
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"], ifs)
    if its:
        A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"], [it["bouCoeffs"] for it in its],
                     [it["intCoeffs"] for it in its])
    else:
        A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
    if s.get("faceWeights") is not None:
        A.set_face_weights(s["faceWeights"])
    return A


import cases
from cases import SYSTEMS


@pytest.fixture(params=list(SYSTEMS))
def system(request):
    return request.param, cases.system(request.param)


def test_amul_family_bit_exact(ctx, system):
    name, s = system
    O = _oracle()
    w = O.World([s])
    A = _matrix(ctx, s)
    rng = np.random.default_rng(7)
    x = rng.standard_normal(s["nCells"])
    assert np.array_equal(A.Amul(x), w.amul(x)[0])
    assert np.array_equal(A.Tmul(x), w.tmul(x)[0])
    assert np.array_equal(A.sumA(), w.sumA()[0])
    assert np.array_equal(A.residual(x, s["source"]), w.residual(x, s["source"])[0])
    # the face-loop operators around the solve: H, H1, faceH
    assert np.array_equal(A.H(x), w.H(x)[0])
    assert np.array_equal(A.H1(), w.H1()[0])
    if s["nFaces"]:
        assert np.array_equal(A.faceH(x), w.faceH(x)[0])
    A.destroy()


@pytest.mark.parametrize("pre", ["none", "diagonal", "DIC", "FDIC", "DILU"])
def test_preconditioners_bit_exact(ctx, system, pre):
    import ldub200
    name, s = system
    if not cases.selectable(s, pre):
        pytest.skip("not in the reference's table for this matrix type")
    O = _oracle()
    w = O.World([s])
    A = _matrix(ctx, s)
    P = ldub200.lduMatrix.preconditioner.New(A, pre)
    assert np.array_equal(P.precondition(s["source"]), w.precondition(pre, s["source"])[0])
    if pre == "DILU":
        assert np.array_equal(P.preconditionT(s["source"]), w.precondition(pre, s["source"], True)[0])
    A.destroy()


@pytest.mark.parametrize("sm", ["GaussSeidel", "symGaussSeidel", "DIC", "DILU", "FDIC",
                                "DICGaussSeidel", "DILUGaussSeidel", "nonBlockingGaussSeidel"])
def test_smoothers_bit_exact(ctx, system, sm):
    import ldub200
    name, s = system
    if not cases.selectable(s, sm):
        pytest.skip("not in the reference's table for this matrix type")
    O = _oracle()
    w = O.World([s])
    A = _matrix(ctx, s)
    rng = np.random.default_rng(3)
    psi0 = rng.standard_normal(s["nCells"])
    want = w.smooth(sm, psi0, s["source"], 3)[0]
    psi = psi0.copy()
    ldub200.lduMatrix.smoother.New("p", A, sm).smooth(psi, s["source"], 3)
    assert np.array_equal(psi, want)
    A.destroy()


@pytest.mark.parametrize("W", ["4", "8", "12", "15", "16"])
@pytest.mark.parametrize("version", ["1", "2", "3", "4", "4u"])
def test_box_sweeps_all_stack_heights(ctx, W, version, monkeypatch):
    """Structured-box sweeps with every stack height (planes per CTA / per warp) and all three kernel
    generations: DIC and DILU applications stay bit-identical to the reference order.
    40 planes: 10 / 5 / 4 / 3 / 3 stacks, the last one ragged; 70 lines: 3 columns, ragged."""
    import ldub200
    if version == "1" and W != "4":
        pytest.skip("stack height only exists in the second and third generation")
    if (version == "2" and W == "12") or (version == "3" and W == "15"):
        pytest.skip("not a stack height of this generation")
    if version in ("4", "4u"):  # chained warps: W selects the number of warps per CTA (2 planes each);
        if W not in ("4", "8", "16"):                       # "4u": the planes of a warp one tick apart (unblocked)
            pytest.skip("generation 4 has 2, 4 or 8 warps of 2 planes")
        if version == "4u" and W != "16":
            pytest.skip("the unblocked layout exists for chains of 8 warps only")
        monkeypatch.setenv("LDU_STENCIL_M", str(int(W) // 2))
        monkeypatch.setenv("LDU_STENCIL_BLK", "1" if version == "4u" else "2")
    monkeypatch.setenv("LDU_STENCIL_W", W)
    monkeypatch.setenv("LDU_STENCIL", version[0])
    O = _oracle()
    for kw, pre in ((dict(nx=37, ny=70, nz=40, variable=True), "DIC"),
                    (dict(nx=3, ny=33, nz=40, variable=True, asym=0.3), "DILU")):
        s = meshes.laplacian_system(**kw)
        w = O.World([s])
        A = _matrix(ctx, s)
        P = ldub200.lduMatrix.preconditioner.New(A, pre)
        for rep in range(2):   # second application reuses rings, tickets and epochs
            assert np.array_equal(P.precondition(s["source"]), w.precondition(pre, s["source"])[0])
        if pre == "DILU":
            assert np.array_equal(P.preconditionT(s["source"]), w.precondition(pre, s["source"], True)[0])
        A.destroy()


def _rtol(controls):
    """short DIC/FDIC/smoother runs keep 1e-12; long unpreconditioned CG runs amplify the
    dot-product reordering"""
    if controls.get("solver") in ("smoothSolver",):
        return REL_TOL_RESIDUAL
    return REL_TOL_RESIDUAL_LONG


SOLVES = [(n, c, _rtol(c)) for n, c in cases.SOLVES]


@pytest.mark.parametrize("case", range(len(SOLVES)))
def test_solves_match_oracle(ctx, case):
    import ldub200
    sysname, controls, rtol = SOLVES[case]
    s = cases.system(sysname)
    O = _oracle()
    w = O.World([s])
    psi_o, perf_o = w.solve(controls, s["psi0"], s["source"], hist_cap=2048)
    A = _matrix(ctx, s)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, controls).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"], (str(perf), perf_o)
    assert perf.converged == perf_o["converged"] and perf.singular == perf_o["singular"]
    assert perf.initialResidual == pytest.approx(perf_o["initialResidual"], rel=1e-13, abs=1e-300)
    assert perf.finalResidual == pytest.approx(perf_o["finalResidual"], rel=rtol, abs=1e-13)
    scale = np.abs(psi_o[0]).max() + 1e-300
    assert np.abs(psi - psi_o[0]).max() / scale < 1e-7
    # same solve with the reductions accumulated in the reference's loop order:
    # everything, every iteration, is bit-identical
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(controls, referenceOrderSums=True)).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"]
    assert perf.initialResidual == perf_o["initialResidual"]
    assert perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(psi, psi_o[0])
    hist = A.residual_history()
    assert np.array_equal(hist, perf_o["history"][:len(hist)])
    A.destroy()


GAMG_CASES = [(n, c) for n, c in cases.GAMG_SOLVES if c["solver"] == "GAMG"]


@pytest.mark.parametrize("case", range(len(GAMG_CASES)))
def test_gamg_hierarchy_and_iterations(ctx, case):
    """north_star: iteration-count parity for GAMG; here also bit-exact coarse matrices."""
    import ldub200
    sysname, controls = GAMG_CASES[case]
    s = cases.system(sysname)
    O = _oracle()
    w = O.World([s])
    lev_o = w.gamg_levels(controls)
    A = _matrix(ctx, s)
    lev = A.gamg_levels(controls)
    assert len(lev) == len(lev_o)
    for a, b in zip(lev, lev_o):
        assert a["nCoarse"] == b["nCoarse"] and a["nFaces"] == b["nFaces"]
        assert np.array_equal(a["restrict"], b["restrict"])
        assert np.array_equal(a["diag"], b["diag"])
        assert np.array_equal(a["upperCoef"], b["upperCoef"])
    psi_o, perf_o = w.solve(controls, s["psi0"], s["source"])
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, controls).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"], (str(perf), perf_o)
    assert perf.finalResidual == pytest.approx(perf_o["finalResidual"], rel=1e-7)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(controls, referenceOrderSums=True)).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"]
    assert perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(psi, psi_o[0])
    A.destroy()


def test_pcg_gamg_preconditioner(ctx):
    import ldub200
    s = meshes.laplacian_system(**SYSTEMS["box12_var"])
    controls = dict(solver="PCG", tolerance=1e-9, relTol=0,
                    preconditioner=dict(preconditioner="GAMG", smoother="GaussSeidel",
                                        agglomerator="faceAreaPair", nCellsInCoarsestLevel=10,
                                        mergeLevels=1, tolerance=1e-5, relTol=0, nVcycles=2))
    O = _oracle()
    psi_o, perf_o = O.World([s]).solve(controls, s["psi0"], s["source"])
    A = _matrix(ctx, s)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, controls).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"]
    assert perf.finalResidual == pytest.approx(perf_o["finalResidual"], rel=1e-7)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(controls, referenceOrderSums=True)).solve(psi, s["source"])
    assert perf.finalResidual == perf_o["finalResidual"] and np.array_equal(psi, psi_o[0])
    A.destroy()


def test_solver_performance_print_format(ctx):
    import ldub200
    s = meshes.laplacian_system(20, 20, 1)
    A = _matrix(ctx, s)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(solver="PCG", preconditioner="DIC",
                                                     tolerance=1e-6, relTol=0)).solve(psi, s["source"])
    txt = str(perf)
    assert txt.startswith("DICPCG:  Solving for p, Initial residual = 1, Final residual = ")
    assert txt.endswith(f"No Iterations {perf.nIterations}")
    A.destroy()


def test_full_size_box_sweeps_match_generic_path(ctx, monkeypatch):
    """BASELINE's 10M-cell box (216^3): the register-stacked sweeps (27 stacks x 7 columns of warps), the
    plane-stacked sweeps (36 stacks x 7 columns of CTAs) and the generic dataflow sweeps are three independent
    implementations of the same DIC application and must agree bit for bit (the comparison with the reference
    itself at this size is tests/test_gpu_parity_fullsize.py)."""
    import ldub200
    n = 216
    s = meshes.laplacian_system(n, n, n, variable=True)
    r = np.sin(0.37 * np.arange(s["nCells"]))
    out = {}
    for ver in ("4", "4u", "3", "2", "0"):
        monkeypatch.setenv("LDU_STENCIL", ver[0])
        monkeypatch.setenv("LDU_STENCIL_BLK", "1" if ver == "4u" else "2")
        A = _matrix(ctx, s)
        P = ldub200.lduMatrix.preconditioner.New(A, "DIC")
        out[ver] = P.precondition(r)
        if ver != "0":
            again = P.precondition(r)      # rings, tickets and epochs are reused
            assert np.array_equal(again, out[ver])
        A.destroy()
    assert np.array_equal(out["4"], out["0"])     # chained register-stacked warps, two planes per tick (default)
    assert np.array_equal(out["4u"], out["0"])    # the same, planes one tick apart
    assert np.array_equal(out["3"], out["0"])     # register-stacked warps
    assert np.array_equal(out["2"], out["0"])     # plane-stacked CTAs
    assert np.isfinite(out["3"]).all() and np.abs(out["3"]).max() > 0


def test_large_box_properties(ctx):
    """Full-size class check without the oracle: linearity of Amul and a PCG
    residual that really is the residual (size-independent properties)."""
    import ldub200
    n = 96
    s = meshes.laplacian_system(n, n, n, variable=True)
    A = _matrix(ctx, s)
    rng = np.random.default_rng(11)
    x, y = rng.standard_normal(s["nCells"]), rng.standard_normal(s["nCells"])
    ax, ay, axy = A.Amul(x), A.Amul(y), A.Amul(x + y)
    assert np.abs(axy - (ax + ay)).max() <= 1e-12 * np.abs(axy).max()
    # symmetric matrix: <Ax, y> == <x, Ay>
    assert abs(ax @ y - x @ ay) <= 1e-10 * abs(ax @ y)
    psi = s["psi0"].copy()
    ctl = dict(solver="PCG", preconditioner="diagonal", tolerance=1e-5, relTol=0, maxIter=2000)
    perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])
    assert perf.converged
    res = A.residual(psi, s["source"])
    nf = 2 * np.abs(s["source"]).sum()
    assert np.abs(res).sum() / nf < 2e-5
    hist = A.residual_history()
    assert len(hist) == perf.nIterations + 1 and hist[-1] == perf.finalResidual
    A.destroy()


# --- cyclic (periodic) patches: interfaces whose neighbour is the region itself; a single process
# --- needs no ldu_comm_connect for them (the library is its own only peer)
@pytest.mark.parametrize("name,axis", cases.CYCLIC_SYSTEMS)
def test_cyclic_operators_and_smoothers_bit_exact(ctx, name, axis):
    import ldub200
    s = cases.cyclic_system(name, axis)
    O = _oracle()
    w = O.World([s])
    A = _matrix(ctx, s)
    x = np.random.default_rng(11).standard_normal(s["nCells"])
    assert np.array_equal(A.Amul(x), w.amul(x)[0])
    assert np.array_equal(A.Tmul(x), w.tmul(x)[0])
    assert np.array_equal(A.sumA(), w.sumA()[0])
    assert np.array_equal(A.residual(x, s["source"]), w.residual(x, s["source"])[0])
    for sm in cases.SMOOTHERS:
        if cases.selectable(s, sm):
            psi = x.copy()
            ldub200.lduMatrix.smoother.New("p", A, sm).smooth(psi, s["source"], 3)
            assert np.array_equal(psi, w.smooth(sm, x, s["source"], 3)[0]), sm
    A.destroy()


@pytest.mark.parametrize("case", range(len(cases.CYCLIC_SOLVES)))
def test_cyclic_solves_bit_exact(ctx, case):
    import ldub200
    name, axis, ctl = cases.CYCLIC_SOLVES[case]
    s = cases.cyclic_system(name, axis)
    O = _oracle()
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
    A = _matrix(ctx, s)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"], (str(perf), perf_o)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"]
    assert perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(psi, psi_o[0])
    A.destroy()


def test_generic_coarsest_solver_bit_exact(ctx, monkeypatch):
    """With interfaces or several ranks the coarsest GAMG level goes to the generic PBiCG+DILU /
    PCG+DIC solver instead of the single-thread kernel.  GAMG's coarse addressing does not sort the
    faces of a cell by neighbour, so DILU's preconditionT (reverse losort walk) needs its own row
    order there: forced here on a matrix without interfaces."""
    import ldub200
    monkeypatch.setenv("LDU_GAMG_GENERIC_COARSEST", "1")
    for name, sm in [("asym10", "GaussSeidel"), ("asym10", "DILU"), ("box12_var", "GaussSeidel")]:
        s = cases.system(name)
        ctl = dict(solver="GAMG", smoother=sm, agglomerator="algebraicPair", nCellsInCoarsestLevel=10,
                   mergeLevels=1, cacheAgglomeration=False, tolerance=1e-8, relTol=0)
        O = _oracle()
        psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
        A = _matrix(ctx, s)
        psi = s["psi0"].copy()
        perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, s["source"])
        assert perf.nIterations == perf_o["nIterations"]
        assert perf.finalResidual == perf_o["finalResidual"], (name, sm)
        assert np.array_equal(psi, psi_o[0]), (name, sm)
        A.destroy()


def test_airfoil_real_unstructured_mesh(ctx):
    """the polyMesh of the reference's airFoil2D tutorial (imported by ldub200.polymesh, committed as
    tests/golden/airfoil2d.npz): operators and solves against the REFERENCE's own committed results"""
    import ldub200
    s, g = cases.airfoil_system()
    A = _matrix(ctx, s)
    x = cases.airfoil_x(s["nCells"])
    assert np.array_equal(cases.digest(A.Amul(x)), g["sha_amul"])
    psi = x.copy()
    ldub200.lduMatrix.smoother.New("p", A, "GaussSeidel").smooth(psi, s["source"], 2)
    assert np.array_equal(cases.digest(psi), g["sha_smooth_GaussSeidel"])
    P = ldub200.lduMatrix.preconditioner.New(A, "DIC")
    assert np.array_equal(cases.digest(P.precondition(s["source"])), g["sha_pre_DIC"])
    for i, ctl in enumerate(cases.AIRFOIL_SOLVES):
        ref = g[f"perf_{i}"]
        psi = s["psi0"].copy()
        perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, s["source"])
        assert perf.nIterations == int(ref[2]), ctl
        assert perf.initialResidual == ref[0] and perf.finalResidual == ref[1], ctl
        assert np.array_equal(cases.digest(psi), g[f"sha_psi_{i}"]), ctl
        psi = s["psi0"].copy()
        perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])   # tree-ordered sums
        assert abs(perf.nIterations - int(ref[2])) <= 1, ctl
    A.destroy()


@pytest.mark.parametrize("case", range(len(cases.EDGE_SOLVES)))
def test_edge_cases_bit_exact(ctx, case):
    """solver front end corner cases against the REFERENCE's committed results (tests/golden/edge_cases.npz):
    diagonal() vs faceless matrices, maxIter 0 and 1, converged / random initial guesses, zero sources"""
    import ldub200
    g = np.load(cases.__file__.replace("cases.py", "golden/edge_cases.npz"))
    s, ctl, psi0, source = cases.edge_case(case)
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
    A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
    if s.get("faceWeights") is not None:
        A.set_face_weights(s["faceWeights"])
    psi = psi0.copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, source)
    ref = g[f"perf_{case}"]
    assert perf.nIterations == int(ref[2]), (str(perf), ref)
    assert perf.initialResidual == ref[0] and perf.finalResidual == ref[1], (str(perf), ref)
    assert perf.converged == bool(ref[3]) and perf.singular == bool(ref[4])
    assert np.array_equal(psi, g[f"psi_{case}"])
    A.destroy()


@pytest.mark.parametrize("case", range(len(cases.SINGULAR_SOLVES)))
def test_singular_matrix(ctx, case):
    """SolverPerformance::checkSingularity: an all-zero matrix -> singular, 0 iterations, psi untouched,
    and the reference's "solution singularity" print"""
    import ldub200
    s, ctl = cases.singular_case(case)
    O = _oracle()
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])
    A = _matrix(ctx, s)
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])
    assert perf.singular and not perf.converged and perf.nIterations == perf_o["nIterations"] == 0
    assert perf.initialResidual == perf_o["initialResidual"] and perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(psi, psi_o[0])
    assert str(perf).endswith("solution singularity")
    A.destroy()


@pytest.mark.parametrize("case", range(len(cases.CACHE_SOLVES)))
def test_cached_agglomeration_across_solves(ctx, case):
    """cacheAgglomeration: the second solve on the mesh (new coefficients) reuses the agglomeration of
    the first, as the reference's MeshObject does -- against the REFERENCE's committed results"""
    import ldub200
    g = np.load(cases.__file__.replace("cases.py", "golden/cache_solves.npz"))
    name, ctl = cases.CACHE_SOLVES[case]
    s = cases.system(name)
    O = _oracle()
    A = _matrix(ctx, s)
    exact = dict(ctl, referenceOrderSums=True)
    psi = s["psi0"].copy()
    perf1 = ldub200.lduMatrix.solver.New("p", A, exact).solve(psi, s["source"])
    d2, u2, l2 = O.second_coeffs(s)
    A.set_coeffs(d2, u2, l2)
    psi = s["psi0"].copy()
    perf2 = ldub200.lduMatrix.solver.New("p", A, exact).solve(psi, s["source"])
    ref = g[f"perf_{case}"]
    assert (perf1.nIterations, perf1.finalResidual) == (int(ref[0]), ref[1])
    assert (perf2.nIterations, perf2.finalResidual) == (int(ref[2]), ref[3])
    assert np.array_equal(psi, g[f"psi_{case}"])
    A.destroy()
