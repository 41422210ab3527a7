"""Colour-ordered renumbering: host logic on the CPU, parity on the GPU.

The renumbered system is an ordinary LDU system, so everything the oracle and the CUDA path do
on it is checked the usual way (bit-exact against the oracle, which is pinned to the reference);
what is specific here: the permutation is valid, the colouring is proper, the dependency depth of
the reference's lexicographic sweeps drops to the number of colours, and the renumbered system is
the same linear system."""
import numpy as np
import pytest

import cases
from ldub200 import meshes, renumber
from oracle import oracle as O


@pytest.mark.parametrize("name", ["cavity20x20", "box12_var", "asym10", "scrambled9", "box9x7x5_dirichlet", "single"])
def test_colour_order_host_logic(name):
    s = cases.system(name)
    c = renumber.colour_order(s)
    n = s["nCells"]
    assert np.array_equal(np.sort(c["perm"]), np.arange(n))
    # LDU invariants survive: l < u, faces sorted by (l, u)
    l, u = c["lower"].astype(np.int64), c["upper"].astype(np.int64)
    assert (l < u).all()
    assert (np.diff(l * n + u) > 0).all() if l.size > 1 else True
    # proper colouring: the new index ranges of the colours never contain both ends of a face,
    # i.e. the forward sweep's dependency depth equals the number of colours
    depth = renumber.sweep_depth(n, c["lower"], c["upper"])
    assert depth == c["nColours"] or (s["nFaces"] == 0 and depth <= 1)
    assert depth <= renumber.sweep_depth(n, s["lower"], s["upper"])
    if name in ("cavity20x20", "box12_var", "asym10", "box9x7x5_dirichlet"):
        assert c["nColours"] == 2           # hex box: red-black
    # the same linear system: A' x' = (A x)' up to the summation order inside a row
    x = np.random.default_rng(3).standard_normal(n)
    y = O.World([s]).amul(x)[0]
    yc = O.World([c]).amul(x[c["inv"]])[0]
    assert np.abs(yc[c["perm"]] - y).max() <= 1e-13 * max(1.0, np.abs(y).max())
    ty = O.World([s]).tmul(x)[0]
    tyc = O.World([c]).tmul(x[c["inv"]])[0]
    assert np.abs(tyc[c["perm"]] - ty).max() <= 1e-13 * max(1.0, np.abs(ty).max())


def test_renumbered_solve_is_the_same_solution():
    s = cases.system("box12_var")
    c = renumber.colour_order(s)
    ctl = dict(solver="PCG", preconditioner="diagonal", tolerance=1e-11, relTol=0, maxIter=2000)
    psi, _ = O.World([s]).solve(ctl, s["psi0"], s["source"])
    psic, _ = O.World([c]).solve(ctl, c["psi0"], c["source"])
    assert np.abs(psic[0][c["perm"]] - psi[0]).max() <= 1e-8 * np.abs(psi[0]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["box12_var", "asym10", "scrambled9", "box40x30x20"])
def test_colour_ordered_parity_gpu(ctx, name):
    """Smoothers, preconditioners and solves on the colour-ordered mesh: bit-exact against the oracle on the
    same mesh (the reference's lexicographic sweeps ARE multi-colour sweeps there)."""
    import ldub200
    s = renumber.colour_order(cases.system(name))
    w = O.World([s])
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
    A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
    rng = np.random.default_rng(4)
    x = rng.standard_normal(s["nCells"])
    assert np.array_equal(A.Amul(x), w.amul(x)[0])
    sym = s["lowerCoef"] is None
    for sm in (["GaussSeidel", "symGaussSeidel", "DIC"] if sym else ["GaussSeidel", "DILU"]):
        psi = x.copy()
        ldub200.lduMatrix.smoother.New("p", A, sm).smooth(psi, s["source"], 2)
        assert np.array_equal(psi, w.smooth(sm, x, s["source"], 2)[0]), sm
    pre = "DIC" if sym else "DILU"
    P = ldub200.lduMatrix.preconditioner.New(A, pre)
    assert np.array_equal(P.precondition(s["source"]), w.precondition(pre, s["source"])[0])
    ctl = dict(solver="PCG" if sym else "PBiCG", preconditioner=pre, tolerance=1e-8, relTol=0,
               referenceOrderSums=True)
    psi_o, perf_o = w.solve(ctl, s["psi0"], s["source"])
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"] and np.array_equal(psi, psi_o[0])
    A.destroy()
