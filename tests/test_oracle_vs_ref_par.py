"""Pin the MULTI-REGION CPU restatement (oracle World: interface updates, rank-ordered global
sums, per-region agglomeration) against the UNMODIFIED reference running as one process per
mesh region, coupled through the shared-memory Pstream seam (oracle/pstream_shm, SURVEY.md 8f
row 2; oracle/_ref/ref_driver_par).  Bit-for-bit.  Skipped where oracle/_ref is absent."""
import numpy as np
import pytest

import cases
from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.ref_par_available(), reason="oracle/_ref (parallel driver) not built")


def same(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name,R,part", [("box6x40x9", 2, "slab"), ("asym4x35x13", 4, "slab"),
                                         ("asym10", 3, "random"), ("cavity20x20", 5, "random")])
def test_operators(name, R, part):
    s, regs = cases.regions(name, R, part)
    w = O.World(regs)
    x = np.random.default_rng(5).standard_normal(s["nCells"])
    xs = [x[r["cells"]] for r in regs]
    assert same(w.amul(xs), O.ref_run_par(regs, "amul", psi=xs)[0])
    assert same(w.tmul(xs), O.ref_run_par(regs, "tmul", psi=xs)[0])
    assert same(w.sumA(), O.ref_run_par(regs, "suma")[0])
    assert same(w.residual(xs, [r["source"] for r in regs]), O.ref_run_par(regs, "residual", psi=xs)[0])


@pytest.mark.parametrize("name,R", [("box12_var", 2), ("asym10", 3)])
@pytest.mark.parametrize("sm", cases.SMOOTHERS)
def test_smoothers(name, R, sm):
    s, regs = cases.regions(name, R, "slab")
    if not cases.selectable(s, sm):
        pytest.skip("not in the reference's table for this matrix type")
    w = O.World(regs)
    x = np.random.default_rng(2).standard_normal(s["nCells"])
    xs = [x[r["cells"]] for r in regs]
    want, _ = O.ref_run_par(regs, "smooth", O.dict_text(dict(smoother=sm)), 3, psi=xs)
    assert same(w.smooth(sm, xs, [r["source"] for r in regs], 3), want)


@pytest.mark.parametrize("case", range(len(cases.MULTI_REGION_SOLVES)))
def test_solves(case):
    name, R, part, ctl = cases.MULTI_REGION_SOLVES[case]
    s, regs = cases.regions(name, R, part)
    psi_o, perf_o = O.World(regs).solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    psi_r, so = O.ref_run_par(regs, "solve", O.dict_text(cases.ref_controls(ctl)))
    perf_r = O.parse_perf(so)
    assert perf_o["nIterations"] == perf_r["nIterations"]
    assert perf_o["initialResidual"] == perf_r["initialResidual"]
    assert perf_o["finalResidual"] == perf_r["finalResidual"]
    assert perf_o["converged"] == perf_r["converged"]
    assert same(psi_o, psi_r)


@pytest.mark.parametrize("name,R,merge,weights", [("box12_var", 2, 1, True), ("box6x40x9", 4, 2, True),
                                                  ("asym10", 3, 1, False)])
def test_agglomeration(name, R, merge, weights):
    s, regs = cases.regions(name, R, "slab")
    ctl = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=10, mergeLevels=merge,
               agglomerator="faceAreaPair" if weights else "algebraicPair")
    w = O.World(regs)
    out, _ = O.ref_run_par(regs, "agglom", O.dict_text(cases.ref_controls(ctl)), ints=True)
    for r in range(R):
        mine = w.gamg_levels(ctl, r)
        n, pos = int(out[r][0]), 1
        assert n == len(mine) and n > 0
        for lev in range(n):
            nf, nc = int(out[r][pos]), int(out[r][pos + 1])
            assert mine[lev]["nCoarse"] == nc
            assert np.array_equal(mine[lev]["restrict"], out[r][pos + 2:pos + 2 + nf])
            pos += 2 + nf


@pytest.mark.parametrize("case", range(len(cases.CYCLIC_REGION_SOLVES)))
def test_regions_with_cyclic_patches(case):
    """processor interfaces between the slabs AND a cyclic pair inside every region"""
    name, R, axis, ctl = cases.CYCLIC_REGION_SOLVES[case]
    s, regs = cases.cyclic_regions(name, R, axis)
    w = O.World(regs)
    x = np.random.default_rng(9).standard_normal(s["nCells"])
    xs = [x[r["cells"]] for r in regs]
    assert same(w.amul(xs), O.ref_run_par(regs, "amul", psi=xs)[0])
    psi_o, perf_o = w.solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    psi_r, so = O.ref_run_par(regs, "solve", O.dict_text(cases.ref_controls(ctl)))
    perf_r = O.parse_perf(so)
    assert perf_o["nIterations"] == perf_r["nIterations"]
    assert perf_o["finalResidual"] == perf_r["finalResidual"]
    assert same(psi_o, psi_r)
