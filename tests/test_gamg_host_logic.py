"""Host side of the GAMG hierarchy (csrc/gamg_host.h: the code gamg.cu runs on the host), on the CPU:
level by level against the oracle's hierarchy, which tests/test_oracle_vs_ref*.py pin to the reference's
pairGAMGAgglomeration (pairGAMGAgglomerate.C:36-198, GAMGAgglomerateLduAddressing.C:34-214), and the flat
coarse addressing against the straightforward version on random graphs."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from ldub200 import meshes
from oracle import oracle as O

import cases

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "openfoam-2.2.x_b200" / "csrc" / "libldu_hosttest.so"


@pytest.fixture(scope="module")
def lib():
    if not LIB.exists():
        pytest.skip("libldu_hosttest.so not built (make -C openfoam-2.2.x_b200/csrc)")
    L = C.CDLL(str(LIB))
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.ldu_hosttest_pair_cluster.argtypes = [C.c_int, C.c_int, ip, ip, dp, ip]
    L.ldu_hosttest_pair_cluster.restype = C.c_int
    L.ldu_hosttest_coarse_addressing.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, C.c_int, ip, ip, ip]
    L.ldu_hosttest_coarse_addressing.restype = C.c_int
    L.ldu_hosttest_agglomerate_interface.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip]
    L.ldu_hosttest_agglomerate_interface.restype = C.c_int
    return L


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def pair_cluster(L, n, lower, upper, w):
    lower, upper = np.ascontiguousarray(lower, np.int32), np.ascontiguousarray(upper, np.int32)
    w = np.ascontiguousarray(w, np.float64)
    cmap = np.empty(n, np.int32)
    nc = L.ldu_hosttest_pair_cluster(n, lower.size, _ip(lower), _ip(upper), w.ctypes.data_as(C.POINTER(C.c_double)),
                                     _ip(cmap))
    return nc, cmap


def coarse_addressing(L, which, nc, lower, upper, cmap):
    lower, upper = np.ascontiguousarray(lower, np.int32), np.ascontiguousarray(upper, np.int32)
    cmap = np.ascontiguousarray(cmap, np.int32)
    nf = lower.size
    fm, co, cn = (np.empty(max(nf, 1), np.int32) for _ in range(3))
    ncf = L.ldu_hosttest_coarse_addressing(which, nc, nf, _ip(lower), _ip(upper), _ip(cmap), cmap.size, _ip(fm),
                                           _ip(co), _ip(cn))
    return fm[:nf].copy(), co[:ncf].copy(), cn[:ncf].copy()


SYSTEMS = [
    ("box 12x10x8", lambda: meshes.laplacian_system(12, 10, 8, variable=True)),
    ("sheet 40x30", lambda: meshes.laplacian_system(40, 30, 1, variable=True)),
    ("scrambled box", lambda: meshes.scramble(meshes.laplacian_system(9, 9, 9, variable=True), 5)),
    ("asymmetric box", lambda: meshes.laplacian_system(8, 7, 6, variable=True, asym=0.3)),
]


@pytest.mark.parametrize("name,make", SYSTEMS, ids=[s[0] for s in SYSTEMS])
@pytest.mark.parametrize("agglomerator", ["algebraicPair", "faceAreaPair"])
def test_host_levels_equal_the_oracle_hierarchy(lib, name, make, agglomerator):
    s = make()
    if agglomerator == "faceAreaPair" and "faceWeights" not in s:
        pytest.skip("system has no face-area weights")
    ctl = dict(cases.ref_controls(dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=4,
                                       mergeLevels=1)), agglomerator=agglomerator)
    want = O.World([s]).gamg_levels(ctl)
    assert len(want) >= 2
    n, lower, upper = s["nCells"], s["lower"], s["upper"]
    w = np.abs(s["upperCoef"]) if agglomerator == "algebraicPair" else np.asarray(s["faceWeights"], float)
    for lev in want:
        nc, cmap = pair_cluster(lib, n, lower, upper, w)
        assert nc == lev["nCoarse"]
        assert np.array_equal(cmap, lev["restrict"])
        fm, co, cn = coarse_addressing(lib, 0, nc, lower, upper, cmap)
        assert np.array_equal(co, lev["lower"]) and np.array_equal(cn, lev["upper"])
        # restrict the weights in face order (GAMGAgglomerationTemplates.C:63-83)
        cw = np.zeros(co.size)
        keep = fm >= 0
        np.add.at(cw, fm[keep], w[keep])
        n, lower, upper, w = nc, co, cn, cw


@pytest.mark.parametrize("seed", range(6))
def test_flat_coarse_addressing_equals_the_straightforward_one(lib, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 400))
    nf = int(rng.integers(1, 6 * n))
    a, b = rng.integers(0, n, nf), rng.integers(0, n, nf)
    keep = a != b
    lower, upper = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
    order = np.lexsort((upper, lower))
    lower, upper = lower[order].astype(np.int32), upper[order].astype(np.int32)
    nc = int(rng.integers(1, n + 1))
    cmap = rng.integers(0, nc, n).astype(np.int32)      # any map, also with empty coarse cells
    got = coarse_addressing(lib, 0, nc, lower, upper, cmap)
    want = coarse_addressing(lib, 1, nc, lower, upper, cmap)
    for g, wv in zip(got, want):
        assert np.array_equal(g, wv)


def agglomerate_interface(L, my_rank, nbr_rank, local, nbr):
    local, nbr = np.ascontiguousarray(local, np.int32), np.ascontiguousarray(nbr, np.int32)
    fc, fr = np.empty(max(local.size, 1), np.int32), np.empty(max(local.size, 1), np.int32)
    n = L.ldu_hosttest_agglomerate_interface(my_rank, nbr_rank, local.size, _ip(local), _ip(nbr), _ip(fc), _ip(fr))
    return fc[:n].copy(), fr[:local.size].copy()


@pytest.mark.skipif(not O.ref_par_available(), reason="ref_driver_par not built")
@pytest.mark.parametrize("n_regions,partition", [(2, "slab"), (3, "slab"), (3, "random")])
def test_coupled_levels_equal_the_compiled_reference(lib, n_regions, partition):
    """Several regions: clustering per region, and the coarse processor interfaces formed from the neighbour's
    restrict map (what gamg.cu exchanges through the halo kernels), against GAMGAgglomeration::New of the
    unmodified reference running one process per region."""
    s, regs = cases.regions("box12_var", n_regions, partition)
    ctl = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=4, mergeLevels=1,
               agglomerator="faceAreaPair")
    ref = O.ref_agglom_full_par(regs, cases.ref_controls(ctl))
    n_lev = len(ref[0])
    assert n_lev >= 2 and all(len(r) == n_lev for r in ref)
    cur = [dict(n=r["nCells"], lower=r["lower"], upper=r["upper"], w=np.asarray(r["faceWeights"], float),
                ifCells=[it["faceCells"] for it in r["interfaces"]]) for r in regs]
    for lev in range(n_lev):
        cmaps = []
        for r, c in enumerate(cur):
            nc, cmap = pair_cluster(lib, c["n"], c["lower"], c["upper"], c["w"])
            assert nc == ref[r][lev]["nCoarse"] and np.array_equal(cmap, ref[r][lev]["restrict"])
            cmaps.append(cmap)
        nxt = []
        for r, c in enumerate(cur):
            fm, co, cn = coarse_addressing(lib, 0, ref[r][lev]["nCoarse"], c["lower"], c["upper"], cmaps[r])
            assert np.array_equal(co, ref[r][lev]["lower"]) and np.array_equal(cn, ref[r][lev]["upper"])
            assert np.array_equal(fm, ref[r][lev]["faceRestrict"])
            if_cells = []
            for p, it in enumerate(regs[r]["interfaces"]):
                q, pq = it["nbrRegion"], it["nbrInterface"]
                local = cmaps[r][c["ifCells"][p]]
                nbr = cmaps[q][cur[q]["ifCells"][pq]]          # the halo exchange of the restrict map
                fc, fr = agglomerate_interface(lib, r, q, local, nbr)
                assert np.array_equal(fc, ref[r][lev]["ifCells"][p])
                assert np.array_equal(fr, ref[r][lev]["ifRestrict"][p])
                if_cells.append(fc)
            cw = np.zeros(co.size)
            keep = fm >= 0
            np.add.at(cw, fm[keep], c["w"][keep])
            nxt.append(dict(n=ref[r][lev]["nCoarse"], lower=co, upper=cn, w=cw, ifCells=if_cells))
        cur = nxt
