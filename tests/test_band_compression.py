"""ldu_band_compression (host): the Cuthill-McKee numbering of renumberMesh's default method, pinned
to the reference's own Foam::bandCompression (libOpenFOAM, driven by oracle/ref_driver), and what it
does to the systems: same solution in the new numbering, smaller bandwidth on scrambled meshes."""
import numpy as np
import pytest

import cases
from ldub200 import renumber
from oracle import oracle as O

NAMES = ["cavity20x20", "box12_var", "asym10", "scrambled9", "scrambled17", "line50", "single"]


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", NAMES + ["airfoil"])
def test_same_numbering_as_the_reference(name):
    s = cases.airfoil_system()[0] if name == "airfoil" else cases.system(name)
    perm = renumber.band_compression(s["nCells"], s["lower"], s["upper"])
    ref_new_to_old = O.ref_run(s, "bandCompression", ints=True)[0]
    mine = np.empty(s["nCells"], dtype=np.int64)
    mine[perm] = np.arange(s["nCells"])
    assert np.array_equal(mine, ref_new_to_old)


def test_two_disconnected_components():
    """every component is started from its lowest-numbered cell of minimal degree"""
    # two chains: 0-1-2 and 3-4
    lower, upper = np.array([0, 1, 3], np.int32), np.array([1, 2, 4], np.int32)
    perm = renumber.band_compression(5, lower, upper)
    assert sorted(perm.tolist()) == [0, 1, 2, 3, 4]
    # isolated cells come first (degree 0), in index order
    perm = renumber.band_compression(4, np.array([1], np.int32), np.array([3], np.int32))
    assert np.argsort(perm).tolist() == [0, 2, 1, 3]


@pytest.mark.parametrize("name", ["scrambled9", "scrambled17"])
def test_bandwidth_shrinks_on_scrambled_meshes(name):
    s = cases.system(name)
    p = renumber.permute(s, renumber.band_compression(s["nCells"], s["lower"], s["upper"]))
    assert renumber.bandwidth(p["lower"], p["upper"]) < 0.5 * renumber.bandwidth(s["lower"], s["upper"])


@pytest.mark.parametrize("name", ["box12_var", "asym10", "scrambled9"])
def test_renumbered_system_has_the_same_solution(name):
    s = cases.system(name)
    p = renumber.permute(s, renumber.band_compression(s["nCells"], s["lower"], s["upper"]))
    ctl = dict(solver="PCG" if s["lowerCoef"] is None else "PBiCG", preconditioner="diagonal", tolerance=1e-12,
               relTol=0)
    x = O.World([s]).solve(ctl, s["psi0"].copy(), s["source"])[0][0]
    y = O.World([p]).solve(ctl, p["psi0"].copy(), p["source"])[0][0]
    assert np.abs(y[p["perm"]] - x).max() < 1e-9 * np.abs(x).max()


def test_large_system_pointers_are_not_truncated():
    """> 100k cells: numpy places such arrays in mmap'd memory above 4 GiB, which a ctypes binding without
    argtypes would truncate to 32 bits (round-1 advisor finding)."""
    from ldub200 import meshes
    n = 60
    s = meshes.laplacian_system(n, n, n)
    perm = renumber.band_compression(s["nCells"], s["lower"], s["upper"])
    assert np.array_equal(np.sort(perm), np.arange(s["nCells"]))
    p = renumber.permute(s, perm)
    # Cuthill-McKee on a box walks diagonal hyperplanes: the profile stays O(n^2)
    assert renumber.bandwidth(p["lower"], p["upper"]) < 2 * n * n
