"""polyMesh import (ldub200.polymesh): OpenFOAM's on-disk mesh / decomposePar
output -> LDU addressing + coupled patches, and a real unstructured mesh (the reference's airFoil2D
tutorial polyMesh) through the oracle against the reference's committed results."""
import gzip
import os
from pathlib import Path

import numpy as np
import pytest

import cases
from ldub200 import polymesh
from oracle import oracle as O


def test_decomposed_case_round_trip(tmp_path):
    """regions -> processorN/constant/polyMesh files -> regions: same addressing, same interfaces"""
    s, regs = cases.cyclic_regions("box12_var", 3, 0)
    polymesh.write_decomposed_case(tmp_path, regs)
    back = polymesh.read_decomposed_case(tmp_path)
    assert len(back) == len(regs)
    for a, b in zip(back, regs):
        assert a["nCells"] == b["nCells"] and a["nFaces"] == b["nFaces"]
        assert np.array_equal(a["lower"], b["lower"]) and np.array_equal(a["upper"], b["upper"])
        assert len(a["interfaces"]) == len(b["interfaces"])
        for x, y in zip(a["interfaces"], b["interfaces"]):
            assert x["nbrRegion"] == y["nbrRegion"] and x["nbrInterface"] == y["nbrInterface"]
            assert np.array_equal(x["faceCells"], y["faceCells"])
    kinds = [p["type"] for p in back[1]["patches"]]
    assert kinds.count("processor") == 2 and kinds.count("cyclic") == 2


def test_imported_regions_solve_like_the_originals(tmp_path):
    """the imported connectivity carries a solve: coefficients attached to the imported regions give
    the oracle world the same bits as the regions they were written from"""
    s, regs = cases.regions("box6x40x9", 4, "random")
    polymesh.write_decomposed_case(tmp_path, regs)
    back = polymesh.read_decomposed_case(tmp_path)
    for a, b in zip(back, regs):
        for key in ("diag", "upperCoef", "lowerCoef", "source", "psi0", "faceWeights"):
            a[key] = b[key]
        for x, y in zip(a["interfaces"], b["interfaces"]):
            x["bouCoeffs"], x["intCoeffs"] = y["bouCoeffs"], y["intCoeffs"]
    ctl = dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)
    p1, f1 = O.World(back).solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    p2, f2 = O.World(regs).solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    assert f1["nIterations"] == f2["nIterations"] and all(np.array_equal(a, b) for a, b in zip(p1, p2))


def test_gz_and_comments(tmp_path):
    d = tmp_path / "constant" / "polyMesh"
    d.mkdir(parents=True)
    head = 'FoamFile\n{\n    version 2.0;\n    format ascii;\n    class labelList;\n    object owner;\n}\n'
    with gzip.open(d / "owner.gz", "wt") as fh:
        fh.write("/* banner */\n" + head + "// note: nCells:3\n5\n(\n0 0 1 // trailing comment\n1\n2)\n")
    (d / "neighbour").write_text(head.replace("owner", "neighbour") + "3(1 2 2)\n")
    (d / "boundary").write_text(head.replace("labelList", "polyBoundaryMesh").replace("owner", "boundary")
                                + "1\n(\nwalls\n{\n    type wall;\n    nFaces 2;\n    startFace 3;\n}\n)\n")
    m = polymesh.read_poly_mesh(d)
    assert m["nCells"] == 3 and m["nFaces"] == 3
    assert m["lower"].tolist() == [0, 0, 1] and m["upper"].tolist() == [1, 2, 2]
    assert m["patches"][0]["name"] == "walls" and m["patches"][0]["faceCells"].tolist() == [1, 2]


def test_refuses_binary_and_inconsistent_files(tmp_path):
    d = tmp_path
    head = 'FoamFile\n{\n    version 2.0;\n    format binary;\n    class labelList;\n    object owner;\n}\n'
    (d / "owner").write_text(head + "2(0 0)")
    with pytest.raises(polymesh.FoamFileError, match="binary"):
        polymesh.read_label_list(d / "owner")
    (d / "owner").write_text(head.replace("binary", "ascii") + "3(0 0)")
    with pytest.raises(polymesh.FoamFileError, match="header says 3"):
        polymesh.read_label_list(d / "owner")
    # internal faces out of upper-triangular order
    a = head.replace("binary", "ascii")
    (d / "owner").write_text(a + "2(1 0)")
    (d / "neighbour").write_text(a + "2(2 1)")
    (d / "boundary").write_text(a + "0()")
    with pytest.raises(polymesh.FoamFileError, match="upper-triangular"):
        polymesh.read_poly_mesh(d)


@pytest.mark.skipif(not os.path.isdir(cases.AIRFOIL_POLYMESH), reason="reference tutorials not present")
def test_airfoil_fixture_is_the_shipped_mesh():
    """tests/golden/airfoil2d.npz == what polymesh reads from the reference's tutorial today"""
    mesh = polymesh.read_poly_mesh(cases.AIRFOIL_POLYMESH, geometry=True)
    assert mesh["nCells"] == 10720 and mesh["nFaces"] == 21254
    assert [p["type"] for p in mesh["patches"]] == ["patch", "patch", "wall", "empty"]
    s = polymesh.laplacian_system(mesh, variable=True)
    f, g = cases.airfoil_system()
    for key in ("lower", "upper", "diag", "upperCoef", "faceWeights", "source"):
        assert np.array_equal(s[key], f[key]), key


def test_airfoil_oracle_matches_reference_results():
    """unstructured real mesh: the oracle against the reference's committed outputs (digests)"""
    s, g = cases.airfoil_system()
    w = O.World([s])
    x = cases.airfoil_x(s["nCells"])
    assert np.array_equal(cases.digest(w.amul(x)[0]), g["sha_amul"])
    assert np.array_equal(cases.digest(w.smooth("GaussSeidel", x, s["source"], 2)[0]), g["sha_smooth_GaussSeidel"])
    assert np.array_equal(cases.digest(w.precondition("DIC", s["source"])[0]), g["sha_pre_DIC"])
    for i, ctl in enumerate(cases.AIRFOIL_SOLVES):
        psi, perf = w.solve(ctl, s["psi0"], s["source"])
        ref = g[f"perf_{i}"]
        assert perf["nIterations"] == int(ref[2]), ctl
        assert perf["initialResidual"] == ref[0] and perf["finalResidual"] == ref[1], ctl
        assert np.array_equal(cases.digest(psi[0]), g[f"sha_psi_{i}"]), ctl
        if i == 0:
            assert np.array_equal(psi[0], g["psi_0"])


@pytest.mark.skipif(not O.ref_par_available(), reason="oracle/_ref (parallel driver) not built")
def test_airfoil_decomposed_against_parallel_reference(tmp_path):
    """the airfoil system cut into 4 regions, written as a decomposed case, imported again and
    solved by the coupled reference and the oracle world"""
    from ldub200 import decompose
    s, _ = cases.airfoil_system()
    proc = (np.arange(s["nCells"]) * 4 // s["nCells"]).astype(np.int32)
    regs = decompose.decompose(s, proc, 4)
    polymesh.write_decomposed_case(tmp_path, regs)
    back = polymesh.read_decomposed_case(tmp_path)
    for a, b in zip(back, regs):
        assert np.array_equal(a["lower"], b["lower"]) and np.array_equal(a["upper"], b["upper"])
        assert [i["nbrRegion"] for i in a["interfaces"]] == [i["nbrRegion"] for i in b["interfaces"]]
        assert [i["nbrInterface"] for i in a["interfaces"]] == [i["nbrInterface"] for i in b["interfaces"]]
    for ctl in (dict(solver="PCG", preconditioner="DIC", tolerance=1e-7, relTol=0),
                dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair", nCellsInCoarsestLevel=10,
                     mergeLevels=1, cacheAgglomeration=False, tolerance=1e-6, relTol=0, maxIter=30)):
        psi_o, perf_o = O.World(regs).solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
        psi_r, so = O.ref_run_par(regs, "solve", O.dict_text(cases.ref_controls(ctl)))
        perf_r = O.parse_perf(so)
        assert perf_o["nIterations"] == perf_r["nIterations"]
        assert perf_o["finalResidual"] == perf_r["finalResidual"]
        assert all(np.array_equal(a, b) for a, b in zip(psi_o, psi_r))
