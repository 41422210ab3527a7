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
    # multi-region: the reference as one process per region over the shared-memory Pstream
    # (oracle/pstream_shm); fields stored gathered into the global cell order
    assert O.ref_par_available(), "oracle/_ref/ref_driver_par missing"
    from ldub200 import decompose
    multi = {}
    for i, (name, R, part, ctl) in enumerate(cases.MULTI_REGION_SOLVES):
        s, regs = cases.regions(name, R, part)
        psi, so = O.ref_run_par(regs, "solve", O.dict_text(cases.ref_controls(ctl)))
        perf = O.parse_perf(so)
        multi[f"psi_{i}"] = decompose.gather_field(regs, psi, s["nCells"])
        multi[f"perf_{i}"] = np.array([perf["initialResidual"], perf["finalResidual"],
                                       perf["nIterations"], perf["converged"], perf["singular"]], dtype=np.float64)
        print("multi-region solve", i, name, R, part, ctl["solver"], perf["nIterations"])
    for name, R, part in [("asym4x35x13", 4, "slab"), ("box12_var", 3, "random")]:
        s, regs = cases.regions(name, R, part)
        x = rng.standard_normal(s["nCells"])
        xs = [x[r["cells"]] for r in regs]
        key = f"{name}_{R}_{part}"
        multi[key + "_x"] = x
        for op in ("amul", "tmul", "suma", "residual"):
            out = O.ref_run_par(regs, op, psi=xs)[0]
            multi[f"{key}_{op}"] = decompose.gather_field(regs, out, s["nCells"])
        for sm in cases.SMOOTHERS:
            if cases.selectable(s, sm):
                out = O.ref_run_par(regs, "smooth", O.dict_text(dict(smoother=sm)), 2, psi=xs)[0]
                multi[f"{key}_smooth_{sm}"] = decompose.gather_field(regs, out, s["nCells"])
    np.savez_compressed(HERE / "multi_region.npz", **multi)
    # cyclic (periodic) patches, one region
    cyc = {}
    for name, axis in cases.CYCLIC_SYSTEMS[:3]:
        s = cases.cyclic_system(name, axis)
        x = rng.standard_normal(s["nCells"])
        key = f"{name}_{axis}"
        cyc[key + "_x"] = x
        for op in ("amul", "tmul", "suma", "residual"):
            cyc[f"{key}_{op}"] = O.ref_run(s, op, psi=x)[0]
        for sm in cases.SMOOTHERS:
            if cases.selectable(s, sm):
                cyc[f"{key}_smooth_{sm}"] = O.ref_run(s, "smooth", O.dict_text(dict(smoother=sm)), 2, psi=x)[0]
    for i, (name, axis, ctl) in enumerate(cases.CYCLIC_SOLVES):
        s = cases.cyclic_system(name, axis)
        psi, perf = O.ref_solve(s, cases.ref_controls(ctl))
        cyc[f"psi_{i}"] = psi
        cyc[f"perf_{i}"] = np.array([perf["initialResidual"], perf["finalResidual"],
                                     perf["nIterations"], perf["converged"], perf["singular"]], dtype=np.float64)
        print("cyclic solve", i, name, ctl["solver"], perf["nIterations"])
    np.savez_compressed(HERE / "cyclic.npz", **cyc)
    # GAMG controls beyond GAMG_SOLVES: digests of the reference's solutions
    gopt = {}
    for i, (name, ctl) in enumerate(cases.GAMG_OPTION_SOLVES):
        s = cases.system(name)
        psi, perf = O.ref_solve(s, cases.ref_controls(ctl))
        gopt[f"sha_psi_{i}"] = cases.digest(psi)
        gopt[f"perf_{i}"] = np.array([perf["initialResidual"], perf["finalResidual"], perf["nIterations"],
                                      perf["converged"]], dtype=np.float64)
        print("GAMG option", i, name, perf["nIterations"])
    np.savez_compressed(HERE / "gamg_options.npz", **gopt)
    # two solves on one mesh with the agglomeration cached (or not) in between
    cache = {}
    for i, (name, ctl) in enumerate(cases.CACHE_SOLVES):
        s = cases.system(name)
        psi, so = O.ref_run(s, "solve2", O.dict_text(cases.ref_controls(ctl)))
        p1, p2 = O.parse_perfs(so)
        cache[f"psi_{i}"] = psi
        cache[f"perf_{i}"] = np.array([p1["nIterations"], p1["finalResidual"], p2["nIterations"],
                                       p2["finalResidual"]], dtype=np.float64)
        print("cached agglomeration", i, name, p1["nIterations"], p2["nIterations"])
    np.savez_compressed(HERE / "cache_solves.npz", **cache)
    # edge cases of the solver front end
    edge = {}
    for i in range(len(cases.EDGE_SOLVES)):
        s, ctl, psi0, source = cases.edge_case(i)
        psi, so = O.ref_run(s, "solve", O.dict_text(cases.ref_controls(ctl)), psi=psi0, source=source)
        perf = O.parse_perf(so)
        edge[f"psi_{i}"] = psi
        edge[f"perf_{i}"] = np.array([perf["initialResidual"], perf["finalResidual"],
                                      perf["nIterations"], perf["converged"], perf["singular"]], dtype=np.float64)
        print("edge", i, cases.EDGE_SOLVES[i][0], perf["solverName"], perf["nIterations"])
    np.savez_compressed(HERE / "edge_cases.npz", **edge)
    # a real unstructured mesh: the polyMesh the reference ships with the airFoil2D tutorial, read by
    # ldub200.polymesh; Laplacian coefficients from its geometry; solved by the reference
    from ldub200 import polymesh
    mesh = polymesh.read_poly_mesh(cases.AIRFOIL_POLYMESH, geometry=True)
    s = polymesh.laplacian_system(mesh, variable=True)
    air = dict(lower=s["lower"], upper=s["upper"], diag=s["diag"], upperCoef=s["upperCoef"],
               faceWeights=s["faceWeights"], source=s["source"])
    # outputs are kept as SHA-256 digests of their bytes (bit-exact comparison, small fixture);
    # the first solution in full
    x = cases.airfoil_x(s["nCells"])
    air["sha_amul"] = cases.digest(O.ref_run(s, "amul", psi=x)[0])
    air["sha_smooth_GaussSeidel"] = cases.digest(
        O.ref_run(s, "smooth", O.dict_text(dict(smoother="GaussSeidel")), 2, psi=x)[0])
    air["sha_pre_DIC"] = cases.digest(O.ref_run(s, "precondition", "DIC")[0])
    for i, ctl in enumerate(cases.AIRFOIL_SOLVES):
        psi, perf = O.ref_solve(s, cases.ref_controls(ctl))
        if i == 0:
            air["psi_0"] = psi
        air[f"sha_psi_{i}"] = cases.digest(psi)
        air[f"perf_{i}"] = np.array([perf["initialResidual"], perf["finalResidual"],
                                     perf["nIterations"], perf["converged"], perf["singular"]], dtype=np.float64)
        print("airFoil2D solve", i, ctl["solver"], perf["nIterations"])
    np.savez_compressed(HERE / "airfoil2d.npz", **air)
    print("multi-region", len(multi))


if __name__ == "__main__":
    main()
