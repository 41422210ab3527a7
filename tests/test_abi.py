"""The C-ABI library loads without a GPU, exports every symbol include/ldu_b200.h
declares, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import ldub200
from ldub200 import api

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "ldu_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(ldu_[a-z0-9_A-Z]+)\s*\(", body)))


def test_header_and_binding_list_agree():
    assert declared_functions() == sorted(api.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(ldub200.library_path()))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/ldu_b200.h but not exported"


def test_no_torch_or_cxx_types_in_signatures():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    assert "std::" not in body and "torch" not in body and "at::" not in body
    assert 'extern "C"' in body


def test_controls_defaults_match_reference():
    """lduMatrixSolver.C:164-169, smoothSolver.C:70-74, GAMGSolver.C:66-76, GAMGPreconditioner.C:74-78"""
    c = ldub200.make_controls({})
    assert (c.maxIter, c.tolerance, c.relTol, c.nSweeps) == (1000, 1e-6, 0.0, 1)
    assert (c.nPreSweeps, c.preSweepsLevelMultiplier, c.maxPreSweeps) == (0, 1, 4)
    assert (c.nPostSweeps, c.postSweepsLevelMultiplier, c.maxPostSweeps, c.nFinestSweeps) == (2, 1, 4, 2)
    assert c.interpolateCorrection == 0 and c.scaleCorrection == -1 and c.nVcycles == 2
    assert c.referenceOrderSums == 0


def test_controls_struct_layout_matches_header():
    fields = re.search(r"typedef struct ldu_controls \{(.*?)\} ldu_controls;", HEADER, flags=re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = re.findall(r"\b(?:int|double)\s+(\w+)\s*;", fields)
    assert names == [f[0] for f in api.Controls._fields_]


def test_dictionary_translation():
    c = ldub200.make_controls(dict(solver="PCG", tolerance=1e-9,
                                   preconditioner=dict(preconditioner="GAMG", smoother="DIC", tolerance=1e-5,
                                                       relTol=0.1, nVcycles=3, mergeLevels=2,
                                                       agglomerator="algebraicPair")))
    assert c.solver == api.SOLVERS["PCG"] and c.preconditioner == api.PRECONDITIONERS["GAMG"]
    assert c.smoother == api.SMOOTHERS["DIC"] and c.nVcycles == 3 and c.mergeLevels == 2
    assert c.precTolerance == 1e-5 and c.precRelTol == 0.1 and c.tolerance == 1e-9
    assert c.useFaceWeights == 0
    c = ldub200.make_controls(dict(solver="ICCG", preconditioner="DIC"))      # ICCG.C:67-86: PCG on the same dictionary
    assert c.solver == api.SOLVERS["PCG"] and c.preconditioner == api.PRECONDITIONERS["DIC"]
    with pytest.raises(ldub200.LduError):
        ldub200.make_controls(dict(solver="notASolver"))


def test_no_cpu_fallback_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(ldub200.LduError, match="no CUDA device"):
        ldub200.Context(0)


def test_solver_performance_print_format():
    """SolverPerformance.C:95-125, the line foamLog parses"""
    p = ldub200.SolverPerformance("DICPCG", "p", 1.0, 3.26718e-07, 32, True, False)
    assert str(p) == "DICPCG:  Solving for p, Initial residual = 1, Final residual = 3.26718e-07, No Iterations 32"
    p = ldub200.SolverPerformance("DICPCG", "p", singular=True)
    assert str(p) == "DICPCG:  Solving for p:  solution singularity"


def test_iccg_biccg_aliases_and_mandatory_preconditioner():
    """ICCG / BICCG = PCG / PBiCG on the same dictionary (ICCG.C:67-86); the preconditioner entry has
    no default (lduMatrixPreconditioner.C:39-58)"""
    import ldub200
    c = ldub200.make_controls(dict(solver="ICCG", preconditioner="diagonal", tolerance=1e-8))
    assert c.solver == 0 and c.preconditioner == ldub200.api.PRECONDITIONERS["diagonal"]
    c = ldub200.make_controls(dict(solver="BICCG", preconditioner="DILU"))
    assert c.solver == 1 and c.preconditioner == ldub200.api.PRECONDITIONERS["DILU"]
    for name in ("PCG", "PBiCG", "ICCG", "BICCG"):
        with pytest.raises(ldub200.LduError, match="preconditioner"):
            ldub200.make_controls(dict(solver=name, tolerance=1e-8))


def test_direct_solve_coarsest_is_refused():
    with pytest.raises(ldub200.LduError, match="directSolveCoarsest"):
        ldub200.make_controls(dict(solver="GAMG", smoother="GaussSeidel", directSolveCoarsest=True))
    ldub200.make_controls(dict(solver="GAMG", smoother="GaussSeidel", directSolveCoarsest=False))


def test_environment_switches_are_documented():
    """every LDU_* variable the library or the plug-in reads is listed in INTEGRATION.md section 4, and the table
    lists nothing that no longer exists"""
    import re
    root = Path(__file__).resolve().parent.parent
    src = ""
    for pat in ("openfoam-2.2.x_b200/csrc/*.cu", "openfoam-2.2.x_b200/csrc/*.cuh", "openfoam-2.2.x_b200/csrc/*.h",
                "openfoam-2.2.x_b200/foam/*.C"):
        for f in root.glob(pat):
            src += f.read_text()
    read = set(re.findall(r'getenv\("(LDU_[A-Z0-9_]+)"\)', src))
    doc = (root / "INTEGRATION.md").read_text()
    table = doc[doc.index("## 4. Environment switches"):]
    listed = set(re.findall(r"`(LDU_[A-Z0-9_]+)", table))
    assert read - listed == set(), f"undocumented: {sorted(read - listed)}"
    assert listed - read == set(), f"documented but not read anywhere: {sorted(listed - read)}"
