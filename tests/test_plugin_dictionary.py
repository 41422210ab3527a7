"""The OpenFOAM plug-in reads its dictionary the way the reference's solver constructors do: the entries
the reference looks up without a default are mandatory, and what the GPU solver cannot do is refused
before any device work.  Runs without a GPU (the errors come first); needs the reference build."""
from pathlib import Path

import pytest

import cases
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent
PLUGIN = ROOT / "openfoam-2.2.x_b200" / "foam" / "libgpuLduSolvers.so"

pytestmark = pytest.mark.skipif(not (PLUGIN.exists() and O.ref_available()),
                                reason="plug-in / reference binaries not built")


def run(text):
    s = cases.system("cavity20x20")
    return O.ref_run(s, "solve", text, extra_env=dict(LDU_REF_LIBS=str(PLUGIN)))


@pytest.mark.parametrize("text,message", [
    # pairGAMGAgglomeration.C:45
    ("solver gpuGAMG; smoother GaussSeidel; agglomerator algebraicPair; nCellsInCoarsestLevel 10;",
     "keyword mergeLevels is undefined"),
    # GAMGAgglomeration.C:77-80
    ("solver gpuGAMG; smoother GaussSeidel; agglomerator algebraicPair; mergeLevels 1;",
     "keyword nCellsInCoarsestLevel is undefined"),
    # lduMatrixSmoother.C:38-66
    ("solver gpuGAMG; agglomerator algebraicPair; mergeLevels 1; nCellsInCoarsestLevel 10;",
     "keyword smoother is undefined"),
    ("solver gpuSmoothSolver; tolerance 1e-8;", "keyword smoother is undefined"),
    # GAMGAgglomeration.C:104-107
    ("solver gpuGAMG; smoother GaussSeidel; mergeLevels 1; nCellsInCoarsestLevel 10;",
     "keyword agglomerator is undefined"),
    # lduMatrixPreconditioner.C:39-58
    ("solver gpuPCG; tolerance 1e-8;", "keyword preconditioner is undefined"),
    ("solver gpuICCG; tolerance 1e-8;", "keyword preconditioner is undefined"),
    ("solver gpuPCG; preconditioner notAPreconditioner;", "Unknown preconditioner notAPreconditioner"),
    ("solver gpuSmoothSolver; smoother notASmoother;", "Unknown smoother notASmoother"),
    # outside the library: said, not ignored
    ("solver gpuGAMG; smoother GaussSeidel; agglomerator algebraicPair; mergeLevels 1; nCellsInCoarsestLevel 10; "
     "directSolveCoarsest on;",
     "directSolveCoarsest is not supported"),
    # the reference's own table look-up (lduMatrixSolver.C:68-84): gpuPBiCG is not a symmetric-matrix solver
    ("solver gpuPBiCG; preconditioner DILU;", "Unknown symmetric matrix solver gpuPBiCG"),
])
def test_dictionary_errors_come_first(text, message):
    with pytest.raises(RuntimeError, match=message):
        run(text)


def test_the_same_dictionaries_fail_the_same_way_in_the_reference():
    for text, message in [
        ("solver GAMG; smoother GaussSeidel; agglomerator algebraicPair; nCellsInCoarsestLevel 10;",
         "keyword mergeLevels is undefined"),
        ("solver GAMG; smoother GaussSeidel; agglomerator algebraicPair; mergeLevels 1;",
         "keyword nCellsInCoarsestLevel is undefined"),
        ("solver GAMG; agglomerator algebraicPair; mergeLevels 1; nCellsInCoarsestLevel 10;",
         "keyword smoother is undefined"),
        ("solver smoothSolver; tolerance 1e-8;", "keyword smoother is undefined"),
        ("solver GAMG; smoother GaussSeidel; mergeLevels 1; nCellsInCoarsestLevel 10;",
         "keyword agglomerator is undefined"),
        ("solver PCG; tolerance 1e-8;", "keyword preconditioner is undefined"),
    ]:
        with pytest.raises(RuntimeError, match=message):
            O.ref_run(cases.system("cavity20x20"), "solve", text)
