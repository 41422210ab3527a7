"""ldub200 — Python mirror of OpenFOAM-2.2.x's lduMatrix solver interface on top of
the B200-native C ABI (include/ldu_b200.h, csrc/libldu_b200.so).

Names follow the reference (src/OpenFOAM/matrices/lduMatrix/lduMatrix/lduMatrix.H):
    lduMatrix(...).Amul / Tmul / sumA / residual
    lduMatrix.solver.New(fieldName, matrix, solverControls).solve(psi, source)
    lduMatrix.smoother.New(...).smooth(psi, source, nSweeps)
    lduMatrix.preconditioner.New(...).precondition(rA)
    SolverPerformance (solverName, fieldName, initialResidual, finalResidual,
                       nIterations, converged, singular) with the reference's print format.

There is no CPU path here: every compute call goes through the CUDA library and
raises LduError when the library or a GPU is missing.
"""
from .api import (  # noqa: F401
    Context,
    DeviceField,
    LduError,
    SolverPerformance,
    lduInterface,
    lduMatrix,
    library,
    library_path,
    launch_count,
    make_controls,
    pinned_array,
)
