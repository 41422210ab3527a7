"""The untested tail of round 1 (VERDICT weak #9), on the GPU:
  * cases.GAMG_OPTION_SOLVES (level multipliers, explicit scaleCorrection, asymmetric interpolateCorrection):
    pinned against the reference on the CPU before, now also CUDA against the oracle, bit for bit;
  * a seeded slice of tests/fuzz_gpu_vs_oracle.py (random LDU graphs, dictionaries, cyclic pairs)."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(len(cases.GAMG_OPTION_SOLVES)))
def test_gamg_option_solves_bit_exact(ctx, case):
    import ldub200
    from oracle import oracle as O
    name, ctl = cases.GAMG_OPTION_SOLVES[case]
    s = cases.system(name)
    psi_o, perf_o = O.World([s]).solve(ctl, s["psi0"], s["source"])
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
    A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
    A.set_face_weights(s["faceWeights"])
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, s["source"])
    assert perf.nIterations == perf_o["nIterations"], (str(perf), perf_o)
    assert perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(psi, psi_o[0])
    psi = s["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, s["source"])     # default sums
    assert perf.nIterations == perf_o["nIterations"]
    A.destroy()


def test_seeded_fuzz_slice(ctx):
    import fuzz_gpu_vs_oracle as G
    bad = []
    for k in range(120):
        for msg in (G.solve_case(ctx, 5000 + k), G.operator_case(ctx, 5000 + k) if k % 4 == 0 else None):
            if msg:
                bad.append(msg)
    assert not bad, "\n".join(bad[:10])
