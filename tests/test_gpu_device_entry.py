"""Row f1: the device-resident hand-off.  ldu_matrix_set_coeffs_device / ldu_solve_device /
ldu_amul_device / ldu_tmul_device / ldu_H_device take fields that already live in HBM: same bits as the
host-pointer entry points (which are the ones the oracle-parity tests exercise), no PCIe traffic.  Also the staged
copy path that host-pointer calls take for PAGEABLE memory above 1 MB (what an application's scalarFields are)."""
import numpy as np
import pytest

import cases
from ldub200 import meshes

pytestmark = pytest.mark.gpu


def _host_matrix(ctx, s):
    import ldub200
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
    A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
    return A


@pytest.mark.parametrize("name", ["box12_var", "asym10", "scrambled9"])
def test_device_entry_points_match_host_entry_points(ctx, name):
    import ldub200
    s = cases.system(name)
    O = __import__("oracle.oracle", fromlist=["x"])
    w = O.World([s])
    n = s["nCells"]
    # coefficients handed over on the device
    B = ldub200.lduMatrix(ctx, n, s["lower"], s["upper"])
    d_diag = ldub200.DeviceField(ctx, n, s["diag"])
    d_up = ldub200.DeviceField(ctx, s["nFaces"], s["upperCoef"])
    d_lo = None if s["lowerCoef"] is None else ldub200.DeviceField(ctx, s["nFaces"], s["lowerCoef"])
    B.set_coeffs_device(d_diag, d_up, d_lo)
    x = np.sin(0.3 * np.arange(n))
    d_x, d_y = ldub200.DeviceField(ctx, n, x), ldub200.DeviceField(ctx, n)
    B.Amul_device(d_y, d_x)
    assert np.array_equal(d_y.download(), w.amul(x)[0])
    B.Tmul_device(d_y, d_x)
    assert np.array_equal(d_y.download(), w.tmul(x)[0])
    B.H_device(d_y, d_x)
    assert np.array_equal(d_y.download(), w.H(x)[0])
    # device-resident solve == the oracle (reference-order sums: bit for bit)
    ctl = (dict(solver="PCG", preconditioner="DIC") if s["lowerCoef"] is None
           else dict(solver="PBiCG", preconditioner="DILU"))
    ctl.update(tolerance=1e-9, relTol=0, referenceOrderSums=True)
    psi_o, perf_o = w.solve(ctl, s["psi0"], s["source"])
    d_psi, d_src = ldub200.DeviceField(ctx, n, s["psi0"]), ldub200.DeviceField(ctx, n, s["source"])
    perf = ldub200.lduMatrix.solver.New("p", B, ctl).solve_device(d_psi, d_src)
    assert perf.nIterations == perf_o["nIterations"] and perf.finalResidual == perf_o["finalResidual"]
    assert np.array_equal(d_psi.download(), psi_o[0])
    # new coefficients on the same addressing, again on the device (the next time step's matrix)
    diag2, up2, lo2 = O.second_coeffs(s)
    d_diag.upload(diag2)
    d_up.upload(up2)
    if d_lo is not None:
        d_lo.upload(lo2)
    B.set_coeffs_device(d_diag, d_up, d_lo)
    w.set_coeffs(0, diag2, up2, lo2)
    B.Amul_device(d_y, d_x)
    assert np.array_equal(d_y.download(), w.amul(x)[0])
    d_psi.upload(s["psi0"])
    psi_o, perf_o = w.solve(ctl, s["psi0"], s["source"])
    perf = ldub200.lduMatrix.solver.New("p", B, ctl).solve_device(d_psi, d_src)
    assert perf.nIterations == perf_o["nIterations"] and np.array_equal(d_psi.download(), psi_o[0])
    B.destroy()


def test_device_coefficients_of_a_coupled_matrix_need_the_interface_coefficients(ctx):
    """round-1 advisor finding: ldu_matrix_set_coeffs_device left bouCoeffs/intCoeffs uninitialised"""
    import ldub200
    s = cases.cyclic_system("box12_var", 0)
    O = __import__("oracle.oracle", fromlist=["x"])
    its = s["interfaces"]
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in its]
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"], ifs)
    d_diag = ldub200.DeviceField(ctx, s["nCells"], s["diag"])
    d_up = ldub200.DeviceField(ctx, s["nFaces"], s["upperCoef"])
    with pytest.raises(ldub200.LduError):
        A.set_coeffs_device(d_diag, d_up)
    A.set_interface_coeffs([it["bouCoeffs"] for it in its], [it["intCoeffs"] for it in its])
    A.set_coeffs_device(d_diag, d_up)
    x = np.cos(0.2 * np.arange(s["nCells"]))
    assert np.array_equal(A.Amul(x), O.World([s]).amul(x)[0])
    A.destroy()


def test_pageable_and_pinned_host_fields_give_the_same_bits(ctx):
    """3 MB fields: above the staging threshold.  Pageable numpy arrays go through the multi-threaded pinned ring,
    ldu_host_alloc'd arrays straight to the copy engine."""
    import ldub200
    s = meshes.laplacian_system(72, 72, 72, variable=True)      # 373k cells, 1.1M faces: 3 MB / 8.7 MB arrays
    n = s["nCells"]
    A = _host_matrix(ctx, s)                                    # pageable diag / upper
    x = np.sin(0.01 * np.arange(n))
    y_pageable = A.Amul(x)
    hx, hd, hu = ldub200.pinned_array(n), ldub200.pinned_array(n), ldub200.pinned_array(s["nFaces"])
    hx[:], hd[:], hu[:] = x, s["diag"], s["upperCoef"]
    A.set_coeffs(hd, hu)
    y_pinned = A.Amul(hx)
    assert np.array_equal(y_pageable, y_pinned)
    # against an independent evaluation of the same operator (scipy CSR, different summation order)
    import scipy.sparse as sp
    lo, up = s["lower"], s["upper"]
    M = sp.coo_matrix((np.concatenate([s["diag"], s["upperCoef"], s["upperCoef"]]),
                       (np.concatenate([np.arange(n), lo, up]), np.concatenate([np.arange(n), up, lo]))),
                      shape=(n, n)).tocsr()
    ref = M @ x
    assert np.abs(y_pageable - ref).max() <= 1e-12 * np.abs(ref).max()
    # in/out field of a solve through pageable memory: psi comes back complete
    psi = np.zeros(n)
    perf = ldub200.lduMatrix.solver.New("p", A, dict(solver="PCG", preconditioner="DIC", tolerance=0, relTol=0,
                                                     maxIter=4)).solve(psi, s["source"])
    assert perf.nIterations == 5 and np.isfinite(psi).all() and np.count_nonzero(psi) > 0.99 * n
    psi2 = ldub200.pinned_array(n)
    psi2[:] = 0
    ldub200.lduMatrix.solver.New("p", A, dict(solver="PCG", preconditioner="DIC", tolerance=0, relTol=0,
                                              maxIter=4)).solve(psi2, s["source"])
    assert np.array_equal(psi, psi2)
    A.destroy()
