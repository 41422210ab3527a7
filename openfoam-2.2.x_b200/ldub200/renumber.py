"""Renumbering of an LDU system.

`permute` applies a cell permutation and restores the upper-triangular face order
(what renumberMesh does to owner/neighbour, applications/utilities/mesh/manipulation/
renumberMesh); `colour_order` computes the colour-ordered permutation with the host
function `ldu_colour_order` of the C ABI.  In the colour-ordered numbering the
reference's lexicographic Gauss-Seidel / DIC / DILU sweeps depend on nColours levels
only, so the dataflow sweeps of the CUDA path finish in a few hops instead of
nx+ny+nz: the reference smoother itself becomes a multi-colour smoother, and parity
with the reference on the renumbered mesh stays bit-exact.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api


def permute(sysd: dict, perm: np.ndarray) -> dict:
    """System renumbered with perm[old cell] = new cell.  Faces that change
    orientation swap their lower/upper coefficients; faces are re-sorted into
    upper-triangular order.  The result carries `perm` (and `inv`, old index of
    a new cell) so fields can be mapped back: x_old = x_new[perm]."""
    n = sysd["nCells"]
    perm = np.asarray(perm, dtype=np.int64)
    assert perm.shape == (n,) and np.array_equal(np.sort(perm), np.arange(n)), "not a permutation"
    lo, up = perm[sysd["lower"]], perm[sysd["upper"]]
    flip = lo > up
    l2, u2 = np.where(flip, up, lo), np.where(flip, lo, up)
    uc = sysd["upperCoef"]
    lc = sysd["lowerCoef"] if sysd["lowerCoef"] is not None else uc
    uc2, lc2 = np.where(flip, lc, uc), np.where(flip, uc, lc)
    order = np.lexsort((u2, l2))
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    out = dict(sysd)
    out.update(lower=l2[order].astype(np.int32), upper=u2[order].astype(np.int32),
               upperCoef=uc2[order].copy(),
               lowerCoef=None if sysd["lowerCoef"] is None else lc2[order].copy(),
               diag=sysd["diag"][inv].copy(), source=sysd["source"][inv].copy(),
               psi0=sysd["psi0"][inv].copy(),
               faceWeights=None if sysd.get("faceWeights") is None else sysd["faceWeights"][order].copy(),
               perm=perm, inv=inv)
    return out


def colour_permutation(nCells: int, lower, upper):
    """(perm, nColours) from ldu_colour_order: perm[old cell] = new cell."""
    L = api.library()
    lo = np.ascontiguousarray(lower, dtype=np.int32)
    up = np.ascontiguousarray(upper, dtype=np.int32)
    perm = np.empty(nCells, dtype=np.int32)
    nc = C.c_int(0)
    rc = L.ldu_colour_order(nCells, lo.size, lo.ctypes.data, up.ctypes.data, perm.ctypes.data, C.addressof(nc))
    if rc != 0:
        raise api.LduError(f"ldu_colour_order failed ({rc}): {L.ldu_last_error().decode()}")
    return perm.astype(np.int64), nc.value


def colour_order(sysd: dict) -> dict:
    """The system in colour-ordered numbering; adds `nColours`."""
    perm, nc = colour_permutation(sysd["nCells"], sysd["lower"], sysd["upper"])
    out = permute(sysd, perm)
    out["nColours"] = nc
    return out


def sweep_depth(nCells: int, lower, upper) -> int:
    """Dependency depth of a forward lexicographic sweep: longest chain of lower neighbours."""
    depth = np.zeros(nCells, dtype=np.int64)
    lower = np.asarray(lower)
    upper = np.asarray(upper)
    # faces are sorted by lower cell: one pass in face order propagates depths
    for l, u in zip(lower.tolist(), upper.tolist()):
        d = depth[l] + 1
        if d > depth[u]:
            depth[u] = d
    return int(depth.max()) + 1 if nCells else 0


def band_compression(nCells: int, lower, upper):
    """Cuthill-McKee numbering of renumberMesh's default method (ldu_band_compression, host):
    returns perm with perm[old cell] = new cell, ready for `permute`."""
    L = api.library()
    lo = np.ascontiguousarray(lower, dtype=np.int32)
    up = np.ascontiguousarray(upper, dtype=np.int32)
    new_to_old = np.empty(nCells, dtype=np.int32)
    rc = L.ldu_band_compression(nCells, lo.size, lo.ctypes.data, up.ctypes.data, new_to_old.ctypes.data)
    if rc != 0:
        raise api.LduError(f"ldu_band_compression failed ({rc}): {L.ldu_last_error().decode()}")
    perm = np.empty(nCells, dtype=np.int64)
    perm[new_to_old] = np.arange(nCells)
    return perm


def bandwidth(lower, upper) -> int:
    """largest |upper - lower| over the faces (the profile the gathers of Amul walk)"""
    lower, upper = np.asarray(lower, dtype=np.int64), np.asarray(upper, dtype=np.int64)
    return int(np.abs(upper - lower).max()) if lower.size else 0
