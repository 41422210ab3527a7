"""Synthetic LDU systems for the lduMatrix hot path.

Hex box nx*ny*nz with lexicographic cell index c=(k*ny+j)*nx+i and faces
generated per cell in the order (+i, +j, +k): this is the upper-triangular
owner/neighbour order blockMesh + fvMeshLduAddressing produce
(reference: src/finiteVolume/fvMesh/fvMeshLduAddressing.H:83-102,
 src/OpenFOAM/matrices/lduMatrix/lduAddressing/lduAddressing.H:36-63).

Coefficients follow fvm::laplacian (gaussLaplacianScheme.C:46-88):
upper = gamma*magSf*deltaCoeffs (> 0), diag = -sum(offdiag) (negSumDiag),
plus either one reference cell (fvMatrix::setReference, fvMatrix.C:509-521)
or Dirichlet wall faces adding to diag.
"""
from __future__ import annotations

import numpy as np


def box_addressing(nx: int, ny: int, nz: int):
    """lowerAddr, upperAddr (int32) of an nx*ny*nz hex box, upper-triangular order."""
    n = nx * ny * nz
    c = np.arange(n, dtype=np.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx * ny)
    nbr = np.empty((n, 3), dtype=np.int64)
    nbr[:, 0] = np.where(i < nx - 1, c + 1, -1)
    nbr[:, 1] = np.where(j < ny - 1, c + nx, -1)
    nbr[:, 2] = np.where(k < nz - 1, c + nx * ny, -1)
    mask = nbr >= 0
    lower = np.repeat(c, 3).reshape(n, 3)[mask].astype(np.int32)
    upper = nbr[mask].astype(np.int32)
    direction = np.tile(np.arange(3, dtype=np.int8), n).reshape(n, 3)[mask]
    return lower, upper, direction


def laplacian_system(nx, ny, nz, *, variable=False, dirichlet=False,
                     ref_cell=0, ref_value=0.0, asym=0.0, seed=None):
    """Return dict(nCells,nFaces,lower,upper,diag,upperCoef,lowerCoef|None,source,psi0,faceWeights)."""
    lower, upper, direction = box_addressing(nx, ny, nz)
    n = nx * ny * nz
    nf = lower.size
    f = np.arange(nf, dtype=np.float64)
    if variable:
        up = 1.0 + 0.5 * np.sin(0.013 * f + 0.3)
    else:
        up = np.ones(nf)
    lo = None
    if asym != 0.0:
        # convection-like skew: lower = upper*(1+asym*s), upper = upper*(1-asym*s)
        s = np.cos(0.021 * f)
        lo = up * (1.0 + asym * s)
        up = up * (1.0 - asym * s)
    # negSumDiag: diag[l] -= lower-coefficient? reference lduMatrixOperations.C:60-83:
    #   Diag[l[face]] -= Lower[face]; Diag[u[face]] -= Upper[face];
    diag = np.zeros(n)
    lo_eff = up if lo is None else lo
    np.subtract.at(diag, lower, lo_eff)
    np.subtract.at(diag, upper, up)
    if asym != 0.0:
        diag -= 0.05  # transient term keeps the asymmetric system well posed
    idx = np.arange(n, dtype=np.float64)
    if seed is None:
        source = np.sin(0.37 * idx)
    else:
        rng = np.random.Generator(np.random.MT19937(seed))
        source = rng.uniform(-1.0, 1.0, n)
    if dirichlet:
        # fixed-value walls on all six sides: internalCoeffs add -2*gamma*h to diag per wall face
        c = np.arange(n)
        i = c % nx
        j = (c // nx) % ny
        k = c // (nx * ny)
        nwall = ((i == 0).astype(float) + (i == nx - 1) + (j == 0) + (j == ny - 1))
        if nz > 1:
            nwall = nwall + (k == 0) + (k == nz - 1)
        diag -= 2.0 * nwall
    else:
        # fvMatrix::setReference: source[c] += diag[c]*value; diag[c] += diag[c]
        source[ref_cell] += diag[ref_cell] * ref_value
        diag[ref_cell] += diag[ref_cell]
    # faceAreaPair weights for a uniform hex box: |Sf/sqrt(magSf) (x) (1,1.01,1.02)|
    fw = np.array([1.0, 1.01, 1.02])[direction]
    return dict(nCells=n, nFaces=nf, lower=lower, upper=upper, diag=diag,
                upperCoef=up, lowerCoef=lo, source=source,
                psi0=np.zeros(n), faceWeights=fw)


def write_problem(path, sysd, psi=None, source=None, weights=False):
    """Flat binary consumed by oracle/ref_driver.C (format documented there)."""
    asym = sysd["lowerCoef"] is not None
    with open(path, "wb") as fh:
        np.array([0x3155444C, sysd["nCells"], sysd["nFaces"], int(asym), int(weights)],
                 dtype=np.int32).tofile(fh)
        sysd["lower"].astype(np.int32).tofile(fh)
        sysd["upper"].astype(np.int32).tofile(fh)
        sysd["diag"].astype(np.float64).tofile(fh)
        sysd["upperCoef"].astype(np.float64).tofile(fh)
        if asym:
            sysd["lowerCoef"].astype(np.float64).tofile(fh)
        (sysd["source"] if source is None else source).astype(np.float64).tofile(fh)
        (sysd["psi0"] if psi is None else psi).astype(np.float64).tofile(fh)
        if weights:
            sysd["faceWeights"].astype(np.float64).tofile(fh)


def scramble(sysd: dict, seed: int = 1) -> dict:
    """Renumber the cells of a system with a random permutation and restore the
    LDU upper-triangular face order: an unstructured-looking matrix with the
    same spectrum (what renumberMesh / an arbitrary mesher would hand over)."""
    rng = np.random.default_rng(seed)
    n = sysd["nCells"]
    perm = rng.permutation(n)                       # new index of old cell
    lo, up = perm[sysd["lower"]], perm[sysd["upper"]]
    flip = lo > up
    l2, u2 = np.where(flip, up, lo), np.where(flip, lo, up)
    uc = sysd["upperCoef"]
    lc = sysd["lowerCoef"] if sysd["lowerCoef"] is not None else uc
    uc2, lc2 = np.where(flip, lc, uc), np.where(flip, uc, lc)
    order = np.lexsort((u2, l2))
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)                        # old index of new cell
    out = dict(sysd)
    out.update(lower=l2[order].astype(np.int32), upper=u2[order].astype(np.int32),
               upperCoef=uc2[order].copy(),
               lowerCoef=None if sysd["lowerCoef"] is None else lc2[order].copy(),
               diag=sysd["diag"][inv].copy(), source=sysd["source"][inv].copy(),
               psi0=sysd["psi0"][inv].copy(),
               faceWeights=None if sysd.get("faceWeights") is None else sysd["faceWeights"][order].copy())
    return out
