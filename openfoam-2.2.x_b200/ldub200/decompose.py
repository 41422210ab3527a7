"""Domain decomposition of an LDU system into one mesh region per GPU.

What decomposePar + processorFvPatch produce in the reference:
every region keeps the faces internal to it (same relative order, so the LDU
upper-triangular ordering survives) and gets one processor interface per
neighbouring region listing faceCells of the cut faces, both sides in the same
(global face) order.  Sign convention (gaussLaplacianScheme.C:73-79,
coupledFvPatchField.C:171-197, processorFvPatchScalarField.C:127):
    Apsi[faceCell] -= bouCoeffs * psiNbr      =>   bouCoeffs = -(off-diagonal)
and the diagonal of a decomposed matrix equals the undecomposed one because
fvMatrix::addBoundaryDiag adds intCoeffs before the solver sees it.
"""
from __future__ import annotations

import numpy as np


def block_partition(nx, ny, nz, px, py, pz):
    """cell -> region for a px*py*pz block split of an nx*ny*nz lexicographic box
    (decomposePar `simple` method)."""
    c = np.arange(nx * ny * nz, dtype=np.int64)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    bi = np.minimum(i * px // nx, px - 1)
    bj = np.minimum(j * py // ny, py - 1)
    bk = np.minimum(k * pz // nz, pz - 1)
    return ((bk * py + bj) * px + bi).astype(np.int32)


def split_for(n_ranks: int):
    """block counts (px, py, pz) used for 1/2/4/8 ranks (2x1x1, 2x2x1, 2x2x2); 16 and 32 regions
    (2x2x4, 2x4x4: the cuts go where the lexicographic numbering is coarsest) serve the many-core CPU arm."""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 16: (2, 2, 4), 32: (2, 4, 4)}[n_ranks]


def decompose(sysd: dict, proc: np.ndarray, n_regions: int):
    """Split a global system (ldub200.meshes dict) by the cell->region map `proc`.
    Returns a list of region dicts with the same keys plus
      interfaces = [dict(nbrRegion, nbrInterface, faceCells, bouCoeffs, intCoeffs)]
      cells = global index of every local cell."""
    lower, upper = sysd["lower"].astype(np.int64), sysd["upper"].astype(np.int64)
    up = sysd["upperCoef"]
    lo = sysd["lowerCoef"] if sysd["lowerCoef"] is not None else up
    asym = sysd["lowerCoef"] is not None
    pl, pu = proc[lower], proc[upper]
    local = np.empty(sysd["nCells"], dtype=np.int64)
    regions = []
    for r in range(n_regions):
        cells = np.nonzero(proc == r)[0]
        local[cells] = np.arange(cells.size)
        regions.append(dict(cells=cells))
    for r in range(n_regions):
        reg = regions[r]
        cells = reg["cells"]
        internal = np.nonzero((pl == r) & (pu == r))[0]
        reg.update(
            nCells=cells.size, nFaces=internal.size,
            lower=local[lower[internal]].astype(np.int32), upper=local[upper[internal]].astype(np.int32),
            diag=sysd["diag"][cells].copy(), upperCoef=up[internal].copy(),
            lowerCoef=lo[internal].copy() if asym else None,
            source=sysd["source"][cells].copy(), psi0=sysd["psi0"][cells].copy(),
            faceWeights=None if sysd.get("faceWeights") is None else sysd["faceWeights"][internal].copy(),
            interfaces=[])
        cut = np.nonzero(((pl == r) | (pu == r)) & (pl != pu))[0]
        other = np.where(pl[cut] == r, pu[cut], pl[cut])
        for s in np.unique(other):
            f = cut[other == s]                   # global face order on both sides
            own_side = pl[f] == r                 # this region holds the owner cell
            mine = np.where(own_side, lower[f], upper[f])
            off = np.where(own_side, up[f], lo[f])       # row coefficient multiplying the remote psi
            offT = np.where(own_side, lo[f], up[f])      # same for the transpose product
            reg["interfaces"].append(dict(nbrRegion=int(s), faceCells=local[mine].astype(np.int32),
                                          bouCoeffs=-off, intCoeffs=-offT))
    for r in range(n_regions):
        for it in regions[r]["interfaces"]:
            nbr = regions[it["nbrRegion"]]["interfaces"]
            it["nbrInterface"] = [k for k, jt in enumerate(nbr) if jt["nbrRegion"] == r][0]
    return regions


def gather_field(regions, fields, n_cells):
    """per-region fields -> global field"""
    out = np.empty(n_cells)
    for reg, f in zip(regions, fields):
        out[reg["cells"]] = f
    return out


def local_box_region(n, rank, n_ranks):
    """Region `rank` of the uniform-coefficient n^3 Laplacian box (the bench
    workload) built directly, without forming the global system.  Same result as
    decompose(laplacian_system(n,n,n), block_partition(...))[rank]."""
    from . import meshes
    px, py, pz = split_for(n_ranks)
    bi, bj, bk = rank % px, (rank // px) % py, rank // (px * py)

    def span(b, p):
        idx = np.arange(n)
        own = np.minimum(idx * p // n, p - 1)
        w = np.nonzero(own == b)[0]
        return int(w[0]), int(w[-1]) + 1

    (i0, i1), (j0, j1), (k0, k1) = span(bi, px), span(bj, py), span(bk, pz)
    nxl, nyl, nzl = i1 - i0, j1 - j0, k1 - k0
    lower, upper, direction = meshes.box_addressing(nxl, nyl, nzl)
    nl = nxl * nyl * nzl
    c = np.arange(nl, dtype=np.int64)
    li, lj, lk = c % nxl, (c // nxl) % nyl, c // (nxl * nyl)
    gi, gj, gk = li + i0, lj + j0, lk + k0
    gcell = (gk * n + gj) * n + gi
    # global negSumDiag of the 7-point Laplacian with unit coefficients
    nnb = ((gi > 0).astype(float) + (gi < n - 1) + (gj > 0) + (gj < n - 1) + (gk > 0) + (gk < n - 1))
    diag = -nnb
    source = np.sin(0.37 * gcell.astype(np.float64))
    ref = np.nonzero(gcell == 0)[0]
    if ref.size:     # fvMatrix::setReference on global cell 0, value 0
        diag[ref[0]] += diag[ref[0]]
    reg = dict(nCells=nl, nFaces=lower.size, lower=lower, upper=upper, diag=diag,
               upperCoef=np.ones(lower.size), lowerCoef=None, source=source, psi0=np.zeros(nl),
               faceWeights=np.array([1.0, 1.01, 1.02])[direction], cells=gcell, interfaces=[])

    def rank_of(a, b, cc):
        return (cc * py + b) * px + a

    # neighbours in ascending rank order, faces in global face order (owner-major)
    cand = []
    for d, (db, lo_side) in enumerate([((-1, 0, 0), True), ((1, 0, 0), False), ((0, -1, 0), True),
                                       ((0, 1, 0), False), ((0, 0, -1), True), ((0, 0, 1), False)]):
        a, b, cc = bi + db[0], bj + db[1], bk + db[2]
        if not (0 <= a < px and 0 <= b < py and 0 <= cc < pz):
            continue
        axis = d // 2
        coord = (li, lj, lk)[axis]
        ext = (nxl, nyl, nzl)[axis]
        mask = coord == (0 if lo_side else ext - 1)
        cells = np.nonzero(mask)[0]          # ascending local == ascending global order of the owner/neighbour
        cand.append((rank_of(a, b, cc), cells))
    cand.sort(key=lambda t: t[0])
    for s, cells in cand:
        reg["interfaces"].append(dict(nbrRegion=int(s), faceCells=cells.astype(np.int32),
                                      bouCoeffs=-np.ones(cells.size), intCoeffs=-np.ones(cells.size)))
    # index of this rank's interface in each neighbour's (rank-sorted) list
    for it in reg["interfaces"]:
        s = it["nbrRegion"]
        sa, sb, sc = s % px, (s // px) % py, s // (px * py)
        nb = []
        for db in [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]:
            a, b, cc = sa + db[0], sb + db[1], sc + db[2]
            if 0 <= a < px and 0 <= b < py and 0 <= cc < pz:
                nb.append(rank_of(a, b, cc))
        it["nbrInterface"] = sorted(nb).index(rank)
    return reg
