"""Shared case tables: the systems and solver dictionaries the oracle is pinned on
(against the compiled reference and the golden fixtures) and the GPU is checked on."""
import numpy as np

from ldub200 import meshes

SYSTEMS = {
    "cavity20x20": dict(nx=20, ny=20, nz=1),
    "box12_var": dict(nx=12, ny=12, nz=12, variable=True),
    "box9x7x5_dirichlet": dict(nx=9, ny=7, nz=5, variable=True, dirichlet=True),
    "line50": dict(nx=50, ny=1, nz=1),
    "asym10": dict(nx=10, ny=10, nz=10, variable=True, asym=0.3),
    "single": dict(nx=1, ny=1, nz=1, dirichlet=True),
    # more than one 32-line tile per plane and a ragged last tile (structured-box sweeps)
    "box7x41x3": dict(nx=7, ny=41, nz=3, variable=True),
    "asym5x70x2": dict(nx=5, ny=70, nz=2, variable=True, asym=0.25),
    # several stacks of k-planes and two columns (plane-stacked sweeps, stencil2.cu)
    "box6x40x9": dict(nx=6, ny=40, nz=9, variable=True),
    "asym4x35x13": dict(nx=4, ny=35, nz=13, variable=True, asym=0.2),
    # more than 8 row blocks of 512 rows: the TMA-staged Amul / residual / sumA kernel
    "box40x30x20": dict(nx=40, ny=30, nz=20, variable=True),
    "asym33x17x11": dict(nx=33, ny=17, nz=11, variable=True, asym=0.2),
    "scrambled17": dict(nx=17, ny=17, nz=17, variable=True, scramble=7),
    # randomly renumbered cells: no box structure, generic dataflow sweeps
    "scrambled9": dict(nx=9, ny=9, nz=9, variable=True, scramble=3),
    "scrambled_asym8": dict(nx=8, ny=8, nz=8, variable=True, asym=0.3, scramble=5),
}


def system(name):
    kw = dict(SYSTEMS[name])
    seed = kw.pop("scramble", None)
    s = meshes.laplacian_system(**kw)
    return meshes.scramble(s, seed) if seed is not None else s


_GAMG = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=10, mergeLevels=1,
             cacheAgglomeration=False)

# (system, controls)
SOLVES = [
    ("cavity20x20", dict(solver="PCG", preconditioner="DIC", tolerance=1e-6, relTol=0)),
    ("cavity20x20", dict(solver="PCG", preconditioner="DIC", tolerance=1e-10, relTol=0)),
    ("cavity20x20", dict(solver="PCG", preconditioner="diagonal", tolerance=1e-6, relTol=0)),
    ("cavity20x20", dict(solver="PCG", preconditioner="none", tolerance=1e-6, relTol=0)),
    ("cavity20x20", dict(solver="PCG", preconditioner="FDIC", tolerance=1e-8, relTol=0)),
    ("box12_var", dict(solver="PCG", preconditioner="DIC", tolerance=1e-9, relTol=0)),
    ("box12_var", dict(solver="PCG", preconditioner="DIC", tolerance=0, relTol=0, maxIter=10)),
    ("box9x7x5_dirichlet", dict(solver="PCG", preconditioner="diagonal", tolerance=1e-8, relTol=0.001)),
    ("asym10", dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)),
    ("asym10", dict(solver="PBiCG", preconditioner="diagonal", tolerance=1e-8, relTol=0)),
    ("asym10", dict(solver="smoothSolver", smoother="GaussSeidel", nSweeps=2, tolerance=1e-7, relTol=0)),
    ("cavity20x20", dict(solver="smoothSolver", smoother="symGaussSeidel", nSweeps=1, tolerance=1e-4,
                         relTol=0, maxIter=200)),
    ("box12_var", dict(solver="smoothSolver", smoother="DICGaussSeidel", nSweeps=1, tolerance=1e-6, relTol=0)),
    ("box12_var", dict(solver="smoothSolver", smoother="GaussSeidel", nSweeps=-4)),
    ("asym10", dict(solver="smoothSolver", smoother="DILUGaussSeidel", nSweeps=2, tolerance=1e-7, relTol=0)),
    ("line50", dict(solver="PCG", preconditioner="DIC", tolerance=1e-12, relTol=0)),
    ("box7x41x3", dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("box7x41x3", dict(solver="PCG", preconditioner="FDIC", tolerance=1e-8, relTol=0)),
    ("asym5x70x2", dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)),
    ("scrambled9", dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("scrambled_asym8", dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)),
    ("scrambled9", dict(solver="smoothSolver", smoother="symGaussSeidel", nSweeps=2, tolerance=1e-6, relTol=0)),
]

GAMG_SOLVES = [
    ("cavity20x20", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0)),
    ("cavity20x20", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-7, relTol=0.01)),
    ("box12_var", dict(_GAMG, agglomerator="faceAreaPair", mergeLevels=2, tolerance=1e-8, relTol=0, nPreSweeps=1)),
    ("box12_var", dict(_GAMG, smoother="DIC", agglomerator="algebraicPair", nCellsInCoarsestLevel=4,
                       tolerance=1e-8, relTol=0, interpolateCorrection=True)),
    ("box9x7x5_dirichlet", dict(_GAMG, smoother="symGaussSeidel", agglomerator="faceAreaPair",
                                nCellsInCoarsestLevel=20, mergeLevels=3, tolerance=1e-9, relTol=0)),
    ("asym10", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0)),
    ("asym10", dict(_GAMG, smoother="DILU", agglomerator="algebraicPair", tolerance=1e-8, relTol=0, nPreSweeps=2)),
    ("scrambled9", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0)),
    ("box12_var", dict(solver="PCG", tolerance=1e-9, relTol=0,
                       preconditioner=dict(preconditioner="GAMG", smoother="GaussSeidel",
                                           agglomerator="faceAreaPair", nCellsInCoarsestLevel=10,
                                           mergeLevels=1, cacheAgglomeration=False, tolerance=1e-5,
                                           relTol=0, nVcycles=2))),
]

# (system, regions, partition, controls): the system cut into one mesh region per rank
# (ldub200.decompose), solved by the coupled reference (one process per region) / the oracle
# World / the multi-GPU path
MULTI_REGION_SOLVES = [
    ("box6x40x9", 2, "slab", dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("box6x40x9", 4, "slab", dict(solver="PCG", preconditioner="FDIC", tolerance=1e-8, relTol=0)),
    ("box6x40x9", 3, "random", dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("cavity20x20", 2, "slab", dict(solver="PCG", preconditioner="diagonal", tolerance=1e-6, relTol=0)),
    ("asym4x35x13", 4, "slab", dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)),
    ("asym10", 2, "random", dict(solver="PBiCG", preconditioner="diagonal", tolerance=1e-8, relTol=0)),
    ("asym4x35x13", 2, "slab", dict(solver="smoothSolver", smoother="GaussSeidel", nSweeps=2, tolerance=1e-6,
                                     relTol=0, maxIter=50)),
    ("box6x40x9", 2, "slab", dict(solver="smoothSolver", smoother="DICGaussSeidel", nSweeps=1, tolerance=1e-6,
                                   relTol=0, maxIter=50)),
    ("box12_var", 3, "slab", dict(solver="smoothSolver", smoother="symGaussSeidel", nSweeps=2, tolerance=1e-6,
                                   relTol=0, maxIter=40)),
    ("box12_var", 2, "slab", dict(solver="smoothSolver", smoother="nonBlockingGaussSeidel", nSweeps=2,
                                   tolerance=1e-6, relTol=0, maxIter=40)),
    ("box6x40x9", 2, "slab", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0)),
    ("box12_var", 4, "slab", dict(_GAMG, agglomerator="faceAreaPair", mergeLevels=2, tolerance=1e-8, relTol=0)),
    ("box40x30x20", 4, "slab", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0)),
    ("asym33x17x11", 4, "slab", dict(_GAMG, smoother="DILU", agglomerator="faceAreaPair", tolerance=1e-8,
                                      relTol=0)),
    ("box12_var", 2, "random", dict(solver="PCG", tolerance=1e-9, relTol=0,
                                     preconditioner=dict(preconditioner="GAMG", smoother="GaussSeidel",
                                                         agglomerator="faceAreaPair", nCellsInCoarsestLevel=10,
                                                         mergeLevels=1, cacheAgglomeration=False, tolerance=1e-5,
                                                         relTol=0, nVcycles=2))),
]


def regions(name, n_regions, partition):
    """system `name` cut into n_regions mesh regions: contiguous cell ranges ("slab") or a
    seeded random cell -> region map ("random": every region touches every other one)."""
    from ldub200 import decompose
    s = system(name)
    n = s["nCells"]
    if partition == "slab":
        proc = (np.arange(n) * n_regions // n).astype(np.int32)
    else:
        proc = np.random.default_rng(n_regions).integers(0, n_regions, n).astype(np.int32)
    return s, decompose.decompose(s, proc, n_regions)


# periodic (cyclic) cases: (system, axis, controls) for one region, (system, regions, axis, controls)
# for z-slab regions that are each periodic along `axis`
CYCLIC_SYSTEMS = [("box12_var", 0), ("asym10", 1), ("cavity20x20", 1), ("box6x40x9", 2)]
CYCLIC_SOLVES = [
    ("box12_var", 0, dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("cavity20x20", 1, dict(solver="PCG", preconditioner="FDIC", tolerance=1e-8, relTol=0)),
    ("asym10", 1, dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)),
    ("box6x40x9", 2, dict(solver="smoothSolver", smoother="symGaussSeidel", nSweeps=2, tolerance=1e-6,
                          relTol=0, maxIter=60)),
    ("box12_var", 0, dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0)),
    ("asym10", 1, dict(_GAMG, smoother="DILU", agglomerator="algebraicPair", tolerance=1e-8, relTol=0)),
    ("box6x40x9", 2, dict(_GAMG, smoother="nonBlockingGaussSeidel", agglomerator="faceAreaPair", mergeLevels=2,
                          tolerance=1e-8, relTol=0)),
]
CYCLIC_REGION_SOLVES = [
    ("box12_var", 3, 0, dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)),
    ("asym10", 2, 1, dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)),
    ("box12_var", 2, 1, dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0)),
    ("asym10", 2, 0, dict(solver="smoothSolver", smoother="nonBlockingGaussSeidel", nSweeps=2, tolerance=1e-6,
                          relTol=0, maxIter=40)),
]


# edge cases of the solver front end (lduMatrixSolver.C:40-136, PCG.C:84-181): (system, controls,
# initial guess, source).  Systems "faceless6" / "diagonal6": six cells without faces, the first with an
# (empty) upper field as fvm::laplacian leaves it -> the selected solver runs; the second without upper
# and lower -> lduMatrix::diagonal() -> diagonalSolver whatever the dictionary says.
_PCG = dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0)
_BICG = dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-8, relTol=0)
EDGE_SOLVES = [
    ("faceless6", _PCG, "zero", "given"),
    ("faceless6", dict(solver="smoothSolver", smoother="GaussSeidel", tolerance=1e-8, relTol=0), "zero", "given"),
    ("diagonal6", _PCG, "zero", "given"),
    ("diagonal6", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0), "random", "given"),
    ("single", _PCG, "random", "given"),
    ("box12_var", dict(_PCG, maxIter=0), "zero", "given"),          # the loop body still runs once
    ("box12_var", dict(_PCG, maxIter=1), "zero", "given"),
    ("box12_var", _PCG, "random", "given"),
    ("box12_var", _PCG, "exact", "given"),                          # converged before the first iteration
    ("box12_var", _PCG, "zero", "zero"),                            # normFactor = 1e-20 only
    ("box12_var", _PCG, "random", "zero"),
    ("box12_var", dict(solver="PCG", preconditioner="DIC", tolerance=0, relTol=0.05), "zero", "given"),
    ("box12_var", dict(solver="PCG", preconditioner="DIC", tolerance=10.0, relTol=0), "zero", "given"),
    ("box12_var", dict(solver="smoothSolver", smoother="GaussSeidel", nSweeps=5, tolerance=1e-12, relTol=0,
                       maxIter=7), "zero", "given"),                # nSweeps does not divide maxIter
    ("box12_var", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0, maxIter=0), "zero", "given"),
    ("asym10", _BICG, "random", "given"),
    ("asym10", _BICG, "zero", "zero"),
    ("asym10", dict(_BICG, maxIter=1), "zero", "given"),
]


# GAMG controls beyond GAMG_SOLVES (GAMGSolver.C:157-181): sweep counts with level multipliers and caps,
# explicit scaleCorrection on either matrix type, interpolateCorrection on an asymmetric matrix, the
# remaining smoothers on the levels, GAMG preconditioner variants.  Pinned on the oracle against the
# reference (CPU); the CUDA path reads the same controls and is checked on these in the next round.
GAMG_OPTION_SOLVES = [
    ("box12_var", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0, nPreSweeps=1,
                       preSweepsLevelMultiplier=2, maxPreSweeps=3)),
    ("box12_var", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0, nPostSweeps=1,
                       postSweepsLevelMultiplier=3, maxPostSweeps=5, nFinestSweeps=3)),
    ("box12_var", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0, scaleCorrection=False)),
    ("asym10", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0, scaleCorrection=True)),
    ("asym10", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0, interpolateCorrection=True,
                    nPreSweeps=1)),
    ("box9x7x5_dirichlet", dict(_GAMG, smoother="DICGaussSeidel", agglomerator="faceAreaPair", tolerance=1e-9,
                                relTol=0, nFinestSweeps=1, nPostSweeps=3)),
    ("asym10", dict(_GAMG, smoother="DILUGaussSeidel", agglomerator="faceAreaPair", tolerance=1e-8, relTol=0,
                    mergeLevels=2)),
    ("cavity20x20", dict(_GAMG, smoother="FDIC", agglomerator="algebraicPair", tolerance=1e-8, relTol=0,
                         nCellsInCoarsestLevel=50)),
    ("scrambled17", dict(_GAMG, smoother="symGaussSeidel", agglomerator="faceAreaPair", tolerance=1e-7, relTol=0,
                         mergeLevels=2, nPreSweeps=2, maxPreSweeps=2)),
    ("box12_var", dict(solver="PCG", tolerance=1e-9, relTol=0,
                       preconditioner=dict(preconditioner="GAMG", smoother="DIC", agglomerator="algebraicPair",
                                           nCellsInCoarsestLevel=10, mergeLevels=1, cacheAgglomeration=False,
                                           tolerance=1e-5, relTol=0, nVcycles=1, nPreSweeps=1))),
]


# two solves on one mesh, coefficients changed in between (oracle.second_coeffs): with
# cacheAgglomeration on, the second GAMG solve reuses the agglomeration of the first (a MeshObject in
# the reference, GAMGSolver.C:70,144-154) -- for algebraicPair that is a different hierarchy than a
# fresh one (18/8 iterations without the cache, 18/13 with it on box12_var)
CACHE_SOLVES = [
    ("box12_var", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0, cacheAgglomeration=True)),
    ("box12_var", dict(_GAMG, agglomerator="algebraicPair", tolerance=1e-8, relTol=0, cacheAgglomeration=False)),
    ("box12_var", dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-8, relTol=0, cacheAgglomeration=True)),
    ("asym10", dict(_GAMG, smoother="DILU", agglomerator="algebraicPair", tolerance=1e-8, relTol=0,
                    cacheAgglomeration=True)),
    ("box12_var", dict(solver="PCG", tolerance=1e-9, relTol=0,
                       preconditioner=dict(preconditioner="GAMG", smoother="GaussSeidel", agglomerator="algebraicPair",
                                           nCellsInCoarsestLevel=10, mergeLevels=1, cacheAgglomeration=True,
                                           tolerance=1e-5, relTol=0, nVcycles=2))),
]


# singular systems (SolverPerformance::checkSingularity, SolverPerformance.C:31-52): an all-zero matrix
# makes wApA vanish in the first iteration -> singular, not converged, no iteration counted
SINGULAR_SOLVES = [
    ("box12_var", dict(solver="PCG", preconditioner="none", tolerance=1e-8, relTol=0)),
    ("asym10", dict(solver="PBiCG", preconditioner="none", tolerance=1e-8, relTol=0)),
]


def singular_case(i):
    name, ctl = SINGULAR_SOLVES[i]
    s = dict(system(name))
    s["diag"] = np.zeros_like(s["diag"])
    s["upperCoef"] = np.zeros_like(s["upperCoef"])
    if s["lowerCoef"] is not None:
        s["lowerCoef"] = np.zeros_like(s["lowerCoef"])
    return s, ctl


def edge_case(i):
    """-> (system dict, controls, psi0, source) of EDGE_SOLVES[i]"""
    name, ctl, guess, src = EDGE_SOLVES[i]
    if name in ("faceless6", "diagonal6"):
        n = 6
        s = dict(nCells=n, nFaces=0, lower=np.zeros(0, np.int32), upper=np.zeros(0, np.int32),
                 diag=-(1.0 + np.arange(n)), upperCoef=np.zeros(0) if name == "faceless6" else None,
                 lowerCoef=None, source=np.sin(np.arange(n) + 1.0), psi0=np.zeros(n), faceWeights=None)
    else:
        s = system(name)
    rng = np.random.default_rng(100 + i)
    source = np.zeros(s["nCells"]) if src == "zero" else s["source"]
    if guess == "random":
        psi0 = rng.standard_normal(s["nCells"])
    elif guess == "exact":
        from oracle import oracle as O
        psi0 = O.World([s]).solve(dict(ctl, tolerance=1e-14), s["psi0"].copy(), source)[0][0]
    else:
        psi0 = np.zeros(s["nCells"])
    return s, ctl, psi0, source


# a real unstructured mesh: the polyMesh shipped with the reference's airFoil2D tutorial (10,720 cells,
# 21,254 internal faces), read by ldub200.polymesh; tests/golden/airfoil2d.npz carries the system built
# from it and the reference's results, so the tests run where /root/reference is absent
AIRFOIL_POLYMESH = "/root/reference/tutorials/incompressible/simpleFoam/airFoil2D/constant/polyMesh"
AIRFOIL_SOLVES = [
    dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0),
    dict(solver="PCG", preconditioner="FDIC", tolerance=1e-6, relTol=0),
    dict(_GAMG, agglomerator="faceAreaPair", tolerance=1e-7, relTol=0, maxIter=60),
    dict(solver="smoothSolver", smoother="symGaussSeidel", nSweeps=2, tolerance=1e-6, relTol=0, maxIter=30),
]


def digest(a):
    """SHA-256 of an array's bytes as a uint8 array (bit-exact comparison in a small fixture)"""
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).digest(), dtype=np.uint8)


def airfoil_x(n):
    return np.sin(0.11 * np.arange(n)) + 0.3 * np.cos(0.013 * np.arange(n))


def airfoil_system():
    """the airFoil2D Laplacian system from the committed fixture (same dict layout as system())"""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "airfoil2d.npz")
    n = g["diag"].size
    return dict(nCells=n, nFaces=g["lower"].size, lower=g["lower"], upper=g["upper"], diag=g["diag"],
                upperCoef=g["upperCoef"], lowerCoef=None, source=g["source"], psi0=np.zeros(n),
                faceWeights=g["faceWeights"]), g


def _add_cyclic(reg, region, gcells, dims, axis, up, lo):
    """append the two halves of a cyclic pair along `axis` to region dict `reg` (cells given by their
    global index gcells in the nx*ny*nz box): end plane 0 is the owner half, plane n-1 the other."""
    nx, ny, nz = dims
    coord = (gcells % nx, (gcells // nx) % ny, gcells // (nx * ny))[axis]
    a = np.nonzero(coord == 0)[0]
    b = np.nonzero(coord == dims[axis] - 1)[0]
    assert a.size == b.size and a.size > 0
    k = 0.6 + 0.3 * np.sin(0.9 * np.arange(a.size))            # coupling coefficient per face pair
    ku, kl = up * k, lo * k
    reg["diag"] = reg["diag"].copy()
    reg["diag"][a] -= ku
    reg["diag"][b] -= ku
    first = len(reg["interfaces"])
    reg["interfaces"].append(dict(nbrRegion=region, nbrInterface=first + 1, faceCells=a.astype(np.int32),
                                  bouCoeffs=-ku, intCoeffs=-kl))
    reg["interfaces"].append(dict(nbrRegion=region, nbrInterface=first, faceCells=b.astype(np.int32),
                                  bouCoeffs=-kl, intCoeffs=-ku))


def cyclic_system(name, axis=0):
    """box system `name` made periodic along `axis`: its two end planes become the halves of a cyclic
    patch pair, i.e. two interfaces of the single region that point at each other."""
    kw = SYSTEMS[name]
    s = dict(system(name))
    s["interfaces"] = []
    asym = s["lowerCoef"] is not None
    _add_cyclic(s, 0, np.arange(s["nCells"]), (kw["nx"], kw["ny"], kw["nz"]), axis, 1.0, 0.8 if asym else 1.0)
    return s


def cyclic_regions(name, n_regions, axis=0):
    """box system `name` cut into z-slabs (one region per rank), every region periodic along `axis`:
    processor interfaces between the slabs plus a cyclic pair inside each region."""
    from ldub200 import decompose
    kw = SYSTEMS[name]
    dims = (kw["nx"], kw["ny"], kw["nz"])
    s = system(name)
    regs = decompose.decompose(s, decompose.block_partition(*dims, 1, 1, n_regions), n_regions)
    asym = s["lowerCoef"] is not None
    for r, reg in enumerate(regs):
        _add_cyclic(reg, r, reg["cells"], dims, axis, 1.0, 0.8 if asym else 1.0)
    return s, regs


PRECONDITIONERS = ["none", "diagonal", "DIC", "FDIC", "DILU"]
SMOOTHERS = ["GaussSeidel", "symGaussSeidel", "DIC", "DILU", "FDIC", "DICGaussSeidel",
             "DILUGaussSeidel", "nonBlockingGaussSeidel"]
SYMMETRIC_ONLY = {"DIC", "FDIC", "DICGaussSeidel"}
ASYMMETRIC_ONLY = {"DILU", "DILUGaussSeidel"}


def selectable(s, name):
    """run-time selection tables of the reference (symMatrix / asymMatrix)"""
    asym = s["lowerCoef"] is not None
    return not ((asym and name in SYMMETRIC_ONLY) or (not asym and name in ASYMMETRIC_ONLY))


def ref_controls(controls):
    """The compiled reference has no faceAreaPair (libfiniteVolume): its driver
    registers `weightedPair`, the same pairGAMGAgglomeration fed with the
    problem file's face weights (oracle/ref_driver.C)."""
    def fix(d):
        d = dict(d)
        if d.get("agglomerator") == "faceAreaPair":
            d["agglomerator"] = "weightedPair"
        if isinstance(d.get("preconditioner"), dict):
            d["preconditioner"] = fix(d["preconditioner"])
        return d
    return fix(controls)
