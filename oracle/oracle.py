"""ctypes front end of oracle/libldu_oracle.so (the CPU restatement) and of the
compiled reference drivers: oracle/_ref/ref_driver (one process) and
oracle/_ref/ref_driver_par (one process per mesh region, coupled through the
shared-memory Pstream of oracle/pstream_shm).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libldu_oracle.so"
REF_DRIVER = HERE / "_ref" / "ref_driver"

SOLVERS = {"PCG": 0, "PBiCG": 1, "smoothSolver": 2, "GAMG": 3, "diagonal": 4}
PRECONDS = {"none": 0, "diagonal": 1, "DIC": 2, "FDIC": 3, "DILU": 4, "GAMG": 5}
SMOOTHERS = {"GaussSeidel": 0, "symGaussSeidel": 1, "DIC": 2, "DILU": 3,
             "DICGaussSeidel": 4, "DILUGaussSeidel": 5, "FDIC": 6,
             "nonBlockingGaussSeidel": 7, "multiColourGaussSeidel": 8}


class Controls(C.Structure):
    _fields_ = [
        ("solver", C.c_int), ("preconditioner", C.c_int), ("smoother", C.c_int),
        ("maxIter", C.c_int), ("tolerance", C.c_double), ("relTol", C.c_double),
        ("nSweeps", C.c_int),
        ("nCellsInCoarsestLevel", C.c_int), ("mergeLevels", C.c_int),
        ("nPreSweeps", C.c_int), ("preSweepsLevelMultiplier", C.c_int), ("maxPreSweeps", C.c_int),
        ("nPostSweeps", C.c_int), ("postSweepsLevelMultiplier", C.c_int), ("maxPostSweeps", C.c_int),
        ("nFinestSweeps", C.c_int), ("interpolateCorrection", C.c_int),
        ("scaleCorrection", C.c_int), ("nVcycles", C.c_int),
        ("precTolerance", C.c_double), ("precRelTol", C.c_double),
        ("useFaceWeights", C.c_int), ("cacheAgglomeration", C.c_int),
    ]


class Perf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("nIterations", C.c_int), ("converged", C.c_int), ("singular", C.c_int)]


def build(force: bool = False) -> Path:
    src = HERE / "ldu_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(
            src.stat().st_mtime, (HERE / "ldu_oracle.h").stat().st_mtime):
        subprocess.check_call(["make", "-s", "-C", str(HERE), "libldu_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        PP = C.POINTER(C.POINTER(C.c_double))
        L.orc_world_new.restype = C.c_void_p
        L.orc_world_new.argtypes = [C.c_int]
        L.orc_world_free.argtypes = [C.c_void_p]
        L.orc_world_set_region.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.orc_world_add_interface.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3
        L.orc_world_add_interface.restype = C.c_int
        L.orc_world_set_face_weights.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_world_set_diagonal.argtypes = [C.c_void_p, C.c_int]
        for name in ("orc_amul", "orc_tmul", "orc_H", "orc_faceH"):
            getattr(L, name).argtypes = [C.c_void_p, PP, PP]
        L.orc_sumA.argtypes = [C.c_void_p, PP]
        L.orc_H1.argtypes = [C.c_void_p, PP]
        L.orc_residual.argtypes = [C.c_void_p, PP, PP, PP]
        L.orc_normFactor.argtypes = [C.c_void_p, PP, PP, PP]
        L.orc_normFactor.restype = C.c_double
        L.orc_gSumProd.argtypes = [C.c_void_p, PP, PP]
        L.orc_gSumProd.restype = C.c_double
        L.orc_gSumMag.argtypes = [C.c_void_p, PP]
        L.orc_gSumMag.restype = C.c_double
        L.orc_precondition.argtypes = [C.c_void_p, C.c_int, PP, PP, C.c_int]
        L.orc_smooth.argtypes = [C.c_void_p, C.c_int, PP, PP, C.c_int]
        L.orc_solve.argtypes = [C.c_void_p, C.POINTER(Controls), PP, PP, C.POINTER(Perf),
                                C.c_void_p, C.c_int]
        L.orc_gamg_build.argtypes = [C.c_void_p, C.POINTER(Controls)]
        for name in ("orc_gamg_nlevels",):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        for name in ("orc_gamg_level_ncells", "orc_gamg_level_nfaces", "orc_gamg_level_nfine"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
        for name in ("orc_gamg_restrict", "orc_gamg_face_restrict", "orc_gamg_level_lower",
                     "orc_gamg_level_upper"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
            getattr(L, name).restype = C.POINTER(C.c_int)
        for name in ("orc_gamg_level_diag", "orc_gamg_level_upperCoef", "orc_gamg_level_lowerCoef"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_int]
            getattr(L, name).restype = C.POINTER(C.c_double)
        _lib = L
    return _lib


def make_controls(d: dict) -> Controls:
    """OpenFOAM-style solver dictionary -> Controls (defaults as in the reference:
    lduMatrixSolver.C:164-169, smoothSolver.C:70-74, GAMGSolver.C:66-76,157-181)."""
    c = Controls()
    # ICCG / BICCG selected at run time are PCG / PBiCG on the same dictionary (ICCG.C:67-86)
    c.solver = SOLVERS[{"ICCG": "PCG", "BICCG": "PBiCG"}.get(d.get("solver", "PCG"), d.get("solver", "PCG"))]
    pre = d.get("preconditioner", "none")
    sub = {}
    if isinstance(pre, dict):
        sub = pre
        pre = pre["preconditioner"]
    c.preconditioner = PRECONDS[pre]
    src = dict(d)
    if c.preconditioner == PRECONDS["GAMG"]:
        src = dict(sub)
        c.precTolerance = float(sub.get("tolerance", 1e-6))
        c.precRelTol = float(sub.get("relTol", 0.0))
    c.smoother = SMOOTHERS[src.get("smoother", "GaussSeidel")]
    c.maxIter = int(d.get("maxIter", 1000))
    c.tolerance = float(d.get("tolerance", 1e-6))
    c.relTol = float(d.get("relTol", 0.0))
    c.nSweeps = int(d.get("nSweeps", 1))
    c.nCellsInCoarsestLevel = int(src.get("nCellsInCoarsestLevel", 10))
    c.mergeLevels = int(src.get("mergeLevels", 1))
    c.nPreSweeps = int(src.get("nPreSweeps", 0))
    c.preSweepsLevelMultiplier = int(src.get("preSweepsLevelMultiplier", 1))
    c.maxPreSweeps = int(src.get("maxPreSweeps", 4))
    c.nPostSweeps = int(src.get("nPostSweeps", 2))
    c.postSweepsLevelMultiplier = int(src.get("postSweepsLevelMultiplier", 1))
    c.maxPostSweeps = int(src.get("maxPostSweeps", 4))
    c.nFinestSweeps = int(src.get("nFinestSweeps", 2))
    c.interpolateCorrection = int(bool(src.get("interpolateCorrection", False)))
    sc = src.get("scaleCorrection", None)
    c.scaleCorrection = -1 if sc is None else int(bool(sc))
    c.nVcycles = int(src.get("nVcycles", 2))
    c.useFaceWeights = int(src.get("agglomerator", "faceAreaPair") == "faceAreaPair")
    c.cacheAgglomeration = int(bool(src.get("cacheAgglomeration", False)))
    return c


def _pp(arrs):
    arr_t = C.POINTER(C.c_double) * len(arrs)
    return arr_t(*[a.ctypes.data_as(C.POINTER(C.c_double)) for a in arrs])


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class World:
    """regions: list of dicts with lower, upper, diag, upperCoef, lowerCoef|None,
    optional faceWeights, optional interfaces = list of dict(nbrRegion,
    nbrInterface, faceCells, bouCoeffs, intCoeffs)."""

    def __init__(self, regions):
        L = lib()
        self.L = L
        self.R = len(regions)
        self._keep = []
        self.w = L.orc_world_new(self.R)
        self.nCells = []
        self.nFaces = []
        self._coeffs = []
        for r, reg in enumerate(regions):
            lo = np.ascontiguousarray(reg["lower"], dtype=np.int32)
            up = np.ascontiguousarray(reg["upper"], dtype=np.int32)
            diag = _f64(reg["diag"])
            uc = _f64(np.zeros(0) if reg["upperCoef"] is None else reg["upperCoef"])
            lc = None if reg.get("lowerCoef") is None else _f64(reg["lowerCoef"])
            # own copies: set_coeffs() overwrites them in place
            diag, uc, lc = diag.copy(), uc.copy(), None if lc is None else lc.copy()
            self._keep += [lo, up, diag, uc, lc]
            self._coeffs.append((diag, uc, lc))
            self.nCells.append(diag.size)
            self.nFaces.append(lo.size)
            L.orc_world_set_region(self.w, r, diag.size, lo.size, lo.ctypes.data, up.ctypes.data,
                                   diag.ctypes.data, uc.ctypes.data,
                                   None if lc is None else lc.ctypes.data)
            if reg.get("faceWeights") is not None:
                fw = _f64(reg["faceWeights"])
                self._keep.append(fw)
                L.orc_world_set_face_weights(self.w, r, fw.ctypes.data)
        # lduMatrix::diagonal(): upper and lower were never set (only possible without faces)
        L.orc_world_set_diagonal(self.w, int(all(reg["upperCoef"] is None and reg.get("lowerCoef") is None
                                                 for reg in regions)))
        for r, reg in enumerate(regions):
            for it in reg.get("interfaces", []):
                fc = np.ascontiguousarray(it["faceCells"], dtype=np.int32)
                bou = _f64(it["bouCoeffs"])
                intc = _f64(it["intCoeffs"])
                self._keep += [fc, bou, intc]
                L.orc_world_add_interface(self.w, r, int(it["nbrRegion"]), int(it["nbrInterface"]),
                                          fc.size, fc.ctypes.data, bou.ctypes.data, intc.ctypes.data)

    def set_coeffs(self, r, diag, upperCoef, lowerCoef=None):
        """new coefficients on the same addressing (the arrays the C side borrows are updated in place)"""
        d, u, lo = self._coeffs[r]
        d[:] = diag
        u[:] = upperCoef
        if lo is not None:
            lo[:] = lowerCoef

    def __del__(self):
        try:
            self.L.orc_world_free(self.w)
        except Exception:
            pass

    def _new(self):
        return [np.zeros(n) for n in self.nCells]

    @staticmethod
    def _as_list(x):
        if isinstance(x, np.ndarray):
            return [_f64(x)]
        return [_f64(a) for a in x]

    def amul(self, psi):
        psi = self._as_list(psi)
        out = self._new()
        self.L.orc_amul(self.w, _pp(out), _pp(psi))
        return out

    def tmul(self, psi):
        psi = self._as_list(psi)
        out = self._new()
        self.L.orc_tmul(self.w, _pp(out), _pp(psi))
        return out

    def sumA(self):
        out = self._new()
        self.L.orc_sumA(self.w, _pp(out))
        return out

    def H(self, psi):
        """lduMatrix::H (lduMatrixTemplates.C:33-65)"""
        psi = self._as_list(psi)
        out = self._new()
        self.L.orc_H(self.w, _pp(out), _pp(psi))
        return out

    def H1(self):
        """lduMatrix::H1 (lduMatrixATmul.C:298-327)"""
        out = self._new()
        self.L.orc_H1(self.w, _pp(out))
        return out

    def faceH(self, psi):
        """lduMatrix::faceH (lduMatrixTemplates.C:79-113): one value per face"""
        psi = self._as_list(psi)
        out = [np.zeros(n) for n in self.nFaces]
        self.L.orc_faceH(self.w, _pp(out), _pp(psi))
        return out

    def residual(self, psi, source):
        psi = self._as_list(psi)
        source = self._as_list(source)
        out = self._new()
        self.L.orc_residual(self.w, _pp(out), _pp(psi), _pp(source))
        return out

    def normFactor(self, psi, source, Apsi):
        return self.L.orc_normFactor(self.w, _pp(self._as_list(psi)), _pp(self._as_list(source)),
                                     _pp(self._as_list(Apsi)))

    def precondition(self, name, rA, transpose=False):
        rA = self._as_list(rA)
        out = self._new()
        rc = self.L.orc_precondition(self.w, PRECONDS[name], _pp(out), _pp(rA), int(transpose))
        assert rc == 0
        return out

    def smooth(self, name, psi, source, nSweeps):
        psi = [a.copy() for a in self._as_list(psi)]
        source = self._as_list(source)
        rc = self.L.orc_smooth(self.w, SMOOTHERS[name], _pp(psi), _pp(source), int(nSweeps))
        assert rc == 0
        return psi

    def solve(self, controls: dict, psi, source, hist_cap=0):
        c = make_controls(controls)
        psi = [a.copy() for a in self._as_list(psi)]
        source = self._as_list(source)
        perf = Perf()
        hist = np.zeros(max(hist_cap, 1))
        rc = self.L.orc_solve(self.w, C.byref(c), _pp(psi), _pp(source), C.byref(perf),
                              hist.ctypes.data if hist_cap else None, hist_cap)
        assert rc == 0, "oracle solve failed"
        out = dict(initialResidual=perf.initialResidual, finalResidual=perf.finalResidual,
                   nIterations=perf.nIterations, converged=bool(perf.converged),
                   singular=bool(perf.singular))
        if hist_cap:
            out["history"] = hist[:min(hist_cap, perf.nIterations + 1)].copy()
        return psi, out

    def gamg_levels(self, controls: dict, r=0):
        c = make_controls(controls)
        rc = self.L.orc_gamg_build(self.w, C.byref(c))
        if rc != 0:
            return []
        L = self.L
        out = []
        for lev in range(L.orc_gamg_nlevels(self.w, r)):
            nf = L.orc_gamg_level_nfine(self.w, r, lev)
            nc = L.orc_gamg_level_ncells(self.w, r, lev)
            nfa = L.orc_gamg_level_nfaces(self.w, r, lev)
            out.append(dict(
                nFine=nf, nCoarse=nc, nFaces=nfa,
                restrict=np.ctypeslib.as_array(L.orc_gamg_restrict(self.w, r, lev), (nf,)).copy(),
                lower=np.ctypeslib.as_array(L.orc_gamg_level_lower(self.w, r, lev), (max(nfa, 1),))[:nfa].copy(),
                upper=np.ctypeslib.as_array(L.orc_gamg_level_upper(self.w, r, lev), (max(nfa, 1),))[:nfa].copy(),
                diag=np.ctypeslib.as_array(L.orc_gamg_level_diag(self.w, r, lev), (nc,)).copy(),
                upperCoef=np.ctypeslib.as_array(L.orc_gamg_level_upperCoef(self.w, r, lev), (max(nfa, 1),))[:nfa].copy(),
            ))
        return out


def world_from_system(sysd) -> World:
    return World([sysd])


# --------------------------------------------------------------------------- #
# the compiled reference itself (oracle/_ref), single region
# --------------------------------------------------------------------------- #

def ref_available() -> bool:
    return REF_DRIVER.exists() and (HERE / "_ref" / "libOpenFOAM.so").exists()


def dict_text(d: dict) -> str:
    """python dict -> OpenFOAM dictionary text."""
    out = []
    for k, v in d.items():
        if isinstance(v, dict):
            out.append(f"{k} {{ {dict_text(v)} }}")
        elif isinstance(v, bool):
            out.append(f"{k} {'on' if v else 'off'};")
        else:
            out.append(f"{k} {v};")
    return " ".join(out)


def _write_problem(path, sysd, psi=None, source=None):
    asym = sysd.get("lowerCoef") is not None
    weights = sysd.get("faceWeights") is not None
    n = np.asarray(sysd["diag"]).size
    with open(path, "wb") as fh:
        diagonal = sysd["upperCoef"] is None
        np.array([0x3155444C, n, np.asarray(sysd["lower"]).size, int(asym), int(weights) | (2 if diagonal else 0)],
                 dtype=np.int32).tofile(fh)
        np.asarray(sysd["lower"], dtype=np.int32).tofile(fh)
        np.asarray(sysd["upper"], dtype=np.int32).tofile(fh)
        _f64(sysd["diag"]).tofile(fh)
        _f64(np.zeros(0) if diagonal else sysd["upperCoef"]).tofile(fh)
        if asym:
            _f64(sysd["lowerCoef"]).tofile(fh)
        _f64(sysd["source"] if source is None else source).tofile(fh)
        _f64(np.zeros(n) if psi is None else psi).tofile(fh)
        if weights:
            _f64(sysd["faceWeights"]).tofile(fh)


def ref_run(sysd, op, *args, psi=None, source=None, ints=False, timeout=3600, extra_env=None):
    """Run the unmodified reference on a single-region system (cyclic interfaces allowed:
    interfaces with nbrRegion 0).  Returns (array, stdout)."""
    env = dict(os.environ, WM_PROJECT_DIR=str(HERE / "_ref"), **(extra_env or {}))
    ld = env.get("LD_LIBRARY_PATH", "")
    env["LD_LIBRARY_PATH"] = str(HERE / "_ref") + (":" + ld if ld else "")
    with tempfile.TemporaryDirectory() as td:
        prob = os.path.join(td, "p.bin")
        outp = os.path.join(td, "o.bin")
        if sysd.get("interfaces"):
            _write_region(prob, sysd, psi=psi, source=source)
        else:
            _write_problem(prob, sysd, psi=psi, source=source)
        r = subprocess.run([str(REF_DRIVER), prob, outp, op, *[str(a) for a in args]],
                           env=env, capture_output=True, text=True, timeout=timeout, cwd=td)
        if r.returncode != 0:
            raise RuntimeError(f"ref_driver failed ({r.returncode}): {r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        out = np.fromfile(outp, dtype=np.int32 if ints else np.float64)
    return out, r.stdout


REF_DRIVER_PAR = HERE / "_ref" / "ref_driver_par"
_SHM_WORLD_BYTES = 1100 << 20    # >= sizeof(lduShm::World) (oracle/pstream_shm/shmWorld.H); sparse


def ref_par_available() -> bool:
    return REF_DRIVER_PAR.exists() and (HERE / "_ref" / "libPstream_shm.so").exists()


def _write_region(path, reg, psi=None, source=None, rank=0):
    """'LDU2' problem file of one mesh region (oracle/ref_driver.C header).  An interface whose
    neighbour is the region itself is one half of a cyclic pair."""
    asym = reg.get("lowerCoef") is not None
    weights = reg.get("faceWeights") is not None
    n = np.asarray(reg["diag"]).size
    its = reg.get("interfaces", [])
    with open(path, "wb") as fh:
        np.array([0x3255444C, n, np.asarray(reg["lower"]).size, int(asym), int(weights), len(its)],
                 dtype=np.int32).tofile(fh)
        np.asarray(reg["lower"], dtype=np.int32).tofile(fh)
        np.asarray(reg["upper"], dtype=np.int32).tofile(fh)
        _f64(reg["diag"]).tofile(fh)
        _f64(reg["upperCoef"]).tofile(fh)
        if asym:
            _f64(reg["lowerCoef"]).tofile(fh)
        _f64(reg["source"] if source is None else source).tofile(fh)
        _f64(np.zeros(n) if psi is None else psi).tofile(fh)
        if weights:
            _f64(reg["faceWeights"]).tofile(fh)
        for it in its:
            fc = np.asarray(it["faceCells"], dtype=np.int32)
            nbr = it["nbrRegion"] if it["nbrRegion"] != rank else -1 - it["nbrInterface"]
            np.array([nbr, fc.size], dtype=np.int32).tofile(fh)
            fc.tofile(fh)
            _f64(it["bouCoeffs"]).tofile(fh)
            _f64(it["intCoeffs"]).tofile(fh)


def ref_run_par(regions, op, *args, psi=None, source=None, ints=False, timeout=3600, extra_env=None):
    """Run the unmodified reference as one process per mesh region, coupled through
    the shared-memory Pstream (oracle/pstream_shm).  psi/source: per-region lists.
    Returns (list of per-region arrays, stdout of rank 0)."""
    n = len(regions)
    env = dict(os.environ, WM_PROJECT_DIR=str(HERE / "_ref"), **(extra_env or {}))
    ld = env.get("LD_LIBRARY_PATH", "")
    env["LD_LIBRARY_PATH"] = str(HERE / "_ref") + (":" + ld if ld else "")
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory() as td, tempfile.NamedTemporaryFile(dir=shm_dir, prefix="ldu_pstream_") as seg:
        seg.truncate(_SHM_WORLD_BYTES)
        seg.flush()
        for r, reg in enumerate(regions):
            _write_region(os.path.join(td, f"p{r}.bin"), reg, rank=r,
                          psi=None if psi is None else psi[r], source=None if source is None else source[r])
        procs = []
        for r in range(n):
            e = dict(env, LDU_PSTREAM_SHM=seg.name, LDU_PSTREAM_RANK=str(r), LDU_PSTREAM_SIZE=str(n))
            procs.append(subprocess.Popen(
                [str(REF_DRIVER_PAR), os.path.join(td, "p%d.bin"), os.path.join(td, "o%d.bin"), op,
                 *[str(a) for a in args]],
                env=e, stdout=subprocess.PIPE, stderr=None if os.environ.get("LDU_PSTREAM_STATS") else subprocess.PIPE, text=True, cwd=td))
        outs = []
        failed = None
        for r, p in enumerate(procs):
            try:
                so, se = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(so)
            if p.returncode != 0 and failed is None:
                failed = (r, p.returncode, so, se)
                for q in procs:          # the others would wait for this rank for ever
                    if q.poll() is None:
                        q.kill()
        if failed:
            r, rc, so, se = failed
            raise RuntimeError(f"ref_driver_par rank {r} failed ({rc}): {so[-2000:]}\n{se[-2000:]}")
        res = [np.fromfile(os.path.join(td, f"o{r}.bin"), dtype=np.int32 if ints else np.float64)
               for r in range(n)]
    return res, outs[0]


def second_coeffs(sysd):
    """the coefficient change of the driver's solve2 op (exact multipliers): (diag, upper, lower)"""
    f = np.arange(np.asarray(sysd["upperCoef"]).size)
    c = np.arange(np.asarray(sysd["diag"]).size)
    m = 1 + 0.25 * ((7 * f) % 5)
    lower = None if sysd.get("lowerCoef") is None else sysd["lowerCoef"] * m
    return sysd["diag"] * (1.5 + 0.25 * (c % 3)), sysd["upperCoef"] * m, lower


def parse_perfs(stdout: str):
    out = []
    for line in stdout.splitlines():
        if line.startswith("PERF "):
            t = line.split()
            out.append(dict(solverName=t[1], initialResidual=float(t[2]), finalResidual=float(t[3]),
                            nIterations=int(t[4]), converged=bool(int(t[5])), singular=bool(int(t[6]))))
    return out


def parse_perf(stdout: str) -> dict:
    for line in stdout.splitlines():
        if line.startswith("PERF "):
            t = line.split()
            return dict(solverName=t[1], initialResidual=float(t[2]), finalResidual=float(t[3]),
                        nIterations=int(t[4]), converged=bool(int(t[5])), singular=bool(int(t[6])))
    raise ValueError("no PERF line in: " + stdout[-500:])


def parse_time(stdout: str) -> float:
    for line in stdout.splitlines():
        if line.startswith("TIME "):
            return float(line.split()[1])
    raise ValueError("no TIME line")


def ref_solve(sysd, controls: dict, psi=None, source=None):
    out, so = ref_run(sysd, "solve", dict_text(controls), psi=psi, source=source)
    return out, parse_perf(so)


def ref_agglom(sysd, controls: dict):
    out, _ = ref_run(sysd, "agglom", dict_text(controls), ints=True)
    n = int(out[0])
    pos = 1
    levels = []
    for _ in range(n):
        nf, nc = int(out[pos]), int(out[pos + 1])
        levels.append(dict(nFine=nf, nCoarse=nc, restrict=out[pos + 2:pos + 2 + nf].copy()))
        pos += 2 + nf
    return levels


def _parse_levels_full(out):
    n = int(out[0])
    pos = 1
    levels = []
    for _ in range(n):
        nf, nc, nff, ncf, nif = (int(x) for x in out[pos:pos + 5])
        pos += 5
        d = dict(nFine=nf, nCoarse=nc)
        d["restrict"] = out[pos:pos + nf].copy(); pos += nf
        d["faceRestrict"] = out[pos:pos + nff].copy(); pos += nff
        d["lower"] = out[pos:pos + ncf].copy(); pos += ncf
        d["upper"] = out[pos:pos + ncf].copy(); pos += ncf
        d["ifCells"], d["ifRestrict"] = [], []
        for _ in range(nif):
            a, b = int(out[pos]), int(out[pos + 1])
            pos += 2
            d["ifCells"].append(out[pos:pos + a].copy()); pos += a
            d["ifRestrict"].append(out[pos:pos + b].copy()); pos += b
        levels.append(d)
    return levels


def ref_agglom_full(sysd, controls: dict):
    """the reference's whole GAMG hierarchy (GAMGAgglomeration::New): what ldu_gamg_set_level takes"""
    out, _ = ref_run(sysd, "agglom_full", dict_text(controls), ints=True)
    return _parse_levels_full(out)


def ref_agglom_full_par(regions, controls: dict):
    outs, _ = ref_run_par(regions, "agglom_full", dict_text(controls), ints=True)
    return [_parse_levels_full(o) for o in outs]
