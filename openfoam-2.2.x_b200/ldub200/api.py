"""ctypes binding of include/ldu_b200.h with the reference's class/method names."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE.parent / "csrc" / "libldu_b200.so"

SOLVERS = {"PCG": 0, "PBiCG": 1, "smoothSolver": 2, "GAMG": 3, "diagonal": 4}
PRECONDITIONERS = {"none": 0, "diagonal": 1, "DIC": 2, "FDIC": 3, "DILU": 4, "GAMG": 5}
SMOOTHERS = {"GaussSeidel": 0, "symGaussSeidel": 1, "DIC": 2, "DILU": 3, "DICGaussSeidel": 4,
             "DILUGaussSeidel": 5, "FDIC": 6, "nonBlockingGaussSeidel": 7, "multiColourGaussSeidel": 8}
HANDLE_BYTES = 64


class LduError(RuntimeError):
    """Raised for every failure of the CUDA library (the reference would FatalError)."""


class Controls(C.Structure):
    _fields_ = [
        ("solver", C.c_int), ("preconditioner", C.c_int), ("smoother", C.c_int),
        ("maxIter", C.c_int), ("tolerance", C.c_double), ("relTol", C.c_double),
        ("nSweeps", C.c_int), ("nCellsInCoarsestLevel", C.c_int), ("mergeLevels", C.c_int),
        ("nPreSweeps", C.c_int), ("preSweepsLevelMultiplier", C.c_int), ("maxPreSweeps", C.c_int),
        ("nPostSweeps", C.c_int), ("postSweepsLevelMultiplier", C.c_int), ("maxPostSweeps", C.c_int),
        ("nFinestSweeps", C.c_int), ("interpolateCorrection", C.c_int), ("scaleCorrection", C.c_int),
        ("nVcycles", C.c_int), ("precTolerance", C.c_double), ("precRelTol", C.c_double),
        ("useFaceWeights", C.c_int), ("cacheAgglomeration", C.c_int), ("checkInterval", C.c_int),
        ("referenceOrderSums", C.c_int),
    ]


class _Perf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("nIterations", C.c_int), ("converged", C.c_int), ("singular", C.c_int)]


_lib = None

# every symbol include/ldu_b200.h declares (tests check the .so exports them all)
ABI_SYMBOLS = [
    "ldu_version", "ldu_last_error", "ldu_launch_count", "ldu_device_count", "ldu_context_create", "ldu_context_destroy",
    "ldu_context_synchronize", "ldu_context_stream", "ldu_comm_window_create", "ldu_comm_connect",
    "ldu_device_alloc", "ldu_device_free", "ldu_copy_h2d", "ldu_copy_d2h", "ldu_device_memset", "ldu_host_alloc",
    "ldu_host_free", "ldu_matrix_create", "ldu_matrix_destroy", "ldu_matrix_set_coeffs",
    "ldu_matrix_set_coeffs_device", "ldu_matrix_set_face_weights", "ldu_colour_order", "ldu_band_compression", "ldu_amul", "ldu_tmul", "ldu_sumA", "ldu_H", "ldu_H1", "ldu_faceH", "ldu_H_device",
    "ldu_residual", "ldu_precondition", "ldu_smooth", "ldu_solve", "ldu_amul_device", "ldu_tmul_device",
    "ldu_solve_device", "ldu_residual_history", "ldu_gamg_build", "ldu_gamg_nlevels",
    "ldu_gamg_level_sizes", "ldu_gamg_level_restrict", "ldu_gamg_level_coeffs", "ldu_controls_default",
    "ldu_matrix_set_interface_coeffs", "ldu_gamg_begin_levels", "ldu_gamg_set_level", "ldu_gamg_end_levels", "ldu_gamg_internal_levels",
]


def library_path() -> Path:
    return _LIB


def library():
    """Load csrc/libldu_b200.so; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB.exists():
            raise LduError(f"{_LIB} is missing: build it with `make -C {_LIB.parent}` "
                           "(or __graft_entry__.build()); there is no fallback path")
        L = C.CDLL(str(_LIB))
        vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
        L.ldu_version.restype = C.c_char_p
        L.ldu_last_error.restype = C.c_char_p
        L.ldu_launch_count.restype = ll
        L.ldu_context_create.argtypes = [i, vp, C.POINTER(vp)]
        L.ldu_context_destroy.argtypes = [vp]
        L.ldu_context_synchronize.argtypes = [vp]
        L.ldu_context_stream.argtypes = [vp]
        L.ldu_context_stream.restype = vp
        L.ldu_comm_window_create.argtypes = [vp, i, i, i, ll, vp]
        L.ldu_comm_connect.argtypes = [vp, vp]
        L.ldu_device_alloc.argtypes = [vp, ll, C.POINTER(vp)]
        L.ldu_device_free.argtypes = [vp, vp]
        L.ldu_copy_h2d.argtypes = [vp, vp, vp, ll]
        L.ldu_copy_d2h.argtypes = [vp, vp, vp, ll]
        L.ldu_device_memset.argtypes = [vp, vp, i, ll]
        L.ldu_host_alloc.argtypes = [ll, C.POINTER(vp)]
        L.ldu_host_free.argtypes = [vp]
        L.ldu_matrix_create.argtypes = [vp, i, i, vp, vp, i, vp, vp, vp, vp, C.POINTER(vp)]
        L.ldu_matrix_destroy.argtypes = [vp]
        L.ldu_matrix_set_coeffs.argtypes = [vp, vp, vp, vp, vp, vp]
        L.ldu_matrix_set_coeffs_device.argtypes = [vp, vp, vp, vp]
        L.ldu_matrix_set_interface_coeffs.argtypes = [vp, vp, vp]
        L.ldu_matrix_set_face_weights.argtypes = [vp, vp]
        L.ldu_amul.argtypes = [vp, vp, vp]
        L.ldu_tmul.argtypes = [vp, vp, vp]
        L.ldu_sumA.argtypes = [vp, vp]
        L.ldu_colour_order.argtypes = [i, i, vp, vp, vp, vp]
        L.ldu_band_compression.argtypes = [i, i, vp, vp, vp]
        L.ldu_H.argtypes = [vp, vp, vp]
        L.ldu_H1.argtypes = [vp, vp]
        L.ldu_faceH.argtypes = [vp, vp, vp]
        L.ldu_H_device.argtypes = [vp, vp, vp]
        L.ldu_residual.argtypes = [vp, vp, vp, vp]
        L.ldu_precondition.argtypes = [vp, i, vp, vp, i]
        L.ldu_smooth.argtypes = [vp, i, vp, vp, i]
        L.ldu_solve.argtypes = [vp, C.POINTER(Controls), vp, vp, C.POINTER(_Perf)]
        L.ldu_amul_device.argtypes = [vp, vp, vp]
        L.ldu_tmul_device.argtypes = [vp, vp, vp]
        L.ldu_solve_device.argtypes = [vp, C.POINTER(Controls), vp, vp, C.POINTER(_Perf)]
        L.ldu_residual_history.argtypes = [vp, vp, i]
        L.ldu_gamg_build.argtypes = [vp, C.POINTER(Controls)]
        L.ldu_gamg_nlevels.argtypes = [vp]
        L.ldu_gamg_begin_levels.argtypes = [vp]
        L.ldu_gamg_set_level.argtypes = [vp, i, i, vp, i, vp, i, i, vp, vp, vp, vp, vp]
        L.ldu_gamg_end_levels.argtypes = [vp]
        L.ldu_gamg_internal_levels.argtypes = [vp]
        L.ldu_gamg_level_sizes.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
        L.ldu_gamg_level_restrict.argtypes = [vp, i, vp]
        L.ldu_gamg_level_coeffs.argtypes = [vp, i, vp, vp, vp]
        L.ldu_controls_default.argtypes = [C.POINTER(Controls)]
        _lib = L
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = library().ldu_last_error().decode(errors="replace")
        raise LduError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(library().ldu_launch_count())


def device_count() -> int:
    """CUDA devices visible to this process (rank r of a parallel run binds r % device_count())."""
    return int(library().ldu_device_count())


def make_controls(d: dict) -> Controls:
    """fvSolution-style solver dictionary -> ldu_controls (reference defaults)."""
    L = library()
    c = Controls()
    L.ldu_controls_default(C.byref(c))
    name = d.get("solver", "PCG")
    if name in ("ICCG", "BICCG"):
        # the table constructors of ICCG / BICCG hand the dictionary to PCG / PBiCG unchanged
        # (ICCG.C:67-86, BICCG.C:67-86): same solver, the dictionary's own preconditioner
        d = dict(d, solver="PCG" if name == "ICCG" else "PBiCG")
        name = d["solver"]
    if "solver" in d and name in ("PCG", "PBiCG") and "preconditioner" not in d:
        # lduMatrixPreconditioner.C:39-58 looks the entry up without a default
        raise LduError("keyword preconditioner is undefined in the solver dictionary")
    if name not in SOLVERS:
        # lduMatrixSolver.C:96-110: unknown solver is a FatalIOError listing the table
        raise LduError(f"Unknown solver {name}; valid solvers are {sorted(SOLVERS)}")
    if d.get("directSolveCoarsest") or (isinstance(d.get("preconditioner"), dict)
                                        and d["preconditioner"].get("directSolveCoarsest")):
        # the dense LU solve of the coarsest level (GAMGSolver.C:74,91-106) is outside this library
        raise LduError("directSolveCoarsest is not supported")
    c.solver = SOLVERS[name]
    pre = d.get("preconditioner", "none")
    sub = d
    if isinstance(pre, dict):   # lduMatrixPreconditioner.C:39-58: word or sub-dictionary
        sub = pre
        pre = pre["preconditioner"]
    if pre not in PRECONDITIONERS:
        raise LduError(f"Unknown preconditioner {pre}; valid are {sorted(PRECONDITIONERS)}")
    c.preconditioner = PRECONDITIONERS[pre]
    if c.preconditioner == PRECONDITIONERS["GAMG"]:
        c.precTolerance = float(sub.get("tolerance", 1e-6))
        c.precRelTol = float(sub.get("relTol", 0.0))
    sm = sub.get("smoother", d.get("smoother", "GaussSeidel"))
    if sm not in SMOOTHERS:
        raise LduError(f"Unknown smoother {sm}; valid are {sorted(SMOOTHERS)}")
    c.smoother = SMOOTHERS[sm]
    for key in ("maxIter", "nSweeps"):
        if key in d:
            setattr(c, key, int(d[key]))
    for key in ("tolerance", "relTol"):
        if key in d:
            setattr(c, key, float(d[key]))
    for key in ("nCellsInCoarsestLevel", "mergeLevels", "nPreSweeps", "preSweepsLevelMultiplier",
                "maxPreSweeps", "nPostSweeps", "postSweepsLevelMultiplier", "maxPostSweeps",
                "nFinestSweeps", "nVcycles"):
        if key in sub:
            setattr(c, key, int(sub[key]))
    if "interpolateCorrection" in sub:
        c.interpolateCorrection = int(bool(sub["interpolateCorrection"]))
    if "scaleCorrection" in sub:
        c.scaleCorrection = int(bool(sub["scaleCorrection"]))
    if "cacheAgglomeration" in sub:
        c.cacheAgglomeration = int(bool(sub["cacheAgglomeration"]))
    c.useFaceWeights = int(sub.get("agglomerator", "faceAreaPair") == "faceAreaPair")
    if "checkInterval" in d:
        c.checkInterval = int(d["checkInterval"])
    if "referenceOrderSums" in d:
        c.referenceOrderSums = int(bool(d["referenceOrderSums"]))
    return c


def gather_handles(mine: bytes) -> bytes:
    """All-gather the opaque exchange-window handles through torch.distributed
    (any backend): rank-major concatenation, HANDLE_BYTES each — the layout
    ldu_comm_connect expects."""
    import torch.distributed as dist
    assert len(mine) == HANDLE_BYTES
    allh = [None] * dist.get_world_size()
    dist.all_gather_object(allh, mine)
    return b"".join(allh)


class SolverPerformance:
    """SolverPerformance<scalar> (matrices/LduMatrix/LduMatrix/SolverPerformance.H:78-137)."""

    def __init__(self, solverName, fieldName, initialResidual=0.0, finalResidual=0.0,
                 nIterations=0, converged=False, singular=False):
        self.solverName = solverName
        self.fieldName = fieldName
        self.initialResidual = initialResidual
        self.finalResidual = finalResidual
        self.nIterations = nIterations
        self.converged = converged
        self.singular = singular

    def __str__(self):  # SolverPerformance.C:95-125
        head = f"{self.solverName}:  Solving for {self.fieldName}"
        if self.singular:
            return head + ":  solution singularity"
        return (f"{head}, Initial residual = {self.initialResidual:.6g}, "
                f"Final residual = {self.finalResidual:.6g}, No Iterations {self.nIterations}")

    __repr__ = __str__


class Context:
    """One GPU + stream (+ peers).  stream: a cudaStream_t handle (int) or None."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.L = library()
        h = C.c_void_p()
        _check(self.L.ldu_context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)),
               "ldu_context_create")
        self.h = h
        self.device = device
        self.rank, self.nRanks = 0, 1

    def synchronize(self):
        _check(self.L.ldu_context_synchronize(self.h), "ldu_context_synchronize")

    @property
    def stream(self) -> int:
        return int(self.L.ldu_context_stream(self.h) or 0)

    # -- multi-GPU: the caller moves the opaque handles between ranks -------------
    def comm_window_create(self, rank, nRanks, maxInterfaces, maxInterfaceFaces) -> bytes:
        buf = (C.c_ubyte * HANDLE_BYTES)()
        _check(self.L.ldu_comm_window_create(self.h, rank, nRanks, maxInterfaces, maxInterfaceFaces, buf),
               "ldu_comm_window_create")
        self.rank, self.nRanks = rank, nRanks
        return bytes(buf)

    def comm_connect(self, all_handles: bytes):
        assert len(all_handles) == HANDLE_BYTES * self.nRanks
        buf = (C.c_ubyte * len(all_handles)).from_buffer_copy(all_handles)
        _check(self.L.ldu_comm_connect(self.h, buf), "ldu_comm_connect")

    def connect_torch_distributed(self, maxInterfaces, maxInterfaceFaces):
        """Exchange the window handles through torch.distributed (plumbing only)."""
        import torch.distributed as dist
        rank, n = dist.get_rank(), dist.get_world_size()
        mine = self.comm_window_create(rank, n, maxInterfaces, maxInterfaceFaces)
        self.comm_connect(gather_handles(mine))
        dist.barrier()

    def close(self):
        if self.h:
            self.L.ldu_context_destroy(self.h)
            self.h = None


class DeviceField:
    """A scalarField resident in HBM."""

    def __init__(self, ctx: Context, n: int, host: np.ndarray | None = None):
        self.ctx, self.n = ctx, int(n)
        p = C.c_void_p()
        _check(ctx.L.ldu_device_alloc(ctx.h, self.n * 8, C.byref(p)), "ldu_device_alloc")
        self.ptr = p
        if host is not None:
            self.upload(host)

    def upload(self, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.n
        _check(self.ctx.L.ldu_copy_h2d(self.ctx.h, self.ptr, host.ctypes.data, self.n * 8), "ldu_copy_h2d")
        self.ctx.synchronize()

    def upload_async(self, host_ptr: int):
        _check(self.ctx.L.ldu_copy_h2d(self.ctx.h, self.ptr, C.c_void_p(host_ptr), self.n * 8), "ldu_copy_h2d")

    def zero(self):
        _check(self.ctx.L.ldu_device_memset(self.ctx.h, self.ptr, 0, self.n * 8), "ldu_device_memset")

    def download(self) -> np.ndarray:
        out = np.empty(self.n)
        _check(self.ctx.L.ldu_copy_d2h(self.ctx.h, out.ctypes.data, self.ptr, self.n * 8), "ldu_copy_d2h")
        return out

    def free(self):
        if self.ptr:
            self.ctx.L.ldu_device_free(self.ctx.h, self.ptr)
            self.ptr = None


def pinned_array(n: int, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (ldu_host_alloc); never freed
    explicitly (process lifetime), meant for long-lived staging buffers."""
    L = library()
    p = C.c_void_p()
    nbytes = int(n) * np.dtype(dtype).itemsize
    _check(L.ldu_host_alloc(nbytes, C.byref(p)), "ldu_host_alloc")
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


class lduInterface:
    """A coupled patch: faceCells + where the other side lives
    (matrices/lduMatrix/lduAddressing/lduInterface/lduInterface.H:54-113)."""

    def __init__(self, faceCells, nbrRank: int, nbrInterface: int):
        self.faceCells = np.ascontiguousarray(faceCells, dtype=np.int32)
        self.nbrRank = int(nbrRank)
        self.nbrInterface = int(nbrInterface)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class lduMatrix:
    """lduMatrix + its lduAddressing on one GPU (lduMatrix.H:77-86, lduAddressing.H:111-199)."""

    def __init__(self, ctx: Context, nCells: int, lowerAddr, upperAddr, interfaces=()):
        self.ctx, self.L = ctx, ctx.L
        self.nCells = int(nCells)
        lo = np.ascontiguousarray(lowerAddr, dtype=np.int32)
        up = np.ascontiguousarray(upperAddr, dtype=np.int32)
        self.nFaces = lo.size
        self.interfaces = list(interfaces)
        n_if = len(self.interfaces)
        sizes = (C.c_int * max(n_if, 1))(*[it.faceCells.size for it in self.interfaces])
        cells = (C.c_void_p * max(n_if, 1))(*[it.faceCells.ctypes.data for it in self.interfaces])
        nbr_rank = (C.c_int * max(n_if, 1))(*[it.nbrRank for it in self.interfaces])
        nbr_if = (C.c_int * max(n_if, 1))(*[it.nbrInterface for it in self.interfaces])
        h = C.c_void_p()
        _check(self.L.ldu_matrix_create(ctx.h, self.nCells, self.nFaces, lo.ctypes.data, up.ctypes.data,
                                        n_if, sizes, cells, nbr_rank, nbr_if, C.byref(h)),
               "ldu_matrix_create")
        self.h = h
        self._symmetric = True

    # -- coefficients --------------------------------------------------------------
    def set_coeffs(self, diag, upper, lower=None, bouCoeffs=(), intCoeffs=()):
        """upper=None (only without faces): the matrix is lduMatrix::diagonal() and every solve goes to
        diagonalSolver; an EMPTY upper array is a faceless matrix whose upper() exists."""
        self._diagonal = upper is None and lower is None
        diag = _f64(diag)
        # a non-NULL pointer also for an empty array: NULL means "never set"
        upper = None if upper is None else _f64(upper if np.size(upper) else np.zeros(1))
        lower = None if lower is None else _f64(lower if np.size(lower) else np.zeros(1))
        assert diag.size == self.nCells and (self.nFaces == 0 or upper.size == self.nFaces)
        n_if = len(self.interfaces)
        bou = [_f64(b) for b in bouCoeffs]
        inc = [_f64(b) for b in intCoeffs]
        assert len(bou) == n_if and len(inc) == n_if
        bp = (C.c_void_p * max(n_if, 1))(*[b.ctypes.data for b in bou])
        ip = (C.c_void_p * max(n_if, 1))(*[b.ctypes.data for b in inc])
        _check(self.L.ldu_matrix_set_coeffs(self.h, diag.ctypes.data,
                                            None if upper is None else upper.ctypes.data,
                                            None if lower is None else lower.ctypes.data, bp, ip),
               "ldu_matrix_set_coeffs")
        self._symmetric = lower is None

    def set_interface_coeffs(self, bouCoeffs, intCoeffs):
        """boundary coefficients of the coupled patches alone (host arrays), for matrices whose diag/upper/lower
        are handed over on the device"""
        n_if = len(self.interfaces)
        bou = [_f64(b) for b in bouCoeffs]
        inc = [_f64(b) for b in intCoeffs]
        assert len(bou) == n_if and len(inc) == n_if
        bp = (C.c_void_p * max(n_if, 1))(*[b.ctypes.data for b in bou])
        ip = (C.c_void_p * max(n_if, 1))(*[b.ctypes.data for b in inc])
        _check(self.L.ldu_matrix_set_interface_coeffs(self.h, bp, ip), "ldu_matrix_set_interface_coeffs")

    def set_coeffs_device(self, d_diag: DeviceField, d_upper: DeviceField, d_lower: DeviceField | None = None):
        _check(self.L.ldu_matrix_set_coeffs_device(self.h, d_diag.ptr, d_upper.ptr,
                                                   None if d_lower is None else d_lower.ptr),
               "ldu_matrix_set_coeffs_device")
        self._symmetric = d_lower is None

    def set_face_weights(self, w):
        w = _f64(w)
        _check(self.L.ldu_matrix_set_face_weights(self.h, w.ctypes.data), "ldu_matrix_set_face_weights")

    def symmetric(self) -> bool:
        return self._symmetric

    # -- operators, host fields ----------------------------------------------------
    def Amul(self, psi) -> np.ndarray:
        psi = _f64(psi)
        out = np.empty(self.nCells)
        _check(self.L.ldu_amul(self.h, out.ctypes.data, psi.ctypes.data), "ldu_amul")
        return out

    def Tmul(self, psi) -> np.ndarray:
        psi = _f64(psi)
        out = np.empty(self.nCells)
        _check(self.L.ldu_tmul(self.h, out.ctypes.data, psi.ctypes.data), "ldu_tmul")
        return out

    def sumA(self) -> np.ndarray:
        out = np.empty(self.nCells)
        _check(self.L.ldu_sumA(self.h, out.ctypes.data), "ldu_sumA")
        return out

    def H(self, psi) -> np.ndarray:
        """lduMatrix::H (lduMatrixTemplates.C:33-65)"""
        psi = _f64(psi)
        out = np.empty(self.nCells)
        _check(self.L.ldu_H(self.h, out.ctypes.data, psi.ctypes.data), "ldu_H")
        return out

    def H1(self) -> np.ndarray:
        """lduMatrix::H1 (lduMatrixATmul.C:298-327)"""
        out = np.empty(self.nCells)
        _check(self.L.ldu_H1(self.h, out.ctypes.data), "ldu_H1")
        return out

    def faceH(self, psi) -> np.ndarray:
        """lduMatrix::faceH (lduMatrixTemplates.C:79-113): one value per face"""
        psi = _f64(psi)
        out = np.empty(self.nFaces)
        _check(self.L.ldu_faceH(self.h, out.ctypes.data, psi.ctypes.data), "ldu_faceH")
        return out

    def residual(self, psi, source) -> np.ndarray:
        psi, source = _f64(psi), _f64(source)
        out = np.empty(self.nCells)
        _check(self.L.ldu_residual(self.h, out.ctypes.data, psi.ctypes.data, source.ctypes.data), "ldu_residual")
        return out

    # -- operators, device fields --------------------------------------------------
    def Amul_device(self, d_Apsi: DeviceField, d_psi: DeviceField):
        _check(self.L.ldu_amul_device(self.h, d_Apsi.ptr, d_psi.ptr), "ldu_amul_device")

    def Tmul_device(self, d_Tpsi: DeviceField, d_psi: DeviceField):
        _check(self.L.ldu_tmul_device(self.h, d_Tpsi.ptr, d_psi.ptr), "ldu_tmul_device")

    def H_device(self, d_Hpsi: DeviceField, d_psi: DeviceField):
        _check(self.L.ldu_H_device(self.h, d_Hpsi.ptr, d_psi.ptr), "ldu_H_device")

    def residual_history(self, cap=4096) -> np.ndarray:
        buf = np.zeros(cap)
        n = self.L.ldu_residual_history(self.h, buf.ctypes.data, cap)
        return buf[:n].copy()

    # -- GAMG hierarchy introspection ------------------------------------------------
    def gamg_levels(self, controls: dict):
        c = make_controls(controls)
        _check(self.L.ldu_gamg_build(self.h, C.byref(c)), "ldu_gamg_build")
        out = []
        for lev in range(self.L.ldu_gamg_nlevels(self.h)):
            nf, nc, ncf = C.c_int(), C.c_int(), C.c_int()
            _check(self.L.ldu_gamg_level_sizes(self.h, lev, C.byref(nf), C.byref(nc), C.byref(ncf)), "sizes")
            r = np.empty(nf.value, dtype=np.int32)
            _check(self.L.ldu_gamg_level_restrict(self.h, lev, r.ctypes.data), "restrict")
            diag = np.empty(nc.value)
            upper = np.empty(ncf.value)
            _check(self.L.ldu_gamg_level_coeffs(self.h, lev, diag.ctypes.data,
                                                upper.ctypes.data if ncf.value else None, None), "coeffs")
            out.append(dict(nFine=nf.value, nCoarse=nc.value, nFaces=ncf.value, restrict=r,
                            diag=diag, upperCoef=upper))
        return out

    def set_gamg_levels(self, levels):
        """Hand over a GAMG hierarchy built by the host (the reference's GAMGAgglomeration): levels = list of
        dict(restrict, faceRestrict, nCoarse, lower, upper[, ifCells, ifRestrict]) from the finest level down
        (ldu_gamg_begin_levels / ldu_gamg_set_level / ldu_gamg_end_levels).  None gives the agglomeration
        back to the library."""
        if levels is None:
            _check(self.L.ldu_gamg_internal_levels(self.h), "ldu_gamg_internal_levels")
            return
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        _check(self.L.ldu_gamg_begin_levels(self.h), "ldu_gamg_begin_levels")
        n_if = len(self.interfaces)
        for lev, d in enumerate(levels):
            r, fr, lo, up = i32(d["restrict"]), i32(d["faceRestrict"]), i32(d["lower"]), i32(d["upper"])
            cells = [i32(a) for a in d.get("ifCells", [])]
            ifr = [i32(a) for a in d.get("ifRestrict", [])]
            assert len(cells) == n_if and len(ifr) == n_if
            sizes = (C.c_int * max(n_if, 1))(*[a.size for a in cells])
            cp = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in cells])
            rp = (C.c_void_p * max(n_if, 1))(*[a.ctypes.data for a in ifr])
            _check(self.L.ldu_gamg_set_level(self.h, lev, r.size, r.ctypes.data, fr.size, fr.ctypes.data,
                                             int(d["nCoarse"]), lo.size, lo.ctypes.data, up.ctypes.data,
                                             sizes if n_if else None, cp if n_if else None, rp if n_if else None),
                   "ldu_gamg_set_level")
        _check(self.L.ldu_gamg_end_levels(self.h), "ldu_gamg_end_levels")

    def destroy(self):
        if self.h:
            self.L.ldu_matrix_destroy(self.h)
            self.h = None

    # -- run-time selection mirrors ------------------------------------------------
    class solver:
        """lduMatrix::solver (lduMatrix.H:91-258); New() = lduMatrixSolver.C:40-136."""

        def __init__(self, fieldName, matrix: "lduMatrix", solverControls: dict):
            self.fieldName = fieldName
            self.matrix = matrix
            self.read(solverControls)

        @classmethod
        def New(cls, fieldName, matrix, solverControls):
            return cls(fieldName, matrix, solverControls)

        def read(self, solverControls: dict):
            self.controlDict = dict(solverControls)
            self.controls = make_controls(self.controlDict)

        def _name(self):
            d = self.controlDict
            s = d.get("solver", "PCG")
            if getattr(self.matrix, "_diagonal", False) and self.matrix.ctx.nRanks == 1:
                return "diagonal"
            s = {"ICCG": "PCG", "BICCG": "PBiCG"}.get(s, s)
            if s in ("PCG", "PBiCG"):   # preconditioner name + typeName (PCG.C:72-77)
                pre = d.get("preconditioner", "none")
                if isinstance(pre, dict):
                    pre = pre["preconditioner"]
                return pre + s
            return s

        def solve(self, psi: np.ndarray, source) -> SolverPerformance:
            """psi is updated in place (initial guess in, solution out)."""
            m = self.matrix
            source = _f64(source)
            if not (isinstance(psi, np.ndarray) and psi.dtype == np.float64 and psi.flags.c_contiguous):
                raise LduError("psi must be a contiguous float64 numpy array (it is updated in place)")
            perf = _Perf()
            _check(m.L.ldu_solve(m.h, C.byref(self.controls), psi.ctypes.data, source.ctypes.data,
                                 C.byref(perf)), "ldu_solve")
            return self._perf(perf)

        def solve_device(self, d_psi: DeviceField, d_source: DeviceField) -> SolverPerformance:
            m = self.matrix
            perf = _Perf()
            _check(m.L.ldu_solve_device(m.h, C.byref(self.controls), d_psi.ptr, d_source.ptr, C.byref(perf)),
                   "ldu_solve_device")
            return self._perf(perf)

        def _perf(self, p):
            return SolverPerformance(self._name(), self.fieldName, p.initialResidual, p.finalResidual,
                                     p.nIterations, bool(p.converged), bool(p.singular))

    class smoother:
        """lduMatrix::smoother (lduMatrix.H:264-400)."""

        def __init__(self, fieldName, matrix, solverControls):
            self.matrix = matrix
            name = solverControls["smoother"] if isinstance(solverControls, dict) else solverControls
            if name not in SMOOTHERS:
                raise LduError(f"Unknown smoother {name}")
            self.kind = SMOOTHERS[name]

        @classmethod
        def New(cls, fieldName, matrix, solverControls):
            return cls(fieldName, matrix, solverControls)

        def smooth(self, psi: np.ndarray, source, nSweeps: int):
            m = self.matrix
            source = _f64(source)
            _check(m.L.ldu_smooth(m.h, self.kind, psi.ctypes.data, source.ctypes.data, int(nSweeps)), "ldu_smooth")

    class preconditioner:
        """lduMatrix::preconditioner (lduMatrix.H:406-506)."""

        def __init__(self, matrix, solverControls):
            self.matrix = matrix
            name = solverControls["preconditioner"] if isinstance(solverControls, dict) else solverControls
            if name not in PRECONDITIONERS or name == "GAMG":
                raise LduError(f"Unknown preconditioner {name}")
            self.kind = PRECONDITIONERS[name]

        @classmethod
        def New(cls, matrix, solverControls):
            return cls(matrix, solverControls)

        def precondition(self, rA) -> np.ndarray:
            return self._apply(rA, 0)

        def preconditionT(self, rT) -> np.ndarray:
            return self._apply(rT, 1)

        def _apply(self, rA, transpose):
            m = self.matrix
            rA = _f64(rA)
            out = np.empty(m.nCells)
            _check(m.L.ldu_precondition(m.h, self.kind, out.ctypes.data, rA.ctypes.data, transpose),
                   "ldu_precondition")
            return out
