"""Micro-benchmark (not a test): time per PCG iteration with DIC vs diagonal on an
nx*ny*nz box; the difference is two triangular sweeps.  usage: perf_sweeps.py nx ny nz [iters]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402


def run(shape, iters=50):
    nx, ny, nz = shape
    s = meshes.laplacian_system(nx, ny, nz)
    stream = torch.cuda.Stream()
    ctx = ldub200.Context(0, stream.cuda_stream)
    A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
    A.set_coeffs(s["diag"], s["upperCoef"])
    d_psi = ldub200.DeviceField(ctx, s["nCells"])
    d_src = ldub200.DeviceField(ctx, s["nCells"], s["source"])
    out = {}
    for pre in ("diagonal", "DIC"):
        solver = ldub200.lduMatrix.solver.New("p", A, dict(solver="PCG", preconditioner=pre, tolerance=0,
                                                           relTol=0, maxIter=iters - 1))
        for rep in range(3):
            d_psi.zero()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            perf = solver.solve_device(d_psi, d_src)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[pre] = dt / iters
    sweep = (out["DIC"] - out["diagonal"]) / 2
    print(f"{nx}x{ny}x{nz}: cells {s['nCells']:9d}  diag {out['diagonal']*1e6:8.1f} us/it  DIC {out['DIC']*1e6:8.1f} us/it"
          f"  sweep {sweep*1e6:8.1f} us  levels {nx+ny+nz-2}  us/level {sweep*1e6/(nx+ny+nz-2):6.3f}")
    A.destroy()
    ctx.close()


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    run(a[:3], a[3] if len(a) > 3 else 50)
