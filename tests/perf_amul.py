"""Micro-benchmark (not a test): Amul alone on an n^3 box, CUDA-event timed, algorithmic GB/s.
usage: perf_amul.py n [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 216
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
s = meshes.laplacian_system(n, n, n)
stream = torch.cuda.Stream()
ctx = ldub200.Context(0, stream.cuda_stream)
A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
A.set_coeffs(s["diag"], s["upperCoef"])
x = ldub200.DeviceField(ctx, s["nCells"], np.sin(0.11 * np.arange(s["nCells"])))
y = ldub200.DeviceField(ctx, s["nCells"])
for _ in range(5):
    A.Amul_device(y, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    A.Amul_device(y, x)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = 24 * s["nCells"] + 16 * s["nFaces"]
print(f"Amul {n}^3: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s algorithmic")
