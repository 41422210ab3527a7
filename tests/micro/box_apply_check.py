"""Debug helper (not a test): one DIC / DILU application on an nx*ny*nz box against the oracle.
usage: box_apply_check.py nx ny nz [DIC|DILU] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import ldub200  # noqa: E402
from ldub200 import meshes  # noqa: E402
from oracle import oracle as O  # noqa: E402

nx, ny, nz = (int(x) for x in sys.argv[1:4])
pre = sys.argv[4] if len(sys.argv) > 4 else "DIC"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
s = meshes.laplacian_system(nx, ny, nz, variable=True, asym=0.3 if pre == "DILU" else 0.0)
ctx = ldub200.Context(0)
A = ldub200.lduMatrix(ctx, s["nCells"], s["lower"], s["upper"])
A.set_coeffs(s["diag"], s["upperCoef"], s["lowerCoef"])
want = O.World([s]).precondition(pre, s["source"])[0]
P = ldub200.lduMatrix.preconditioner.New(A, pre)
import os
if os.environ.get("LDU_S3_ONLY"):
    # determinism check of one sweep alone: every repetition must give the same bits
    outs = [P.precondition(s["source"]) for _ in range(reps)]
    keys = [o.tobytes() for o in outs]
    from collections import Counter
    mode = Counter(keys).most_common(1)[0][0]
    want = outs[keys.index(mode)]
    for rep, o in enumerate(outs):
        bad = np.nonzero(o != want)[0]
        if bad.size:
            bi, bj, bk = bad % nx, (bad // nx) % ny, bad // (nx * ny)
            print(f"rep {rep}: {bad.size} differ from the mode; i {bi.min()}..{bi.max()} j {bj.min()}..{bj.max()} k {bk.min()}..{bk.max()}")
    print("distinct results:", len(set(keys)))
    sys.exit(0)
for rep in range(reps):
    got = P.precondition(s["source"])
    bad = np.nonzero(got != want)[0]
    if bad.size or reps <= 8:
        print(f"rep {rep}: {bad.size} of {want.size} cells differ")
    if bad.size:
        bi, bj, bk = bad % nx, (bad // nx) % ny, bad // (nx * ny)
        k0 = bk.max(); j0 = bj[bk == k0].max(); i0 = bi[(bk == k0) & (bj == j0)].max()
        def at(i, j, k):
            c = (k * ny + j) * nx + i
            return f"({i},{j},{k}) got {got[c]!r} want {want[c]!r}"
        print("  origin:", at(i0, j0, k0), "|", at(i0 - 1, j0, k0), "|", at(i0, j0 - 1, k0), "|", at(i0, j0 + 1, k0) if j0 + 1 < ny else "")
        print("  i range", bi.min(), bi.max(), " cells at max k,j:", sorted(bi[(bk == bk.max()) & (bj == bj[bk == bk.max()].max())].tolist())[-5:])
        c = bad[:10]
        print("  first bad cells (i, j, k):", [(int(x % nx), int((x // nx) % ny), int(x // (nx * ny))) for x in c])
        k = bad // (nx * ny)
        j = (bad // nx) % ny
        print("  k range", k.min(), k.max(), " j range", j.min(), j.max(), " distinct k:", np.unique(k)[:20])
A.destroy()
ctx.close()
