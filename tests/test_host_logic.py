"""Host-side logic: synthetic meshes, domain decomposition, and the N>1 path on
CPU with two gloo ranks (halo exchange conventions + handle plumbing)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

import cases
from ldub200 import decompose, meshes
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent


def test_box_addressing_is_upper_triangular():
    lo, up, d = meshes.box_addressing(5, 4, 3)
    assert lo.size == 4 * 4 * 3 + 5 * 3 * 3 + 5 * 4 * 2
    assert np.all(lo < up)
    assert np.all(np.diff(lo) >= 0)
    same = np.diff(lo) == 0
    assert np.all(np.diff(up)[same] > 0)       # neighbours ascending within an owner
    s = meshes.laplacian_system(20, 20, 1)
    assert (s["nCells"], s["nFaces"]) == (400, 760)     # the icoFoam cavity tutorial sizes


def test_negsumdiag_and_reference_cell():
    s = meshes.laplacian_system(6, 5, 4, variable=True)
    rows = O.World([s]).sumA()[0]
    # row sums vanish except at the pinned reference cell (fvMatrix::setReference doubles its diag)
    assert np.abs(np.delete(rows, 0)).max() < 1e-12 and rows[0] < 0


@pytest.mark.parametrize("R", [2, 4, 8])
def test_decomposition_matches_global(R):
    n = 8
    g = meshes.laplacian_system(n, n, n, variable=True)
    px, py, pz = decompose.split_for(R)
    regs = decompose.decompose(g, decompose.block_partition(n, n, n, px, py, pz), R)
    assert sum(r["nCells"] for r in regs) == g["nCells"]
    cut = sum(it["faceCells"].size for r in regs for it in r["interfaces"])
    assert sum(r["nFaces"] for r in regs) + cut // 2 == g["nFaces"]
    for r, reg in enumerate(regs):
        assert np.all(reg["lower"] < reg["upper"]) and np.all(np.diff(reg["lower"]) >= 0)
        for it in reg["interfaces"]:
            back = regs[it["nbrRegion"]]["interfaces"][it["nbrInterface"]]
            assert back["nbrRegion"] == r and back["faceCells"].size == it["faceCells"].size
            # symmetric matrix: both sides carry the same coefficient (bouCoeffs = -upper_cut)
            assert np.array_equal(back["bouCoeffs"], it["bouCoeffs"])
    w, wg = O.World(regs), O.World([g])
    x = np.random.default_rng(1).standard_normal(g["nCells"])
    y = decompose.gather_field(regs, w.amul([x[r["cells"]] for r in regs]), g["nCells"])
    assert np.abs(y - wg.amul(x)[0]).max() < 1e-13
    rs = decompose.gather_field(regs, w.residual([x[r["cells"]] for r in regs], [r["source"] for r in regs]),
                                g["nCells"])
    assert np.abs(rs - wg.residual(x, g["source"])[0]).max() < 1e-13


@pytest.mark.parametrize("R", [2, 4, 8, 16, 32])
def test_local_box_region_equals_decompose(R):
    n = 8       # 16 and 32 regions (2x2x4, 2x4x4): the decompositions of the many-core reference arm
    g = meshes.laplacian_system(n, n, n)
    px, py, pz = decompose.split_for(R)
    regs = decompose.decompose(g, decompose.block_partition(n, n, n, px, py, pz), R)
    for r in range(R):
        a, b = regs[r], decompose.local_box_region(n, r, R)
        for k in ("lower", "upper", "diag", "upperCoef", "source"):
            assert np.array_equal(a[k], b[k])
        assert len(a["interfaces"]) == len(b["interfaces"])
        for ia, ib in zip(a["interfaces"], b["interfaces"]):
            assert (ia["nbrRegion"], ia["nbrInterface"]) == (ib["nbrRegion"], ib["nbrInterface"])
            assert np.array_equal(ia["faceCells"], ib["faceCells"])
            assert np.array_equal(ia["bouCoeffs"], ib["bouCoeffs"])


def test_asymmetric_decomposition_transpose_coefficients():
    g = cases.system("asym10")
    regs = decompose.decompose(g, decompose.block_partition(10, 10, 10, 2, 1, 1), 2)
    w, wg = O.World(regs), O.World([g])
    x = np.random.default_rng(3).standard_normal(g["nCells"])
    t = decompose.gather_field(regs, w.tmul([x[r["cells"]] for r in regs]), g["nCells"])
    assert np.abs(t - wg.tmul(x)[0]).max() < 1e-13


def test_multi_region_solvers_converge_to_the_global_solution():
    n = 8
    g = meshes.laplacian_system(n, n, n, variable=True)
    regs = decompose.decompose(g, decompose.block_partition(n, n, n, 2, 2, 1), 4)
    w, wg = O.World(regs), O.World([g])
    for ctl in (dict(solver="PCG", preconditioner="DIC", tolerance=1e-10, relTol=0),
                dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair",
                     nCellsInCoarsestLevel=4, mergeLevels=1, tolerance=1e-10, relTol=0)):
        psi, perf = w.solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
        pg, _ = wg.solve(ctl, g["psi0"], g["source"])
        assert perf["converged"]
        assert np.abs(decompose.gather_field(regs, psi, g["nCells"]) - pg[0]).max() < 1e-6


_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["LDU_ROOT"]); sys.path.insert(0, os.path.join(os.environ["LDU_ROOT"], "openfoam-2.2.x_b200"))
from ldub200 import decompose, meshes
from ldub200.api import gather_handles, HANDLE_BYTES
from oracle import oracle as O

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["LDU_PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
n = 8
reg = decompose.local_box_region(n, rank, 2)
x = np.sin(0.05 * reg["cells"])                      # a global field sampled on my cells
# halo exchange exactly as the device path orders it: psi[faceCells] per interface
y = reg["diag"] * x
np.add.at(y, reg["upper"], reg["upperCoef"] * x[reg["lower"]])
np.add.at(y, reg["lower"], reg["upperCoef"] * x[reg["upper"]])
for it in reg["interfaces"]:
    send = torch.from_numpy(x[it["faceCells"]].copy())
    recv = torch.empty_like(send)
    ops = [dist.P2POp(dist.isend, send, it["nbrRegion"]), dist.P2POp(dist.irecv, recv, it["nbrRegion"])]
    for q in dist.batch_isend_irecv(ops):
        q.wait()
    np.subtract.at(y, it["faceCells"], it["bouCoeffs"] * recv.numpy())     # Apsi[fc] -= bou*psiNbr
# check against the in-process oracle world
g = meshes.laplacian_system(n, n, n)
regs = decompose.decompose(g, decompose.block_partition(n, n, n, 2, 1, 1), 2)
want = O.World(regs).amul([np.sin(0.05 * r["cells"]) for r in regs])[rank]
assert np.abs(y - want).max() < 1e-13, np.abs(y - want).max()
# global dot product = sum of rank partials (2 ranks: order cannot matter)
t = torch.tensor([float(x @ y)], dtype=torch.float64)
dist.all_reduce(t)
full = sum(float(np.sin(0.05 * r["cells"]) @ O.World(regs).amul([np.sin(0.05 * q["cells"]) for q in regs])[i])
           for i, r in enumerate(regs))
assert abs(t.item() - full) < 1e-10
# window-handle plumbing: rank-major concatenation of 64-byte handles
mine = bytes([rank + 1]) * HANDLE_BYTES
allh = gather_handles(mine)
assert len(allh) == 2 * HANDLE_BYTES and allh[0] == 1 and allh[HANDLE_BYTES] == 2
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_halo_exchange_gloo(tmp_path):
    import socket
    import subprocess
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LDU_PORT=str(port), LDU_ROOT=str(ROOT))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"rank {r} ok" in o
