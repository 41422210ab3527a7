"""One rank of a multi-region run through the CUDA path (launched by
tests/test_gpu_multi.py, one process per region; with fewer GPUs than ranks the
ranks share devices, the peer exchange then goes through CUDA IPC on one GPU)."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    import ldub200
    from ldub200 import decompose, meshes
    from oracle import oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["LDU_PORT"],
                            rank=rank, world_size=world)
    dev = rank % torch.cuda.device_count()
    n = int(os.environ.get("LDU_N", "12"))
    g = meshes.laplacian_system(n, n, n, variable=True)
    px, py, pz = decompose.split_for(world)
    regs = decompose.decompose(g, decompose.block_partition(n, n, n, px, py, pz), world)
    reg = regs[rank]
    ctx = ldub200.Context(dev)
    max_if = max(it["faceCells"].size for r in regs for it in r["interfaces"])
    ctx.connect_torch_distributed(8, int(max_if))
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in reg["interfaces"]]
    A = ldub200.lduMatrix(ctx, reg["nCells"], reg["lower"], reg["upper"], ifs)
    A.set_coeffs(reg["diag"], reg["upperCoef"], None, [it["bouCoeffs"] for it in reg["interfaces"]],
                 [it["intCoeffs"] for it in reg["interfaces"]])
    A.set_face_weights(reg["faceWeights"])
    w = O.World(regs)
    xs = [np.sin(0.05 * r["cells"]) for r in regs]
    results = {}
    # Amul / residual with halos: bit-exact against the oracle world
    results["amul"] = bool(np.array_equal(A.Amul(xs[rank]), w.amul(xs)[rank]))
    results["residual"] = bool(np.array_equal(A.residual(xs[rank], reg["source"]),
                                              w.residual(xs, [r["source"] for r in regs])[rank]))
    results["sumA"] = bool(np.array_equal(A.sumA(), w.sumA()[rank]))
    sm = w.smooth("GaussSeidel", xs, [r["source"] for r in regs], 2)[rank]
    psi = xs[rank].copy()
    ldub200.lduMatrix.smoother.New("p", A, "GaussSeidel").smooth(psi, reg["source"], 2)
    results["gs"] = bool(np.array_equal(psi, sm))
    # every smoother with interface updates (nonBlockingGaussSeidel consumes the halo after the
    # cells below the first coupled cell: another rounding order on the coupled rows)
    for name in ("symGaussSeidel", "nonBlockingGaussSeidel", "DIC", "FDIC", "DICGaussSeidel", "multiColourGaussSeidel"):
        sm = w.smooth(name, xs, [r["source"] for r in regs], 3)[rank]
        psi = xs[rank].copy()
        ldub200.lduMatrix.smoother.New("p", A, name).smooth(psi, reg["source"], 3)
        results["smooth_" + name] = bool(np.array_equal(psi, sm))
    # GAMG with the non-blocking smoother on every level
    ctl = dict(solver="GAMG", smoother="nonBlockingGaussSeidel", agglomerator="faceAreaPair",
               nCellsInCoarsestLevel=4, mergeLevels=1, tolerance=1e-8, relTol=0)
    psi_o, perf_o = w.solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    psi = reg["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, reg["source"])
    results["gamg_nbgs"] = bool(perf.nIterations == perf_o["nIterations"] and np.array_equal(psi, psi_o[rank]))
    # GAMG with the multi-colour smoother on every level of every region
    ctl = dict(solver="GAMG", smoother="multiColourGaussSeidel", agglomerator="faceAreaPair",
               nCellsInCoarsestLevel=4, mergeLevels=1, tolerance=1e-8, relTol=0)
    psi_o, perf_o = w.solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
    psi = reg["psi0"].copy()
    perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=True)).solve(psi, reg["source"])
    results["gamg_mcgs"] = bool(perf.nIterations == perf_o["nIterations"] and np.array_equal(psi, psi_o[rank]))
    solves = [dict(solver="PCG", preconditioner="DIC", tolerance=1e-8, relTol=0),
              dict(solver="PCG", preconditioner="diagonal", tolerance=1e-7, relTol=0),
              dict(solver="GAMG", smoother="GaussSeidel", agglomerator="faceAreaPair", nCellsInCoarsestLevel=4,
                   mergeLevels=1, tolerance=1e-8, relTol=0)]
    for i, ctl in enumerate(solves):
        psi_o, perf_o = w.solve(ctl, [r["psi0"] for r in regs], [r["source"] for r in regs])
        for exact in (False, True):
            psi = reg["psi0"].copy()
            perf = ldub200.lduMatrix.solver.New("p", A, dict(ctl, referenceOrderSums=exact)).solve(psi, reg["source"])
            key = f"solve{i}_{'exact' if exact else 'fast'}"
            results[key + "_iters"] = [perf.nIterations, perf_o["nIterations"]]
            results[key + "_res"] = [perf.finalResidual, perf_o["finalResidual"]]
            results[key + "_psi"] = float(np.abs(psi - psi_o[rank]).max())
    A.destroy()
    dist.barrier()
    ctx.close()
    print("RESULT " + json.dumps(results))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
