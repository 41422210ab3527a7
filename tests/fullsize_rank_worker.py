"""One rank of the full-size decomposed bench workload through the CUDA path (launched by
tests/test_gpu_parity_fullsize.py): region RANK of the LDU_N^3 box cut into WORLD_SIZE blocks, PCG+DIC for
each iteration count in LDU_ITERS, with reference-order sums ("exact") and with the default tree sums.
Writes r<rank>.json = {mode: {iters: [nIterations, initialResidual, finalResidual]}} and the psi of the
longest solve of both modes into LDU_OUT."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))


def main():
    import torch
    import torch.distributed as dist
    import ldub200
    from ldub200 import decompose

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["LDU_PORT"],
                            rank=rank, world_size=world)
    n = int(os.environ["LDU_N"])
    out = Path(os.environ["LDU_OUT"])
    iters = [int(x) for x in os.environ["LDU_ITERS"].split(",")]
    reg = decompose.local_box_region(n, rank, world)
    ctx = ldub200.Context(rank % torch.cuda.device_count())
    max_if = max(it["faceCells"].size for it in reg["interfaces"])
    ctx.connect_torch_distributed(8, int(max_if))
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in reg["interfaces"]]
    A = ldub200.lduMatrix(ctx, reg["nCells"], reg["lower"], reg["upper"], ifs)
    A.set_coeffs(reg["diag"], reg["upperCoef"], None, [it["bouCoeffs"] for it in reg["interfaces"]],
                 [it["intCoeffs"] for it in reg["interfaces"]])
    res = {"exact": {}, "tree": {}}
    for mode, exact in (("exact", True), ("tree", False)):
        for it in iters:
            psi = reg["psi0"].copy()
            ctl = dict(solver="PCG", preconditioner="DIC", tolerance=0.0, relTol=0.0, maxIter=it - 1,
                       referenceOrderSums=exact)
            perf = ldub200.lduMatrix.solver.New("p", A, ctl).solve(psi, reg["source"])
            res[mode][str(it)] = [perf.nIterations, perf.initialResidual, perf.finalResidual]
        np.save(out / f"psi_{mode}_{rank}.npy", psi)
    (out / f"r{rank}.json").write_text(json.dumps(res))
    A.destroy()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
