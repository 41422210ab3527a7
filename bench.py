#!/usr/bin/env python3
"""bench.py — PCG iterations/s and Amul GB/s (fp64) on the 10M-cell box (216^3).

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU path)
    torchrun ... bench.py --gpus N ...                       (N > 1: one rank per GPU)

Workload (BASELINE.json metric / SURVEY.md §8d): icoFoam-style pressure matrix on
a 216^3 hex box (10,077,696 cells, 30,093,120 faces): symmetric 7-point Laplacian
as fvm::laplacian builds it, one reference cell, source sin(0.37 i), psi0 = 0,
solved with `solver PCG; preconditioner DIC` (the cavity tutorial's p solver).
N > 1: the same mesh split into 2x1x1 / 2x2x1 / 2x2x2 blocks, one region per GPU
(strong scaling), processor-patch halos + scalar all-reduces between GPUs.

A "step" is one solver call running exactly ITERS iterations (tolerance 0,
maxIter ITERS-1: the reference's do/while runs maxIter+1 iterations,
PCG.C:174-178).  value = K*ITERS / time = PCG iterations per second with all
fields resident in HBM.  e2e is the same through the host-pointer C ABI
(ldu_matrix_set_coeffs + ldu_solve: coefficients, source and psi cross PCIe
every step).  roofline is the Amul kernel alone: algorithmic bytes
(24 N + 16 F) / CUDA-event time.  Inputs (0.72 GB per Amul, >1 GB
per iteration) exceed the 126 MB L2, so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))
sys.path.insert(0, str(ROOT))

METRIC = "PCG iterations/sec (fp64, DIC) on 10M-cell box; Amul GB/s in roofline"
UNIT = "iterations/s"


def controls(precond, iters):
    return dict(solver="PCG", preconditioner=precond, tolerance=0.0, relTol=0.0, maxIter=iters - 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            t = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(t[0]))
                mx = max(mx, float(t[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------- #
# reference arm / CPU baseline: the unmodified reference compiled into oracle/_ref
# --------------------------------------------------------------------------- #
def reference_cores():
    """Host cores the reference arm uses: the reference is single-threaded per rank, so the arm
    runs it the way OpenFOAM is run on a multi-core host — the box decomposed into one mesh region
    per core, one reference process per region, coupled through processor interfaces — on up to
    32 cores, all the host offers.  This image has no MPI: the reference's Pstream seam is served by
    oracle/pstream_shm (shared memory) instead of src/Pstream/mpi."""
    try:
        c = len(os.sched_getaffinity(0))
    except Exception:
        c = os.cpu_count() or 1
    c = int(os.environ.get("LDU_REF_CORES", c))
    for k in (32, 16, 8, 4, 2):       # the region counts ldub200.decompose.split_for knows
        if c >= k:
            return k
    return 1


def reference_mode(cores):
    from oracle import oracle as O
    if cores == 1:
        return "single"
    return "coupled" if O.ref_par_available() else "blocks"


def reference_rate(n, precond, it_a, it_b, keep=None, cores=None):
    """Steady-state PCG iterations/s of the reference on the n^3 box on `cores` host cores.
    cores == 1: the plain single-rank reference.  cores > 1: the box decomposed into `cores`
    regions, one unmodified reference process per region, coupled (halo exchange through the
    processor interfaces, global sums through the Pstream seam) exactly as a decomposePar'd
    case runs; block-local DIC as in the reference.  Fallback when the multi-rank driver did not
    travel: the same blocks run concurrently but uncoupled (an upper bound of the coupled rate).
    Returns (rate, kind, cpu seconds, cores)."""
    from concurrent.futures import ThreadPoolExecutor
    from ldub200 import meshes, decompose
    from oracle import oracle as O
    cores = cores or reference_cores()
    key = ("sys", cores)
    if keep is not None and key in keep:
        blocks = keep[key]
    else:
        blocks = ([meshes.laplacian_system(n, n, n)] if cores == 1
                  else [decompose.local_box_region(n, r, cores) for r in range(cores)])
        if keep is not None:
            keep[key] = blocks
    ctl = O.dict_text(dict(solver="PCG", preconditioner=precond))

    def iters_line(so):
        t = [x for x in so.splitlines() if x.startswith("ITERS")][0].split()
        return int(t[1]), float(t[2]), int(t[3]), float(t[4])

    def rate(ia, ta, ib, tb):
        # difference of the two fixed-iteration solves; on systems so small that the difference drowns
        # in timer noise, the longer solve alone (set-up included)
        return (ib - ia) / (tb - ta) if tb - ta > 0.25 * tb else ib / max(tb, 1e-9)

    if reference_mode(cores) == "coupled":
        out, so = O.ref_run_par(blocks, "time_iters", ctl, it_a - 1, it_b - 1)
        ia, ta, ib, tb = iters_line(so)
        if keep is not None:
            keep["last"] = dict(perf=O.parse_perfs(so)[-1], psi=out)
        return rate(ia, ta, ib, tb), "reference", (tb + ta) * cores, cores

    def one(s):
        if O.ref_available():
            out, so = O.ref_run({k: v for k, v in s.items() if k != "interfaces"},
                                "time_iters", ctl, it_a - 1, it_b - 1)
            ia, ta, ib, tb = iters_line(so)
            if keep is not None and cores == 1:
                keep["last"] = dict(perf=O.parse_perfs(so)[-1], psi=[out])
            return rate(ia, ta, ib, tb), "reference", tb + ta
        # restatement (oracle/ldu_oracle.c) when the compiled reference did not travel
        s = {k: v for k, v in s.items() if k != "interfaces"}
        w = O.World([s])
        t0 = time.perf_counter()
        w.solve(controls(precond, it_a), s["psi0"], s["source"])
        t1 = time.perf_counter()
        w.solve(controls(precond, it_b), s["psi0"], s["source"])
        t2 = time.perf_counter()
        return (it_b - it_a) / ((t2 - t1) - (t1 - t0)), "port", t2 - t0

    with ThreadPoolExecutor(max_workers=cores) as ex:
        res = list(ex.map(one, blocks))
    return min(r[0] for r in res), res[0][1], sum(r[2] for r in res), cores


def workload_name(n, precond):
    """config.workload: the SAME string in both arms (how each arm cuts the mesh is config.decomposition)"""
    return f"box{n} PCG+{precond}, {n**3} cells"


def decomposition_name(regions, what):
    from ldub200 import decompose
    px, py, pz = decompose.split_for(regions)
    if regions == 1:
        return f"undecomposed (1 region, global DIC), {what}"
    return f"{px}x{py}x{pz} blocks = {regions} regions, processor-patch halos + global sums, block-local DIC, {what}"


def reference_sample(n, precond, iters, cores):
    mode = reference_mode(cores)
    if mode == "single":
        return f"{iters} steady-state iterations of the plain single-rank reference on the whole {n}^3 system"
    if mode == "coupled":
        return (f"{iters} steady-state iterations of the {n}^3 system decomposed into {cores} regions, one "
                f"unmodified reference process per region and core, coupled through processor interfaces over "
                f"the shared-memory Pstream (oracle/pstream_shm; no MPI in this image), block-local {precond}; "
                f"two fixed-iteration solves, the difference removes set-up")
    return (f"{iters} steady-state iterations of the {n}^3 system split into {cores} blocks, one reference "
            f"process per block and core, concurrently, no halo exchange (upper bound of the coupled rate)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    keep = {}
    rates = []
    total = args.warmup + args.steps
    cores = reference_cores()
    for i in range(total):
        r, kind, _, cores = reference_rate(args.n, args.precond, 2, 2 + args.ref_iters, keep)
        if i >= args.warmup:
            rates.append(r)
    value = float(np.mean(rates))
    sample = "per step: " + reference_sample(args.n, args.precond, args.ref_iters, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.ref_iters / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.n, args.precond), "precond": args.precond,
                   "decomposition": decomposition_name(cores, f"one reference process per region on {cores} host cores"),
                   "iters_per_step": args.ref_iters, "timing": "reference clockTime"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "last" in keep:
        line["final_residual"] = {"iterations": keep["last"]["perf"]["nIterations"],
                                  "value": keep["last"]["perf"]["finalResidual"]}
    print(json.dumps(line))


# --------------------------------------------------------------------------- #
# our arm
# --------------------------------------------------------------------------- #
def run_ours(args):
    import torch
    import ldub200
    from ldub200 import decompose

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    ctx = ldub200.Context(local_rank, stream.cuda_stream)
    n = args.n
    reg = decompose.local_box_region(n, rank, world)
    nC, nF = reg["nCells"], reg["nFaces"]
    if world > 1:
        max_if = max((it["faceCells"].size for it in reg["interfaces"]), default=1)
        ctx.connect_torch_distributed(8, int(max_if))
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in reg["interfaces"]]
    A = ldub200.lduMatrix(ctx, nC, reg["lower"], reg["upper"], ifs)
    bou = [it["bouCoeffs"] for it in reg["interfaces"]]
    inc = [it["intCoeffs"] for it in reg["interfaces"]]
    A.set_coeffs(reg["diag"], reg["upperCoef"], None, bou, inc)
    d_psi = ldub200.DeviceField(ctx, nC)
    d_src = ldub200.DeviceField(ctx, nC, reg["source"])
    d_tmp = ldub200.DeviceField(ctx, nC)
    solver = ldub200.lduMatrix.solver.New("p", A, controls(args.precond, args.iters))

    def timed(fn, reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident solve --------------------------------------------------
    last = {}

    def step_device():
        d_psi.zero()
        perf = solver.solve_device(d_psi, d_src)
        assert perf.nIterations == args.iters, str(perf)
        last["perf"] = perf

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ldub200.launch_count()
    ms = timed(step_device, args.steps)
    launches = ldub200.launch_count() - l0
    value = args.steps * args.iters / (ms * 1e-3)

    final_residual = last["perf"].finalResidual
    psi_dev = d_psi.download()       # psi of the last timed step (this rank's region)

    # ---- the HBM-bound curve beside it: the same solve with the diagonal preconditioner ------------
    solver_diag = ldub200.lduMatrix.solver.New("p", A, controls("diagonal", args.iters))

    def step_diag():
        d_psi.zero()
        solver_diag.solve_device(d_psi, d_src)

    for _ in range(3):
        step_diag()
    ms_diag = timed(step_diag, args.steps)
    value_diag = args.steps * args.iters / (ms_diag * 1e-3)

    # ---- Amul alone (roofline) ------------------------------------------------------
    d_src2 = ldub200.DeviceField(ctx, nC, np.sin(0.11 * np.arange(nC)))
    amul_reps = 100

    def time_amul():
        for _ in range(5):
            A.Amul_device(d_tmp, d_src2)
        return timed(lambda: A.Amul_device(d_tmp, d_src2), amul_reps) / amul_reps

    ms_amul = time_amul()            # the kernel the solve uses (box row kernel on blockMesh boxes)
    os.environ["LDU_AMUL_BOX"] = "0"
    ms_amul_generic = time_amul()    # the generic LDU row kernel (any mesh) on the same matrix
    del os.environ["LDU_AMUL_BOX"]
    clocks = sampler.stop() if rank == 0 else None
    # algorithmic bytes of one Amul on this rank's region: 24 N + 16 F (+20 P halo)
    nP = sum(it["faceCells"].size for it in reg["interfaces"])
    amul_bytes = 24 * nC + 16 * nF + 20 * nP
    amul_gbs = amul_bytes / (ms_amul * 1e-3) / 1e9
    amul_generic_gbs = amul_bytes / (ms_amul_generic * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- end to end through the host-pointer ABI ----------------------------------------
    # What the reference-facing plug-in does per solve (foam/gpuLduSolvers.C): ldu_matrix_set_coeffs with the
    # application's diag/upper arrays + ldu_solve with its psi/source -- ORDINARY (pageable) host memory, which
    # the library stages through its pinned ring with several host threads (csrc/context.cu copy_h2d).  The
    # headline e2e uses exactly that; `pinned` is the same with page-locked buffers (ldu_host_alloc).
    # psi is in/out: every step needs its own initial guess psi0 = 0, zeroed BEFORE the timed region (an
    # application's psi is simply there, it is not cleared inside the solve call), each used once.
    e2e_steps = max(1, min(args.steps, 5))

    def run_e2e(alloc):
        h_diag = alloc(nC); h_diag[:] = reg["diag"]
        h_upper = alloc(nF); h_upper[:] = reg["upperCoef"]
        h_src = alloc(nC); h_src[:] = reg["source"]
        h_psis = [alloc(nC) for _ in range(e2e_steps + 1)]
        for h in h_psis:
            h[:] = 0.0

        def step_e2e(h_psi):
            A.set_coeffs(h_diag, h_upper, None, bou, inc)
            perf = solver.solve(h_psi, h_src)
            assert perf.nIterations == args.iters

        step_e2e(h_psis[0])
        barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            step_e2e(h_psis[k + 1])
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return e2e_steps * args.iters / dt

    e2e_value = run_e2e(lambda k: np.empty(k))
    e2e_pinned = run_e2e(ldub200.pinned_array)
    h2d = 8 * (nC + nF + nC + nC) + 16 * nP
    d2h = 8 * nC

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, args.precond),
                   "decomposition": decomposition_name(world, "one region per B200"),
                   "precond": args.precond, "iters_per_step": args.iters,
                   "l2": "inputs larger than L2 (0.72 GB per Amul), no flush",
                   "pcg_alg_bytes_per_cell_iter": 352 if args.precond == "DIC" else 208},
        "roofline": {"bound": "hbm", "kernel": "box_row_kernel<0,6> (Amul on a blockMesh box: addressing implicit)",
                     "achieved": amul_gbs, "peak": peak, "unit": "GB/s", "frac": amul_gbs / peak, "traffic": None,
                     "peak_source": peak_src, "amul_ms": ms_amul, "alg_bytes_per_launch": amul_bytes,
                     "note": "algorithmic bytes = LDU-minimal 24N+16F (SURVEY 8d); the box kernel reads no "
                             "addressing (24N+8F from DRAM), hence frac may exceed 1; generic = the row kernel "
                             "for arbitrary LDU addressing on the same matrix",
                     "generic": {"kernel": "row_kernel<0,1,8>", "achieved": amul_generic_gbs,
                                 "frac": amul_generic_gbs / peak, "amul_ms": ms_amul_generic, "traffic": None}},
        # the whole PCG iteration against the same roofline: unfused algorithmic bytes of SURVEY.md 8d
        # (352 B/cell/iteration with DIC, 208 with diagonal) x cells of all regions x iterations/s, per GPU
        "pcg_roofline": {"alg_bytes_per_cell_iter": 352 if args.precond == "DIC" else 208,
                         "achieved_gbs_per_gpu": value * (352 if args.precond == "DIC" else 208) * n ** 3 / 1e9 / world,
                         "frac": value * (352 if args.precond == "DIC" else 208) * n ** 3 / 1e9 / world / peak,
                         "bound": "dependency chain of the DIC sweeps (nx+ny+nz-2 hyperplanes per sweep), not HBM"
                                  if args.precond == "DIC" else "hbm"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "host_memory": "pageable (ordinary arrays, as the OpenFOAM plug-in passes them; "
                                                   "staged by the library through a multi-threaded pinned ring)",
                "pinned": {"value": e2e_pinned, "unit": UNIT,
                           "host_memory": "page-locked buffers from ldu_host_alloc"}},
        "pcg_diagonal": {"value": value_diag, "unit": UNIT, "ms_per_step": ms_diag / args.steps,
                         "alg_bytes_per_cell_iter": 208,
                         "frac_of_hbm_peak": value_diag * 208 * n ** 3 / 1e9 / world / peak,
                         "note": "same solve with `preconditioner diagonal`: no dependency chain, the HBM-bound "
                                 "curve of the PCG loop (Amul + fused BLAS-1 kernels)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    traffic_file = ROOT / "profiles" / "amul_dram_bytes.json"
    if traffic_file.exists() and world == 1 and n == 216:   # captured on the 216^3 single-region launches
        try:
            t = json.loads(traffic_file.read_text())
            line["roofline"]["traffic"] = t.get("box_dram_bytes_per_launch")
            line["roofline"]["generic"]["traffic"] = t.get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- parity: the unmodified reference on the SAME decomposition, the same number of iterations ----
    if not args.no_parity:
        par = None
        if rank == 0:
            try:
                from oracle import oracle as O
                keep = {}
                t0 = time.perf_counter()
                rate1, kind, _, _ = reference_rate(n, args.precond, 2, args.iters, keep, cores=world)
                pr = keep["last"]["perf"]
                psi_ref = keep["last"]["psi"][0]
                par = {"mode": "default (fixed-shape tree sums; referenceOrderSums off), the timed configuration",
                       "reference": f"oracle/_ref ({kind}), {world} region(s), same decomposition, "
                                    f"{pr['nIterations']} iterations",
                       "iterations": [int(last["perf"].nIterations), int(pr["nIterations"])],
                       "iterations_equal": bool(last["perf"].nIterations == pr["nIterations"]),
                       "final_residual": final_residual, "reference_final_residual": pr["finalResidual"],
                       "rel_diff_final_residual": abs(final_residual - pr["finalResidual"]) / pr["finalResidual"],
                       "max_rel_diff_psi_region0": float(np.abs(psi_dev - psi_ref).max() / np.abs(psi_ref).max()),
                       "exact_mode": "referenceOrderSums on is bit-identical at this size "
                                     "(tests/test_gpu_parity_fullsize.py)",
                       "reference_rate_same_decomposition": rate1,
                       "seconds": time.perf_counter() - t0}
            except Exception as e:
                par = {"failed": str(e)[:300]}
        if dist is not None:
            dist.barrier()
        if rank == 0:
            line["parity"] = par

    # ---- the second workload, compact (BASELINE configs 3/5; the full line: --workload gamg) ------------
    if world == 1 and not args.no_gamg:
        try:
            g = measure_gamg(ldub200, torch, ctx, stream, dist, gamg_mesh(args.gamg_mesh), rank, world,
                             ["GaussSeidel", "multiColourGaussSeidel"], args.gamg_cycles, barrier, with_reference=False)
            g["note"] = ("GaussSeidel = the reference's lexicographic smoother (bit-identical path); "
                         "multiColourGaussSeidel = the same sweep colour by colour, per GAMG level; reference "
                         "V-cycles/s and iteration counts beside them: profiles/r02_gamg_*.json (--workload gamg)")
            line["gamg"] = g
        except Exception as e:
            line["gamg"] = {"failed": str(e)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, kind, spent, cores = reference_rate(n, args.precond, 2, 2 + args.ref_iters)
            if line.get("parity") and "reference_rate_same_decomposition" in line["parity"]:
                v1, spent1 = line["parity"]["reference_rate_same_decomposition"], line["parity"]["seconds"]
            else:
                v1, _, spent1, _ = reference_rate(n, args.precond, 2, 2 + args.ref_iters, cores=1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": reference_sample(n, args.precond, args.ref_iters, cores)
                                              + f" ({spent:.1f} s of CPU work)",
                                    "single_rank": {"value": v1, "cores": 1,
                                                    "sample": reference_sample(n, args.precond, args.ref_iters, 1)
                                                              + f" ({spent1:.1f} s of CPU work)"}}
        except Exception as e:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": reference_cores(), "kind": "reference",
                                    "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line))
    A.destroy()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()



# --------------------------------------------------------------------------- #
# second workload: GAMG (BASELINE configs 3 and 5)
# --------------------------------------------------------------------------- #
GAMG_METRIC = "GAMG V-cycles/sec (fp64, GaussSeidel smoother) on a ~2M-cell mesh"
GAMG_UNIT = "V-cycles/s"
# fixed-cycle timing runs: a tolerance the cycles never reach.  Not 0: GAMG hands its tolerance to the solver of
# the coarsest level (GAMGSolverSolve.C:430-487), which would then burn its 1000 iterations in every cycle
GAMG_FIXED_TOL = 1e-14


def gamg_controls(smoother, **kw):
    """the p solver of the simpleFoam tutorials (pitzDaily / motorBike system/fvSolution:19-31)"""
    d = dict(solver="GAMG", smoother=smoother, agglomerator="faceAreaPair", nCellsInCoarsestLevel=10, mergeLevels=1,
             cacheAgglomeration=True, nPreSweeps=0, nPostSweeps=2, nFinestSweeps=2)
    d.update(kw)
    return d


def gamg_mesh(name):
    """box128: 128^3 hex box (2,097,152 cells, the 3-D stand-in for a ~2M-cell case); sheet1448: 1448 x 1448 x 1
    (2,096,704 cells: pitzDaily is 2-D, one cell thick); boxN / sheetN for other sizes"""
    if name.startswith("sheet"):
        n = int(name[5:])
        return n, n, 1
    n = int(name[3:])
    return n, n, n


def gamg_region(shape, rank, world):
    from ldub200 import meshes, decompose
    nx, ny, nz = shape
    if world == 1:
        s = meshes.laplacian_system(nx, ny, nz)
        s["interfaces"] = []
        return s
    if not (nx == ny == nz):
        raise SystemExit("the multi-GPU GAMG workload is the cubic box (decompose.local_box_region)")
    return decompose.local_box_region(nx, rank, world)


def reference_gamg(shape, ctl, cores, cycles):
    """(V-cycles/s, iterations of the tolerance-driven solve or None) of the unmodified reference on `cores` cores:
    cores == 1 the plain single-rank reference, else the box cut into `cores` regions, one coupled process each"""
    from oracle import oracle as O
    txt = O.dict_text(dict(ctl, agglomerator="weightedPair"))    # = faceAreaPair on this mesh (oracle/ref_driver.C)
    if cores == 1:
        s = gamg_region(shape, 0, 1)
        s = {k: v for k, v in s.items() if k != "interfaces"}
        _, so = O.ref_run(s, "time_iters", txt, 2, 2 + cycles, GAMG_FIXED_TOL)
    else:
        blocks = [gamg_region(shape, r, cores) for r in range(cores)]
        _, so = O.ref_run_par(blocks, "time_iters", txt, 2, 2 + cycles, GAMG_FIXED_TOL)
    t = [x for x in so.splitlines() if x.startswith("ITERS")][0].split()
    ia, ta, ib, tb = int(t[1]), float(t[2]), int(t[3]), float(t[4])
    return (ib - ia) / (tb - ta)


def measure_gamg(ldub200, torch, ctx, stream, dist, shape, rank, world, smoothers, cycles, barrier, with_reference):
    """V-cycle time of our GAMG (difference of two fixed-cycle solves: set-up and prologue cancel), the iteration
    count of a tolerance-driven solve per smoother, and (rank 0, N = 1) the reference beside it."""
    reg = gamg_region(shape, rank, world)
    nC = reg["nCells"]
    ifs = [ldub200.lduInterface(it["faceCells"], it["nbrRegion"], it["nbrInterface"]) for it in reg["interfaces"]]
    A = ldub200.lduMatrix(ctx, nC, reg["lower"], reg["upper"], ifs)
    A.set_coeffs(reg["diag"], reg["upperCoef"], None, [it["bouCoeffs"] for it in reg["interfaces"]],
                 [it["intCoeffs"] for it in reg["interfaces"]])
    A.set_face_weights(reg["faceWeights"])
    d_psi = ldub200.DeviceField(ctx, nC)
    d_src = ldub200.DeviceField(ctx, nC, reg["source"])
    out = {"cells": int(np.prod(shape)), "regions": world, "settings": "GAMG, faceAreaPair, nCellsInCoarsestLevel 10, "
           "mergeLevels 1, nPreSweeps 0, nPostSweeps 2, nFinestSweeps 2 (simpleFoam tutorials)", "smoothers": {}}

    def timed_solve(ctl):
        solver = ldub200.lduMatrix.solver.New("p", A, ctl)
        best = None
        for _ in range(3):
            d_psi.zero()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            perf = solver.solve_device(d_psi, d_src)
            e1.record(stream)
            barrier()
            ms = e0.elapsed_time(e1)
            if dist is not None:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = ms if best is None else min(best, ms)
        return best, perf

    for sm in smoothers:
        l0 = ldub200.launch_count()
        ms_a, _ = timed_solve(gamg_controls(sm, tolerance=GAMG_FIXED_TOL, relTol=0, maxIter=2))
        ms_b, perf_b = timed_solve(gamg_controls(sm, tolerance=GAMG_FIXED_TOL, relTol=0, maxIter=2 + cycles))
        ms_cycle = (ms_b - ms_a) / cycles
        _, perf_tol = timed_solve(gamg_controls(sm, tolerance=1e-7, relTol=0, maxIter=200))
        # the stopping cycle of the default (tree-sum) mode can differ by one from the reference's when a residual
        # lands within rounding of the tolerance; with reference-order sums every cycle is bit-identical
        solver = ldub200.lduMatrix.solver.New("p", A, gamg_controls(sm, tolerance=1e-7, relTol=0, maxIter=200,
                                                                    referenceOrderSums=True))
        d_psi.zero()
        perf_exact = solver.solve_device(d_psi, d_src)
        out["smoothers"][sm] = {"ms_per_vcycle": ms_cycle, "vcycles_per_s": 1e3 / ms_cycle,
                                "levels": len(A.gamg_levels(gamg_controls(sm))) if world == 1 else None,
                                "iterations_to_1e-7": int(perf_tol.nIterations),
                                "final_residual": perf_tol.finalResidual,
                                "iterations_to_1e-7_reference_order_sums": int(perf_exact.nIterations),
                                "final_residual_reference_order_sums": perf_exact.finalResidual,
                                "residual_after_%d_cycles" % (2 + cycles): perf_b.finalResidual}
    if with_reference and rank == 0:
        try:
            from oracle import oracle as O
            ctl = gamg_controls("GaussSeidel")
            ref_cycles = 6
            r1 = reference_gamg(shape, ctl, 1, ref_cycles)
            out["reference"] = {"single_rank": {"vcycles_per_s": r1, "cores": 1}}
            s1 = gamg_region(shape, 0, 1)
            s1 = {k: v for k, v in s1.items() if k != "interfaces"}
            _, so = O.ref_run(s1, "solve", O.dict_text(dict(gamg_controls("GaussSeidel", tolerance=1e-7, relTol=0,
                                                                          maxIter=200), agglomerator="weightedPair")))
            pr = O.parse_perf(so)
            out["reference"]["iterations_to_1e-7"] = pr["nIterations"]
            out["reference"]["final_residual"] = pr["finalResidual"]
            cores = reference_cores()
            if cores > 1 and shape[0] == shape[1] == shape[2]:
                out["reference"]["all_cores"] = {"vcycles_per_s": reference_gamg(shape, ctl, cores, ref_cycles),
                                                 "cores": cores}
        except Exception as e:
            out["reference"] = {"failed": str(e)[:300]}
    A.destroy()
    return out


def run_gamg(args):
    import torch
    import ldub200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    ctx = ldub200.Context(local_rank, stream.cuda_stream)
    shape = gamg_mesh(args.gamg_mesh)
    if world > 1:
        ctx.connect_torch_distributed(8, int(shape[0] * shape[0]))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = measure_gamg(ldub200, torch, ctx, stream, dist, shape, rank, world,
                       ["GaussSeidel", "multiColourGaussSeidel"], args.gamg_cycles, barrier,
                       with_reference=(world == 1 and not args.no_cpu_baseline))
    clocks = sampler.stop() if rank == 0 else None
    lex, mc = res["smoothers"]["GaussSeidel"], res["smoothers"]["multiColourGaussSeidel"]
    line = {"metric": GAMG_METRIC, "value": mc["vcycles_per_s"], "unit": GAMG_UNIT, "n_gpus": world,
            "steps": args.gamg_cycles, "warmup": 2, "ms_per_step": mc["ms_per_vcycle"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.gamg_mesh} GAMG, {res['cells']} cells", "regions": world,
                       "smoother": "multiColourGaussSeidel (value); GaussSeidel = the reference's lexicographic "
                                   "smoother, bit-identical path, beside it"},
            "gamg": res, "clocks": clocks, "gpu_launches": int(ldub200.launch_count())}
    if "reference" in res and "single_rank" in res["reference"]:
        best = res["reference"].get("all_cores", res["reference"]["single_rank"])
        line["cpu_baseline"] = {"value": best["vcycles_per_s"], "unit": GAMG_UNIT, "cores": best["cores"],
                                "kind": "reference", "sample": "6 steady-state V-cycles (difference of two "
                                "fixed-cycle solves) of the same mesh and settings, lexicographic GaussSeidel"}
        line["parity"] = {"iterations_to_1e-7": {"reference": res["reference"].get("iterations_to_1e-7"),
                                                 "GaussSeidel": lex["iterations_to_1e-7"],
                                                 "GaussSeidel_reference_order_sums":
                                                     lex["iterations_to_1e-7_reference_order_sums"],
                                                 "final_residual_bit_identical":
                                                     lex["final_residual_reference_order_sums"]
                                                     == res["reference"].get("final_residual"),
                                                 "multiColourGaussSeidel": mc["iterations_to_1e-7"]}}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=216, help="box edge (216 -> 10,077,696 cells)")
    ap.add_argument("--precond", default="DIC")
    ap.add_argument("--iters", type=int, default=50, help="PCG iterations per step")
    ap.add_argument("--ref-iters", type=int, default=50, help="reference iterations timed per step")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference run behind the parity object")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="pcg", choices=["pcg", "gamg"],
                    help="pcg: the headline (BASELINE metric); gamg: V-cycles/s on a ~2M-cell mesh (configs 3, 5)")
    ap.add_argument("--gamg-mesh", default="box128", help="box<N> (N^3 hex box) or sheet<N> (N x N x 1, 2-D)")
    ap.add_argument("--gamg-cycles", type=int, default=10)
    ap.add_argument("--no-gamg", action="store_true", help="pcg workload: skip the compact GAMG measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "gamg":
        run_gamg(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
