"""Differential fuzzing of the CPU restatement against the compiled reference (not collected by
pytest; `python tests/fuzz_oracle_vs_ref.py [nCases] [seed]`).  Random LDU graphs (not meshes: any
upper-triangular addressing), symmetric or asymmetric coefficients, random solver dictionaries,
initial guesses and partitions into 1-3 regions; everything must agree bit for bit.
tests/test_fuzz_seeded.py runs a short fixed-seed slice of the same generator."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "openfoam-2.2.x_b200"), str(ROOT / "tests")]

import cases  # noqa: E402
from ldub200 import decompose  # noqa: E402
from oracle import oracle as O  # noqa: E402


MAX_CELLS = int(__import__("os").environ.get("FUZZ_MAX_CELLS", "70"))   # larger: deeper GAMG hierarchies


def random_system(rng):
    n = int(rng.integers(1, MAX_CELLS))
    # a spanning chain (so GAMG can agglomerate) plus random extra faces
    pairs = set()
    order = rng.permutation(n)
    for a, b in zip(order[:-1], order[1:]):
        if rng.random() < 0.9:
            pairs.add((min(a, b), max(a, b)))
    for _ in range(int(rng.integers(0, 3 * n + 1))):
        a, b = rng.integers(0, n, 2)
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    faces = sorted(pairs)
    lower = np.array([p[0] for p in faces], dtype=np.int32)
    upper = np.array([p[1] for p in faces], dtype=np.int32)
    nf = lower.size
    asym = bool(rng.random() < 0.4)
    up = rng.uniform(0.2, 1.5, nf)
    lo = up * rng.uniform(0.6, 1.4, nf) if asym else None
    diag = np.zeros(n)
    np.subtract.at(diag, lower, up if lo is None else lo)
    np.subtract.at(diag, upper, up)
    diag -= rng.uniform(0.05, 0.5, n)            # strictly dominant
    return dict(nCells=n, nFaces=nf, lower=lower, upper=upper, diag=diag, upperCoef=up, lowerCoef=lo,
                source=rng.standard_normal(n), psi0=np.zeros(n), faceWeights=rng.uniform(0.5, 2.0, nf))


def random_controls(rng, s, multi):
    asym = s["lowerCoef"] is not None
    smoothers = [x for x in cases.SMOOTHERS if cases.selectable(s, x)]
    kind = rng.choice(["krylov", "smooth", "gamg", "krylov_gamg"], p=[0.4, 0.2, 0.3, 0.1])
    tol = float(rng.choice([1e-4, 1e-7, 1e-10]))
    common = dict(tolerance=tol, relTol=float(rng.choice([0, 0, 0.01])), maxIter=int(rng.choice([0, 3, 40, 1000])))
    gamg = dict(smoother=str(rng.choice(smoothers)), agglomerator=str(rng.choice(["algebraicPair", "faceAreaPair"])),
                nCellsInCoarsestLevel=int(rng.choice([2, 4, 10])), mergeLevels=int(rng.choice([1, 1, 2, 3])),
                cacheAgglomeration=False, nPreSweeps=int(rng.choice([0, 0, 1, 2])),
                nPostSweeps=int(rng.choice([1, 2, 3])), nFinestSweeps=int(rng.choice([1, 2])),
                interpolateCorrection=bool(rng.random() < 0.2))
    if rng.random() < 0.3:
        gamg["scaleCorrection"] = bool(rng.random() < 0.5)
    if kind == "krylov":
        pres = [p for p in cases.PRECONDITIONERS if cases.selectable(s, p)]
        return dict(common, solver="PBiCG" if asym else "PCG", preconditioner=str(rng.choice(pres)))
    if kind == "smooth":
        return dict(common, solver="smoothSolver", smoother=str(rng.choice(smoothers)),
                    nSweeps=int(rng.choice([1, 2, 3, -2])))
    if kind == "gamg" or asym:
        return dict(common, solver="GAMG", **gamg)
    return dict(common, solver="PCG", preconditioner=dict(gamg, preconditioner="GAMG", tolerance=1e-4, relTol=0,
                                                          nVcycles=int(rng.choice([1, 2]))))


def add_random_cyclic(rng, reg, region):
    """two disjoint random cell sets of the region coupled as a cyclic pair"""
    n = reg["nCells"]
    k = int(rng.integers(1, max(2, n // 3)))
    cells = rng.permutation(n)
    a, b = np.sort(cells[:k]), np.sort(cells[k:2 * k])
    if b.size < k:
        return
    asym = reg["lowerCoef"] is not None
    ku = rng.uniform(0.05, 0.3, k)
    kl = ku * rng.uniform(0.7, 1.3, k) if asym else ku
    reg["diag"] = reg["diag"].copy()
    np.subtract.at(reg["diag"], a, ku)
    np.subtract.at(reg["diag"], b, ku)
    first = len(reg.setdefault("interfaces", []))
    reg["interfaces"].append(dict(nbrRegion=region, nbrInterface=first + 1, faceCells=a.astype(np.int32),
                                  bouCoeffs=-ku, intCoeffs=-kl))
    reg["interfaces"].append(dict(nbrRegion=region, nbrInterface=first, faceCells=b.astype(np.int32),
                                  bouCoeffs=-kl, intCoeffs=-ku))


def one_operator_case(seed):
    """operators, preconditioners, smoothers, agglomeration maps and band compression on a random system
    (with a random cyclic pair every other time); -> None or a description of the difference"""
    rng = np.random.default_rng(10_000_000 + seed)
    s = random_system(rng)
    if s["nCells"] >= 6 and rng.random() < 0.5:
        add_random_cyclic(rng, s, 0)
    w = O.World([s])
    x = rng.standard_normal(s["nCells"])
    src = s["source"]
    for op, mine in (("amul", lambda: w.amul(x)[0]), ("tmul", lambda: w.tmul(x)[0]), ("suma", lambda: w.sumA()[0]),
                     ("residual", lambda: w.residual(x, src)[0])):
        if not np.array_equal(mine(), O.ref_run(s, op, psi=x)[0]):
            return f"operator seed {seed}: {op} differs"
    if not s.get("interfaces"):
        for op, mine in (("H", lambda: w.H(x)[0]), ("H1", lambda: w.H1()[0])):
            if not np.array_equal(mine(), O.ref_run(s, op, psi=x)[0]):
                return f"operator seed {seed}: {op} differs"
        if s["nFaces"] and not np.array_equal(w.faceH(x)[0], O.ref_run(s, "faceH", psi=x)[0]):
            return f"operator seed {seed}: faceH differs"
    for pre in cases.PRECONDITIONERS:
        if cases.selectable(s, pre):
            if not np.array_equal(w.precondition(pre, src)[0], O.ref_run(s, "precondition", pre)[0]):
                return f"operator seed {seed}: preconditioner {pre} differs"
    if s["lowerCoef"] is not None:
        if not np.array_equal(w.precondition("DILU", src, True)[0], O.ref_run(s, "preconditionT", "DILU")[0]):
            return f"operator seed {seed}: DILU preconditionT differs"
    nsw = int(rng.integers(1, 4))
    for sm in cases.SMOOTHERS:
        if cases.selectable(s, sm):
            want = O.ref_run(s, "smooth", O.dict_text(dict(smoother=sm)), nsw, psi=x)[0]
            if not np.array_equal(w.smooth(sm, x, src, nsw)[0], want):
                return f"operator seed {seed}: smoother {sm} x{nsw} differs"
    if not s.get("interfaces") and s["nCells"] >= 4:
        ctl = dict(solver="GAMG", smoother="GaussSeidel", nCellsInCoarsestLevel=int(rng.choice([2, 3, 5])),
                   mergeLevels=int(rng.choice([1, 2, 3])),
                   agglomerator=str(rng.choice(["algebraicPair", "faceAreaPair"])))
        mine = w.gamg_levels(ctl)
        try:
            ref = O.ref_agglom(s, cases.ref_controls(ctl))
        except RuntimeError:
            ref = []
        if len(mine) != len(ref) or any(a["nCoarse"] != b["nCoarse"] or not np.array_equal(a["restrict"], b["restrict"])
                                        for a, b in zip(mine, ref)):
            return f"operator seed {seed}: agglomeration maps differ ({len(mine)} vs {len(ref)} levels) {ctl}"
    if not s.get("interfaces"):
        from ldub200 import renumber
        perm = renumber.band_compression(s["nCells"], s["lower"], s["upper"])
        if not np.array_equal(np.argsort(perm), O.ref_run(s, "bandCompression", ints=True)[0]):
            return f"operator seed {seed}: bandCompression differs"
    return None


def one_multi_region_operator_case(seed):
    """operators and smoothers of 2-4 coupled regions (random partition, random cyclic pairs)"""
    if not O.ref_par_available():
        return None
    rng = np.random.default_rng(20_000_000 + seed)
    s = random_system(rng)
    if s["nCells"] < 8:
        return None
    R = int(rng.integers(2, 5))
    proc = rng.integers(0, R, s["nCells"]).astype(np.int32)
    proc[:R] = np.arange(R)
    regs = decompose.decompose(s, proc, R)
    for r, reg in enumerate(regs):
        if reg["nCells"] >= 6 and rng.random() < 0.3:
            add_random_cyclic(rng, reg, r)
    w = O.World(regs)
    x = rng.standard_normal(s["nCells"])
    xs = [x[r["cells"]] for r in regs]
    srcs = [r["source"] for r in regs]

    def same(a, b):
        return all(np.array_equal(p, q) for p, q in zip(a, b))

    for op, mine in (("amul", lambda: w.amul(xs)), ("tmul", lambda: w.tmul(xs)), ("suma", lambda: w.sumA()),
                     ("residual", lambda: w.residual(xs, srcs))):
        if not same(mine(), O.ref_run_par(regs, op, psi=xs, timeout=120)[0]):
            return f"multi-region operator seed {seed}: {op} differs (R={R})"
    nsw = int(rng.integers(1, 4))
    for sm in cases.SMOOTHERS:
        if cases.selectable(s, sm):
            want = O.ref_run_par(regs, "smooth", O.dict_text(dict(smoother=sm)), nsw, psi=xs, timeout=120)[0]
            if not same(w.smooth(sm, xs, srcs, nsw), want):
                return f"multi-region operator seed {seed}: smoother {sm} x{nsw} differs (R={R})"
    return None


def one_polymesh_case(seed):
    """random regions (processor + cyclic patches) -> processorN/constant/polyMesh files -> regions"""
    import tempfile
    from ldub200 import polymesh
    rng = np.random.default_rng(40_000_000 + seed)
    s = random_system(rng)
    if s["nCells"] < 8:
        return None
    R = int(rng.integers(1, 6))
    proc = rng.integers(0, R, s["nCells"]).astype(np.int32)
    proc[:R] = np.arange(R)
    regs = decompose.decompose(s, proc, R)
    for r, reg in enumerate(regs):
        if reg["nCells"] >= 6 and rng.random() < 0.4:
            add_random_cyclic(rng, reg, r)
    with tempfile.TemporaryDirectory() as td:
        polymesh.write_decomposed_case(td, regs)
        back = polymesh.read_decomposed_case(td)
    for r, (a, b) in enumerate(zip(back, regs)):
        if not (np.array_equal(a["lower"], b["lower"]) and np.array_equal(a["upper"], b["upper"])):
            return f"polymesh seed {seed}: addressing of region {r} differs"
        if [(i["nbrRegion"], i["nbrInterface"]) for i in a["interfaces"]] != \
                [(i["nbrRegion"], i["nbrInterface"]) for i in b["interfaces"]]:
            return f"polymesh seed {seed}: interface table of region {r} differs"
        for x, y in zip(a["interfaces"], b["interfaces"]):
            if not np.array_equal(x["faceCells"], y["faceCells"]):
                return f"polymesh seed {seed}: faceCells of region {r} differ"
    return None


def one_cache_case(seed):
    """two solves with changed coefficients in between, cacheAgglomeration on or off (driver op solve2)"""
    rng = np.random.default_rng(30_000_000 + seed)
    s = random_system(rng)
    if s["nCells"] < 12:
        return None
    smoothers = [x for x in cases.SMOOTHERS if cases.selectable(s, x)]
    ctl = dict(solver="GAMG", smoother=str(rng.choice(smoothers)),
               agglomerator=str(rng.choice(["algebraicPair", "faceAreaPair"])),
               nCellsInCoarsestLevel=int(rng.choice([2, 4])), mergeLevels=int(rng.choice([1, 2])),
               cacheAgglomeration=bool(rng.random() < 0.7), tolerance=1e-8, relTol=0, maxIter=60)
    w = O.World([s])
    try:
        w.solve(ctl, s["psi0"].copy(), s["source"])
        w.set_coeffs(0, *O.second_coeffs(s))
        psi2, perf2 = w.solve(ctl, s["psi0"].copy(), s["source"])
    except AssertionError:
        psi2 = None
    try:
        pr, so = O.ref_run(s, "solve2", O.dict_text(cases.ref_controls(ctl)))
        ref2 = O.parse_perfs(so)[1]
    except RuntimeError:
        pr = None
    if psi2 is None and pr is None:
        return None
    if psi2 is None or pr is None:
        return f"cache seed {seed}: one side refused {ctl}"
    if perf2["nIterations"] != ref2["nIterations"] or perf2["finalResidual"] != ref2["finalResidual"] \
            or not np.array_equal(psi2[0], pr):
        return f"cache seed {seed}: second solve differs {perf2} vs {ref2} {ctl}"
    return None


def one_case(seed):
    """-> None if oracle and reference agree (or both refuse), else a description of the difference"""
    rng = np.random.default_rng(seed)
    s = random_system(rng)
    R = int(rng.choice([1, 1, 2, 3])) if s["nCells"] >= 6 and O.ref_par_available() else 1
    ctl = random_controls(rng, s, R > 1)
    psi0 = rng.standard_normal(s["nCells"]) if rng.random() < 0.5 else np.zeros(s["nCells"])
    if R == 1:
        regs = [s]
        if s["nCells"] >= 6 and rng.random() < 0.25:
            add_random_cyclic(rng, s, 0)
    else:
        proc = rng.integers(0, R, s["nCells"]).astype(np.int32)
        proc[:R] = np.arange(R)                       # no empty region
        regs = decompose.decompose(s, proc, R)
        for r, reg in enumerate(regs):
            if reg["nCells"] >= 6 and rng.random() < 0.25:
                add_random_cyclic(rng, reg, r)
    if rng.random() < 0.1:                            # zero source: normFactor is 1e-20 plus |A psi| terms
        s["source"] = np.zeros(s["nCells"])
        for r in regs:
            r["source"] = np.zeros(r["nCells"])
    psis = [psi0[r["cells"]] for r in regs] if R > 1 else [psi0]
    srcs = [r["source"] for r in regs]
    try:
        po, perf = O.World(regs).solve(ctl, [p.copy() for p in psis], srcs)
        mine = (perf, po)
    except AssertionError:
        mine = None
    try:
        if R == 1:
            pr, so = O.ref_run(s, "solve", O.dict_text(cases.ref_controls(ctl)), psi=psi0)
            pr = [pr]
        else:
            pr, so = O.ref_run_par(regs, "solve", O.dict_text(cases.ref_controls(ctl)), psi=psis, timeout=120)
        ref = (O.parse_perf(so), pr)
    except RuntimeError as e:
        ref = None
        err = str(e)
    if mine is None and ref is None:
        return None
    if mine is None or ref is None:
        return f"seed {seed}: one side refused (oracle ok={mine is not None}, reference ok={ref is not None}) " \
               f"n={s['nCells']} R={R} ctl={ctl}" + ("" if ref else " :: " + err[-300:])
    for key in ("initialResidual", "finalResidual", "nIterations", "converged", "singular"):
        a, b = mine[0][key], ref[0][key]
        if not (a == b or (a != a and b != b)):
            return f"seed {seed}: {key} {a} != {b}  n={s['nCells']} nf={s['nFaces']} R={R} ctl={ctl}"
    for a, b in zip(mine[1], ref[1]):
        if not np.array_equal(a, b, equal_nan=True):
            return f"seed {seed}: psi differs (max {np.abs(a - b).max():.3e}) n={s['nCells']} R={R} ctl={ctl}"
    return None


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for k in range(n):
        for msg in (one_case(seed0 + k), one_operator_case(seed0 + k) if k % 2 == 0 else None,
                    one_multi_region_operator_case(seed0 + k) if k % 8 == 1 else None,
                    one_cache_case(seed0 + k) if k % 8 == 2 else None,
                    one_polymesh_case(seed0 + k) if k % 4 == 3 else None):
            if msg:
                bad += 1
                print(msg, flush=True)
    print(f"{n} cases from seed {seed0}: {bad} differences")
