#!/usr/bin/env python3
"""Build the UNMODIFIED reference application chain (OpenFOAM-2.2.x) into oracle/_ref/:

    libfileFormats  libsurfMesh  libtriSurface  libmeshTools  libfiniteVolume   -> icoFoam
    [--blockMesh]   libextrudeModel  libdynamicMesh  libblockMesh               -> blockMesh

TEST INFRASTRUCTURE ONLY (see build_ref.py).  This is what SURVEY.md 8c lists as
config 1's job ("icoFoam cavity, reference plumbing"): the real application, so
that `libs ("libgpuLduSolvers.so");` in system/controlDict can be shown to pick
the CUDA solvers up from an unmodified icoFoam.

Like build_ref.py it restates `wmake libso` / `wmake` for each Make/files with
the flags of wmake/rules/linux64Gcc; sources are compiled where they lie under
/root/reference through per-library flat include directories of symlinks
(wmakeLnInclude); nothing is copied into the repository.  Needs
oracle/_ref/libOpenFOAM.so (build_ref.py) first.

Portability fix for g++ 13, applied to a scratch copy of the header (same policy as build_ref.py):
  * oscillatingFixedValueFvPatchField.H:213-233: four accessors return the autoPtr<DataEntry<scalar>>
    members amplitude_/frequency_ as scalar.  They can never be instantiated (old g++ did not look at
    uninstantiated template members); they are dropped from the scratch copy.

The two flex sources of the chain (STL ASCII readers, surfMesh/…/STLsurfaceFormatASCII.L
and triSurface/…/readSTLASCII.L) cannot be generated here (no flex): their two entry
points are provided by oracle/foam_stubs/stlStubs.C, which raises FatalError when an
ASCII STL is actually read (no case in tests/ reads one).
"""
import os
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from build_ref import (REF, HERE, OUT, BUILD, CXX, CXXFLAGS, expand_make_files,  # noqa: E402
                       make_lninclude, prepare_lninclude)

# library -> (source dir under src/, libraries whose lnInclude it needs, libraries it links)
LIBS = {
    "fileFormats": ("fileFormats", [], []),
    "surfMesh": ("surfMesh", ["fileFormats"], ["fileFormats"]),
    "triSurface": ("triSurface", ["fileFormats", "surfMesh"], ["fileFormats", "surfMesh"]),
    "meshTools": ("meshTools", ["triSurface", "fileFormats"], ["triSurface", "fileFormats"]),
    "finiteVolume": ("finiteVolume", ["triSurface", "meshTools"], ["triSurface", "meshTools"]),
    # blockMesh chain (optional)
    "extrudeModel": ("mesh/extrudeModel", ["meshTools", "dynamicMesh"], ["meshTools"]),
    "dynamicMesh": ("dynamicMesh", ["finiteVolume", "meshTools", "triSurface", "extrudeModel"],
                    ["finiteVolume", "triSurface", "meshTools"]),
    "blockMesh": ("mesh/blockMesh", ["meshTools", "dynamicMesh"], ["meshTools", "dynamicMesh"]),
}
CHAIN_ICO = ["fileFormats", "surfMesh", "triSurface", "meshTools", "finiteVolume"]
CHAIN_BLOCKMESH = ["dynamicMesh", "extrudeModel", "blockMesh"]

APPS = {
    "icoFoam": ("applications/solvers/incompressible/icoFoam", ["icoFoam.C"], ["finiteVolume", "meshTools"],
                ["finiteVolume", "meshTools", "triSurface", "surfMesh", "fileFormats"]),
    "blockMesh": ("applications/utilities/mesh/generation/blockMesh", ["blockMeshApp.C"],
                  ["blockMesh", "meshTools", "dynamicMesh", "finiteVolume"],
                  ["blockMesh", "dynamicMesh", "extrudeModel", "finiteVolume", "meshTools", "triSurface",
                   "surfMesh", "fileFormats"]),
}


def ln_dir(lib):
    return BUILD / f"lnInclude_{lib}"


def header_fixes():
    """scratch copies of headers g++ 13 rejects (see the module docstring)"""
    import re
    name = "oscillatingFixedValueFvPatchField.H"
    link = ln_dir("finiteVolume") / name
    src = (REF / "src/finiteVolume/fields/fvPatchFields/derived/oscillatingFixedValue" / name).read_text()
    fixed = re.sub(r"//- Return amplitude\n\s*scalar amplitude\(\) const.*?scalar& frequency\(\)\s*\{[^}]*\}", "",
                   src, flags=re.S)
    assert fixed != src
    if link.is_symlink() or link.exists():
        link.unlink()
    link.write_text(fixed)


def build_lib(lib, jobs):
    srcdir, incs, links = LIBS[lib]
    src = REF / "src" / srcdir
    out = OUT / f"lib{lib}.so"
    if out.exists() and "--force" not in sys.argv:
        print(f"[build_ref_fv] {out.name} present")
        return 0
    ln = ln_dir(lib)
    make_lninclude(ln, [src])
    objdir = BUILD / f"obj_{lib}"
    objdir.mkdir(parents=True, exist_ok=True)
    inc = " ".join(f"-I{ln_dir(x)}" for x in [lib] + incs) + f" -I{BUILD / 'lnInclude'}"
    tus = []
    for rel in expand_make_files(src / "Make" / "files"):
        if rel.startswith("LIB"):
            continue
        if rel.endswith(".L"):
            continue            # flex source: served by foam_stubs/stlStubs.C
        tus.append(ln / os.path.basename(rel))
    extra = []
    if lib in ("surfMesh", "triSurface"):
        extra.append((HERE / "foam_stubs" / "stlStubs.C", f"-DSTUB_{lib}"))
    mk = [f"CXX={CXX}", f"CXXFLAGS={CXXFLAGS} {inc}", ""]
    objs, rules = [], []
    for i, s in enumerate(tus):
        o = objdir / f"{s.stem}_{i}.o"
        objs.append(str(o))
        rules.append(f"{o}: {s}\n\t@$(CXX) $(CXXFLAGS) -c {s} -o {o}\n")
    for s, fl in extra:
        o = objdir / f"stub_{s.stem}.o"
        objs.append(str(o))
        rules.append(f"{o}: {s}\n\t@$(CXX) $(CXXFLAGS) {fl} -c {s} -o {o}\n")
    dep_so = " ".join(str(OUT / f"lib{x}.so") for x in links if (OUT / f"lib{x}.so").exists() or x in CHAIN_ICO)
    link_flags = " ".join(f"-l{x}" for x in links)
    mk.append(f"{out}: " + " ".join(objs))
    mk.append(f"\t@echo linking {out}; $(CXX) -shared -o {out} " + " ".join(objs) +
              f" -L{OUT} {link_flags} -lOpenFOAM -Wl,-rpath,'$$ORIGIN'\n")
    mk.extend(rules)
    mkf = BUILD / f"Makefile_{lib}"
    mkf.write_text("\n".join(mk))
    print(f"[build_ref_fv] lib{lib}: {len(tus)} translation units, -j{jobs}", flush=True)
    del dep_so
    return subprocess.run(["make", "-f", str(mkf), f"-j{jobs}", "-k", str(out)], cwd=BUILD).returncode


def build_app(app):
    srcdir, srcs, incs, links = APPS[app]
    out = OUT / app
    src = REF / srcdir
    inc = f"-I{src} " + " ".join(f"-I{ln_dir(x)}" for x in incs) + f" -I{BUILD / 'lnInclude'}"
    cmd = (f"{CXX} {CXXFLAGS} {inc} " + " ".join(str(src / s) for s in srcs) + f" -o {out} -L{OUT} "
           "-Wl,--no-as-needed " + " ".join(f"-l{x}" for x in links) + " -lOpenFOAM -Wl,-rpath,$ORIGIN -ldl -lm")
    r = subprocess.run(cmd.split())
    print(f"[build_ref_fv] {app} -> exit {r.returncode}")
    return r.returncode


def main():
    if not REF.exists():
        print(f"[build_ref_fv] {REF} absent: using prebuilt oracle/_ref if any")
        return 0
    if not (OUT / "libOpenFOAM.so").exists():
        print("[build_ref_fv] run build_ref.py first")
        return 1
    if not (BUILD / "lnInclude" / "lduMatrix.H").exists():
        prepare_lninclude()
    jobs = int(os.environ.get("LDU_REF_JOBS", os.cpu_count() or 4))
    chain = list(CHAIN_ICO)
    if "--blockMesh" in sys.argv:
        # extrudeModel and dynamicMesh include each other's headers: flat include dirs first
        for lib in CHAIN_BLOCKMESH:
            make_lninclude(ln_dir(lib), [REF / "src" / LIBS[lib][0]])
        chain += CHAIN_BLOCKMESH
    for lib in chain:
        make_lninclude(ln_dir(lib), [REF / "src" / LIBS[lib][0]])
    header_fixes()
    for lib in chain:
        rc = build_lib(lib, jobs)
        if rc:
            return rc
    rc = build_app("icoFoam")
    if rc == 0 and "--blockMesh" in sys.argv:
        rc = build_app("blockMesh")
        # blockMesh needs the real cell-model table (hex, prism, ...): a data file of the reference's etc/,
        # placed beside the built binaries (oracle/_ref is build output, not repository content)
        import shutil
        shutil.copyfile(REF / "etc" / "cellModels", OUT / "etc" / "cellModels")
    return rc


if __name__ == "__main__":
    sys.exit(main())
