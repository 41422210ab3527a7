#!/usr/bin/env python3
"""Build the UNMODIFIED reference libOpenFOAM.so (OpenFOAM-2.2.x) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may execute what this script produces.

The reference's own build system (wmake) is NOT run (it needs flex, absent in
this image).  This script restates what `wmake libso` does for the three
libraries the lduMatrix path needs:

  src/OpenFOAM/Make/files          (415 translation units -> libOpenFOAM)
  src/OSspecific/POSIX/Make/files  (libOSspecific.o, linked into libOpenFOAM,
                                    see src/OpenFOAM/Make/options)
  src/Pstream/dummy/Make/files     (serial Pstream stubs)

and, for multi-rank runs of the reference without MPI (absent in this image),
  oracle/_ref/libPstream_shm.so    oracle/pstream_shm/shmPstream.C: the Pstream seam over shared
                                   memory.  Like the reference's own libPstream (dummy vs mpi,
                                   picked by LD_LIBRARY_PATH) it interposes on the stubs:
                                   ref_driver_par lists it before libOpenFOAM.so, which stays
                                   the unmodified serial build.

with the flags of wmake/rules/linux64Gcc/{c++,c++Opt,general}:
  g++ -m64 -Dlinux64 -DWM_DP -DNoRepository -ftemplate-depth-100 -O3 -fPIC

Sources are compiled where they lie under /root/reference (through a flat
directory of symlinks, which is what wmakeLnInclude produces); no reference
source is copied into the repository.  Scratch objects go to a build directory
outside the repo (default /tmp/ldu_b200_refbuild); the only outputs kept are
  oracle/_ref/libOpenFOAM.so
  oracle/_ref/etc/controlDict      (a minimal global controlDict we author)
  oracle/_ref/lnInclude.txt        (where the flat include dir lives, for the
                                    driver / plug-in builds in this container)

Portability fixes for g++ 13 / glibc >= 2.34 (all applied to scratch copies):
  * PackedBoolList.H:203  operator=(const UList<bool>&) names the private
    base's injected class name; qualified as Foam::UList<bool>.
  * global.Cver: VERSION_STRING/BUILD_STRING substituted (wmake/rules/General/version).
  * cachedRandom.H is listed as a source in Make/files: compiled with -x c++.
  * sigFpe.C compiled with -Ulinux (no __malloc_hook in modern glibc).
"""
import os
import re
import subprocess
import sys
from pathlib import Path

REF = Path(os.environ.get("LDU_REFERENCE", "/root/reference"))
HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
BUILD = Path(os.environ.get("LDU_REF_BUILD", "/tmp/ldu_b200_refbuild"))

CXX = "g++"
CXXFLAGS = ("-std=gnu++98 -m64 -Dlinux64 -DWM_DP -DNoRepository "
            "-ftemplate-depth-100 -O3 -fPIC -w")


def expand_make_files(path: Path, defines=()):
    """Expand a wmake Make/files: $(var) substitution, /* */ comments, cpp #if."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    var = {}
    out = []
    skip_stack = []
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("#"):
            m = re.match(r"#\s*if\s+!defined\((\w+)\)", line)
            if m:
                skip_stack.append(m.group(1) in defines)
                continue
            m = re.match(r"#\s*ifdef\s+(\w+)", line)
            if m:
                skip_stack.append(m.group(1) not in defines)
                continue
            if re.match(r"#\s*else", line):
                skip_stack[-1] = not skip_stack[-1]
                continue
            if re.match(r"#\s*endif", line):
                skip_stack.pop()
                continue
            continue
        if any(skip_stack):
            continue
        m = re.match(r"(\w+)\s*=\s*(.*)$", line)
        if m:
            val = m.group(2)
            val = re.sub(r"\$\((\w+)\)", lambda mm: var.get(mm.group(1), ""), val)
            var[m.group(1)] = val
            continue
        line = re.sub(r"\$\((\w+)\)", lambda mm: var.get(mm.group(1), ""), line)
        out.append(line)
    return out


def make_lninclude(dst: Path, roots):
    dst.mkdir(parents=True, exist_ok=True)
    pat = re.compile(r".*\.([CHh]|[ch]xx|[ch]pp|type)$")
    prune = {"lnInclude", "Make", "config", "noLink"}
    for root in roots:
        for dirpath, dirnames, filenames in os.walk(root):
            dirnames[:] = [d for d in dirnames if d not in prune]
            for fn in filenames:
                if pat.match(fn):
                    link = dst / fn
                    if not link.exists() and not link.is_symlink():
                        link.symlink_to(Path(dirpath) / fn)


def prepare_lninclude() -> Path:
    """The flat include directory (scratch, symlinks into the reference) + the one header fix."""
    ln = BUILD / "lnInclude"
    make_lninclude(ln, [REF / "src/OpenFOAM", REF / "src/OSspecific/POSIX"])
    pbl = ln / "PackedBoolList.H"
    src = (REF / "src/OpenFOAM/containers/Lists/PackedList/PackedBoolList.H").read_text()
    src = src.replace("PackedBoolList& operator=(const UList<bool>&);",
                      "PackedBoolList& operator=(const Foam::UList<bool>&);")
    if pbl.is_symlink() or pbl.exists():
        pbl.unlink()
    pbl.write_text(src)
    return ln


def main():
    if not REF.exists():
        print(f"[build_ref] {REF} absent: using prebuilt oracle/_ref if any")
        return 0 if (OUT / "libOpenFOAM.so").exists() else 1
    if (OUT / "libOpenFOAM.so").exists() and "--force" not in sys.argv:
        print("[build_ref] oracle/_ref/libOpenFOAM.so present (use --force to rebuild)")
        srcs = [HERE / "ref_driver.C"] + sorted((HERE / "pstream_shm").glob("*"))
        newest = max(p.stat().st_mtime for p in srcs)
        outs = [OUT / "ref_driver", OUT / "ref_driver_par", OUT / "libPstream_shm.so"]
        if not (BUILD / "lnInclude" / "lduMatrix.H").exists():
            # scratch directory gone (new container): the plug-in build needs the flat include directory
            (OUT / "lnInclude.txt").write_text(str(prepare_lninclude()) + "\n")
        if all(o.exists() and o.stat().st_mtime >= newest for o in outs):
            return 0
        return build_driver(prepare_lninclude())

    OUT.mkdir(parents=True, exist_ok=True)
    (BUILD / "obj").mkdir(parents=True, exist_ok=True)
    ln = prepare_lninclude()

    gver = (REF / "src/OpenFOAM/global/global.Cver").read_text()
    gver = gver.replace("VERSION_STRING", "2.2.x").replace("BUILD_STRING", "2.2.x-1f35a0ff")
    (ln / "global_ver.C").write_text(gver)

    # -- translation units ----------------------------------------------------
    tus = []  # (source path as seen by the compiler, extra flags)
    for rel in expand_make_files(REF / "src/OpenFOAM/Make/files"):
        if rel.startswith("LIB"):
            continue
        base = os.path.basename(rel)
        if rel.endswith(".Cver"):
            tus.append((ln / "global_ver.C", f"-I{REF}/src/OpenFOAM/global"))
        elif rel.endswith(".H"):
            tus.append((ln / base, "-x c++"))
        else:
            tus.append((ln / base, ""))
    for rel in expand_make_files(REF / "src/OSspecific/POSIX/Make/files"):
        if rel.startswith("LIB"):
            continue
        base = os.path.basename(rel)
        tus.append((ln / base, "-Ulinux" if base == "sigFpe.C" else ""))
    # Pstream/dummy sources have the same basenames as files in lnInclude
    # (UPstream.C ...): compile them from their own directory.
    for rel in expand_make_files(REF / "src/Pstream/dummy/Make/files"):
        if rel.startswith("LIB"):
            continue
        tus.append((REF / "src/Pstream/dummy" / rel, "@pstream"))

    # check basenames are unique inside lnInclude-compiled set
    mk = [f"CXX={CXX}", f"CXXFLAGS={CXXFLAGS} -I{ln}", "OBJS=", ""]
    objs = []
    rules = []
    for i, (srcp, extra) in enumerate(tus):
        tag = "ps_" if extra == "@pstream" else ""
        obj = BUILD / "obj" / f"{tag}{srcp.stem}_{i}.o"
        ex = "" if extra == "@pstream" else extra
        objs.append(str(obj))
        rules.append(f"{obj}: {srcp}\n\t@$(CXX) $(CXXFLAGS) {ex} -c {srcp} -o {obj}\n")
    lib = OUT / "libOpenFOAM.so"
    mk.append(f"{lib}: " + " ".join(objs))
    mk.append(f"\t@echo linking {lib}; $(CXX) -shared -o {lib} " + " ".join(objs) + " -lz -ldl -lm\n")
    mk.extend(rules)
    (BUILD / "Makefile").write_text("\n".join(mk))

    jobs = os.cpu_count() or 4
    print(f"[build_ref] compiling {len(tus)} translation units with -j{jobs}")
    r = subprocess.run(["make", "-f", str(BUILD / "Makefile"), f"-j{jobs}", str(lib)],
                       cwd=BUILD)
    if r.returncode != 0:
        return r.returncode
    (OUT / "lnInclude.txt").write_text(str(ln) + "\n")
    write_control_dict(OUT / "etc" / "controlDict")
    print(f"[build_ref] built {lib}")
    return build_driver(ln)


def build_driver(ln: Path):
    """Compile oracle/ref_driver.C against the reference headers + library."""
    cmd = (f"{CXX} {CXXFLAGS} -I{ln} {HERE / 'ref_driver.C'} -o {OUT / 'ref_driver'} "
           f"-L{OUT} -lOpenFOAM -Wl,-rpath,$ORIGIN -ldl -lm")
    r = subprocess.run(cmd.split())
    print(f"[build_ref] ref_driver -> exit {r.returncode}")
    if r.returncode:
        return r.returncode
    # the shared-memory Pstream and the multi-rank driver: same source, the seam listed first
    shm = HERE / "pstream_shm"
    cmd = f"{CXX} {CXXFLAGS} -shared -I{shm} -I{ln} {shm / 'shmPstream.C'} -o {OUT / 'libPstream_shm.so'}"
    r = subprocess.run(cmd.split())
    print(f"[build_ref] libPstream_shm.so -> exit {r.returncode}")
    if r.returncode:
        return r.returncode
    cmd = (f"{CXX} {CXXFLAGS} -I{ln} {HERE / 'ref_driver.C'} -o {OUT / 'ref_driver_par'} "
           f"-L{OUT} -Wl,--no-as-needed -lPstream_shm -lOpenFOAM -Wl,-rpath,$ORIGIN -ldl -lm")
    r = subprocess.run(cmd.split())
    print(f"[build_ref] ref_driver_par -> exit {r.returncode}")
    return r.returncode


def write_control_dict(path: Path):
    """A minimal global controlDict (our own file, values are the SI constants
    and the OptimisationSwitches the shipped etc/controlDict:47-66 sets)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    path.write_text(CONTROL_DICT)
    # cellModeller's static initialiser insists on etc/cellModels
    # (meshes/meshShapes/cellModeller/globalCellModeller.C:39); the lduMatrix
    # path uses no cell shapes, so an empty PtrList is enough.
    (path.parent / "cellModels").write_text(
        "// empty cell-model list written by oracle/build_ref.py\n0()\n")


CONTROL_DICT = r"""
// minimal global controlDict written by oracle/build_ref.py (test infrastructure)
Documentation { docBrowser "none"; doxyDocDirs (); doxySourceFileExts (); }
InfoSwitches { writePrecision 6; writeJobInfo 0; writeDictionaries 0; writeOptionalEntries 0; allowSystemOperations 0; }
OptimisationSwitches
{
    fileModificationSkew 10;
    fileModificationChecking timeStampMaster;
    commsType       nonBlocking;
    floatTransfer   0;
    nProcsSimpleSum 0;
    writeNowSignal  -1;
    stopAtWriteNowSignal -1;
}
DebugSwitches { lduMatrix 0; SolverPerformance 0; GAMG 0; GAMGAgglomeration 0; }
DimensionedConstants
{
    unitSet SI;
    SICoeffs
    {
        universal { c c [0 1 -1 0 0 0 0] 2.99792e+08; G G [-1 3 -2 0 0 0 0] 6.67429e-11; h h [1 2 -1 0 0 0 0] 6.62607e-34; }
        electromagnetic { e e [0 0 1 0 0 1 0] 1.60218e-19; }
        atomic { me me [1 0 0 0 0 0 0] 9.10938e-31; mp mp [1 0 0 0 0 0 0] 1.67262e-27; }
        physicoChemical { mu mu [1 0 0 0 0 0 0] 1.66054e-27; k k [1 2 -2 -1 0 0 0] 1.38065e-23; }
        standard { Pstd Pstd [1 -1 -2 0 0 0 0] 100000; Tstd Tstd [0 0 0 1 0 0 0] 298.15; }
    }
}
DimensionSets
{
    unitSet SI;
    SICoeffs
    {
        kg kg [1 0 0 0 0 0 0] 1.0;
        m  m  [0 1 0 0 0 0 0] 1.0;
        s  s  [0 0 1 0 0 0 0] 1.0;
        K  K  [0 0 0 1 0 0 0] 1.0;
        mol mol [0 0 0 0 1 0 0] 1.0;
        A  A  [0 0 0 0 0 1 0] 1.0;
        Cd Cd [0 0 0 0 0 0 1] 1.0;
    }
    writeUnits (kg m s K mol A Cd);
}
"""

if __name__ == "__main__":
    sys.exit(main())
