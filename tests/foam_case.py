"""OpenFOAM case directories for the compiled reference applications in oracle/_ref (blockMesh, icoFoam):
test infrastructure for BASELINE config 1 -- the real application picking the CUDA solvers up from
`libs ("libgpuLduSolvers.so");` in system/controlDict.  The lid-driven cavity of the icoFoam tutorial
(tutorials/incompressible/icoFoam/cavity: 20x20x1 cells, nu = 0.01, lid speed 1, deltaT 0.005, PISO with 2
correctors, p: PCG + DIC to 1e-6, U: PBiCG + DILU to 1e-5) written with any resolution and solver dictionaries."""
import os
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
PLUGIN = ROOT / "openfoam-2.2.x_b200" / "foam" / "libgpuLduSolvers.so"

HEADER = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    object      {obj};
}}
"""


def available():
    return all((REF / f).exists() for f in ("icoFoam", "blockMesh", "libfiniteVolume.so", "etc/cellModels"))


def dict_text(d, indent="    "):
    out = []
    for k, v in d.items():
        if isinstance(v, dict):
            out.append(f"{indent}{k}\n{indent}{{\n{dict_text(v, indent + '    ')}{indent}}}\n")
        elif isinstance(v, bool):
            out.append(f"{indent}{k} {'on' if v else 'off'};\n")
        else:
            out.append(f"{indent}{k} {v};\n")
    return "".join(out)


def write_cavity(case, nx=20, ny=20, nz=1, end_time=0.05, delta_t=0.005, p=None, U=None, libs=(),
                 write_precision=12, write_interval=100000):
    """nz == 1: the tutorial's 2-D cavity (front and back `empty`); nz > 1: a 3-D lid-driven box with walls."""
    case = Path(case)
    p = p or dict(solver="PCG", preconditioner="DIC", tolerance=1e-06, relTol=0)
    U = U or dict(solver="PBiCG", preconditioner="DILU", tolerance=1e-05, relTol=0)
    for d in ("system", "constant/polyMesh", "0"):
        (case / d).mkdir(parents=True, exist_ok=True)
    three_d = nz > 1
    depth = 1.0 if three_d else 0.1
    fb = "wall" if three_d else "empty"
    (case / "constant/polyMesh/blockMeshDict").write_text(HEADER.format(cls="dictionary", obj="blockMeshDict") + f"""
convertToMeters 0.1;
vertices ( (0 0 0) (1 0 0) (1 1 0) (0 1 0) (0 0 {depth}) (1 0 {depth}) (1 1 {depth}) (0 1 {depth}) );
blocks ( hex (0 1 2 3 4 5 6 7) ({nx} {ny} {nz}) simpleGrading (1 1 1) );
edges ( );
boundary
(
    movingWall {{ type wall; faces ( (3 7 6 2) ); }}
    fixedWalls {{ type wall; faces ( (0 4 7 3) (2 6 5 1) (1 5 4 0) ); }}
    frontAndBack {{ type {fb}; faces ( (0 3 2 1) (4 5 6 7) ); }}
);
mergePatchPairs ( );
""")
    (case / "constant/transportProperties").write_text(HEADER.format(cls="dictionary", obj="transportProperties")
                                                       + "\nnu nu [ 0 2 -1 0 0 0 0 ] 0.01;\n")
    fbU = "type fixedValue; value uniform (0 0 0);" if three_d else "type empty;"
    fbp = "type zeroGradient;" if three_d else "type empty;"
    (case / "0/U").write_text(HEADER.format(cls="volVectorField", obj="U") + f"""
dimensions [0 1 -1 0 0 0 0];
internalField uniform (0 0 0);
boundaryField
{{
    movingWall {{ type fixedValue; value uniform (1 0 0); }}
    fixedWalls {{ type fixedValue; value uniform (0 0 0); }}
    frontAndBack {{ {fbU} }}
}}
""")
    (case / "0/p").write_text(HEADER.format(cls="volScalarField", obj="p") + f"""
dimensions [0 2 -2 0 0 0 0];
internalField uniform 0;
boundaryField
{{
    movingWall {{ type zeroGradient; }}
    fixedWalls {{ type zeroGradient; }}
    frontAndBack {{ {fbp} }}
}}
""")
    libs_line = ("libs ( " + " ".join(f'"{x}"' for x in libs) + " );\n") if libs else ""
    (case / "system/controlDict").write_text(HEADER.format(cls="dictionary", obj="controlDict") + f"""
application icoFoam;
{libs_line}startFrom startTime;
startTime 0;
stopAt endTime;
endTime {end_time};
deltaT {delta_t};
writeControl timeStep;
writeInterval {write_interval};
purgeWrite 0;
writeFormat ascii;
writePrecision {write_precision};
writeCompression off;
timeFormat general;
timePrecision 6;
runTimeModifiable false;
""")
    (case / "system/fvSchemes").write_text(HEADER.format(cls="dictionary", obj="fvSchemes") + """
ddtSchemes { default Euler; }
gradSchemes { default Gauss linear; grad(p) Gauss linear; }
divSchemes { default none; div(phi,U) Gauss linear; }
laplacianSchemes { default none; laplacian(nu,U) Gauss linear orthogonal; laplacian((1|A(U)),p) Gauss linear orthogonal; }
interpolationSchemes { default linear; interpolate(HbyA) linear; }
snGradSchemes { default orthogonal; }
fluxRequired { default no; p ; }
""")
    (case / "system/fvSolution").write_text(HEADER.format(cls="dictionary", obj="fvSolution") + "\nsolvers\n{\n"
                                            + dict_text(dict(p=p, U=U))
                                            + "}\nPISO { nCorrectors 2; nNonOrthogonalCorrectors 0; pRefCell 0; pRefValue 0; }\n")
    return case


_PROJECT = None


def project_dir():
    """WM_PROJECT_DIR for the applications: oracle/_ref/etc with `SolverPerformance 1`, the value OpenFOAM's own
    etc/controlDict ships with (etc/controlDict:269) -- the application then prints the `Solving for` lines; the
    global switch cannot be overridden per case (SolverPerformance<Type>::debug is not a registered switch)."""
    global _PROJECT
    if _PROJECT is None:
        import tempfile
        _PROJECT = Path(tempfile.mkdtemp(prefix="ldu_foam_project_"))
        (_PROJECT / "etc").mkdir()
        text = (REF / "etc" / "controlDict").read_text().replace("SolverPerformance 0;", "SolverPerformance 1;")
        (_PROJECT / "etc" / "controlDict").write_text(text)
        (_PROJECT / "etc" / "cellModels").write_text((REF / "etc" / "cellModels").read_text())
    return _PROJECT


def run(app, case, env=None, timeout=3600):
    e = dict(os.environ, WM_PROJECT_DIR=str(project_dir()), **(env or {}))
    ld = e.get("LD_LIBRARY_PATH", "")
    e["LD_LIBRARY_PATH"] = str(REF) + (":" + ld if ld else "")
    r = subprocess.run([str(REF / app), "-case", str(case)], env=e, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"{app} failed ({r.returncode}):\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}")
    return r.stdout


def solver_lines(log):
    """the `<solver>:  Solving for <field>, Initial residual = ..., Final residual = ..., No Iterations n` lines"""
    return [x.strip() for x in log.splitlines() if "Solving for" in x]


def field_text(case, time_name, name):
    """the internalField block of a written field file (header stripped)"""
    t = (Path(case) / time_name / name).read_text()
    return t[t.index("internalField"):]
