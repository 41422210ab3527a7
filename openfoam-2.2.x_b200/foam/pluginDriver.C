/*---------------------------------------------------------------------------*\
  pluginDriver — a stand-in for icoFoam's pEqn.solve(): an OpenFOAM-2.2.x host
  program that (1) loads libgpuLduSolvers.so the way any application does
  (`libs` entry -> dlLibraryTable::open, Time.C:343 / dlLibraryTable.C:188-200),
  (2) builds an lduMatrix from a flat problem file, (3) solves it through
  lduMatrix::solver::New with the dictionary given on the command line, once as
  written (the reference's CPU solver) and once with the solver name replaced by
  its gpu* counterpart, and (4) prints both SolverPerformance lines and the
  largest difference of the two solutions.

  usage: pluginDriver <plugin.so> <problem.bin> "<solver dictionary>"
  exit status 0 iff both runs took the same number of iterations.
\*---------------------------------------------------------------------------*/

#include "lduMatrix.H"
#include "lduPrimitiveMesh.H"
#include "Time.H"
#include "IStringStream.H"
#include "dlLibraryTable.H"
#include "clockTime.H"

#include <cstdio>
#include <cstdlib>

using namespace Foam;

class registryLduMesh
:
    public lduPrimitiveMesh
{
    const Time& time_;
public:
    registryLduMesh
    (
        const label nCells, const labelUList& l, const labelUList& u,
        const labelListList& pa, lduInterfacePtrsList interfaces,
        const lduSchedule& ps, const Time& t
    )
    :
        lduPrimitiveMesh(nCells, l, u, pa, interfaces, ps),
        time_(t)
    {}
    virtual const objectRegistry& thisDb() const { return time_; }
};

static dictionary dictFromText(const std::string& text)
{
    IStringStream is(text);
    return dictionary(is);
}

static void readOrDie(void* p, size_t sz, size_t n, FILE* f)
{
    if (n && fread(p, sz, n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char* argv[])
{
    if (argc < 4)
    {
        fprintf(stderr, "usage: pluginDriver plugin.so problem.bin \"dict\"\n");
        return 2;
    }
    FILE* f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 2; }
    int hdr[5];
    readOrDie(hdr, sizeof(int), 5, f);
    const label nCells = hdr[1], nFaces = hdr[2];
    const bool asym = hdr[3];
    labelList l(nFaces), u(nFaces);
    readOrDie(l.begin(), sizeof(label), nFaces, f);
    readOrDie(u.begin(), sizeof(label), nFaces, f);
    scalarField diag(nCells), upper(nFaces), lower(asym ? nFaces : 0);
    readOrDie(diag.begin(), sizeof(scalar), nCells, f);
    readOrDie(upper.begin(), sizeof(scalar), nFaces, f);
    if (asym) readOrDie(lower.begin(), sizeof(scalar), nFaces, f);
    scalarField source(nCells), psi0(nCells);
    readOrDie(source.begin(), sizeof(scalar), nCells, f);
    readOrDie(psi0.begin(), sizeof(scalar), nCells, f);
    fclose(f);

    Time runTime(fileName("."), fileName("."));

    // what `libs ("libgpuLduSolvers.so");` in system/controlDict does
    dictionary libsDict(dictFromText(std::string("libs (\"") + argv[1] + "\");"));
    if (!runTime.libs().open(libsDict, "libs"))
    {
        fprintf(stderr, "could not load %s\n", argv[1]);
        return 3;
    }

    labelListList patchAddr(0);
    lduInterfacePtrsList meshInterfaces(0);
    lduSchedule schedule(0);
    registryLduMesh mesh(nCells, l, u, patchAddr, meshInterfaces, schedule, runTime);
    lduMatrix A(mesh);
    A.diag() = diag;
    A.upper() = upper;
    if (asym) A.lower() = lower;
    FieldField<Field, scalar> bouCoeffs(0), intCoeffs(0);
    lduInterfaceFieldPtrsList interfaces(0);

    dictionary cpuDict(dictFromText(argv[3]));
    dictionary gpuDict(cpuDict);
    const word cpuName(cpuDict.lookup("solver"));
    const word gpuName
    (
        cpuName == "PCG" ? "gpuPCG"
      : cpuName == "PBiCG" ? "gpuPBiCG"
      : cpuName == "GAMG" ? "gpuGAMG"
      : cpuName == "smoothSolver" ? "gpuSmoothSolver"
      : cpuName
    );
    gpuDict.add("solver", gpuName, true);

    scalarField psiCpu(psi0), psiGpu(psi0);
    clockTime t0;
    solverPerformance spCpu = lduMatrix::solver::New
    (
        "p", A, bouCoeffs, intCoeffs, interfaces, cpuDict
    )->solve(psiCpu, source);
    const double tCpu = t0.elapsedTime();
    clockTime t1;
    solverPerformance spGpu = lduMatrix::solver::New
    (
        "p", A, bouCoeffs, intCoeffs, interfaces, gpuDict
    )->solve(psiGpu, source);
    const double tGpu = t1.elapsedTime();

    Info<< "reference: "; spCpu.print(Info);
    Info<< "plug-in  : "; spGpu.print(Info);
    scalar maxDiff = 0, maxMag = 0;
    forAll(psiCpu, i)
    {
        maxDiff = max(maxDiff, mag(psiCpu[i] - psiGpu[i]));
        maxMag = max(maxMag, mag(psiCpu[i]));
    }
    printf
    (
        "RESULT iters %d %d final %.17g %.17g maxRelDiff %.3e time %.4f %.4f\n",
        int(spCpu.nIterations()), int(spGpu.nIterations()),
        spCpu.finalResidual(), spGpu.finalResidual(),
        maxDiff/(maxMag + 1e-300), tCpu, tGpu
    );
    return spCpu.nIterations() == spGpu.nIterations() ? 0 : 1;
}
