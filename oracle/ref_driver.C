/*---------------------------------------------------------------------------*\
  ref_driver — TEST INFRASTRUCTURE (not product code).

  Drives the UNMODIFIED reference lduMatrix path (oracle/_ref/libOpenFOAM.so,
  built from /root/reference by oracle/build_ref.py) on a flat-binary LDU
  problem, so that the C restatement (oracle/ldu_oracle.c) and the CUDA path can
  be pinned against the reference's own arithmetic.

  Usage:
      ref_driver <problem.bin> <out.bin> <op> [args...]
  ops:
      amul | tmul | suma | residual | H | H1   -> out = field[nCells]
      faceH                                    -> out = field[nFaces]
      precondition <name>                      -> out = M^-1 source
      smooth "<dict text>" <nSweeps>           -> out = psi after sweeps
      solve  "<dict text>"                     -> out = psi; PERF line on stdout
      agglom "<dict text>"                     -> out = int32 stream
                                                  nLevels, then per level
                                                  nFine, nCoarse, restrict[nFine]
      time_amul <reps>                         -> TIME line (seconds per Amul)
      time_solve "<dict text>"                 -> TIME + PERF lines

  Problem file (little endian):
      int32 magic(0x3155444c 'LDU1') nCells nFaces asym hasWeights
      int32 lower[nFaces] upper[nFaces]
      f64   diag[nCells] upper[nFaces] (lower[nFaces] if asym)
      f64   source[nCells] psi0[nCells] (faceWeights[nFaces] if hasWeights)

  Reference entry points exercised (all in /root/reference/src/OpenFOAM):
      lduMatrix::Amul/Tmul/sumA/residual   matrices/lduMatrix/lduMatrix/lduMatrixATmul.C:34-295
      lduMatrix::solver::New               matrices/lduMatrix/lduMatrix/lduMatrixSolver.C:40-136
      lduMatrix::preconditioner::New       matrices/lduMatrix/lduMatrix/lduMatrixPreconditioner.C:39-152
      lduMatrix::smoother::New             matrices/lduMatrix/lduMatrix/lduMatrixSmoother.C:38-156
      GAMGAgglomeration::New               matrices/lduMatrix/solvers/GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomeration.C:91-192
\*---------------------------------------------------------------------------*/

#include "lduMatrix.H"
#include "lduPrimitiveMesh.H"
#include "Time.H"
#include "IStringStream.H"
#include "clockTime.H"
#include "GAMGAgglomeration.H"
#include "pairGAMGAgglomeration.H"
#include "addToRunTimeSelectionTable.H"
#include "PCG.H"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace Foam;

// An lduPrimitiveMesh that owns an object registry (GAMG stores its
// agglomeration there: GAMGAgglomeration.C:97-102).
class registryLduMesh
:
    public lduPrimitiveMesh
{
    const Time& time_;

public:

    registryLduMesh
    (
        const label nCells,
        const labelUList& l,
        const labelUList& u,
        const labelListList& pa,
        lduInterfacePtrsList interfaces,
        const lduSchedule& ps,
        const Time& t
    )
    :
        lduPrimitiveMesh(nCells, l, u, pa, interfaces, ps),
        time_(t)
    {}

    virtual const objectRegistry& thisDb() const
    {
        return time_;
    }
};


// Face weights handed in through the problem file (stands in for
// faceAreaPairGAMGAgglomeration, which lives in libfiniteVolume and needs
// fvMesh::Sf(); same pairGAMGAgglomeration::agglomerate underneath:
// finiteVolume/fvMatrices/solvers/GAMGSymSolver/GAMGAgglomerations/
// faceAreaPairGAMGAgglomeration/faceAreaPairGAMGAgglomeration.C:48-73).
static scalarField* gFaceWeights = NULL;

namespace Foam
{
class weightedPairGAMGAgglomeration
:
    public pairGAMGAgglomeration
{
public:
    TypeName("weightedPair");

    weightedPairGAMGAgglomeration
    (
        const lduMesh& mesh,
        const dictionary& controlDict
    )
    :
        pairGAMGAgglomeration(mesh, controlDict)
    {
        agglomerate(mesh, *gFaceWeights);
    }
};

defineTypeNameAndDebug(weightedPairGAMGAgglomeration, 0);
addToRunTimeSelectionTable
(
    GAMGAgglomeration,
    weightedPairGAMGAgglomeration,
    lduMesh
);
}


static dictionary dictFromText(const std::string& text)
{
    IStringStream is(text);
    return dictionary(is);
}


static void readOrDie(void* p, size_t sz, size_t n, FILE* f)
{
    if (n && fread(p, sz, n, f) != n)
    {
        fprintf(stderr, "ref_driver: short read\n");
        exit(2);
    }
}


static void printPerf(const solverPerformance& sp)
{
    printf
    (
        "PERF %s %.17g %.17g %d %d %d\n",
        sp.solverName().c_str(),
        sp.initialResidual(),
        sp.finalResidual(),
        int(sp.nIterations()),
        int(sp.converged()),
        int(sp.singular())
    );
}


int main(int argc, char* argv[])
{
    if (argc < 4)
    {
        fprintf(stderr, "usage: ref_driver problem out op [args]\n");
        return 2;
    }
    const char* probFile = argv[1];
    const char* outFile = argv[2];
    const std::string op(argv[3]);

    FILE* f = fopen(probFile, "rb");
    if (!f) { perror(probFile); return 2; }
    int hdr[5];
    readOrDie(hdr, sizeof(int), 5, f);
    if (hdr[0] != 0x3155444c) { fprintf(stderr, "bad magic\n"); return 2; }
    const label nCells = hdr[1];
    const label nFaces = hdr[2];
    const bool asym = hdr[3];
    const bool hasWeights = hdr[4];

    labelList l(nFaces), u(nFaces);
    readOrDie(l.begin(), sizeof(label), nFaces, f);
    readOrDie(u.begin(), sizeof(label), nFaces, f);

    scalarField diag(nCells), upper(nFaces), lower(asym ? nFaces : 0);
    readOrDie(diag.begin(), sizeof(scalar), nCells, f);
    readOrDie(upper.begin(), sizeof(scalar), nFaces, f);
    if (asym) readOrDie(lower.begin(), sizeof(scalar), nFaces, f);
    scalarField source(nCells), psi(nCells);
    readOrDie(source.begin(), sizeof(scalar), nCells, f);
    readOrDie(psi.begin(), sizeof(scalar), nCells, f);
    scalarField faceWeights(hasWeights ? nFaces : 0);
    if (hasWeights)
    {
        readOrDie(faceWeights.begin(), sizeof(scalar), nFaces, f);
        gFaceWeights = &faceWeights;
    }
    fclose(f);

    Time runTime(fileName("."), fileName("."));

    labelListList patchAddr(0);
    lduInterfacePtrsList meshInterfaces(0);
    lduSchedule schedule(0);

    registryLduMesh mesh
    (
        nCells, l, u, patchAddr, meshInterfaces, schedule, runTime
    );

    lduMatrix A(mesh);
    A.diag() = diag;
    A.upper() = upper;
    if (asym) A.lower() = lower;

    FieldField<Field, scalar> bouCoeffs(0);
    FieldField<Field, scalar> intCoeffs(0);
    lduInterfaceFieldPtrsList interfaces(0);

    scalarField out(nCells, 0.0);
    std::vector<int> outInts;
    bool intsOut = false;

    if (op == "amul")
    {
        A.Amul(out, psi, bouCoeffs, interfaces, 0);
    }
    else if (op == "tmul")
    {
        A.Tmul(out, psi, intCoeffs, interfaces, 0);
    }
    else if (op == "suma")
    {
        A.sumA(out, bouCoeffs, interfaces);
    }
    else if (op == "residual")
    {
        A.residual(out, psi, source, bouCoeffs, interfaces, 0);
    }
    else if (op == "H")        // lduMatrixTemplates.C:33-65
    {
        out = A.H(psi)();
    }
    else if (op == "H1")       // lduMatrixATmul.C:298-327
    {
        out = A.H1()();
    }
    else if (op == "faceH")    // lduMatrixTemplates.C:79-113: one value per face
    {
        out = A.faceH(psi)();
    }
    else if (op == "precondition" || op == "preconditionT")
    {
        // a solver object is needed to construct a preconditioner
        dictionary d
        (
            dictFromText
            (
                std::string("solver PCG; preconditioner ")
              + argv[4] + "; tolerance 0; relTol 0;"
            )
        );
        PCG dummy("p", A, bouCoeffs, intCoeffs, interfaces, d);
        autoPtr<lduMatrix::preconditioner> pre =
            lduMatrix::preconditioner::New(dummy, d);
        if (op == "precondition")
        {
            pre->precondition(out, source, 0);
        }
        else
        {
            pre->preconditionT(out, source, 0);
        }
    }
    else if (op == "smooth")
    {
        dictionary d(dictFromText(argv[4]));
        const label nSweeps = atoi(argv[5]);
        autoPtr<lduMatrix::smoother> sm = lduMatrix::smoother::New
        (
            "p", A, bouCoeffs, intCoeffs, interfaces, d
        );
        sm->smooth(psi, source, 0, nSweeps);
        out = psi;
    }
    else if (op == "solve" || op == "time_solve")
    {
        dictionary d(dictFromText(argv[4]));
        clockTime timer;
        solverPerformance sp = lduMatrix::solver::New
        (
            "p", A, bouCoeffs, intCoeffs, interfaces, d
        )->solve(psi, source);
        const double t = timer.elapsedTime();
        printPerf(sp);
        if (op == "time_solve")
        {
            printf("TIME %.9g\n", t);
        }
        out = psi;
    }
    else if (op == "time_iters")
    {
        // steady-state cost of one solver iteration: two fixed-iteration solves
        // (tolerance 0; maxIter a and b run a+1 and b+1 iterations, PCG.C:174-178)
        // timed separately, the difference removes construction + prologue
        const int a = atoi(argv[5]);
        const int b = atoi(argv[6]);
        double t[2];
        int its[2];
        for (int pass = 0; pass < 2; pass++)
        {
            dictionary d(dictFromText(argv[4]));
            d.add("maxIter", pass ? b : a, true);
            d.add("tolerance", 0.0, true);
            d.add("relTol", 0.0, true);
            scalarField x(psi);
            clockTime timer;
            solverPerformance sp = lduMatrix::solver::New
            (
                "p", A, bouCoeffs, intCoeffs, interfaces, d
            )->solve(x, source);
            t[pass] = timer.elapsedTime();
            its[pass] = sp.nIterations();
            printPerf(sp);
            if (pass) out = x;
        }
        printf("ITERS %d %.9g %d %.9g\n", its[0], t[0], its[1], t[1]);
    }
    else if (op == "time_amul")
    {
        const int reps = atoi(argv[4]);
        A.Amul(out, psi, bouCoeffs, interfaces, 0);
        clockTime timer;
        for (int r = 0; r < reps; r++)
        {
            A.Amul(out, psi, bouCoeffs, interfaces, 0);
        }
        printf("TIME %.9g\n", timer.elapsedTime()/reps);
    }
    else if (op == "agglom")
    {
        dictionary d(dictFromText(argv[4]));
        const GAMGAgglomeration& agg = GAMGAgglomeration::New(A, d);
        intsOut = true;
        outInts.push_back(agg.size());
        for (label lev = 0; lev < agg.size(); lev++)
        {
            const labelField& r = agg.restrictAddressing(lev);
            outInts.push_back(r.size());
            outInts.push_back(agg.meshLevel(lev + 1).lduAddr().size());
            for (label i = 0; i < r.size(); i++) outInts.push_back(r[i]);
        }
    }
    else
    {
        fprintf(stderr, "unknown op %s\n", op.c_str());
        return 2;
    }

    FILE* g = fopen(outFile, "wb");
    if (!g) { perror(outFile); return 2; }
    if (intsOut)
    {
        fwrite(&outInts[0], sizeof(int), outInts.size(), g);
    }
    else
    {
        fwrite(out.begin(), sizeof(scalar), out.size(), g);
    }
    fclose(g);
    return 0;
}
