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
      solve2 "<dict text>"                     -> two solves on the same mesh: as given, then with
                                                  upper/lower[f] *= 1 + 0.25*((7 f) % 5) and
                                                  diag[c] *= 1.5 + 0.25*(c % 3), both from the file's
                                                  psi; out = second psi; two PERF lines
                                                  (cacheAgglomeration: GAMGSolver.C:70,144-154)
      agglom "<dict text>"                     -> out = int32 stream
                                                  nLevels, then per level
                                                  nFine, nCoarse, restrict[nFine]
      bandCompression                          -> out = int32 newOrder[nCells] of Foam::bandCompression on
                                                  the cell-cell addressing of the internal faces
      time_amul <reps>                         -> TIME line (seconds per Amul)
      time_solve "<dict text>"                 -> TIME + PERF lines
  Environment: LDU_REF_LIBS=<lib.so> is opened through the Time's dlLibraryTable before the
  operation (the `libs` entry of a case's controlDict), so a dictionary may select a plug-in.

  Problem file (little endian):
      int32 magic(0x3155444c 'LDU1') nCells nFaces asym flags(1: face weights, 2: diagonal matrix,
                                                              upper() never touched)
      int32 lower[nFaces] upper[nFaces]
      f64   diag[nCells] upper[nFaces] (lower[nFaces] if asym)
      f64   source[nCells] psi0[nCells] (faceWeights[nFaces] if hasWeights)

  Parallel mode (binary ref_driver_par, linked against oracle/_ref/libOpenFOAM_par.so =
  the same objects with oracle/pstream_shm/ in place of src/Pstream/dummy): when
  LDU_PSTREAM_SIZE is set, one process per mesh region is started; "%d" in the
  problem / output paths is the rank, and the problem file (magic 'LDU2') carries
  after the header one extra int32 nInterfaces and, at the end,
      per interface: int32 nbrRank, int32 size, int32 faceCells[size],
                     f64 bouCoeffs[size], f64 intCoeffs[size]
  nbrRank >= 0: a processor interface (classes below) doing what processorFvPatch /
  processorFvPatchField<scalar> do in libfiniteVolume
  (finiteVolume/fields/fvPatchFields/constraint/processor/processorFvPatchScalarField.C:33-116).
  nbrRank < 0: one half of a cyclic pair inside this region whose other half is interface
  -1 - nbrRank (cyclicFvPatch / cyclicFvPatchField<scalar>::updateInterfaceMatrix,
  finiteVolume/fields/fvPatchFields/constraint/cyclic/cyclicFvPatchField.C:174-197).
  The serial binary reads 'LDU2' files too (cyclic interfaces only).

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
#include "GAMGInterface.H"
#include "pairGAMGAgglomeration.H"
#include "addToRunTimeSelectionTable.H"
#include "PCG.H"
#include "processorLduInterface.H"
#include "processorLduInterfaceField.H"
#include "cyclicLduInterface.H"
#include "cyclicLduInterfaceField.H"
#include "IPstream.H"
#include "OPstream.H"
#include "dlLibraryTable.H"
#include "bandCompression.H"

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



// Processor interface of a flat LDU region: faceCells of the cut faces, the
// neighbouring rank, raw transfers through processorLduInterface::send/receive.
// Type name "processor" so that GAMGInterface::New / GAMGInterfaceField::New
// select processorGAMGInterface(Field) on the coarse levels
// (GAMGInterface/GAMGInterfaceNew.C:32-62, GAMGInterfaceField/GAMGInterfaceFieldNew.C:31-56).
namespace Foam
{
class flatProcessorInterface
:
    public lduInterface,
    public processorLduInterface
{
    labelList faceCells_;
    int nbr_;
    tensorField noTransform_;

public:
    TypeName("processor");

    flatProcessorInterface(const labelList& fc, const int nbr)
    :
        faceCells_(fc),
        nbr_(nbr),
        noTransform_(0)
    {}

    virtual const labelUList& faceCells() const { return faceCells_; }
    virtual int myProcNo() const { return Pstream::myProcNo(); }
    virtual int neighbProcNo() const { return nbr_; }
    virtual const tensorField& forwardT() const { return noTransform_; }
    virtual int tag() const { return Pstream::msgType(); }

    virtual tmp<labelField> interfaceInternalField(const labelUList& iF) const
    {
        tmp<labelField> tf(new labelField(faceCells_.size()));
        labelField& pf = tf();
        forAll(pf, i) pf[i] = iF[faceCells_[i]];
        return tf;
    }

    virtual void initInternalFieldTransfer
    (
        const Pstream::commsTypes commsType,
        const labelUList& iF
    ) const
    {
        send(commsType, interfaceInternalField(iF)());
    }

    virtual tmp<labelField> internalFieldTransfer
    (
        const Pstream::commsTypes commsType,
        const labelUList&
    ) const
    {
        return receive<label>(commsType, faceCells_.size());
    }
};

defineTypeNameAndDebug(flatProcessorInterface, 0);


class flatProcessorInterfaceField
:
    public lduInterfaceField,
    public processorLduInterfaceField
{
    const flatProcessorInterface& patch_;
    mutable scalarField sendBuf_;
    mutable scalarField recvBuf_;
    mutable label recvRequest_;

public:
    TypeName("processor");

    flatProcessorInterfaceField(const flatProcessorInterface& p)
    :
        lduInterfaceField(p),
        patch_(p),
        recvRequest_(-1)
    {}

    virtual int myProcNo() const { return patch_.myProcNo(); }
    virtual int neighbProcNo() const { return patch_.neighbProcNo(); }
    virtual bool doTransform() const { return false; }
    virtual const tensorField& forwardT() const { return patch_.forwardT(); }
    virtual int rank() const { return 0; }

    // psi next to the cut goes to the neighbour ...
    virtual void initInterfaceMatrixUpdate
    (
        scalarField&,
        const scalarField& psiInternal,
        const scalarField&,
        const direction,
        const Pstream::commsTypes commsType
    ) const
    {
        const labelUList& fc = patch_.faceCells();
        sendBuf_.setSize(fc.size());
        forAll(fc, i) sendBuf_[i] = psiInternal[fc[i]];

        if (commsType == Pstream::nonBlocking)
        {
            recvBuf_.setSize(fc.size());
            recvRequest_ = UPstream::nRequests();
            IPstream::read
            (
                commsType, patch_.neighbProcNo(),
                reinterpret_cast<char*>(recvBuf_.begin()), recvBuf_.byteSize(), patch_.tag()
            );
            OPstream::write
            (
                commsType, patch_.neighbProcNo(),
                reinterpret_cast<const char*>(sendBuf_.begin()), sendBuf_.byteSize(), patch_.tag()
            );
        }
        else
        {
            patch_.send(commsType, sendBuf_);
        }
        const_cast<flatProcessorInterfaceField&>(*this).updatedMatrix() = false;
    }

    // ... and the neighbour's values come back: result[faceCell] -= coeff*psiNbr
    virtual void updateInterfaceMatrix
    (
        scalarField& result,
        const scalarField&,
        const scalarField& coeffs,
        const direction,
        const Pstream::commsTypes commsType
    ) const
    {
        if (updatedMatrix()) return;
        const labelUList& fc = patch_.faceCells();
        if (commsType == Pstream::nonBlocking)
        {
            if (recvRequest_ >= 0 && recvRequest_ < UPstream::nRequests())
            {
                UPstream::waitRequest(recvRequest_);
            }
            recvRequest_ = -1;
        }
        else
        {
            recvBuf_.setSize(fc.size());
            patch_.receive(commsType, recvBuf_);
        }
        forAll(fc, i) result[fc[i]] -= coeffs[i]*recvBuf_[i];
        const_cast<flatProcessorInterfaceField&>(*this).updatedMatrix() = true;
    }
};

defineTypeNameAndDebug(flatProcessorInterfaceField, 0);
}


// One half of a cyclic (periodic) pair of a flat LDU region, no transformation.
// Type name "cyclic": the coarse levels get cyclicGAMGInterface(Field)
// (cyclicGAMGInterface.C:47-119, cyclicGAMGInterfaceField.C:66-89).
namespace Foam
{
class flatCyclicInterface
:
    public lduInterface,
    public cyclicLduInterface
{
    labelList faceCells_;
    label index_;
    label nbrIndex_;
    const lduInterfacePtrsList& all_;
    tensorField noTransform_;

public:
    TypeName("cyclic");

    flatCyclicInterface
    (
        const labelList& fc,
        const label index,
        const label nbrIndex,
        const lduInterfacePtrsList& all
    )
    :
        faceCells_(fc),
        index_(index),
        nbrIndex_(nbrIndex),
        all_(all),
        noTransform_(0)
    {}

    virtual const labelUList& faceCells() const { return faceCells_; }
    virtual label neighbPatchID() const { return nbrIndex_; }
    virtual bool owner() const { return index_ < nbrIndex_; }
    virtual const cyclicLduInterface& neighbPatch() const
    {
        return refCast<const cyclicLduInterface>(all_[nbrIndex_]);
    }
    virtual const tensorField& forwardT() const { return noTransform_; }
    virtual const tensorField& reverseT() const { return noTransform_; }

    virtual tmp<labelField> interfaceInternalField(const labelUList& iF) const
    {
        tmp<labelField> tf(new labelField(faceCells_.size()));
        labelField& pf = tf();
        forAll(pf, i) pf[i] = iF[faceCells_[i]];
        return tf;
    }

    // the other half's cells, same region: no transfer needed
    virtual tmp<labelField> internalFieldTransfer
    (
        const Pstream::commsTypes,
        const labelUList& iF
    ) const
    {
        return all_[nbrIndex_].interfaceInternalField(iF);
    }
};

defineTypeNameAndDebug(flatCyclicInterface, 0);


class flatCyclicInterfaceField
:
    public lduInterfaceField,
    public cyclicLduInterfaceField
{
    const flatCyclicInterface& patch_;

public:
    TypeName("cyclic");

    flatCyclicInterfaceField(const flatCyclicInterface& p)
    :
        lduInterfaceField(p),
        patch_(p)
    {}

    virtual bool doTransform() const { return false; }
    virtual const tensorField& forwardT() const { return patch_.forwardT(); }
    virtual const tensorField& reverseT() const { return patch_.reverseT(); }
    virtual int rank() const { return 0; }

    // result[faceCell] -= coeff*psi[cell behind the other half]
    virtual void updateInterfaceMatrix
    (
        scalarField& result,
        const scalarField& psiInternal,
        const scalarField& coeffs,
        const direction,
        const Pstream::commsTypes
    ) const
    {
        const labelUList& nbrCells =
            dynamic_cast<const lduInterface&>(patch_.neighbPatch()).faceCells();
        const labelUList& fc = patch_.faceCells();
        forAll(fc, i) result[fc[i]] -= coeffs[i]*psiInternal[nbrCells[i]];
    }
};

defineTypeNameAndDebug(flatCyclicInterfaceField, 0);
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
    const bool parallel = getenv("LDU_PSTREAM_SIZE") != NULL;
    if (parallel)
    {
        UPstream::init(argc, argv);     // oracle/pstream_shm/shmPstream.C
    }
    char probFile[4096], outFile[4096];
    snprintf(probFile, sizeof(probFile), argv[1], int(Pstream::myProcNo()));
    snprintf(outFile, sizeof(outFile), argv[2], int(Pstream::myProcNo()));
    const std::string op(argv[3]);

    FILE* f = fopen(probFile, "rb");
    if (!f) { perror(probFile); return 2; }
    int hdr[5];
    readOrDie(hdr, sizeof(int), 5, f);
    int nInterfaces = 0;
    if (hdr[0] == 0x3255444c)           // 'LDU2': processor interfaces follow the fields
    {
        readOrDie(&nInterfaces, sizeof(int), 1, f);
    }
    else if (hdr[0] != 0x3155444c) { fprintf(stderr, "bad magic\n"); return 2; }
    const label nCells = hdr[1];
    const label nFaces = hdr[2];
    const bool asym = hdr[3];
    const bool hasWeights = hdr[4] & 1;
    const bool diagonalOnly = hdr[4] & 2;    // leave upper/lower unallocated: lduMatrix::diagonal()

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
    PtrList<lduInterface> patches(nInterfaces);
    PtrList<lduInterfaceField> patchFields(nInterfaces);
    labelListList patchAddr(nInterfaces);
    lduInterfacePtrsList meshInterfaces(nInterfaces);
    lduSchedule schedule(2*nInterfaces);
    FieldField<Field, scalar> bouCoeffs(nInterfaces);
    FieldField<Field, scalar> intCoeffs(nInterfaces);
    lduInterfaceFieldPtrsList interfaces(nInterfaces);
    for (int i = 0; i < nInterfaces; i++)
    {
        int ih[2];
        readOrDie(ih, sizeof(int), 2, f);
        patchAddr[i].setSize(ih[1]);
        readOrDie(patchAddr[i].begin(), sizeof(label), ih[1], f);
        bouCoeffs.set(i, new scalarField(ih[1]));
        intCoeffs.set(i, new scalarField(ih[1]));
        readOrDie(bouCoeffs[i].begin(), sizeof(scalar), ih[1], f);
        readOrDie(intCoeffs[i].begin(), sizeof(scalar), ih[1], f);
        if (ih[0] >= 0)
        {
            flatProcessorInterface* pp = new flatProcessorInterface(patchAddr[i], ih[0]);
            patches.set(i, pp);
            patchFields.set(i, new flatProcessorInterfaceField(*pp));
        }
        else
        {
            flatCyclicInterface* cp =
                new flatCyclicInterface(patchAddr[i], i, -1 - ih[0], meshInterfaces);
            patches.set(i, cp);
            patchFields.set(i, new flatCyclicInterfaceField(*cp));
        }
        meshInterfaces.set(i, &patches[i]);
        interfaces.set(i, &patchFields[i]);
        schedule[2*i].patch = i;
        schedule[2*i].init = true;
        schedule[2*i + 1].patch = i;
        schedule[2*i + 1].init = false;
    }
    fclose(f);

    Time runTime(fileName("."), fileName("."));

    // LDU_REF_LIBS=<lib.so>: what `libs ("lib.so");` in system/controlDict does (Time.C:343 ->
    // dlLibraryTable::open): the solver dictionary can then name run-time selected plug-ins
    if (getenv("LDU_REF_LIBS"))
    {
        dictionary libsDict(dictFromText(std::string("libs (\"") + getenv("LDU_REF_LIBS") + "\");"));
        if (!runTime.libs().open(libsDict, "libs"))
        {
            fprintf(stderr, "ref_driver: could not load %s\n", getenv("LDU_REF_LIBS"));
            return 3;
        }
    }

    registryLduMesh mesh
    (
        nCells, l, u, patchAddr, meshInterfaces, schedule, runTime
    );

    lduMatrix A(mesh);
    A.diag() = diag;
    if (!diagonalOnly) A.upper() = upper;
    if (asym) A.lower() = lower;

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
    else if (op == "solve2")
    {
        dictionary d(dictFromText(argv[4]));
        const scalarField psi0(psi);
        for (int pass = 0; pass < 2; pass++)
        {
            if (pass)
            {
                // exact multipliers (quarters), the same in oracle/oracle.py::second_coeffs
                forAll(A.upper(), f)
                {
                    const scalar m = 1 + 0.25*((7*f) % 5);
                    A.upper()[f] *= m;
                    if (asym) A.lower()[f] *= m;
                }
                forAll(A.diag(), c) A.diag()[c] *= 1.5 + 0.25*(c % 3);
            }
            psi = psi0;
            solverPerformance sp = lduMatrix::solver::New
            (
                "p", A, bouCoeffs, intCoeffs, interfaces, d
            )->solve(psi, source);
            printPerf(sp);
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
        // optional 7th argument: the tolerance of the fixed-iteration solves.  0 (default) for the Krylov
        // solvers; GAMG hands its tolerance to the coarsest-level solver (GAMGSolverSolve.C:430-487), which
        // would then run its 1000 iterations per cycle: a tolerance no cycle count reaches is given instead
        const double fixedTol = argc > 7 ? atof(argv[7]) : 0.0;
        double t[2];
        int its[2];
        for (int pass = 0; pass < 2; pass++)
        {
            dictionary d(dictFromText(argv[4]));
            d.add("maxIter", pass ? b : a, true);
            d.add("tolerance", fixedTol, true);
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
    else if (op == "bandCompression")
    {
        // cell-cell addressing as decompositionMethod::calcCellCells builds it: face order
        labelList nNbrs(nCells, 0);
        forAll(l, f) { nNbrs[l[f]]++; nNbrs[u[f]]++; }
        labelListList cellCells(nCells);
        forAll(cellCells, c) { cellCells[c].setSize(nNbrs[c]); nNbrs[c] = 0; }
        forAll(l, f)
        {
            cellCells[l[f]][nNbrs[l[f]]++] = u[f];
            cellCells[u[f]][nNbrs[u[f]]++] = l[f];
        }
        const labelList order(bandCompression(cellCells));
        intsOut = true;
        forAll(order, i) outInts.push_back(order[i]);
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
    else if (op == "agglom_full")
    {
        // everything a host needs to hand the hierarchy to another solver (include/ldu_b200.h,
        // ldu_gamg_set_level): per level restrictAddressing, faceRestrictAddressing, the coarse
        // lduAddressing and, per coupled patch, the GAMGInterface's faceCells + faceRestrictAddressing
        dictionary d(dictFromText(argv[4]));
        const GAMGAgglomeration& agg = GAMGAgglomeration::New(A, d);
        intsOut = true;
        outInts.push_back(agg.size());
        for (label lev = 0; lev < agg.size(); lev++)
        {
            const labelField& r = agg.restrictAddressing(lev);
            const labelList& fr = agg.faceRestrictAddressing(lev);
            const lduAddressing& ca = agg.meshLevel(lev + 1).lduAddr();
            const lduInterfacePtrsList& ifs = agg.interfaceLevel(lev + 1);
            label nIf = 0;
            forAll(ifs, i) if (ifs.set(i)) nIf++;
            outInts.push_back(r.size());
            outInts.push_back(ca.size());
            outInts.push_back(fr.size());
            outInts.push_back(ca.lowerAddr().size());
            outInts.push_back(nIf);
            forAll(r, i) outInts.push_back(r[i]);
            forAll(fr, i) outInts.push_back(fr[i]);
            forAll(ca.lowerAddr(), i) outInts.push_back(ca.lowerAddr()[i]);
            forAll(ca.upperAddr(), i) outInts.push_back(ca.upperAddr()[i]);
            forAll(ifs, i)
            {
                if (!ifs.set(i)) continue;
                const GAMGInterface& gi = refCast<const GAMGInterface>(ifs[i]);
                outInts.push_back(gi.faceCells().size());
                outInts.push_back(gi.faceRestrictAddressing().size());
                forAll(gi.faceCells(), k) outInts.push_back(gi.faceCells()[k]);
                forAll(gi.faceRestrictAddressing(), k) outInts.push_back(gi.faceRestrictAddressing()[k]);
            }
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
