/*---------------------------------------------------------------------------*\
  gpuLduSolvers — see gpuLduSolvers.H.  Thin shim: OpenFOAM objects in, POD out.
\*---------------------------------------------------------------------------*/

#include "gpuLduSolvers.H"
#include "addToRunTimeSelectionTable.H"
#include "Switch.H"
#include "processorLduInterface.H"
#include "cyclicLduInterface.H"
#include "cyclicLduInterfaceField.H"
#include "processorLduInterfaceField.H"
#include "PstreamReduceOps.H"
#include "UIPstream.H"
#include "UOPstream.H"
#include "GAMGAgglomeration.H"
#include "GAMGInterface.H"

#include "../../include/ldu_b200.h"

#include <cstdlib>
#include <map>

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
    defineTypeNameAndDebug(gpuPCG, 0);
    defineTypeNameAndDebug(gpuPBiCG, 0);
    defineTypeNameAndDebug(gpuICCG, 0);
    defineTypeNameAndDebug(gpuBICCG, 0);
    defineTypeNameAndDebug(gpuSmoothSolver, 0);
    defineTypeNameAndDebug(gpuGAMG, 0);

    // symmetric matrices: PCG, GAMG, smoothSolver (PCG.C:34, GAMGSolver.C:34, smoothSolver.C:34)
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuPCG>
        addgpuPCGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuICCG>
        addgpuICCGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuGAMG>
        addgpuGAMGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuSmoothSolver>
        addgpuSmoothSolverSymMatrixConstructorToTable_;

    // asymmetric matrices: PBiCG, GAMG, smoothSolver (PBiCG.C:34, GAMGSolver.C:37, smoothSolver.C:37)
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuPBiCG>
        addgpuPBiCGAsymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuBICCG>
        addgpuBICCGAsymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuGAMG>
        addgpuGAMGAsymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuSmoothSolver>
        addgpuSmoothSolverAsymMatrixConstructorToTable_;

    // smoothers: which table each reference smoother is in (GaussSeidelSmoother.C:34-39, DICSmoother.C:34-36,
    // DILUSmoother.C:34-36, ...)
    #define registerGpuSmoother(Class, sym, asym)                                                  \
        defineTypeNameAndDebug(Class, 0);                                                         \
        struct Class##Registrar                                                                   \
        {                                                                                         \
            Class##Registrar()                                                                    \
            {                                                                                     \
                if (sym)                                                                          \
                {                                                                                 \
                    static lduMatrix::smoother::addsymMatrixConstructorToTable<Class> s_;         \
                }                                                                                 \
                if (asym)                                                                         \
                {                                                                                 \
                    static lduMatrix::smoother::addasymMatrixConstructorToTable<Class> a_;        \
                }                                                                                 \
            }                                                                                     \
        } Class##Registrar_
    registerGpuSmoother(gpuGaussSeidelSmoother, true, true);
    registerGpuSmoother(gpuSymGaussSeidelSmoother, true, true);
    registerGpuSmoother(gpuNonBlockingGaussSeidelSmoother, true, true);
    registerGpuSmoother(gpuMultiColourGaussSeidelSmoother, true, true);
    registerGpuSmoother(gpuDICSmoother, true, false);
    registerGpuSmoother(gpuFDICSmoother, true, false);
    registerGpuSmoother(gpuDICGaussSeidelSmoother, true, false);
    registerGpuSmoother(gpuDILUSmoother, false, true);
    registerGpuSmoother(gpuDILUGaussSeidelSmoother, false, true);

    #define registerGpuPreconditioner(Class, sym, asym)                                            \
        defineTypeNameAndDebug(Class, 0);                                                         \
        struct Class##Registrar                                                                   \
        {                                                                                         \
            Class##Registrar()                                                                    \
            {                                                                                     \
                if (sym)                                                                          \
                {                                                                                 \
                    static lduMatrix::preconditioner::addsymMatrixConstructorToTable<Class> s_;   \
                }                                                                                 \
                if (asym)                                                                         \
                {                                                                                 \
                    static lduMatrix::preconditioner::addasymMatrixConstructorToTable<Class> a_;  \
                }                                                                                 \
            }                                                                                     \
        } Class##Registrar_
    registerGpuPreconditioner(gpuDiagonalPreconditioner, true, true);
    registerGpuPreconditioner(gpuDICPreconditioner, true, false);
    registerGpuPreconditioner(gpuFDICPreconditioner, true, false);
    registerGpuPreconditioner(gpuDILUPreconditioner, false, true);

    //- With LDU_GPU_OVERRIDE=1 the reference's own names are re-pointed at the
    //  GPU classes (HashTable::set replaces, insert would refuse a duplicate:
    //  runTimeSelectionTables.H:85-91), so unmodified fvSolution files use them.
    class gpuLduSolverOverride
    {
    public:
        gpuLduSolverOverride()
        {
            const char* e = ::getenv("LDU_GPU_OVERRIDE");
            if (!e || e[0] == '0') return;
            typedef lduMatrix::solver S;
            S::symMatrixConstructorTablePtr_->set
            (
                "PCG", S::addsymMatrixConstructorToTable<gpuPCG>::New
            );
            S::symMatrixConstructorTablePtr_->set
            (
                "ICCG", S::addsymMatrixConstructorToTable<gpuICCG>::New
            );
            S::symMatrixConstructorTablePtr_->set
            (
                "GAMG", S::addsymMatrixConstructorToTable<gpuGAMG>::New
            );
            S::symMatrixConstructorTablePtr_->set
            (
                "smoothSolver",
                S::addsymMatrixConstructorToTable<gpuSmoothSolver>::New
            );
            S::asymMatrixConstructorTablePtr_->set
            (
                "PBiCG", S::addasymMatrixConstructorToTable<gpuPBiCG>::New
            );
            S::asymMatrixConstructorTablePtr_->set
            (
                "BICCG", S::addasymMatrixConstructorToTable<gpuBICCG>::New
            );
            S::asymMatrixConstructorTablePtr_->set
            (
                "GAMG", S::addasymMatrixConstructorToTable<gpuGAMG>::New
            );
            S::asymMatrixConstructorTablePtr_->set
            (
                "smoothSolver",
                S::addasymMatrixConstructorToTable<gpuSmoothSolver>::New
            );
        }
    };
    gpuLduSolverOverride gpuLduSolverOverride_;
}


// * * * * * * * * * * * * * * * Local Functions * * * * * * * * * * * * * * //

namespace
{

using namespace Foam;

// CUDA/ABI failures follow the reference's error convention
void check(const int rc, const char* what)
{
    if (rc != LDU_OK)
    {
        FatalErrorIn("gpuLduSolver")
            << what << " failed (code " << rc << "): "
            << ldu_last_error() << abort(FatalError);
    }
}

// One context per process; in a parallel run rank r takes device r % nDevices
// (LDU_DEVICE overrides), the usual one-rank-per-GPU binding.
ldu_context* context()
{
    static ldu_context* ctx = NULL;
    if (!ctx)
    {
        const char* e = ::getenv("LDU_DEVICE");
        int device = e ? ::atoi(e) : 0;
        if (!e && Pstream::parRun())
        {
            const int n = ldu_device_count();
            device = n > 0 ? Pstream::myProcNo() % n : 0;
        }
        check(ldu_context_create(device, NULL, &ctx), "ldu_context_create");
    }
    return ctx;
}

// The coupled patches of a matrix as the C ABI wants them.  Processor patches: the
// other side lives on another rank (another GPU), faces in the same order on both
// sides (processorLduInterface.H:88-97).  Cyclic patches: the other half is a patch
// of this very region (cyclicLduInterface.H:60-75), i.e. the neighbour rank is this
// rank.  Transforming cyclics (rotational periodicity) do nothing to a scalar
// (transformCoupleField with rank 0), which is all this solver sees.
struct coupledPatches
{
    DynamicList<label> patchIDs;      // interfaces_ index of compact interface i
    DynamicList<int> sizes;
    DynamicList<int> nbrRank;
    DynamicList<label> nbrPatchID;    // cyclic: interfaces_ index of the other half, else -1
    DynamicList<const int*> faceCells;
    bool anyProcessor;

    coupledPatches() : anyProcessor(false) {}
};

void findCoupledPatches
(
    const lduMatrix& A,
    const lduInterfaceFieldPtrsList& interfaces,
    coupledPatches& cp
)
{
    forAll(interfaces, patchi)
    {
        if (!interfaces.set(patchi)) continue;
        const lduInterface& li = interfaces[patchi].interface();
        const labelUList& fc = A.lduAddr().patchAddr(patchi);
        // a transforming couple (rotational cyclic / processorCyclic) scales the neighbour values of a
        // vector or tensor COMPONENT by pow(diag(forwardT).component(cmpt), rank) (cyclicLduInterfaceField.H:108-122);
        // the device interfaces are identity couples: right for scalars (rank 0) and untransformed patches only
        if
        (
            (isA<cyclicLduInterfaceField>(interfaces[patchi])
          && refCast<const cyclicLduInterfaceField>(interfaces[patchi]).doTransform()
          && refCast<const cyclicLduInterfaceField>(interfaces[patchi]).rank() > 0)
         || (isA<processorLduInterfaceField>(interfaces[patchi])
          && refCast<const processorLduInterfaceField>(interfaces[patchi]).doTransform()
          && refCast<const processorLduInterfaceField>(interfaces[patchi]).rank() > 0)
        )
        {
            FatalErrorIn("gpuLduSolver::solve")
                << "coupled patch " << patchi << " transforms the components of a vector/tensor field"
                << " (rotational cyclic): not supported by the GPU solver" << exit(FatalError);
        }
        if (isA<processorLduInterface>(li))
        {
            cp.nbrRank.append(refCast<const processorLduInterface>(li).neighbProcNo());
            cp.nbrPatchID.append(-1);
            cp.anyProcessor = true;
        }
        else if (isA<cyclicLduInterface>(li))
        {
            cp.nbrRank.append(Pstream::myProcNo());
            cp.nbrPatchID.append(refCast<const cyclicLduInterface>(li).neighbPatchID());
        }
        else
        {
            FatalErrorIn("gpuLduSolver::solve")
                << "coupled patch " << patchi << " of type " << li.type()
                << ": only processor and cyclic patches are handed to the GPU solver"
                << exit(FatalError);
        }
        cp.patchIDs.append(patchi);
        cp.sizes.append(fc.size());
        cp.faceCells.append(fc.begin());
    }
}

// Exchange window between the GPUs of a parallel run (ldu_b200.h: ldu_comm_*): the
// 64-byte handles travel once over the application's own Pstream, after that halos
// and sums go GPU to GPU.  Collective: every rank gets here in its first coupled solve.
void connectDevices(const coupledPatches& cp)
{
    static bool connected = false;
    static label windowInterfaces = 0, windowFaces = 0;
    label nIfs = cp.sizes.size(), nFaces = 1;
    forAll(cp.sizes, i) nFaces = max(nFaces, label(cp.sizes[i]));
    if (connected)
    {
        if (nIfs > windowInterfaces || nFaces > windowFaces)
        {
            FatalErrorIn("gpuLduSolver::solve")
                << "matrix with more/larger coupled patches than the exchange window"
                   " was created for" << exit(FatalError);
        }
        return;
    }
    reduce(nIfs, maxOp<label>());
    reduce(nFaces, maxOp<label>());
    const label me = Pstream::myProcNo(), n = Pstream::nProcs();
    List<char> all(n*LDU_COMM_HANDLE_BYTES);
    unsigned char* mine = reinterpret_cast<unsigned char*>(&all[me*LDU_COMM_HANDLE_BYTES]);
    check
    (
        ldu_comm_window_create(context(), me, n, nIfs, nFaces, mine),
        "ldu_comm_window_create"
    );
    for (label proc = 0; proc < n; proc++)
    {
        if (proc != me)
        {
            UOPstream::write
            (
                Pstream::blocking, proc,
                reinterpret_cast<const char*>(mine), LDU_COMM_HANDLE_BYTES
            );
        }
    }
    for (label proc = 0; proc < n; proc++)
    {
        if (proc != me)
        {
            UIPstream::read
            (
                Pstream::blocking, proc,
                &all[proc*LDU_COMM_HANDLE_BYTES], LDU_COMM_HANDLE_BYTES
            );
        }
    }
    check
    (
        ldu_comm_connect(context(), reinterpret_cast<const unsigned char*>(all.begin())),
        "ldu_comm_connect"
    );
    label sync = 0;
    reduce(sync, sumOp<label>());    // nobody starts before every window is mapped
    windowInterfaces = nIfs;
    windowFaces = nFaces;
    connected = true;
}

// Device copy of the addressing, built once per lduAddressing and kept for the
// life of the process (the reference caches GAMGAgglomeration on the mesh in the
// same spirit, GAMGAgglomeration.H:59-62).  Coefficients are refreshed per solve.
struct cachedMatrix
{
    ldu_matrix* m;
    label nCells;
    label nFaces;
    // fingerprint of the addressing (sizes, coupled patches, 4096 evenly spaced owner/neighbour pairs): an
    // lduAddressing destroyed and another one allocated at the same address must not inherit the device copy
    unsigned long long fingerprint;
    unsigned long long lastUse;
    // the GAMGAgglomeration (a MeshObject when cacheAgglomeration is on) whose levels the device holds
    const GAMGAgglomeration* agglomeration;
};

unsigned long long addressingFingerprint(const lduAddressing& addr, const coupledPatches& cp)
{
    unsigned long long h = 1469598103934665603ull;
    #define LDU_MIX(v) h = (h ^ (unsigned long long)(v))*1099511628211ull
    const labelUList& l = addr.lowerAddr();
    const labelUList& u = addr.upperAddr();
    LDU_MIX(addr.size());
    LDU_MIX(l.size());
    const label step = max(label(1), l.size()/4096);
    for (label f = 0; f < l.size(); f += step) { LDU_MIX(l[f]); LDU_MIX(u[f]); }
    if (l.size()) { LDU_MIX(l[l.size() - 1]); LDU_MIX(u[l.size() - 1]); }
    forAll(cp.sizes, i)
    {
        LDU_MIX(cp.sizes[i]);
        LDU_MIX(cp.nbrRank[i]);
        if (cp.sizes[i]) { LDU_MIX(cp.faceCells[i][0]); LDU_MIX(cp.faceCells[i][cp.sizes[i] - 1]); }
    }
    #undef LDU_MIX
    return h;
}

cachedMatrix& deviceMatrix(const lduMatrix& A, const coupledPatches& cp)
{
    static std::map<const lduAddressing*, cachedMatrix> cache;
    static unsigned long long useClock = 0;
    // the meshes of one application: its fvMesh(es) and, per GAMG agglomeration kept on a mesh, a couple of dozen
    // coarse levels when the reference's own GAMG calls gpu smoothers.  Beyond that the least recently used go
    // (serial runs only: creating a matrix is collective in a parallel run, eviction must not differ per rank)
    const label maxCached = 96;
    const lduAddressing& addr = A.lduAddr();
    const label nCells = addr.size();
    const label nFaces = addr.lowerAddr().size();
    const unsigned long long fp = addressingFingerprint(addr, cp);

    std::map<const lduAddressing*, cachedMatrix>::iterator it = cache.find(&addr);
    if (it != cache.end())
    {
        if (it->second.nCells == nCells && it->second.nFaces == nFaces && it->second.fingerprint == fp)
        {
            it->second.lastUse = ++useClock;
            return it->second;
        }
        ldu_matrix_destroy(it->second.m);   // another mesh under the same address
        cache.erase(it);
    }
    if (!Pstream::parRun() && label(cache.size()) >= maxCached)
    {
        std::map<const lduAddressing*, cachedMatrix>::iterator oldest = cache.begin();
        for (it = cache.begin(); it != cache.end(); ++it)
        {
            if (it->second.lastUse < oldest->second.lastUse) oldest = it;
        }
        ldu_matrix_destroy(oldest->second.m);
        cache.erase(oldest);
    }

    // index of the matching interface in the neighbour's (compact) list: each side
    // tells the other (patches towards one neighbour are in the same order on both sides)
    List<int> nbrInterface(cp.sizes.size(), 0);
    if (Pstream::parRun())
    {
        // collective, also for a rank whose own patches are all cyclic
        connectDevices(cp);
    }
    forAll(nbrInterface, i)
    {
        if (cp.nbrPatchID[i] >= 0)
        {
            forAll(cp.patchIDs, j)
            {
                if (cp.patchIDs[j] == cp.nbrPatchID[i]) nbrInterface[i] = j;
            }
        }
        else
        {
            const int mine = i;
            UOPstream::write
            (
                Pstream::blocking, cp.nbrRank[i],
                reinterpret_cast<const char*>(&mine), sizeof(int)
            );
        }
    }
    forAll(nbrInterface, i)
    {
        if (cp.nbrPatchID[i] < 0)
        {
            UIPstream::read
            (
                Pstream::blocking, cp.nbrRank[i],
                reinterpret_cast<char*>(&nbrInterface[i]), sizeof(int)
            );
        }
    }

    cachedMatrix c;
    c.nCells = nCells;
    c.nFaces = nFaces;
    c.m = NULL;
    c.agglomeration = NULL;
    c.fingerprint = fp;
    c.lastUse = ++useClock;
    check
    (
        ldu_matrix_create
        (
            context(), nCells, nFaces,
            addr.lowerAddr().begin(), addr.upperAddr().begin(),
            cp.sizes.size(),
            cp.sizes.size() ? cp.sizes.begin() : NULL,
            cp.sizes.size() ? cp.faceCells.begin() : NULL,
            cp.sizes.size() ? cp.nbrRank.begin() : NULL,
            cp.sizes.size() ? nbrInterface.begin() : NULL,
            &c.m
        ),
        "ldu_matrix_create"
    );
    cache[&addr] = c;
    return cache[&addr];
}

// The GAMG hierarchy is the reference's own: GAMGAgglomeration::New (GAMGAgglomeration.C:91-198) selects the
// agglomerator named in the dictionary from the reference's tables -- algebraicPair (libOpenFOAM), faceAreaPair
// (libfiniteVolume, needs fvMesh::Sf), MGridGen, ... -- and keeps it on the mesh as a MeshObject when
// cacheAgglomeration is on.  Its levels are handed to the device (ldu_b200.h: ldu_gamg_set_level); only the
// coefficients are agglomerated there, per solve.  Exactly as GAMGSolver does (GAMGSolver.C:52-154): looked up or
// built per solver object, deleted again afterwards unless cached.
void uploadAgglomeration
(
    cachedMatrix& cm,
    const lduMatrix& A,
    const dictionary& gamgDict,
    const coupledPatches& cp
)
{
    const bool cacheAgglomeration = gamgDict.lookupOrDefault<Switch>("cacheAgglomeration", false);
    const GAMGAgglomeration& agg = GAMGAgglomeration::New(A, gamgDict);

    if (!(cacheAgglomeration && cm.agglomeration == &agg))
    {
        check(ldu_gamg_begin_levels(cm.m), "ldu_gamg_begin_levels");
        for (label lev = 0; lev < agg.size(); lev++)
        {
            const labelField& restrictAddr = agg.restrictAddressing(lev);
            const labelList& faceRestrictAddr = agg.faceRestrictAddressing(lev);
            const lduAddressing& coarse = agg.meshLevel(lev + 1).lduAddr();
            const lduInterfacePtrsList& ifs = agg.interfaceLevel(lev + 1);
            List<int> sizes(cp.patchIDs.size());
            List<const int*> cells(cp.patchIDs.size()), ifRestrict(cp.patchIDs.size());
            forAll(cp.patchIDs, i)
            {
                const GAMGInterface& gi = refCast<const GAMGInterface>(ifs[cp.patchIDs[i]]);
                sizes[i] = gi.faceCells().size();
                cells[i] = gi.faceCells().begin();
                ifRestrict[i] = gi.faceRestrictAddressing().begin();
            }
            check
            (
                ldu_gamg_set_level
                (
                    cm.m, lev,
                    restrictAddr.size(), restrictAddr.begin(),
                    faceRestrictAddr.size(), faceRestrictAddr.begin(),
                    coarse.size(), coarse.lowerAddr().size(),
                    coarse.lowerAddr().begin(), coarse.upperAddr().begin(),
                    sizes.size() ? sizes.begin() : NULL,
                    sizes.size() ? cells.begin() : NULL,
                    sizes.size() ? ifRestrict.begin() : NULL
                ),
                "ldu_gamg_set_level"
            );
        }
        check(ldu_gamg_end_levels(cm.m), "ldu_gamg_end_levels");
        cm.agglomeration = &agg;
    }

    if (!cacheAgglomeration)
    {
        delete &agg;          // GAMGSolver::~GAMGSolver, GAMGSolver.C:144-154
        cm.agglomeration = NULL;
    }
}

// addressing (cached) + this object's coefficients on the device
cachedMatrix& uploadMatrix
(
    const lduMatrix& A,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    coupledPatches& cp
)
{
    findCoupledPatches(A, interfaces, cp);
    cachedMatrix& cm = deviceMatrix(A, cp);

    // interfaceBouCoeffs_/interfaceIntCoeffs_ of the coupled patches (lduMatrix.H:97-104)
    List<const double*> bou(cp.patchIDs.size()), intc(cp.patchIDs.size());
    forAll(cp.patchIDs, i)
    {
        bou[i] = interfaceBouCoeffs[cp.patchIDs[i]].begin();
        intc[i] = interfaceIntCoeffs[cp.patchIDs[i]].begin();
    }
    const bool hasUpper = A.hasUpper() || A.hasLower();
    check
    (
        ldu_matrix_set_coeffs
        (
            cm.m,
            A.diag().begin(),
            hasUpper ? A.upper().begin() : NULL,
            A.asymmetric() ? A.lower().begin() : NULL,
            bou.size() ? bou.begin() : NULL,
            intc.size() ? intc.begin() : NULL
        ),
        "ldu_matrix_set_coeffs"
    );
    return cm;
}

int preconditionerKind(const word& name)
{
    if (name == "none") return LDU_PRECOND_NONE;
    if (name == "diagonal") return LDU_PRECOND_DIAGONAL;
    if (name == "DIC") return LDU_PRECOND_DIC;
    if (name == "FDIC") return LDU_PRECOND_FDIC;
    if (name == "DILU") return LDU_PRECOND_DILU;
    if (name == "GAMG") return LDU_PRECOND_GAMG;
    FatalErrorIn("gpuLduSolver") << "Unknown preconditioner " << name
        << exit(FatalError);
    return -1;
}

int smootherKind(const word& name)
{
    if (name == "GaussSeidel") return LDU_SMOOTHER_GS;
    if (name == "symGaussSeidel") return LDU_SMOOTHER_SYMGS;
    if (name == "nonBlockingGaussSeidel") return LDU_SMOOTHER_NBGS;
    if (name == "multiColourGaussSeidel") return LDU_SMOOTHER_MCGS;   // extension, see ldu_b200.h
    if (name == "DIC") return LDU_SMOOTHER_DIC;
    if (name == "DILU") return LDU_SMOOTHER_DILU;
    if (name == "FDIC") return LDU_SMOOTHER_FDIC;
    if (name == "DICGaussSeidel") return LDU_SMOOTHER_DICGS;
    if (name == "DILUGaussSeidel") return LDU_SMOOTHER_DILUGS;
    FatalErrorIn("gpuLduSolver") << "Unknown smoother " << name
        << exit(FatalError);
    return -1;
}

// GAMG keys (GAMGSolver.C:157-181, GAMGAgglomeration.C:77-80, pairGAMGAgglomeration.C:45)
void readGamgControls(const dictionary& d, ldu_controls& c)
{
    // the dense LU solve of the coarsest level (GAMGSolver.C:74,91-106) is outside this library:
    // say so instead of silently running the iterative coarsest-level solver
    if (d.lookupOrDefault<Switch>("directSolveCoarsest", false))
    {
        FatalErrorIn("gpuLduSolver") << "directSolveCoarsest is not supported by the GPU GAMG solver"
            << exit(FatalError);
    }
    // mandatory in the reference: GAMGAgglomeration.C:77-80 reads it without a default
    c.nCellsInCoarsestLevel = readLabel(d.lookup("nCellsInCoarsestLevel"));
    // mandatory in the reference: pairGAMGAgglomeration.C:45 reads it without a default
    c.mergeLevels = readLabel(d.lookup("mergeLevels"));
    c.nPreSweeps = d.lookupOrDefault<label>("nPreSweeps", 0);
    c.preSweepsLevelMultiplier = d.lookupOrDefault<label>("preSweepsLevelMultiplier", 1);
    c.maxPreSweeps = d.lookupOrDefault<label>("maxPreSweeps", 4);
    c.nPostSweeps = d.lookupOrDefault<label>("nPostSweeps", 2);
    c.postSweepsLevelMultiplier = d.lookupOrDefault<label>("postSweepsLevelMultiplier", 1);
    c.maxPostSweeps = d.lookupOrDefault<label>("maxPostSweeps", 4);
    c.nFinestSweeps = d.lookupOrDefault<label>("nFinestSweeps", 2);
    c.interpolateCorrection = d.lookupOrDefault<Switch>("interpolateCorrection", false);
    if (d.found("scaleCorrection"))
    {
        c.scaleCorrection = Switch(d.lookup("scaleCorrection"));
    }
    c.cacheAgglomeration = d.lookupOrDefault<Switch>("cacheAgglomeration", false);
    c.nVcycles = d.lookupOrDefault<label>("nVcycles", 2);
    // mandatory in the reference (lduMatrixSmoother.C:38-66, GAMGAgglomeration.C:104-107)
    c.smoother = smootherKind(word(d.lookup("smoother")));
    // mandatory in the reference (GAMGAgglomeration.C:104-107).  The agglomeration itself is the reference's:
    // uploadAgglomeration hands its levels to the device, whatever the agglomerator
    const word agglomerator(d.lookup("agglomerator"));
    c.useFaceWeights = 0;
}

} // End anonymous namespace


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::gpuLduSolver::gpuLduSolver
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces,
    const dictionary& solverControls
)
:
    lduMatrix::solver
    (
        fieldName,
        matrix,
        interfaceBouCoeffs,
        interfaceIntCoeffs,
        interfaces,
        solverControls
    )
{}


// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

void Foam::gpuLduSolver::fillControls(ldu_controls& c) const
{
    ldu_controls_default(&c);
    c.solver = solverKind();
    c.maxIter = maxIter_;
    c.tolerance = tolerance_;
    c.relTol = relTol_;
    c.nSweeps = controlDict_.lookupOrDefault<label>("nSweeps", 1);
    c.referenceOrderSums =
        controlDict_.lookupOrDefault<Switch>("referenceOrderSums", false);

    if (solverKind() == LDU_SOLVER_SMOOTH)
    {
        // mandatory in the reference (lduMatrixSmoother.C:38-66)
        c.smoother = smootherKind(word(controlDict_.lookup("smoother")));
    }

    if (solverKind() == LDU_SOLVER_PCG || solverKind() == LDU_SOLVER_PBICG)
    {
        // word or sub-dictionary (lduMatrixPreconditioner.C:39-58)
        c.preconditioner =
            preconditionerKind(lduMatrix::preconditioner::getName(controlDict_));
        const entry& e = controlDict_.lookupEntry("preconditioner", false, false);
        if (e.isDict() && c.preconditioner == LDU_PRECOND_GAMG)
        {
            const dictionary& pd = e.dict();
            readGamgControls(pd, c);
            c.precTolerance = pd.lookupOrDefault<scalar>("tolerance", 1e-6);
            c.precRelTol = pd.lookupOrDefault<scalar>("relTol", 0);
        }
    }
    else if (solverKind() == LDU_SOLVER_GAMG)
    {
        readGamgControls(controlDict_, c);
    }
}


Foam::solverPerformance Foam::gpuLduSolver::solve
(
    scalarField& psi,
    const scalarField& source,
    const direction
) const
{
    // dictionary errors first, as the reference's constructors raise them (readControls)
    ldu_controls c;
    fillControls(c);

    coupledPatches cp;
    cachedMatrix& cm = uploadMatrix(matrix_, interfaceBouCoeffs_, interfaceIntCoeffs_, interfaces_, cp);
    ldu_matrix* m = cm.m;
    if (c.solver == LDU_SOLVER_GAMG)
    {
        uploadAgglomeration(cm, matrix_, controlDict_, cp);
    }
    else if (c.preconditioner == LDU_PRECOND_GAMG)
    {
        uploadAgglomeration
        (
            cm, matrix_, controlDict_.lookupEntry("preconditioner", false, false).dict(), cp
        );
    }

    ldu_solver_performance p;
    check(ldu_solve(m, &c, psi.begin(), source.begin(), &p), "ldu_solve");

    // matrix.diagonal() goes to diagonalSolver in the reference before any table
    // look-up (lduMatrixSolver.C:52-66); the library does the same internally
    return solverPerformance
    (
        performanceName(),
        fieldName_,
        p.initialResidual,
        p.finalResidual,
        p.nIterations,
        p.converged,
        p.singular
    );
}


// * * * * * * * * * * * * * smoothers / preconditioners * * * * * * * * * * * //

Foam::gpuLduSmoother::gpuLduSmoother
(
    const word& fieldName,
    const lduMatrix& matrix,
    const FieldField<Field, scalar>& interfaceBouCoeffs,
    const FieldField<Field, scalar>& interfaceIntCoeffs,
    const lduInterfaceFieldPtrsList& interfaces
)
:
    lduMatrix::smoother(fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces),
    m_(NULL)
{
    coupledPatches cp;
    m_ = uploadMatrix(matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces, cp).m;
}


void Foam::gpuLduSmoother::smooth
(
    scalarField& psi,
    const scalarField& source,
    const direction,
    const label nSweeps
) const
{
    check
    (
        ldu_smooth(static_cast<ldu_matrix*>(m_), smootherKind(), psi.begin(), source.begin(), nSweeps),
        "ldu_smooth"
    );
}


Foam::gpuLduPreconditioner::gpuLduPreconditioner(const lduMatrix::solver& sol, const dictionary&)
:
    lduMatrix::preconditioner(sol),
    m_(NULL)
{
    coupledPatches cp;
    m_ = uploadMatrix
    (
        sol.matrix(), sol.interfaceBouCoeffs(), sol.interfaceIntCoeffs(), sol.interfaces(), cp
    ).m;
}


void Foam::gpuLduPreconditioner::precondition(scalarField& wA, const scalarField& rA, const direction) const
{
    check
    (
        ldu_precondition(static_cast<ldu_matrix*>(m_), preconditionerKind(), wA.begin(), rA.begin(), 0),
        "ldu_precondition"
    );
}


void Foam::gpuLduPreconditioner::preconditionT(scalarField& wT, const scalarField& rT, const direction) const
{
    check
    (
        ldu_precondition(static_cast<ldu_matrix*>(m_), preconditionerKind(), wT.begin(), rT.begin(), 1),
        "ldu_precondition"
    );
}


// ************************************************************************* //
