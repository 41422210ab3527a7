/*---------------------------------------------------------------------------*\
  gpuLduSolvers — see gpuLduSolvers.H.  Thin shim: OpenFOAM objects in, POD out.
\*---------------------------------------------------------------------------*/

#include "gpuLduSolvers.H"
#include "addToRunTimeSelectionTable.H"
#include "Switch.H"

#include "../../include/ldu_b200.h"

#include <cstdlib>
#include <map>

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
    defineTypeNameAndDebug(gpuPCG, 0);
    defineTypeNameAndDebug(gpuPBiCG, 0);
    defineTypeNameAndDebug(gpuSmoothSolver, 0);
    defineTypeNameAndDebug(gpuGAMG, 0);

    // symmetric matrices: PCG, GAMG, smoothSolver (PCG.C:34, GAMGSolver.C:34, smoothSolver.C:34)
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuPCG>
        addgpuPCGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuGAMG>
        addgpuGAMGSymMatrixConstructorToTable_;
    lduMatrix::solver::addsymMatrixConstructorToTable<gpuSmoothSolver>
        addgpuSmoothSolverSymMatrixConstructorToTable_;

    // asymmetric matrices: PBiCG, GAMG, smoothSolver (PBiCG.C:34, GAMGSolver.C:37, smoothSolver.C:37)
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuPBiCG>
        addgpuPBiCGAsymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuGAMG>
        addgpuGAMGAsymMatrixConstructorToTable_;
    lduMatrix::solver::addasymMatrixConstructorToTable<gpuSmoothSolver>
        addgpuSmoothSolverAsymMatrixConstructorToTable_;

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

ldu_context* context()
{
    static ldu_context* ctx = NULL;
    if (!ctx)
    {
        const char* e = ::getenv("LDU_DEVICE");
        check(ldu_context_create(e ? ::atoi(e) : 0, NULL, &ctx), "ldu_context_create");
    }
    return ctx;
}

// Device copy of the addressing, built once per lduAddressing and kept for the
// life of the process (the reference caches GAMGAgglomeration on the mesh in the
// same spirit, GAMGAgglomeration.H:59-62).  Coefficients are refreshed per solve.
struct cachedMatrix
{
    ldu_matrix* m;
    label nCells;
    label nFaces;
};

ldu_matrix* deviceMatrix(const lduMatrix& A)
{
    static std::map<const lduAddressing*, cachedMatrix> cache;
    const lduAddressing& addr = A.lduAddr();
    const label nCells = addr.size();
    const label nFaces = addr.lowerAddr().size();

    std::map<const lduAddressing*, cachedMatrix>::iterator it = cache.find(&addr);
    if (it != cache.end())
    {
        if (it->second.nCells == nCells && it->second.nFaces == nFaces)
        {
            return it->second.m;
        }
        ldu_matrix_destroy(it->second.m);   // mesh changed under the same address
        cache.erase(it);
    }

    cachedMatrix c;
    c.nCells = nCells;
    c.nFaces = nFaces;
    c.m = NULL;
    check
    (
        ldu_matrix_create
        (
            context(), nCells, nFaces,
            addr.lowerAddr().begin(), addr.upperAddr().begin(),
            0, NULL, NULL, NULL, NULL, &c.m
        ),
        "ldu_matrix_create"
    );
    cache[&addr] = c;
    return c.m;
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
    c.nCellsInCoarsestLevel = d.lookupOrDefault<label>("nCellsInCoarsestLevel", 10);
    c.mergeLevels = d.lookupOrDefault<label>("mergeLevels", 1);
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
    if (d.found("smoother"))
    {
        c.smoother = smootherKind(word(d.lookup("smoother")));
    }
    // faceAreaPair needs fvMesh::Sf() (libfiniteVolume); until the weights are
    // handed over the algebraic pair agglomerator is used for every name
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

    if (controlDict_.found("smoother"))
    {
        c.smoother = smootherKind(word(controlDict_.lookup("smoother")));
    }

    if (solverKind() == LDU_SOLVER_PCG || solverKind() == LDU_SOLVER_PBICG)
    {
        // word or sub-dictionary (lduMatrixPreconditioner.C:39-58)
        c.preconditioner =
            preconditionerKind(lduMatrix::preconditioner::getName(controlDict_));
        const entry& e = controlDict_.lookupEntry("preconditioner", false, false);
        if (e.isDict())
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
    forAll(interfaces_, patchi)
    {
        if (interfaces_.set(patchi))
        {
            FatalErrorIn("gpuLduSolver::solve")
                << "coupled patches (processor/cyclic) reach the GPU solver through"
                   " ldu_matrix_create's interface arguments, which this shim does"
                   " not fill yet; run the case undecomposed"
                << exit(FatalError);
        }
    }

    ldu_matrix* m = deviceMatrix(matrix_);

    check
    (
        ldu_matrix_set_coeffs
        (
            m,
            matrix_.diag().begin(),
            matrix_.upper().begin(),
            matrix_.asymmetric() ? matrix_.lower().begin() : NULL,
            NULL,
            NULL
        ),
        "ldu_matrix_set_coeffs"
    );

    ldu_controls c;
    fillControls(c);

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


// ************************************************************************* //
