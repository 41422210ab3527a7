/*
 * ldu_b200.h — C ABI of the B200-native lduMatrix linear-solver path.
 *
 * Drop-in boundary for OpenFOAM-2.2.x's lduMatrix::solver run-time-selection
 * API: the host-side shim (openfoam-2.2.x_b200/foam/gpuLduSolvers.C, a
 * lduMatrix::solver subclass registered in the reference's own tables) and the
 * Python mirror (openfoam-2.2.x_b200/ldub200) translate OpenFOAM objects into
 * the plain pointers + sizes below.  No C++/torch types cross this boundary.
 * All integers are int32 (Foam::label), all reals are fp64 (Foam::scalar with
 * WM_DP).  Reference paths are relative to /root/reference/src/OpenFOAM/.
 *
 * Every function returns 0 on success, a negative LDU_E* code on failure;
 * ldu_last_error() gives the message (the C++ shim turns it into FatalError,
 * matching the reference's error convention, SURVEY.md §8b "Errors").
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with LDU_ENODEVICE.
 */
#ifndef LDU_B200_H
#define LDU_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define LDU_OK 0
#define LDU_ENODEVICE (-1)   /* no usable CUDA device / driver */
#define LDU_ECUDA (-2)       /* a CUDA runtime call failed */
#define LDU_EINVAL (-3)      /* bad argument */
#define LDU_ECOMM (-4)       /* multi-GPU exchange failed / timed out */
#define LDU_EUNSUPPORTED (-5)

typedef struct ldu_context ldu_context; /* one GPU + stream (+ peers)          */
typedef struct ldu_matrix ldu_matrix;   /* one mesh region: lduAddressing +    */
                                        /* lduMatrix coefficients + interfaces */

/* run-time selection names of the reference, as enums ------------------------ */
/* solvers: matrices/lduMatrix/solvers/{PCG,PBiCG,smoothSolver,GAMG,diagonalSolver} */
enum { LDU_SOLVER_PCG = 0, LDU_SOLVER_PBICG = 1, LDU_SOLVER_SMOOTH = 2,
       LDU_SOLVER_GAMG = 3, LDU_SOLVER_DIAGONAL = 4 };
/* preconditioners: matrices/lduMatrix/preconditioners/ */
enum { LDU_PRECOND_NONE = 0, LDU_PRECOND_DIAGONAL = 1, LDU_PRECOND_DIC = 2,
       LDU_PRECOND_FDIC = 3, LDU_PRECOND_DILU = 4, LDU_PRECOND_GAMG = 5 };
/* smoothers: matrices/lduMatrix/smoothers/ */
enum { LDU_SMOOTHER_GS = 0, LDU_SMOOTHER_SYMGS = 1, LDU_SMOOTHER_DIC = 2,
       LDU_SMOOTHER_DILU = 3, LDU_SMOOTHER_DICGS = 4, LDU_SMOOTHER_DILUGS = 5,
       LDU_SMOOTHER_FDIC = 6, LDU_SMOOTHER_NBGS = 7,
       /* `smoother multiColourGaussSeidel;` -- an EXTENSION, not a reference smoother: GaussSeidelSmoother.C:66-187
        * with the cells of every matrix (every GAMG level its own) visited colour by colour of a greedy colouring
        * instead of in cell order, all cells of a colour in parallel.  It is the reference's Gauss-Seidel on the
        * mesh renumbered by colour (ldu_colour_order): other iterates than the lexicographic smoother, same
        * smoothing property; the dependency chain of a sweep is the number of colours (2 on a hex box). */
       LDU_SMOOTHER_MCGS = 8 };

/*
 * Solver controls = the keys the reference's solver constructors read from the
 * fvSolution sub-dictionary:
 *   lduMatrix::solver::readControls   matrices/lduMatrix/lduMatrix/lduMatrixSolver.C:164-169
 *   smoothSolver::readControls        matrices/lduMatrix/solvers/smoothSolver/smoothSolver.C:70-74
 *   GAMGSolver ctor / readControls    matrices/lduMatrix/solvers/GAMG/GAMGSolver.C:66-76,157-181
 *   GAMGAgglomeration / pair          .../GAMGAgglomeration.C:77-80, pairGAMGAgglomeration.C:45
 *   GAMGPreconditioner::readControls  matrices/lduMatrix/preconditioners/GAMGPreconditioner/GAMGPreconditioner.C:74-78
 * Fill with ldu_controls_default() first, then override.
 */
typedef struct ldu_controls {
    int solver;
    int preconditioner;
    int smoother;
    int maxIter;               /* 1000 */
    double tolerance;          /* 1e-6 */
    double relTol;             /* 0 */
    int nSweeps;               /* 1 */
    int nCellsInCoarsestLevel; /* 10 */
    int mergeLevels;           /* 1 */
    int nPreSweeps;            /* 0 */
    int preSweepsLevelMultiplier; /* 1 */
    int maxPreSweeps;          /* 4 */
    int nPostSweeps;           /* 2 */
    int postSweepsLevelMultiplier; /* 1 */
    int maxPostSweeps;         /* 4 */
    int nFinestSweeps;         /* 2 */
    int interpolateCorrection; /* 0 */
    int scaleCorrection;       /* -1 = matrix.symmetric() */
    int nVcycles;              /* 2 (GAMG as preconditioner) */
    double precTolerance;      /* tolerance / relTol of a GAMG preconditioner sub-dict */
    double precRelTol;
    int useFaceWeights;        /* 1: agglomerate on ldu_matrix_set_face_weights (faceAreaPair)
                                  0: on mag(upper) (algebraicPair) */
    int cacheAgglomeration;    /* keep the GAMG hierarchy addressing across solves */
    int checkInterval;         /* iterations enqueued between host convergence polls
                                  (device decides the exact stopping iteration) */
    int referenceOrderSums;    /* 0 (default): dot products / norms by a fixed parallel tree.
                                  1: accumulate them strictly left to right like the reference's
                                  sumProd/sumMag (FieldFunctions.C:363-385,422-434): every iterate
                                  becomes bit-identical to the reference; O(nCells) dependent adds,
                                  meant for verification on small/medium systems */
} ldu_controls;

/* SolverPerformance<scalar>: matrices/LduMatrix/LduMatrix/SolverPerformance.H:78-137 */
typedef struct ldu_solver_performance {
    double initialResidual;
    double finalResidual;
    int nIterations;
    int converged;
    int singular;
} ldu_solver_performance;

/* ---- library / context ---------------------------------------------------- */
const char* ldu_version(void);
const char* ldu_last_error(void);
/* number of CUDA kernels this library has launched in this process */
long long ldu_launch_count(void);

/* CUDA devices visible to this process (0 when there is none); a multi-rank host binds rank r to
 * device r % count, the way mpirun-launched applications pick a GPU per rank */
int ldu_device_count(void);
/* device: CUDA ordinal; stream: a cudaStream_t to enqueue on (NULL = create one) */
int ldu_context_create(int device, void* stream, ldu_context** out);
int ldu_context_destroy(ldu_context* ctx);
int ldu_context_synchronize(ldu_context* ctx);
void* ldu_context_stream(ldu_context* ctx);

/*
 * Multi-GPU (one process per GPU, one mesh region per GPU): replaces libPstream
 * on the hot path (src/Pstream/mpi/UPstream.C:174-204 reduce(sumOp),
 * UIPread.C:280-300 / UOPwrite.C:97-110 non-blocking halo send/recv).
 * Peers put halos and reduction partials straight into each other's exchange
 * window over NVLink.  The host runtime (torch.distributed / MPI / anything)
 * only carries the opaque 64-byte window handles between ranks once.
 *   1. every rank: ldu_comm_window_create -> handle[64]
 *   2. host all-gathers the handles (rank-major, 64 B each)
 *   3. every rank: ldu_comm_connect(allHandles)
 */
#define LDU_COMM_HANDLE_BYTES 64
/* maxInterfaces: most coupled patches any rank has (finest or coarse level);
 * maxInterfaceFaces: most faces any single coupled patch has */
int ldu_comm_window_create(ldu_context* ctx, int rank, int nRanks, int maxInterfaces,
                           long long maxInterfaceFaces, unsigned char* handleOut);
int ldu_comm_connect(ldu_context* ctx, const unsigned char* allHandles);

/* ---- device memory (for callers that keep fields resident in HBM) ---------- */
int ldu_device_alloc(ldu_context* ctx, long long bytes, void** dptr);
int ldu_device_free(ldu_context* ctx, void* dptr);
int ldu_copy_h2d(ldu_context* ctx, void* dst, const void* src, long long bytes);
int ldu_copy_d2h(ldu_context* ctx, void* dst, const void* src, long long bytes);
int ldu_device_memset(ldu_context* ctx, void* dptr, int byteValue, long long bytes);
/* page-locked host staging buffers (optional, speeds up the host entry points) */
int ldu_host_alloc(long long bytes, void** hptr);
int ldu_host_free(void* hptr);

/* ---- matrix = lduAddressing + lduMatrix + interfaces ---------------------- */
/*
 * Addressing, built once per mesh (lduAddressing: matrices/lduMatrix/lduAddressing/
 * lduAddressing.H:111-199; losort/ownerStart/losortStart: lduAddressing.C:31-169).
 * lowerAddr/upperAddr: owner/neighbour of every internal face, upper-triangular
 * order (lower[f] < upper[f], faces sorted by lower).
 * Interfaces = the coupled patches (lduInterface::faceCells, lduInterface.H:54-113):
 * interface i has ifaceSizes[i] faces, faceCells[i][k] = local cell of face k,
 * nbrRank[i] = rank owning the other side, nbrInterface[i] = index of the
 * matching interface there (both sides list faces in the same order).
 */
int ldu_matrix_create(ldu_context* ctx, int nCells, int nFaces,
                      const int* lowerAddr, const int* upperAddr,
                      int nInterfaces, const int* ifaceSizes,
                      const int* const* faceCells, const int* nbrRank,
                      const int* nbrInterface, ldu_matrix** out);
int ldu_matrix_destroy(ldu_matrix* m);
/*
 * Coefficients, refreshed every solve (lduMatrix.H:77-86; lower == NULL means
 * symmetric: lduMatrix.C:198-215).  bouCoeffs/intCoeffs: one array per interface
 * (interfaceBouCoeffs_/interfaceIntCoeffs_, lduMatrix.H:97-104).  Host pointers.
 * upper == NULL and lower == NULL (allowed only when nFaces == 0) is lduMatrix::diagonal()
 * (lduMatrix.H:547-550): ldu_solve then runs diagonalSolver whatever the controls say
 * (lduMatrixSolver.C:52-66).  A faceless matrix given a non-NULL upper is NOT diagonal() — the
 * state fvm::laplacian leaves on a one-cell mesh — and goes to the selected solver.
 */
int ldu_matrix_set_coeffs(ldu_matrix* m, const double* diag, const double* upper,
                          const double* lower, const double* const* bouCoeffs,
                          const double* const* intCoeffs);
/* same, from device-resident arrays (no PCIe traffic).  A matrix with coupled patches needs their
 * boundary coefficients too: give them (host pointers, one array per interface, a few KB) with
 * ldu_matrix_set_interface_coeffs before the first ldu_matrix_set_coeffs_device, and again whenever
 * they change; without them the call fails with LDU_EINVAL. */
int ldu_matrix_set_coeffs_device(ldu_matrix* m, const double* d_diag, const double* d_upper,
                                 const double* d_lower);
int ldu_matrix_set_interface_coeffs(ldu_matrix* m, const double* const* bouCoeffs,
                                    const double* const* intCoeffs);
/* faceAreaPair agglomeration weights (finiteVolume/.../faceAreaPairGAMGAgglomeration.C:48-73) */
int ldu_matrix_set_face_weights(ldu_matrix* m, const double* weights);

/* Colour-ordered renumbering, host only (stands where renumberMesh stands in the reference
 * tool chain, applications/utilities/mesh/manipulation/renumberMesh): greedy colouring of the cell graph, cells
 * sorted by (colour, old index).  newIndexOfOldCell[nCells] out; *nColours out (may be NULL).  In that numbering
 * the reference's lexicographic Gauss-Seidel / DIC sweeps have a dependency depth of nColours. */
int ldu_colour_order(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr,
                     int* newIndexOfOldCell, int* nColours);

/* Cuthill-McKee band compression, host only: the numbering renumberMesh's default method gives
 * (meshes/bandCompression/bandCompression.C:42-148 on the cell-cell addressing of the internal faces,
 * restated with its quirks so that the result is the reference's).  oldCellOfNewCell[nCells] out. */
int ldu_band_compression(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr,
                         int* oldCellOfNewCell);

/* ---- operators: host in / host out ---------------------------------------- */
/* lduMatrix::Amul  matrices/lduMatrix/lduMatrix/lduMatrixATmul.C:34-92 */
int ldu_amul(ldu_matrix* m, double* Apsi, const double* psi);
/* lduMatrix::Tmul  lduMatrixATmul.C:95-151 */
int ldu_tmul(ldu_matrix* m, double* Tpsi, const double* psi);
/* lduMatrix::sumA  lduMatrixATmul.C:154-200 */
int ldu_sumA(ldu_matrix* m, double* sumA);
/* lduMatrix::residual  lduMatrixATmul.C:203-281 */
int ldu_residual(ldu_matrix* m, double* rA, const double* psi, const double* source);
/* lduMatrix::H  matrices/lduMatrix/lduMatrix/lduMatrixTemplates.C:33-65: Hpsi = -(A - diag) psi over the
 * internal faces (what fvMatrix::H / HbyA = rAU*UEqn.H() of the PISO/SIMPLE step call, icoFoam.C:69-72) */
int ldu_H(ldu_matrix* m, double* Hpsi, const double* psi);
/* lduMatrix::H1  lduMatrixATmul.C:298-327: H1 = -rowsum(A - diag) */
int ldu_H1(ldu_matrix* m, double* H1);
/* lduMatrix::faceH  lduMatrixTemplates.C:79-113: faceHpsi[nFaces] = upper*psi[u] - lower*psi[l]
 * (fvMatrix::flux); LDU_EINVAL for a matrix without off-diagonal coefficients (the reference aborts) */
int ldu_faceH(ldu_matrix* m, double* faceHpsi, const double* psi);
/* lduMatrix::preconditioner::precondition / preconditionT  lduMatrix.H:482-505 */
int ldu_precondition(ldu_matrix* m, int preconditioner, double* wA, const double* rA, int transpose);
/* lduMatrix::smoother::smooth  lduMatrix.H:391-397 (psi in/out) */
int ldu_smooth(ldu_matrix* m, int smoother, double* psi, const double* source, int nSweeps);
/* lduMatrix::solver::solve  lduMatrix.H:242-247 (psi in/out) */
int ldu_solve(ldu_matrix* m, const ldu_controls* controls, double* psi,
              const double* source, ldu_solver_performance* perf);

/* ---- operators: fields resident in HBM (ldu_device_alloc'd pointers) ------- */
int ldu_amul_device(ldu_matrix* m, double* d_Apsi, const double* d_psi);
int ldu_tmul_device(ldu_matrix* m, double* d_Tpsi, const double* d_psi);
int ldu_H_device(ldu_matrix* m, double* d_Hpsi, const double* d_psi);
int ldu_solve_device(ldu_matrix* m, const ldu_controls* controls, double* d_psi,
                     const double* d_source, ldu_solver_performance* perf);
/* normalised residual after every iteration of the LAST solve on this matrix
 * (entry 0 = initial); returns the number of entries written */
int ldu_residual_history(ldu_matrix* m, double* hist, int capacity);

/* ---- GAMG hierarchy built by the host ------------------------------------- */
/*
 * The reference builds its hierarchy on the host and keeps it on the mesh:
 * GAMGAgglomeration::New(matrix|mesh, dict)  solvers/GAMG/GAMGAgglomerations/GAMGAgglomeration/GAMGAgglomeration.C:91-198
 * (run-time selected: algebraicPair in libOpenFOAM, faceAreaPair in libfiniteVolume,
 * finiteVolume/fvMatrices/solvers/GAMGSymSolver/GAMGAgglomerations/faceAreaPairGAMGAgglomeration/faceAreaPairGAMGAgglomeration.C:46-73,
 * MGridGen ...; cached as a MeshObject with `cacheAgglomeration on`).  A host that has that object hands its levels
 * over instead of letting the library agglomerate (which it can only do for the pair agglomerators):
 *   ldu_gamg_begin_levels(m);
 *   for level = 0 .. agglomeration.size()-1:      (level l maps mesh level l onto mesh level l+1)
 *     ldu_gamg_set_level(m, level,
 *         nFine, restrictAddressing(level),                 GAMGAgglomeration.H:206-209
 *         nFineFaces, faceRestrictAddressing(level),        GAMGAgglomeration.H:212-215 (>= 0: coarse face, < 0: -1 - coarse cell)
 *         nCoarse, nCoarseFaces, lowerAddr, upperAddr of meshLevel(level+1).lduAddr(),
 *         per coupled patch of the matrix, in ldu_matrix_create's order, the GAMGInterface of
 *         interfaceLevel(level+1): size, faceCells() and faceRestrictAddressing()  (GAMGInterface.H:158-187));
 *   ldu_gamg_end_levels(m);
 * After that every GAMG solve / preconditioner on m uses these levels as they are (only the coefficients are
 * agglomerated per solve, GAMGSolverAgglomerateMatrix.C:31-207) until ldu_gamg_begin_levels is called again or
 * ldu_gamg_internal_levels(m) gives the agglomeration back to the library.
 */
int ldu_gamg_begin_levels(ldu_matrix* m);
int ldu_gamg_set_level(ldu_matrix* m, int level, int nFine, const int* restrictAddr, int nFineFaces,
                       const int* faceRestrictAddr, int nCoarse, int nCoarseFaces, const int* coarseLower,
                       const int* coarseUpper, const int* coarseIfSizes, const int* const* coarseIfFaceCells,
                       const int* const* ifRestrictAddr);
int ldu_gamg_end_levels(ldu_matrix* m);
int ldu_gamg_internal_levels(ldu_matrix* m);

/* ---- GAMG hierarchy introspection (parity tests) -------------------------- */
int ldu_gamg_build(ldu_matrix* m, const ldu_controls* controls);
int ldu_gamg_nlevels(ldu_matrix* m);
int ldu_gamg_level_sizes(ldu_matrix* m, int level, int* nFine, int* nCoarse, int* nCoarseFaces);
int ldu_gamg_level_restrict(ldu_matrix* m, int level, int* restrictAddr /* [nFine] */);
int ldu_gamg_level_coeffs(ldu_matrix* m, int level, double* diag, double* upper, double* lower);

void ldu_controls_default(ldu_controls* c);

#ifdef __cplusplus
}
#endif
#endif
