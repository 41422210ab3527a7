// Internal declarations of libldu_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/ldu_b200.h"

namespace ldu {

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define LDU_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return ldu::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define LDU_TRY(call)                 \
    do {                              \
        int rc__ = (call);            \
        if (rc__ != LDU_OK) return rc__; \
    } while (0)

extern long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

// ---------------------------------------------------------------------------
// device-side solver scalars (one block per solve, lives in HBM)
// ---------------------------------------------------------------------------
// Mirrors the locals of PCG::solve / PBiCG::solve (solvers/PCG/PCG.C:65-182):
// nothing but the final SolverPerformance crosses back to the host.
struct SolverScalars {
    double wArA;        // <wA, rA>   (PBiCG: <wA, rT>)
    double wArAold;
    double wApA;        // <wA, pA>   (PBiCG: <wA, pT>)
    double alpha;
    double beta;
    double normFactor;
    double initialResidual;
    double finalResidual;
    double sumPsi;      // normFactor: gAverage numerator
    double avgPsi;
    double tolerance;
    double relTol;
    double scaleNum, scaleDen, sf;   // GAMG scale
    int nIterations;
    int maxIter;
    int converged;
    int singular;
    int done;           // set on device: later launches of this solve are no-ops
    int histCount;
    int commError;
    int pad;
    double* hist;       // [kMaxHist] normalised residual after every iteration
};

constexpr int kRowBlock = 512;     // rows per block of the TMA-staged row kernel (amul.cu)
constexpr int kMaxRed = 4;         // values reduced together per kernel
constexpr int kMaxHist = 4096;     // residual history entries kept on device

// ---------------------------------------------------------------------------
// multi-GPU exchange window (peer-mapped, see comm.cu)
// ---------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
constexpr int kRedSlots = 8;       // doubles per reduction message

struct Comm {
    int rank = 0, nRanks = 1;
    bool connected = false;
    bool selfOnly = false;                        // window made by the library for cyclic-only matrices (comm.cu)
    unsigned char* window = nullptr;              // my window (device)
    size_t windowBytes = 0;
    int maxInterfaces = 0;                        // interface slots per rank
    long long slotStride = 0;                     // doubles per interface slot
    unsigned char* peer[kMaxRanks] = {nullptr};   // peer windows mapped here (peer[rank]==window)
    unsigned char** d_peer = nullptr;             // device copy of peer[]
};

}  // namespace ldu

// ---------------------------------------------------------------------------
// opaque ABI objects
// ---------------------------------------------------------------------------
struct ldu_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int smCount = 148;
    // deterministic reduction scratch
    double* d_partials = nullptr;     // [maxBlocks * kMaxRed]
    unsigned int* d_ticket = nullptr; // last-block election counter
    double* d_red = nullptr;          // [kMaxRed] final values of the last reduction
    int maxBlocks = 0;
    // pinned staging for scalar read-back
    ldu::SolverScalars* h_scalars = nullptr;
    ldu::Comm comm;
    // pinned staging ring for large copies from / to PAGEABLE host memory (what an application's
    // scalarFields are): kStageThreads workers x 2 chunks, each worker with its own stream
    void* stage = nullptr;            // ldu::StagePool, context.cu
};

namespace ldu {

struct Interface {
    int n = 0;
    int nbrRank = 0;
    int nbrInterface = 0;
    int offset = 0;        // start in the concatenated interface arrays
    int nbrOffset = 0;     // start of the matching segment in the neighbour's arrays
};

// Rows grouped by dependency depth for the order-dependent sweeps
// (DIC/DILU/FDIC forward/backward, Gauss-Seidel): rows of one level are
// independent; levels run in order.  Arithmetic per row is done in the
// reference's face order, so results are bit-identical to the sequential loops.
struct Schedule {
    int nLevels = 0;
    std::vector<int> levelStart;   // host, [nLevels+1]
    int* d_rows = nullptr;         // device, rows sorted by (level, row)
    int* d_levelStart = nullptr;   // device copy
    int maxLevelSize = 0;
};

struct GamgLevel;

}  // namespace ldu

struct ldu_matrix {
    ldu_context* ctx = nullptr;
    int nCells = 0, nFaces = 0;
    bool symmetric = true;
    bool haveCoeffs = false;
    bool diagonalOnly = false;     // lduMatrix::diagonal(): set_coeffs was given neither upper nor lower
    // addressing (device)
    int* d_l = nullptr;
    int* d_u = nullptr;
    int* d_ownerStart = nullptr;   // [nCells+1] CSR of faces by owner (upper part of a row)
    int* d_losortStart = nullptr;  // [nCells+1] CSR of losort by neighbour (lower part of a row)
    int* d_losort = nullptr;       // [nFaces] faces sorted by upper cell
    int* d_lowerCol = nullptr;     // [nFaces] l[losort[k]]: column of the k-th lower entry
    // k-th lower entry as ONE word: (column << 5) | position of its face among the faces the
    // column owns, i.e. face = ownerStart[column] + (word & 31).  Null when a cell owns more
    // than 32 faces or nCells >= 2^26 (then losort/lowerCol are used).
    int* d_lowerPacked = nullptr;
    // row blocks of the TMA-staged row kernel (amul.cu): per block of kRowBlock rows the
    // 16-byte aligned face / lower-entry ranges {f0a, nFa, k0a, nKa}; capacities of a stage
    void* d_rowBlocks = nullptr;
    int nRowBlocks = 0, rowFaceCap = 0, rowLowerCap = 0;
    // host copies (schedules, agglomeration)
    std::vector<int> h_l, h_u, h_ownerStart, h_losortStart, h_losort;
    // coefficients (device)
    double* d_diag = nullptr;
    double* d_upper = nullptr;
    double* d_lower = nullptr;     // == d_upper when symmetric
    bool ownLower = false;
    std::vector<double> h_faceWeights;
    // interfaces
    std::vector<ldu::Interface> ifs;
    int nIfFaces = 0;              // total coupled faces
    int* d_ifCells = nullptr;      // [nIfFaces] faceCells, interface-major
    double* d_bou = nullptr;       // [nIfFaces]
    double* d_int = nullptr;       // [nIfFaces]
    double* d_recv = nullptr;      // [nIfFaces] neighbour values (single rank: unused)
    void* d_ifTable = nullptr;     // per-interface {offset, n, nbrRank, nbrInterface} for the halo kernels
    // boundary rows: cells touched by interfaces, entries in reference order
    int nBRows = 0;
    int* d_bRowCell = nullptr;     // [nBRows]
    int* d_bRowStart = nullptr;    // [nBRows+1]
    int* d_bEntry = nullptr;       // [nIfFaces] index into the concatenated arrays
    // faces of every owner sorted by neighbour (upper-triangular order)?  If not (GAMG coarse levels),
    // the reverse-losort walk of DILU's preconditionT needs its own per-row order (sweeps.cu)
    bool nbrSorted = true;
    int* d_ownerByNbrDesc = nullptr;   // [nFaces] faces of each owner range by descending neighbour (lazy)
    int ifBlockStart = 0;          // first cell on any interface (nonBlockingGaussSeidelSmoother.C:66-81)
    int* d_cellBRow = nullptr;     // [nCells] boundary row of a cell or -1 (lazy, nonBlockingGaussSeidel only)
    // multiColourGaussSeidel (sweeps.cu): rows grouped by greedy colour (lazy)
    int* d_mcRows = nullptr;
    std::vector<int> mcStart;      // [nColours+1]
    // sweep schedules (lazy)
    ldu::Schedule fwd, bwd;
    bool haveSchedules = false;
    // dataflow sweeps (flow.cu): level-ordered chunk tables, published {value, epoch}
    // words and chunk counters
    void* d_ll = nullptr;
    void* flowDir[2] = {nullptr, nullptr};   // forward / backward tables (flow.cu)
    int flowEpoch = 0;
    // structured hex box detected from the addressing (nx, ny, nz; 0 = not a box) and
    // the line-pipelined sweep state (stencil.cu)
    int box[3] = {0, 0, 0};
    int boxDivOk = 0;         // box row kernel: 0 = not checked yet, 1 = usable, -1 = use the generic row kernel
    void* stencil = nullptr;
    void* stencil2 = nullptr;  // plane-stacked second-generation sweeps (stencil2.cu)
    long long coefGen = 0;    // bumped whenever diag/upper/lower change
    long long sweepGen = 0;   // bumped whenever a reciprocal diagonal is recomputed
    // work vectors owned by the matrix (allocated lazily, reused across solves)
    std::vector<double*> work;
    ldu::SolverScalars* d_scalars = nullptr;
    double* d_hist = nullptr;
    std::vector<double> lastHistory;
    // GAMG hierarchy
    std::vector<ldu::GamgLevel*> levels;
    bool haveIfCoeffs = false;        // bouCoeffs / intCoeffs of the coupled patches have been given
    bool hierarchyValid = false;
    bool externalHierarchy = false;   // levels handed over through ldu_gamg_set_level: never rebuilt by the library
    bool precondHierarchyReady = false;   // GAMG-as-preconditioner: coarse coefficients current
    ldu_controls hierarchyControls;
    bool isCoarse = false;
    int graphLaunches = 0;             // kernels in one captured GAMG cycle (launch accounting)
    bool referenceOrderSums = false;   // reductions accumulate in the reference's loop order
};

namespace ldu {

struct GamgLevel {
    int nFine = 0, nCoarse = 0, nFineFaces = 0, nCoarseFaces = 0;
    std::vector<int> h_restrict;      // [nFine] fine cell -> coarse cell
    std::vector<int> h_faceRestrict;  // [nFineFaces]
    int* d_restrict = nullptr;
    // gather form of restrictField (coarse cell -> its fine cells, ascending)
    int* d_cellStart = nullptr;       // [nCoarse+1]
    int* d_cellFine = nullptr;        // [nFine]
    // gather form of agglomerateMatrix
    int* d_faceStart = nullptr;       // [nCoarseFaces+1] coarse face -> fine faces (ascending)
    int* d_faceFine = nullptr;        // [#fine faces mapped to coarse faces]; sign bit = flipped
    int* d_intStart = nullptr;        // [nCoarse+1] coarse cell -> interior fine faces (ascending)
    int* d_intFine = nullptr;
    // coarse interface agglomeration
    std::vector<std::vector<int>> h_ifRestrict;  // per interface: fine if-face -> coarse if-face
    int* d_ifStart = nullptr;         // [coarse nIfFaces+1] gather lists over fine concatenated index
    int* d_ifFine = nullptr;
    ldu_matrix* coarse = nullptr;     // the coarse-level matrix
    double* d_corr = nullptr;         // coarseCorrFields[level]
    double* d_src = nullptr;          // coarseSources[level]
};

// ---- helpers implemented across the .cu files -----------------------------
double* work_vec(ldu_matrix* m, int idx);
// host <-> device copies of whole fields through the C ABI.  Page-locked host memory (ldu_host_alloc,
// cudaHostRegister) goes straight to the copy engine on the context's stream; pageable memory above
// 1 MB is cut into chunks that several host threads stage through pinned buffers, each on its own
// stream (a plain cudaMemcpy from pageable memory is one thread memcpy-ing into the driver's bounce
// buffer).  Both return with the copy COMPLETE and ordered after earlier work on the context's stream.
int copy_h2d(ldu_context* ctx, void* dst, const void* src, size_t bytes);
int copy_d2h(ldu_context* ctx, void* dst, const void* src, size_t bytes);
void stage_free(ldu_context* ctx);   // lazily allocated nCells-sized scratch
int ensure_scalars(ldu_matrix* m);
int greedy_colouring(int nCells, int nFaces, const int* lowerAddr, const int* upperAddr, std::vector<int>& colour,
                     int* nColours);
int build_schedules(ldu_matrix* m);

// kernels.cu
// guarded: the launch is a no-op once m->d_scalars->done is raised on device
int k_amul(ldu_matrix* m, double* Apsi, const double* psi, bool transpose, bool guarded = false);
int k_interfaces(ldu_matrix* m, double* result, const double* psi, int whichCoeffs, double sign,
                 bool guarded = false);
int k_sumA(ldu_matrix* m, double* sumA);
// PCG on a single-region box: Apsi = A psi fused with <Apsi, psi> and the alpha epilogue (guarded)
bool k_amul_dot_available(ldu_matrix* m);
int k_amul_dot(ldu_matrix* m, double* Apsi, const double* psi);
int k_H(ldu_matrix* m, double* Hpsi, const double* psi);        // lduMatrix::H
int k_H1(ldu_matrix* m, double* H1);                             // lduMatrix::H1
int k_faceH(ldu_matrix* m, double* faceHpsi, const double* psi);  // lduMatrix::faceH (face-sized output)
int k_offdiag(ldu_matrix* m, double* y, const double* x);   // (A - diag) x, GAMGSolverInterpolate.C:50-76
int k_residual(ldu_matrix* m, double* rA, const double* psi, const double* source, bool guarded = false);

// solvers.cu
int solve_device(ldu_matrix* m, const ldu_controls* c, double* d_psi, const double* d_source,
                 ldu_solver_performance* perf);
int precondition_device(ldu_matrix* m, int precond, double* d_wA, const double* d_rA, int transpose);
int smooth_device(ldu_matrix* m, int smoother, double* d_psi, const double* d_source, int nSweeps);

// gamg.cu
int gamg_build(ldu_matrix* m, const ldu_controls* c);
void gamg_free(ldu_matrix* m);
int gamg_solve(ldu_matrix* m, const ldu_controls* c, double* d_psi, const double* d_source,
               ldu_solver_performance* perf);

// comm.cu
int comm_allreduce(ldu_context* ctx, double* d_vals, int n);          // in-place sum over ranks
int comm_halo_put(ldu_matrix* m, const double* d_psi, bool guarded);       // psi[faceCells] -> neighbours' windows
int comm_halo_recv(ldu_matrix* m, bool guarded);                           // wait, fills m->d_recv
int comm_halo_exchange(ldu_matrix* m, const double* d_psi, bool guarded);  // put + recv

}  // namespace ldu
