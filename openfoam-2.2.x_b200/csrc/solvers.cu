// Device-resident Krylov and smooth solvers.
//
//   PCG::solve           matrices/lduMatrix/solvers/PCG/PCG.C:65-182
//   PBiCG::solve         matrices/lduMatrix/solvers/PBiCG/PBiCG.C:65-198
//   smoothSolver::solve  matrices/lduMatrix/solvers/smoothSolver/smoothSolver.C:77-180
//   diagonalSolver       matrices/lduMatrix/solvers/diagonalSolver/diagonalSolver.C:62-81
//   solver::normFactor   matrices/lduMatrix/lduMatrix/lduMatrixSolver.C:179-197
//   SolverPerformance    matrices/LduMatrix/LduMatrix/SolverPerformance.C:31-91
//
// The whole do/while loop of the reference lives on the device: the scalars the
// reference keeps in host locals (wArA, wApA, alpha, beta, residual, iteration
// count, converged/singular) sit in a SolverScalars block in HBM and are updated
// by the epilogue of the reducing kernels.  Once the device decides the loop has
// ended it raises `done`; kernels enqueued after that point return immediately,
// so the host can enqueue iterations in batches and poll rarely while the
// iteration count and the iterates stay exactly the reference's.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "epilogue.cuh"
#include "reduce.cuh"
#include "sweeps.h"

namespace ldu {

struct EpiAvgPsi {  // gAverage(psi): FieldFunctions.C:514-533
    __device__ void operator()(SolverScalars* S, const double* t) const
    {
        S->sumPsi = t[0];
        S->avgPsi = __ddiv_rn(t[0], t[1]);
    }
};

struct EpiNormFactor {  // lduMatrixSolver.C:191-193 + PCG.C:108-112
    __device__ void operator()(SolverScalars* S, const double* t) const
    {
        S->normFactor = __dadd_rn(t[0], kSmall);
        S->initialResidual = __ddiv_rn(t[1], S->normFactor);
        S->finalResidual = S->initialResidual;
        push_history(S);
        if (check_convergence(S)) S->done = 1;
    }
};

// ---------------------------------------------------------------------------
// element-wise maps
// ---------------------------------------------------------------------------
struct InitResidualMap {  // rA = source - wA (PCG.C:95); sum(psi), count
    double* rA;
    const double* source;
    const double* wA;
    const double* psi;
    double* rT;           // PBiCG: rT = source - wT
    const double* wT;
    __device__ void operator()(int i, double (&acc)[2]) const
    {
        rA[i] = __dsub_rn(source[i], wA[i]);
        if (rT) rT[i] = __dsub_rn(source[i], wT[i]);
        acc[0] = __dadd_rn(acc[0], psi[i]);
        acc[1] = __dadd_rn(acc[1], 1.0);
    }
};

struct NormFactorMap {  // lduMatrixSolver.C:187-193
    const SolverScalars* S;
    const double* sumA;
    const double* Apsi;
    const double* source;
    const double* rA;
    __device__ void operator()(int i, double (&acc)[2]) const
    {
        const double t = __dmul_rn(sumA[i], S->avgPsi);
        acc[0] = __dadd_rn(acc[0], __dadd_rn(fabs(__dsub_rn(Apsi[i], t)), fabs(__dsub_rn(source[i], t))));
        acc[1] = __dadd_rn(acc[1], fabs(rA[i]));
    }
};

struct DotMap {
    const double* a;
    const double* b;
    __device__ void operator()(int i, double (&acc)[1]) const
    {
        acc[0] = __dadd_rn(acc[0], __dmul_rn(a[i], b[i]));
    }
};

struct CopyDotMap {  // noPreconditioner.C:58-74 fused with <wA, rA>
    double* wA;
    const double* rA;
    const double* rT;  // dot partner (rA for PCG, rT for PBiCG)
    __device__ void operator()(int i, double (&acc)[1]) const
    {
        const double w = rA[i];
        wA[i] = w;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(w, rT[i]));
    }
};

struct DiagPrecondDotMap {  // diagonalPreconditioner.C:70-87 fused with <wA, rA>
    double* wA;
    const double* rD;
    const double* rA;
    const double* rT;
    __device__ void operator()(int i, double (&acc)[1]) const
    {
        const double w = __dmul_rn(rD[i], rA[i]);
        wA[i] = w;
        acc[0] = __dadd_rn(acc[0], __dmul_rn(w, rT[i]));
    }
};

struct ScaleMap {  // wT = rD*rT
    double* w;
    const double* rD;
    const double* r;
    __device__ void operator()(int i) const { w[i] = rD ? __dmul_rn(rD[i], r[i]) : r[i]; }
};

struct PUpdateMap {  // PCG.C:134-149
    const SolverScalars* S;
    double* pA;
    const double* wA;
    double* pT;
    const double* wT;
    __device__ void operator()(int i) const
    {
        if (S->nIterations == 0) {
            pA[i] = wA[i];
            if (pT) pT[i] = wT[i];
        } else {
            const double beta = S->beta;
            pA[i] = __dadd_rn(wA[i], __dmul_rn(beta, pA[i]));
            if (pT) pT[i] = __dadd_rn(wT[i], __dmul_rn(beta, pT[i]));
        }
    }
};

struct XRUpdateMap {  // PCG.C:166-172 fused with gSumMag(rA)
    const SolverScalars* S;
    double* psi;
    double* rA;
    const double* pA;
    const double* wA;
    double* rT;
    const double* wT;
    __device__ void operator()(int i, double (&acc)[1]) const
    {
        const double alpha = S->alpha;
        psi[i] = __dadd_rn(psi[i], __dmul_rn(alpha, pA[i]));
        const double r = __dsub_rn(rA[i], __dmul_rn(alpha, wA[i]));
        rA[i] = r;
        if (rT) rT[i] = __dsub_rn(rT[i], __dmul_rn(alpha, wT[i]));
        acc[0] = __dadd_rn(acc[0], fabs(r));
    }
};

struct SumMagMap {
    const double* a;
    __device__ void operator()(int i, double (&acc)[1]) const { acc[0] = __dadd_rn(acc[0], fabs(a[i])); }
};

struct MulMap {  // rA *= rD
    double* a;
    const double* b;
    __device__ void operator()(int i) const { a[i] = __dmul_rn(a[i], b[i]); }
};

struct AddMap {  // psi += rA
    double* a;
    const double* b;
    __device__ void operator()(int i) const { a[i] = __dadd_rn(a[i], b[i]); }
};

struct CopyMap {
    double* a;
    const double* b;
    __device__ void operator()(int i) const { a[i] = b[i]; }
};

struct DivMap {  // diagonalSolver.C:62-81
    double* psi;
    const double* source;
    const double* diag;
    __device__ void operator()(int i) const { psi[i] = __ddiv_rn(source[i], diag[i]); }
};

__global__ void init_scalars_kernel(SolverScalars* S, double tol, double relTol, int maxIter, double* hist)
{
    S->wArA = kGreat;
    S->wArAold = kGreat;
    S->wApA = 0;
    S->alpha = 0;
    S->beta = 0;
    S->normFactor = 0;
    S->initialResidual = 0;
    S->finalResidual = 0;
    S->sumPsi = 0;
    S->avgPsi = 0;
    S->tolerance = tol;
    S->relTol = relTol;
    S->nIterations = 0;
    S->maxIter = maxIter;
    S->converged = 0;
    S->singular = 0;
    S->done = 0;
    S->histCount = 0;
    S->commError = 0;
    S->hist = hist;
}

int init_scalars(ldu_matrix* m, const ldu_controls* c)
{
    LDU_TRY(ensure_scalars(m));
    stencil2_invalidate(m);
    init_scalars_kernel<<<1, 1, 0, m->ctx->stream>>>(m->d_scalars, c->tolerance, c->relTol, c->maxIter, m->d_hist);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

int read_scalars(ldu_matrix* m, SolverScalars* out)
{
    ldu_context* ctx = m->ctx;
    LDU_CUDA(cudaMemcpyAsync(ctx->h_scalars, m->d_scalars, sizeof(SolverScalars), cudaMemcpyDeviceToHost,
                             ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = *ctx->h_scalars;
    if (out->commError) {
        set_error("multi-GPU exchange timed out (peer did not arrive)");
        return LDU_ECOMM;
    }
    return LDU_OK;
}

int fetch_performance(ldu_matrix* m, ldu_solver_performance* perf)
{
    SolverScalars s;
    LDU_TRY(read_scalars(m, &s));
    if (perf) {
        perf->initialResidual = s.initialResidual;
        perf->finalResidual = s.finalResidual;
        perf->nIterations = s.nIterations;
        perf->converged = s.converged;
        perf->singular = s.singular;
    }
    const int nh = s.histCount < kMaxHist ? s.histCount : kMaxHist;
    m->lastHistory.resize(nh);
    if (nh)
        LDU_CUDA(cudaMemcpy(m->lastHistory.data(), m->d_hist, nh * sizeof(double), cudaMemcpyDeviceToHost));
    return LDU_OK;
}

int solve_prologue(ldu_matrix* m, double* psi, const double* source, double* wA, double* rA, double* tmp)
{
    const int n = m->nCells;
    LDU_TRY(k_amul(m, wA, psi, false));
    LDU_TRY((launch_map_reduce<2, false>(m, n, InitResidualMap{rA, source, wA, psi, nullptr, nullptr}, EpiAvgPsi())));
    LDU_TRY(k_sumA(m, tmp));
    LDU_TRY((launch_map_reduce<2, false>(m, n, NormFactorMap{m->d_scalars, tmp, wA, source, rA}, EpiNormFactor())));
    return LDU_OK;
}

// ---------------------------------------------------------------------------
// preconditioners
// ---------------------------------------------------------------------------
// run-time selection tables of the reference: DIC/FDIC are registered for
// symmetric matrices only, DILU for asymmetric only (DICPreconditioner.C:35,
// FDICPreconditioner.C:35, DILUPreconditioner.C:35; smoothers likewise)
static int check_tables(const ldu_matrix* m, bool symOnly, bool asymOnly, const char* what, const char* name)
{
    if ((symOnly && !m->symmetric) || (asymOnly && m->symmetric)) {
        set_error(std::string("Unknown ") + (m->symmetric ? "symmetric" : "asymmetric") + " matrix " + what
                  + " " + name);
        return LDU_EINVAL;
    }
    return LDU_OK;
}

int precond_setup(ldu_matrix* m, int kind, Precond& p, int rDSlot)
{
    p = Precond();
    p.kind = kind;
    if (kind == LDU_PRECOND_DIC) LDU_TRY(check_tables(m, true, false, "preconditioner", "DIC"));
    if (kind == LDU_PRECOND_FDIC) LDU_TRY(check_tables(m, true, false, "preconditioner", "FDIC"));
    if (kind == LDU_PRECOND_DILU) LDU_TRY(check_tables(m, false, true, "preconditioner", "DILU"));
    switch (kind) {
    case LDU_PRECOND_NONE:
        return LDU_OK;
    case LDU_PRECOND_DIAGONAL:
        p.rD = work_vec(m, rDSlot);
        return calc_reciprocal_diag(m, p.rD);
    case LDU_PRECOND_DIC:
        p.rD = work_vec(m, rDSlot);
        return calc_reciprocal_D(m, p.rD, false);
    case LDU_PRECOND_DILU:
        p.rD = work_vec(m, rDSlot);
        return calc_reciprocal_D(m, p.rD, true);
    case LDU_PRECOND_FDIC:
        p.rD = work_vec(m, rDSlot);
        LDU_TRY(calc_reciprocal_D(m, p.rD, false));
        LDU_CUDA(cudaMalloc((void**)&p.rDuUpper, (m->nFaces > 0 ? m->nFaces : 1) * sizeof(double)));
        LDU_CUDA(cudaMalloc((void**)&p.rDlUpper, (m->nFaces > 0 ? m->nFaces : 1) * sizeof(double)));
        return calc_fdic_coeffs(m, p.rD, p.rDuUpper, p.rDlUpper);
    default:
        set_error("unknown preconditioner");
        return LDU_EINVAL;
    }
}

void precond_release(Precond& p)
{
    cudaFree(p.rDuUpper);
    cudaFree(p.rDlUpper);
    p.rDuUpper = p.rDlUpper = nullptr;
}

int precond_apply(ldu_matrix* m, const Precond& p, double* wA, const double* rA, bool transpose)
{
    switch (p.kind) {
    case LDU_PRECOND_NONE:
        return launch_map<true>(m, m->nCells, ScaleMap{wA, nullptr, rA});
    case LDU_PRECOND_DIAGONAL:
        return launch_map<true>(m, m->nCells, ScaleMap{wA, p.rD, rA});
    case LDU_PRECOND_DIC:  // DICPreconditioner.C:87-123
        return sweep_pair(m, p.rD, m->d_upper, m->d_upper, false, rA, wA, true);
    case LDU_PRECOND_FDIC:  // FDICPreconditioner.C:88-125
        return sweep_pair(m, p.rD, p.rDuUpper, p.rDlUpper, true, rA, wA, true);
    case LDU_PRECOND_DILU:
        if (!transpose)  // DILUPreconditioner.C:88-135
            return sweep_pair(m, p.rD, m->d_lower, m->d_upper, false, rA, wA, true);
        // preconditionT: DILUPreconditioner.C:138-185.  Its second loop walks the faces in reverse
        // LOSORT order: on a matrix whose owner ranges are not sorted by neighbour (GAMG coarse
        // levels) that is not the reverse face order of the other backward sweeps
        if (!m->nbrSorted) {
            LDU_TRY(sweep_forward(m, p.rD, m->d_upper, false, rA, wA, true));
            return sweep_backward_losort(m, p.rD, m->d_lower, wA);
        }
        return sweep_pair(m, p.rD, m->d_upper, m->d_lower, false, rA, wA, true);
    }
    return LDU_EINVAL;
}

// ---------------------------------------------------------------------------
// smoothers
// ---------------------------------------------------------------------------
int smoother_setup(ldu_matrix* m, int kind, Smoother& s)
{
    s = Smoother();
    s.kind = kind;
    switch (kind) {
    case LDU_SMOOTHER_GS:
    case LDU_SMOOTHER_SYMGS:
    case LDU_SMOOTHER_NBGS:
    case LDU_SMOOTHER_MCGS:
        return LDU_OK;
    case LDU_SMOOTHER_DIC:
    case LDU_SMOOTHER_DICGS:
        LDU_TRY(check_tables(m, true, false, "smoother", kind == LDU_SMOOTHER_DIC ? "DIC" : "DICGaussSeidel"));
        return precond_setup(m, LDU_PRECOND_DIC, s.dic, W_SRD);
    case LDU_SMOOTHER_DILU:
    case LDU_SMOOTHER_DILUGS:
        LDU_TRY(check_tables(m, false, true, "smoother", kind == LDU_SMOOTHER_DILU ? "DILU" : "DILUGaussSeidel"));
        return precond_setup(m, LDU_PRECOND_DILU, s.dic, W_SRD);
    case LDU_SMOOTHER_FDIC:
        LDU_TRY(check_tables(m, true, false, "smoother", "FDIC"));
        return precond_setup(m, LDU_PRECOND_FDIC, s.dic, W_SRD);
    }
    set_error("unknown smoother");
    return LDU_EINVAL;
}

void smoother_release(Smoother& s) { precond_release(s.dic); }

// GaussSeidelSmoother.C:66-187 / symGaussSeidelSmoother.C:66-216
static int gs_apply(ldu_matrix* m, double* psi, const double* source, int nSweeps, bool sym)
{
    for (int sweep = 0; sweep < nSweeps; sweep++) {
        const double* bPrime = source;
        if (m->nIfFaces) {
            // bPrime = source; bPrime[fc] -= (-bouCoeffs)*psiNbr  (GaussSeidelSmoother.C:110-145)
            double* bp = work_vec(m, W_BPRIME);
            LDU_TRY(launch_map<true>(m, m->nCells, CopyMap{bp, source}));
            LDU_TRY(k_interfaces(m, bp, psi, 0, -1.0, true));
            bPrime = bp;
        }
        LDU_TRY(gs_sweep(m, bPrime, sym ? work_vec(m, W_BLOWER) : nullptr, psi, sym));
    }
    return LDU_OK;
}

// multiColourGaussSeidel (extension, include/ldu_b200.h LDU_SMOOTHER_MCGS): the boundary treatment of GaussSeidel
// (bPrime = source + Jacobi-coupled interface terms), the cells colour by colour
static int mcgs_apply(ldu_matrix* m, double* psi, const double* source, int nSweeps)
{
    for (int sweep = 0; sweep < nSweeps; sweep++) {
        const double* bPrime = source;
        if (m->nIfFaces) {
            double* bp = work_vec(m, W_BPRIME);
            LDU_TRY(launch_map<true>(m, m->nCells, CopyMap{bp, source}));
            LDU_TRY(k_interfaces(m, bp, psi, 0, -1.0, true));
            bPrime = bp;
        }
        LDU_TRY(mcgs_sweep(m, bPrime, psi));
    }
    return LDU_OK;
}

// nonBlockingGaussSeidelSmoother.C:128-217.  Without interfaces it is GaussSeidel; with them the
// interface terms enter a coupled row between the lower entries from cells below blockStart and
// the rest (a different rounding order), see nbgs_level_kernel.
static int nbgs_apply(ldu_matrix* m, double* psi, const double* source, int nSweeps)
{
    if (!m->nIfFaces) return gs_apply(m, psi, source, nSweeps, false);
    for (int sweep = 0; sweep < nSweeps; sweep++) {
        LDU_TRY(comm_halo_put(m, psi, true));
        LDU_TRY(comm_halo_recv(m, true));
        LDU_TRY(nbgs_sweep(m, source, psi));
    }
    return LDU_OK;
}

// DICSmoother.C:67-116, DILUSmoother.C:67-119, FDICSmoother.C:98-146
static int dic_apply(ldu_matrix* m, const Smoother& s, double* psi, const double* source, int nSweeps)
{
    double* rA = work_vec(m, W_STMP);
    const Precond& p = s.dic;
    for (int sweep = 0; sweep < nSweeps; sweep++) {
        LDU_TRY(k_residual(m, rA, psi, source, true));
        LDU_TRY(launch_map<true>(m, m->nCells, MulMap{rA, p.rD}));
        if (p.kind == LDU_PRECOND_FDIC)
            LDU_TRY(sweep_pair(m, p.rD, p.rDuUpper, p.rDlUpper, true, nullptr, rA, false));
        else if (p.kind == LDU_PRECOND_DILU)
            LDU_TRY(sweep_pair(m, p.rD, m->d_lower, m->d_upper, false, nullptr, rA, false));
        else
            LDU_TRY(sweep_pair(m, p.rD, m->d_upper, m->d_upper, false, nullptr, rA, false));
        LDU_TRY(launch_map<true>(m, m->nCells, AddMap{psi, rA}));
    }
    return LDU_OK;
}

int smoother_apply(ldu_matrix* m, const Smoother& s, double* psi, const double* source, int nSweeps)
{
    switch (s.kind) {
    case LDU_SMOOTHER_GS:
        return gs_apply(m, psi, source, nSweeps, false);
    case LDU_SMOOTHER_NBGS:
        return nbgs_apply(m, psi, source, nSweeps);
    case LDU_SMOOTHER_MCGS:
        return mcgs_apply(m, psi, source, nSweeps);
    case LDU_SMOOTHER_SYMGS:
        return gs_apply(m, psi, source, nSweeps, true);
    case LDU_SMOOTHER_DIC:
    case LDU_SMOOTHER_DILU:
    case LDU_SMOOTHER_FDIC:
        return dic_apply(m, s, psi, source, nSweeps);
    case LDU_SMOOTHER_DICGS:   // DICGaussSeidelSmoother.C:79-89
    case LDU_SMOOTHER_DILUGS:
        LDU_TRY(dic_apply(m, s, psi, source, nSweeps));
        return gs_apply(m, psi, source, nSweeps, false);
    }
    return LDU_EINVAL;
}

// ---------------------------------------------------------------------------
// PCG / PBiCG
// ---------------------------------------------------------------------------
int gamg_precondition(ldu_matrix* m, const ldu_controls* c, double* wA, const double* rA);

static int krylov_solve(ldu_matrix* m, const ldu_controls* c, double* psi, const double* source, bool bicg)
{
    const int n = m->nCells;
    double* pA = work_vec(m, W_PA);
    double* wA = work_vec(m, W_WA);
    double* rA = work_vec(m, W_RA);
    double* pT = bicg ? work_vec(m, W_PT) : nullptr;
    double* wT = bicg ? work_vec(m, W_WT) : nullptr;
    double* rT = bicg ? work_vec(m, W_RT) : nullptr;
    if (!pA || !wA || !rA || (bicg && (!pT || !wT || !rT))) {
        set_error("out of device memory for solver work fields");
        return LDU_ECUDA;
    }
    LDU_TRY(init_scalars(m, c));
    m->precondHierarchyReady = false;
    // prologue (PCG.C:88-112 / PBiCG.C:97-121)
    LDU_TRY(k_amul(m, wA, psi, false));
    if (bicg) LDU_TRY(k_amul(m, wT, psi, true));
    LDU_TRY((launch_map_reduce<2, false>(m, n, InitResidualMap{rA, source, wA, psi, rT, wT}, EpiAvgPsi())));
    LDU_TRY(k_sumA(m, pA));
    LDU_TRY((launch_map_reduce<2, false>(m, n, NormFactorMap{m->d_scalars, pA, wA, source, rA}, EpiNormFactor())));

    SolverScalars hs;
    LDU_TRY(read_scalars(m, &hs));
    if (hs.done) return LDU_OK;

    const bool useGamg = (c->preconditioner == LDU_PRECOND_GAMG);
    if (bicg && useGamg) {
        // GAMGPreconditioner has no preconditionT: the reference stops in the first PBiCG iteration with
        // "Not implemented" (lduMatrix.H:492-505 called from PBiCG.C:139)
        set_error("PBiCG with preconditioner GAMG: GAMG::preconditionT is not implemented (as in the reference)");
        return LDU_EINVAL;
    }
    Precond pre;
    if (!useGamg) LDU_TRY(precond_setup(m, c->preconditioner, pre, W_RD));
    const bool cheap = (c->preconditioner == LDU_PRECOND_NONE || c->preconditioner == LDU_PRECOND_DIAGONAL);
    // the fused kernels sum in tile order: not for the reference-order verification mode
    static const bool fuseOff = getenv("LDU_PCG_FUSE") && getenv("LDU_PCG_FUSE")[0] == '0';
    const bool amulDot = !bicg && !fuseOff && !m->referenceOrderSums && k_amul_dot_available(m);
    const bool fusedBox = !bicg && !useGamg && !fuseOff && !m->referenceOrderSums && stencil_version(m) == 2
                          && (pre.kind == LDU_PRECOND_DIC || pre.kind == LDU_PRECOND_FDIC);
    int interval = c->checkInterval > 0 ? c->checkInterval : (cheap ? 32 : (useGamg ? 1 : 8));
    int enqueued = 0;   // never enqueue more than the maxIter+1 iterations the loop can run
    const double* dotPartner = bicg ? rT : rA;

    int rc = LDU_OK;
    for (;;) {
        for (int it = 0; it < interval && rc == LDU_OK && enqueued <= c->maxIter; it++, enqueued++) {
            // wA = M^-1 rA ; wArA = <wA, rA>        (PCG.C:129-132)
            if (useGamg) {
                rc = gamg_precondition(m, c, wA, rA);
                if (rc == LDU_OK) rc = launch_map_reduce<1, true>(m, n, DotMap{wA, dotPartner}, EpiWArA());
            } else if (pre.kind == LDU_PRECOND_NONE) {
                rc = launch_map_reduce<1, true>(m, n, CopyDotMap{wA, rA, dotPartner}, EpiWArA());
                if (rc == LDU_OK && bicg) rc = precond_apply(m, pre, wT, rT, true);
            } else if (pre.kind == LDU_PRECOND_DIAGONAL) {
                rc = launch_map_reduce<1, true>(m, n, DiagPrecondDotMap{wA, pre.rD, rA, dotPartner}, EpiWArA());
                if (rc == LDU_OK && bicg) rc = precond_apply(m, pre, wT, rT, true);
            } else if (fusedBox) {
                // blockMesh box, DIC / FDIC: both substitutions, the way back from the tile layout and
                // <wA, rA> in one pass (stencil2.cu)
                rc = stencil2_apply_dot(m, pre.rD, m->d_upper, m->d_upper, rA, wA, rA);
            } else {
                rc = precond_apply(m, pre, wA, rA, false);
                if (rc == LDU_OK && bicg) rc = precond_apply(m, pre, wT, rT, true);
                if (rc == LDU_OK) rc = launch_map_reduce<1, true>(m, n, DotMap{wA, dotPartner}, EpiWArA());
            }
            if (rc != LDU_OK) break;
            // pA = wA + beta pA                      (PCG.C:134-149)
            rc = launch_map<true>(m, n, PUpdateMap{m->d_scalars, pA, wA, pT, wT});
            if (rc != LDU_OK) break;
            // wA = A pA ; wApA = <wA, pA>            (PCG.C:153-155)
            if (amulDot) {
                rc = k_amul_dot(m, wA, pA);   // box, one region: wA = A pA and <wA, pA> in one pass
            } else {
                rc = k_amul(m, wA, pA, false, true);
                if (rc == LDU_OK && bicg) rc = k_amul(m, wT, pT, true, true);
                if (rc == LDU_OK) rc = launch_map_reduce<1, true>(m, n, DotMap{wA, bicg ? pT : pA}, EpiWApA());
            }
            if (rc != LDU_OK) break;
            // psi += alpha pA ; rA -= alpha wA ; residual ; loop test   (PCG.C:166-178)
            if (fusedBox)   // ... and the next iteration's rD*rA goes to the tile layout in the same pass
                rc = stencil2_xr_pack(m, pre.rD, psi, rA, pA, wA);
            else
                rc = launch_map_reduce<1, true>(m, n, XRUpdateMap{m->d_scalars, psi, rA, pA, wA, rT, wT},
                                                EpiResidual<true>{1});
        }
        if (rc != LDU_OK) break;
        rc = read_scalars(m, &hs);
        if (rc != LDU_OK || hs.done) break;
    }
    precond_release(pre);
    return rc;
}

// ---------------------------------------------------------------------------
// smoothSolver / diagonalSolver
// ---------------------------------------------------------------------------
static int smooth_solve(ldu_matrix* m, const ldu_controls* c, double* psi, const double* source)
{
    const int n = m->nCells;
    LDU_TRY(init_scalars(m, c));
    Smoother sm;
    if (c->nSweeps < 0) {  // smoothSolver.C:91-111: fixed number of sweeps, no residual
        LDU_TRY(smoother_setup(m, c->smoother, sm));
        int rc = smoother_apply(m, sm, psi, source, -c->nSweeps);
        smoother_release(sm);
        LDU_TRY(rc);
        SolverScalars* S = m->d_scalars;
        const int inc = -c->nSweeps;
        LDU_CUDA(cudaMemcpyAsync(&S->nIterations, &inc, sizeof(int), cudaMemcpyHostToDevice, m->ctx->stream));
        LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
        return LDU_OK;
    }
    double* Apsi = work_vec(m, W_APSI);
    double* res = work_vec(m, W_RES);
    double* tmp = work_vec(m, W_TMP);
    LDU_TRY(solve_prologue(m, psi, source, Apsi, res, tmp));
    SolverScalars hs;
    LDU_TRY(read_scalars(m, &hs));
    if (hs.done) return LDU_OK;
    LDU_TRY(smoother_setup(m, c->smoother, sm));
    const int interval = c->checkInterval > 0 ? c->checkInterval : 2;
    int rc = LDU_OK;
    for (;;) {
        for (int it = 0; it < interval && rc == LDU_OK; it++) {
            rc = smoother_apply(m, sm, psi, source, c->nSweeps);
            if (rc == LDU_OK) rc = k_residual(m, res, psi, source, true);
            if (rc == LDU_OK)
                rc = launch_map_reduce<1, true>(m, n, SumMagMap{res}, EpiResidual<false>{c->nSweeps});
        }
        if (rc != LDU_OK) break;
        rc = read_scalars(m, &hs);
        if (rc != LDU_OK || hs.done) break;
    }
    smoother_release(sm);
    return rc;
}

static int diagonal_solve(ldu_matrix* m, const ldu_controls* c, double* psi, const double* source)
{
    LDU_TRY(init_scalars(m, c));
    LDU_TRY(launch_map<false>(m, m->nCells, DivMap{psi, source, m->d_diag}));
    // diagonalSolver.C:71-80: SolverPerformance(typeName, fieldName, 0, 0, 0, true, false)
    const int one = 1;
    LDU_CUDA(cudaMemcpyAsync(&m->d_scalars->converged, &one, sizeof(int), cudaMemcpyHostToDevice, m->ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return LDU_OK;
}

// lduMatrix::solver::New dispatch: lduMatrixSolver.C:40-136
int solve_device(ldu_matrix* m, const ldu_controls* c, double* d_psi, const double* d_source,
                 ldu_solver_performance* perf)
{
    if (!m || !c || !d_psi || !d_source) {
        set_error("ldu_solve: bad argument");
        return LDU_EINVAL;
    }
    if (!m->haveCoeffs) {
        set_error("ldu_solve: coefficients not set");
        return LDU_EINVAL;
    }
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    m->referenceOrderSums = (c->referenceOrderSums != 0);
    int rc;
    // matrix.diagonal() -> diagonalSolver whatever the dictionary says (lduMatrixSolver.C:52-66).
    // diagonal() means "no upper and no lower field" (lduMatrix.H:547-550), not "no faces": a faceless
    // mesh whose upper() exists (fvm::laplacian always touches it) goes to the selected solver
    if (c->solver == LDU_SOLVER_DIAGONAL || (m->diagonalOnly && m->ctx->comm.nRanks == 1)) {
        rc = diagonal_solve(m, c, d_psi, d_source);
    } else {
        switch (c->solver) {
        case LDU_SOLVER_PCG:
            if (!m->symmetric) {
                set_error("PCG selected for an asymmetric matrix (not in the symMatrix table)");
                return LDU_EINVAL;
            }
            rc = krylov_solve(m, c, d_psi, d_source, false);
            break;
        case LDU_SOLVER_PBICG:
            if (m->symmetric) {
                set_error("PBiCG selected for a symmetric matrix (not in the asymMatrix table)");
                return LDU_EINVAL;
            }
            rc = krylov_solve(m, c, d_psi, d_source, true);
            break;
        case LDU_SOLVER_SMOOTH:
            rc = smooth_solve(m, c, d_psi, d_source);
            break;
        case LDU_SOLVER_GAMG:
            rc = gamg_solve(m, c, d_psi, d_source, perf);
            break;
        default:
            set_error("unknown solver");
            return LDU_EINVAL;
        }
    }
    LDU_TRY(rc);
    return fetch_performance(m, perf);
}

int precondition_device(ldu_matrix* m, int precond, double* d_wA, const double* d_rA, int transpose)
{
    ldu_controls c;
    ldu_controls_default(&c);
    LDU_TRY(init_scalars(m, &c));
    Precond p;
    LDU_TRY(precond_setup(m, precond, p, W_RD));
    int rc = precond_apply(m, p, d_wA, d_rA, transpose != 0);
    if (rc == LDU_OK) rc = cudaStreamSynchronize(m->ctx->stream) == cudaSuccess ? LDU_OK : LDU_ECUDA;
    precond_release(p);
    return rc;
}

int smooth_device(ldu_matrix* m, int smoother, double* d_psi, const double* d_source, int nSweeps)
{
    ldu_controls c;
    ldu_controls_default(&c);
    LDU_TRY(init_scalars(m, &c));
    Smoother s;
    LDU_TRY(smoother_setup(m, smoother, s));
    int rc = smoother_apply(m, s, d_psi, d_source, nSweeps);
    if (rc == LDU_OK) rc = cudaStreamSynchronize(m->ctx->stream) == cudaSuccess ? LDU_OK : LDU_ECUDA;
    smoother_release(s);
    return rc;
}

}  // namespace ldu

// ---------------------------------------------------------------------------
// C ABI: host-pointer and device-pointer entry points
// ---------------------------------------------------------------------------
using namespace ldu;

namespace {

struct Staged {  // device copies of host fields for one call
    ldu_matrix* m;
    std::vector<double*> bufs;
    double* in(int slot, const double* h)
    {
        double* d = work_vec(m, slot);
        if (d && h && copy_h2d(m->ctx, d, h, (size_t)m->nCells * sizeof(double)) != LDU_OK) return nullptr;
        return d;
    }
    int out(double* h, const double* d) { return copy_d2h(m->ctx, h, d, (size_t)m->nCells * sizeof(double)); }
};

int check_matrix(ldu_matrix* m, const char* who)
{
    if (!m) {
        set_error(std::string(who) + ": null matrix");
        return LDU_EINVAL;
    }
    if (!m->haveCoeffs) {
        set_error(std::string(who) + ": coefficients not set");
        return LDU_EINVAL;
    }
    if (cudaSetDevice(m->ctx->device) != cudaSuccess) return LDU_ECUDA;
    return LDU_OK;
}

}  // namespace

extern "C" {

int ldu_amul(ldu_matrix* m, double* Apsi, const double* psi)
{
    LDU_TRY(check_matrix(m, "ldu_amul"));
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* dy = work_vec(m, W_OUT);
    LDU_TRY(k_amul(m, dy, dx, false));
    return s.out(Apsi, dy);
}

int ldu_tmul(ldu_matrix* m, double* Tpsi, const double* psi)
{
    LDU_TRY(check_matrix(m, "ldu_tmul"));
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* dy = work_vec(m, W_OUT);
    LDU_TRY(k_amul(m, dy, dx, true));
    return s.out(Tpsi, dy);
}

int ldu_sumA(ldu_matrix* m, double* sumA)
{
    LDU_TRY(check_matrix(m, "ldu_sumA"));
    Staged s{m};
    double* dy = work_vec(m, W_OUT);
    LDU_TRY(k_sumA(m, dy));
    return s.out(sumA, dy);
}

int ldu_residual(ldu_matrix* m, double* rA, const double* psi, const double* source)
{
    LDU_TRY(check_matrix(m, "ldu_residual"));
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* db = s.in(W_SRC, source);
    double* dy = work_vec(m, W_OUT);
    LDU_TRY(k_residual(m, dy, dx, db));
    return s.out(rA, dy);
}

int ldu_H(ldu_matrix* m, double* Hpsi, const double* psi)
{
    LDU_TRY(check_matrix(m, "ldu_H"));
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* dy = work_vec(m, W_OUT);
    LDU_TRY(k_H(m, dy, dx));
    return s.out(Hpsi, dy);
}

int ldu_H1(ldu_matrix* m, double* H1)
{
    LDU_TRY(check_matrix(m, "ldu_H1"));
    Staged s{m};
    double* dy = work_vec(m, W_OUT);
    LDU_TRY(k_H1(m, dy));
    return s.out(H1, dy);
}

int ldu_faceH(ldu_matrix* m, double* faceHpsi, const double* psi)
{
    LDU_TRY(check_matrix(m, "ldu_faceH"));
    if (m->nFaces <= 0) {   // the reference aborts: "the matrix does not have any off-diagonal coefficients"
        set_error("ldu_faceH: the matrix does not have any off-diagonal coefficients");
        return LDU_EINVAL;
    }
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* dy = nullptr;
    cudaStream_t st = m->ctx->stream;
    LDU_CUDA(cudaMallocAsync((void**)&dy, (size_t)m->nFaces * sizeof(double), st));
    int rc = k_faceH(m, dy, dx);
    if (rc == LDU_OK
        && cudaMemcpyAsync(faceHpsi, dy, (size_t)m->nFaces * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess)
        rc = LDU_ECUDA;
    cudaFreeAsync(dy, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = LDU_ECUDA;
    return rc;
}

int ldu_H_device(ldu_matrix* m, double* d_Hpsi, const double* d_psi)
{
    LDU_TRY(check_matrix(m, "ldu_H_device"));
    return k_H(m, d_Hpsi, d_psi);
}

int ldu_precondition(ldu_matrix* m, int preconditioner, double* wA, const double* rA, int transpose)
{
    LDU_TRY(check_matrix(m, "ldu_precondition"));
    Staged s{m};
    double* dr = s.in(W_SRC, rA);
    double* dw = work_vec(m, W_OUT);
    LDU_TRY(precondition_device(m, preconditioner, dw, dr, transpose));
    return s.out(wA, dw);
}

int ldu_smooth(ldu_matrix* m, int smoother, double* psi, const double* source, int nSweeps)
{
    LDU_TRY(check_matrix(m, "ldu_smooth"));
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* db = s.in(W_SRC, source);
    LDU_TRY(smooth_device(m, smoother, dx, db, nSweeps));
    return s.out(psi, dx);
}

int ldu_solve(ldu_matrix* m, const ldu_controls* controls, double* psi, const double* source,
              ldu_solver_performance* perf)
{
    LDU_TRY(check_matrix(m, "ldu_solve"));
    Staged s{m};
    double* dx = s.in(W_PSI, psi);
    double* db = s.in(W_SRC, source);
    LDU_TRY(solve_device(m, controls, dx, db, perf));
    return s.out(psi, dx);
}

int ldu_amul_device(ldu_matrix* m, double* d_Apsi, const double* d_psi)
{
    LDU_TRY(check_matrix(m, "ldu_amul_device"));
    return k_amul(m, d_Apsi, d_psi, false);
}

int ldu_tmul_device(ldu_matrix* m, double* d_Tpsi, const double* d_psi)
{
    LDU_TRY(check_matrix(m, "ldu_tmul_device"));
    return k_amul(m, d_Tpsi, d_psi, true);
}

int ldu_solve_device(ldu_matrix* m, const ldu_controls* controls, double* d_psi, const double* d_source,
                     ldu_solver_performance* perf)
{
    LDU_TRY(check_matrix(m, "ldu_solve_device"));
    return solve_device(m, controls, d_psi, d_source, perf);
}

int ldu_residual_history(ldu_matrix* m, double* hist, int capacity)
{
    if (!m || !hist) return 0;
    const int n = (int)m->lastHistory.size() < capacity ? (int)m->lastHistory.size() : capacity;
    memcpy(hist, m->lastHistory.data(), n * sizeof(double));
    return n;
}

}  // extern "C"
