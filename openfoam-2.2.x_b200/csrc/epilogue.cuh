// Scalar epilogues shared by the solvers: SolverPerformance logic on device.
#pragma once

#include "ldu_internal.h"

namespace ldu {

constexpr double kGreat = 1.0e+20;   // SolverPerformance.H:261-275
constexpr double kSmall = 1.0e-20;
constexpr double kVSmall = 1.0e-300;

// ---------------------------------------------------------------------------
// scalar epilogues (run by one thread after a reduction has been finalised)
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool check_convergence(SolverScalars* S)
{
    // SolverPerformance.C:59-91
    const bool conv = (S->finalResidual < S->tolerance)
        || (S->relTol > kSmall && S->finalResidual < __dmul_rn(S->relTol, S->initialResidual));
    S->converged = conv ? 1 : 0;
    return conv;
}

__device__ __forceinline__ void push_history(SolverScalars* S)
{
    if (S->hist && S->histCount < kMaxHist) S->hist[S->histCount] = S->finalResidual;
    S->histCount++;
}

struct EpiWArA {  // PCG.C:126-132
    __device__ void operator()(SolverScalars* S, const double* t) const
    {
        S->wArAold = S->wArA;
        S->wArA = t[0];
        S->beta = __ddiv_rn(S->wArA, S->wArAold);
    }
};

struct EpiWApA {  // PCG.C:155-164
    __device__ void operator()(SolverScalars* S, const double* t) const
    {
        S->wApA = t[0];
        // checkSingularity(mag(wApA)/normFactor): SolverPerformance.C:31-43
        if (__ddiv_rn(fabs(S->wApA), S->normFactor) < kVSmall) {
            S->singular = 1;
            S->done = 1;
        } else {
            S->singular = 0;
            S->alpha = __ddiv_rn(S->wArA, S->wApA);
        }
    }
};

// `while (nIterations++ < maxIter && !checkConvergence)` with `inc` added per
// pass (PCG.C:174-178: inc = 1 post-increment; smoothSolver.C:166-170 and
// GAMGSolverSolve.C:109-113 pre-increment)
template <bool POST_INCREMENT>
struct EpiResidual {
    int inc;
    __device__ void operator()(SolverScalars* S, const double* t) const
    {
        S->finalResidual = __ddiv_rn(t[0], S->normFactor);
        push_history(S);
        bool cont;
        if (POST_INCREMENT) {
            cont = S->nIterations < S->maxIter;
            S->nIterations += inc;
        } else {
            S->nIterations += inc;
            cont = S->nIterations < S->maxIter;
        }
        if (cont) cont = !check_convergence(S);
        if (!cont) S->done = 1;
    }
};

}  // namespace ldu
