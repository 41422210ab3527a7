// Level-scheduled sweep primitives (sweeps.cu) and solver building blocks.
#pragma once

#include "ldu_internal.h"

namespace ldu {

// w = forward substitution of r (init) or of w itself; coef per face, multiplied
// by rD[row] unless `pre` (FDIC's precomputed rDuUpper)
int sweep_forward(ldu_matrix* m, const double* rD, const double* coef, bool pre, const double* r,
                  double* w, bool init);
int sweep_backward(ldu_matrix* m, const double* rD, const double* coef, bool pre, double* w);
// backward substitution that consumes the faces of a row in REVERSE LOSORT order (descending
// neighbour): DILUPreconditioner::preconditionT's second loop (DILUPreconditioner.C:172-183).
// Same as sweep_backward when m->nbrSorted.
int sweep_backward_losort(ldu_matrix* m, const double* rD, const double* coef, double* w);
int calc_reciprocal_D(ldu_matrix* m, double* rD, bool dilu);
int calc_reciprocal_diag(ldu_matrix* m, double* rD);
int calc_fdic_coeffs(ldu_matrix* m, const double* rD, double* rDuUpper, double* rDlUpper);
int gs_sweep(ldu_matrix* m, const double* bPrime, double* bLower, double* psi, bool sym);
// one multiColourGaussSeidel sweep: colours in order, the rows of a colour in one launch
int mcgs_sweep(ldu_matrix* m, const double* bPrime, double* psi);
// one nonBlockingGaussSeidel sweep of a region WITH interfaces; m->d_recv already holds the halo
int nbgs_sweep(ldu_matrix* m, const double* source, double* psi);

// dataflow (single persistent kernel per sweep) versions, flow.cu
bool flow_enabled();
void flow_free(ldu_matrix* m);
int flow_forward(ldu_matrix* m, const double* rD, const double* coef, bool pre, const double* r, double* w,
                 bool init);
int flow_backward(ldu_matrix* m, const double* rD, const double* coef, bool pre, double* w);
int flow_rD(ldu_matrix* m, double* rD, const double* upper, const double* lower);
int flow_gs(ldu_matrix* m, const double* bPrime, double* bLower, double* psi, bool sym);

// line-pipelined sweeps for structured boxes, stencil.cu
void stencil_free(ldu_matrix* m);
int stencil_forward(ldu_matrix* m, const double* rD, const double* coef, const double* r, double* w, bool init);
int stencil_backward(ldu_matrix* m, const double* rD, const double* coef, double* w);
// second generation (stencil2.cu): both substitutions of one application, tile layout inside
int stencil_version(const ldu_matrix* m);   // 0 = not a box / disabled, 1 = stencil.cu, 2 = stencil2.cu
void stencil2_free(ldu_matrix* m);
int stencil2_apply(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, const double* r,
                   double* w, bool init);

// PCG on a box (solvers.cu): the application fused with <w, dotWith> (EpiWArA), and the x/r update
// fused with |r|_1 (EpiResidual) and with the packing of rD*rA for the next application
int stencil2_apply_dot(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, const double* r,
                       double* w, const double* dotWith);
// the recurrence of the DIC / DILU diagonal on a blockMesh box (without the final reciprocal)
bool stencil2_rD_available(ldu_matrix* m);
int stencil2_rD(ldu_matrix* m, double* rD, const double* upper, const double* lower);
int stencil2_xr_pack(ldu_matrix* m, const double* rD, double* psi, double* rA, const double* pA, const double* wA);
void stencil2_invalidate(ldu_matrix* m);   // the tile image of rD*rA is stale (a new solve starts)

// forward then backward substitution: w = B^-1 F^-1 (init ? rD*r : w)
int sweep_pair(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, bool pre,
               const double* r, double* w, bool init);

// work-vector slots of a matrix (cell-sized scratch, allocated on first use)
enum {
    W_PA = 0, W_WA, W_RA, W_PT, W_WT, W_RT, W_RD, W_TMP, W_BPRIME, W_BLOWER,
    W_APSI, W_CORR, W_RES, W_SRD, W_STMP, W_PSI, W_SRC, W_OUT, W_PRECOND_A, W_PRECOND_B,
    W_COUNT
};

struct Precond {
    int kind = LDU_PRECOND_NONE;
    double* rD = nullptr;
    double* rDuUpper = nullptr;   // FDIC, face-sized (owned)
    double* rDlUpper = nullptr;
};

int precond_setup(ldu_matrix* m, int kind, Precond& p, int rDSlot);
void precond_release(Precond& p);
int precond_apply(ldu_matrix* m, const Precond& p, double* wA, const double* rA, bool transpose);

struct Smoother {
    int kind = LDU_SMOOTHER_GS;
    Precond dic;   // DIC / DILU / FDIC part
};

int smoother_setup(ldu_matrix* m, int kind, Smoother& s);
void smoother_release(Smoother& s);
int smoother_apply(ldu_matrix* m, const Smoother& s, double* psi, const double* source, int nSweeps);

int init_scalars(ldu_matrix* m, const ldu_controls* c);
int read_scalars(ldu_matrix* m, SolverScalars* out);
// wA = A psi; rA = source - wA; normFactor; initial residual; convergence test
int solve_prologue(ldu_matrix* m, double* psi, const double* source, double* wA, double* rA, double* tmp);
int fetch_performance(ldu_matrix* m, ldu_solver_performance* perf);

}  // namespace ldu
