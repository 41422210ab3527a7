/*
 * ldu_oracle.c — CPU restatement of the OpenFOAM-2.2.x lduMatrix solver path.
 *
 * TEST INFRASTRUCTURE ONLY (see ldu_oracle.h).  Plain C, sequential, written to
 * reproduce the reference's floating-point operation order exactly (build with
 * -ffp-contract=off: the reference's gcc -O3 x86-64 build has no FMA).
 *
 * Every function cites the reference file:line it restates; paths are relative
 * to /root/reference/src/OpenFOAM/matrices/lduMatrix/ ("LM/").
 */
#include "ldu_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* SolverPerformance.H:261-275 / SolverPerformance.C:31-33 */
#define ORC_GREAT 1.0e+20
#define ORC_SMALL 1.0e-20
#define ORC_VSMALL 1.0e-300
#define ORC_MAXLEVELS 50 /* GAMGAgglomeration.C:74 */
#define ORC_SCALAR_GREAT 1.0e+15 /* primitives/Scalar/doubleScalar/doubleScalar.H:54 */

typedef struct {
    int nbrRegion;
    int nbrInterface;
    int n;
    int* faceCells;       /* owned when level > 0, else borrowed */
    double* bou;          /* idem */
    double* intc;
    int owned;
    /* GAMG: fine interface face -> coarse interface face (level >= 1) */
    int* faceRestrict;
    int nFineFaces;
} orc_iface;

typedef struct orc_matrix {
    int nCells, nFaces;
    const int* l;
    const int* u;
    const double* diag;
    const double* upper;
    const double* lower;  /* == upper when symmetric */
    int symmetric;
    int* ownerStart;      /* [nCells+1] LM/lduAddressing/lduAddressing.C:92-123 */
    int* losort;          /* [nFaces]   LM/lduAddressing/lduAddressing.C:31-89 */
    int nIf;
    orc_iface* ifs;
    /* storage owned by coarse levels */
    int* own_l; int* own_u; double* own_diag; double* own_upper; double* own_lower;
} orc_matrix;

typedef struct {
    int nLevels;                         /* number of coarse levels */
    int* restrictAddr[ORC_MAXLEVELS];    /* [nFine(level)] */
    int* faceRestrictAddr[ORC_MAXLEVELS];/* [nFineFaces(level)] */
    int nFine[ORC_MAXLEVELS];
    int nFineFaces[ORC_MAXLEVELS];
    int nCoarse[ORC_MAXLEVELS];
    orc_matrix* level[ORC_MAXLEVELS];    /* coarse matrices: level[i] = matrixLevels_[i] */
    double* faceWeights;                 /* scratch */
} orc_hierarchy;

struct orc_world {
    int R;
    orc_matrix* m;            /* [R] finest level */
    const double** faceWeights;
    orc_hierarchy* h;         /* [R] or NULL */
    int hBuilt;
    int diagonalOnly;         /* lduMatrix::diagonal(): no upper and no lower coefficients were ever set */
    /* agglomeration kept from an earlier solve with cacheAgglomeration on (GAMGAgglomeration is a
     * MeshObject: GAMGAgglomeration.C:97-140 looks it up before building one) */
    int cached;
    int failed;               /* a solve stopped where the reference raises a FatalError */
};

/* lduMatrix::H: LM/lduMatrix/lduMatrixTemplates.C:33-65 (off-diagonal product, negated;
 * no interface terms: fvMatrix::H adds the boundary contributions itself) */
void orc_H(orc_world* w, double** Hpsi, double** psi)
{
    int r, c, f;
    for (r = 0; r < w->R; r++) {
        const orc_matrix* m = &w->m[r];
        double* H = Hpsi[r];
        const double* x = psi[r];
        for (c = 0; c < m->nCells; c++) H[c] = 0.0;
        for (f = 0; f < m->nFaces; f++) {
            H[m->u[f]] -= m->lower[f] * x[m->l[f]];
            H[m->l[f]] -= m->upper[f] * x[m->u[f]];
        }
    }
}

/* lduMatrix::H1: LM/lduMatrix/lduMatrixATmul.C:298-327 */
void orc_H1(orc_world* w, double** H1)
{
    int r, c, f;
    for (r = 0; r < w->R; r++) {
        const orc_matrix* m = &w->m[r];
        double* H = H1[r];
        for (c = 0; c < m->nCells; c++) H[c] = 0.0;
        for (f = 0; f < m->nFaces; f++) {
            H[m->u[f]] -= m->lower[f];
            H[m->l[f]] -= m->upper[f];
        }
    }
}

/* lduMatrix::faceH: LM/lduMatrix/lduMatrixTemplates.C:79-113, one value per face */
void orc_faceH(orc_world* w, double** faceHpsi, double** psi)
{
    int r, f;
    for (r = 0; r < w->R; r++) {
        const orc_matrix* m = &w->m[r];
        const double* x = psi[r];
        for (f = 0; f < m->nFaces; f++)
            faceHpsi[r][f] = m->upper[f] * x[m->u[f]] - m->lower[f] * x[m->l[f]];
    }
}

/* ------------------------------------------------------------------------- */
/* addressing                                                                 */
/* ------------------------------------------------------------------------- */

static void calc_addressing(orc_matrix* m)
{
    const int n = m->nCells, nf = m->nFaces;
    int i, f;
    /* ownerStart: LM/lduAddressing/lduAddressing.C:92-123 (CSR of faces by owner) */
    m->ownerStart = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    for (i = 0; i <= n; i++) m->ownerStart[i] = nf;
    m->ownerStart[0] = 0;
    {
        int nOwnStart = 0;
        i = 1;
        for (f = 0; f < nf; f++) {
            int curOwn = m->l[f];
            if (curOwn > nOwnStart) {
                while (i <= curOwn) m->ownerStart[i++] = f;
                nOwnStart = curOwn;
            }
        }
    }
    /* losort: LM/lduAddressing/lduAddressing.C:31-89 — faces grouped by upper
     * cell, ascending face index inside a group */
    m->losort = (int*)malloc(sizeof(int) * (size_t)(nf > 0 ? nf : 1));
    {
        int* cnt = (int*)calloc((size_t)(n + 1), sizeof(int));
        int* start = (int*)malloc(sizeof(int) * (size_t)(n + 1));
        for (f = 0; f < nf; f++) cnt[m->u[f]]++;
        start[0] = 0;
        for (i = 0; i < n; i++) start[i + 1] = start[i] + cnt[i];
        for (i = 0; i < n; i++) cnt[i] = 0;
        for (f = 0; f < nf; f++) {
            int c = m->u[f];
            m->losort[start[c] + cnt[c]] = f;
            cnt[c]++;
        }
        free(cnt);
        free(start);
    }
}

static void matrix_clear(orc_matrix* m)
{
    int i;
    free(m->ownerStart);
    free(m->losort);
    for (i = 0; i < m->nIf; i++) {
        if (m->ifs[i].owned) {
            free(m->ifs[i].faceCells);
            free(m->ifs[i].bou);
            free(m->ifs[i].intc);
        }
        free(m->ifs[i].faceRestrict);
    }
    free(m->ifs);
    free(m->own_l); free(m->own_u); free(m->own_diag); free(m->own_upper); free(m->own_lower);
    memset(m, 0, sizeof(*m));
}

static void hierarchy_clear(orc_hierarchy* h)
{
    int i;
    for (i = 0; i < h->nLevels; i++) {
        free(h->restrictAddr[i]);
        free(h->faceRestrictAddr[i]);
        if (h->level[i]) { matrix_clear(h->level[i]); free(h->level[i]); }
    }
    memset(h, 0, sizeof(*h));
}

/* lduMatrix::diagonal() (lduMatrix.H:547-550) is about which coefficient fields EXIST, not about the
 * number of faces: a mesh without internal faces whose upper() was touched (fvm::laplacian always
 * does) is not "diagonal" and goes to the selected solver */
void orc_world_set_diagonal(orc_world* w, int flag) { w->diagonalOnly = flag; }

orc_world* orc_world_new(int nRegions)
{
    orc_world* w = (orc_world*)calloc(1, sizeof(orc_world));
    w->R = nRegions;
    w->m = (orc_matrix*)calloc((size_t)nRegions, sizeof(orc_matrix));
    w->faceWeights = (const double**)calloc((size_t)nRegions, sizeof(double*));
    w->h = (orc_hierarchy*)calloc((size_t)nRegions, sizeof(orc_hierarchy));
    return w;
}

void orc_world_free(orc_world* w)
{
    int r;
    if (!w) return;
    for (r = 0; r < w->R; r++) {
        hierarchy_clear(&w->h[r]);
        matrix_clear(&w->m[r]);
    }
    free(w->h);
    free(w->m);
    free((void*)w->faceWeights);
    free(w);
}

void orc_world_set_region(orc_world* w, int r, int nCells, int nFaces,
                          const int* lowerAddr, const int* upperAddr,
                          const double* diag, const double* upper,
                          const double* lower)
{
    orc_matrix* m = &w->m[r];
    matrix_clear(m);
    m->nCells = nCells;
    m->nFaces = nFaces;
    m->l = lowerAddr;
    m->u = upperAddr;
    m->diag = diag;
    m->upper = upper;
    /* lduMatrix.C:198-215: lower() of a symmetric matrix returns upper() */
    m->symmetric = (lower == NULL);
    m->lower = lower ? lower : upper;
    calc_addressing(m);
}

int orc_world_add_interface(orc_world* w, int r, int nbrRegion, int nbrInterface,
                            int nFaces, const int* faceCells,
                            const double* bouCoeffs, const double* intCoeffs)
{
    orc_matrix* m = &w->m[r];
    orc_iface* it;
    m->ifs = (orc_iface*)realloc(m->ifs, sizeof(orc_iface) * (size_t)(m->nIf + 1));
    it = &m->ifs[m->nIf];
    memset(it, 0, sizeof(*it));
    it->nbrRegion = nbrRegion;
    it->nbrInterface = nbrInterface;
    it->n = nFaces;
    it->faceCells = (int*)faceCells;
    it->bou = (double*)bouCoeffs;
    it->intc = (double*)intCoeffs;
    return m->nIf++;
}

void orc_world_set_face_weights(orc_world* w, int r, const double* weights)
{
    w->faceWeights[r] = weights;
}

/* ------------------------------------------------------------------------- */
/* interface update (processor patches, in-process)                           */
/* ------------------------------------------------------------------------- */

/*
 * LM/lduMatrix/lduMatrixUpdateMatrixInterfaces.C:30-266 drives, per coupled
 * patch, processorFvPatchField<scalar>::initInterfaceMatrixUpdate /
 * updateInterfaceMatrix (finiteVolume/fields/fvPatchFields/constraint/processor/
 * processorFvPatchScalarField.C:36-144) on the finest level and
 * processorGAMGInterfaceField (LM/solvers/GAMG/interfaceFields/
 * processorGAMGInterfaceField/processorGAMGInterfaceField.C:73-174) on coarse
 * levels:  send psi[faceCells]; result[faceCells[i]] -= coeffs[i]*recv[i].
 * In-process the "receive buffer" is the neighbour region's psi at its own
 * faceCells.  sign = -1 restates the negated coefficients used by residual()
 * and the Gauss-Seidel smoothers.  which: 0 = bouCoeffs, 1 = intCoeffs.
 */
static void update_interfaces(orc_matrix* ms, int R, int r, double* result,
                              double** psi, int which, double sign)
{
    orc_matrix* m = &ms[r];
    int p, i;
    (void)R;
    for (p = 0; p < m->nIf; p++) {
        const orc_iface* it = &m->ifs[p];
        const orc_iface* nb = &ms[it->nbrRegion].ifs[it->nbrInterface];
        const double* coeffs = which ? it->intc : it->bou;
        const double* psiNbr = psi[it->nbrRegion];
        for (i = 0; i < it->n; i++) {
            double c = sign * coeffs[i]; /* exact: negate() flips the sign bit */
            result[it->faceCells[i]] -= c * psiNbr[nb->faceCells[i]];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* lduMatrix::Amul / Tmul / sumA / residual                                   */
/* ------------------------------------------------------------------------- */

/* LM/lduMatrix/lduMatrixATmul.C:34-92 */
static void amul_levels(orc_matrix* ms, int R, double** Apsi, double** psi)
{
    int r, c, f;
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        double* A = Apsi[r];
        const double* x = psi[r];
        for (c = 0; c < m->nCells; c++) A[c] = m->diag[c] * x[c];
        for (f = 0; f < m->nFaces; f++) {
            A[m->u[f]] += m->lower[f] * x[m->l[f]];
            A[m->l[f]] += m->upper[f] * x[m->u[f]];
        }
    }
    for (r = 0; r < R; r++) update_interfaces(ms, R, r, Apsi[r], psi, 0, 1.0);
}

void orc_amul(orc_world* w, double** Apsi, double** psi)
{
    amul_levels(w->m, w->R, Apsi, psi);
}

/* LM/lduMatrix/lduMatrixATmul.C:95-151 */
void orc_tmul(orc_world* w, double** Tpsi, double** psi)
{
    int r, c, f;
    for (r = 0; r < w->R; r++) {
        const orc_matrix* m = &w->m[r];
        double* T = Tpsi[r];
        const double* x = psi[r];
        for (c = 0; c < m->nCells; c++) T[c] = m->diag[c] * x[c];
        for (f = 0; f < m->nFaces; f++) {
            T[m->u[f]] += m->upper[f] * x[m->l[f]];
            T[m->l[f]] += m->lower[f] * x[m->u[f]];
        }
    }
    for (r = 0; r < w->R; r++) update_interfaces(w->m, w->R, r, Tpsi[r], psi, 1, 1.0);
}

/* LM/lduMatrix/lduMatrixATmul.C:154-200 */
static void sumA_levels(orc_matrix* ms, int R, double** sumA)
{
    int r, c, f, p, i;
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        double* s = sumA[r];
        for (c = 0; c < m->nCells; c++) s[c] = m->diag[c];
        for (f = 0; f < m->nFaces; f++) {
            s[m->u[f]] += m->lower[f];
            s[m->l[f]] += m->upper[f];
        }
        for (p = 0; p < m->nIf; p++) {
            const orc_iface* it = &m->ifs[p];
            for (i = 0; i < it->n; i++) s[it->faceCells[i]] -= it->bou[i];
        }
    }
}

void orc_sumA(orc_world* w, double** sumA) { sumA_levels(w->m, w->R, sumA); }

/* LM/lduMatrix/lduMatrixATmul.C:203-281 */
static void residual_levels(orc_matrix* ms, int R, double** rA, double** psi, double** source)
{
    int r, c, f;
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        double* res = rA[r];
        const double* x = psi[r];
        const double* b = source[r];
        for (c = 0; c < m->nCells; c++) res[c] = b[c] - m->diag[c] * x[c];
        for (f = 0; f < m->nFaces; f++) {
            res[m->u[f]] -= m->lower[f] * x[m->l[f]];
            res[m->l[f]] -= m->upper[f] * x[m->u[f]];
        }
    }
    for (r = 0; r < R; r++) update_interfaces(ms, R, r, rA[r], psi, 0, -1.0);
}

void orc_residual(orc_world* w, double** rA, double** psi, double** source)
{
    residual_levels(w->m, w->R, rA, psi, source);
}

/* ------------------------------------------------------------------------- */
/* global reductions: fields/Fields/Field/FieldFunctions.C:363-385,422-434,    */
/* 477-533 — naive left-to-right local sums, then reduce(sumOp) over ranks     */
/* (summed here in rank order)                                                 */
/* ------------------------------------------------------------------------- */

static double gSumProd_levels(orc_matrix* ms, int R, double** a, double** b)
{
    double g = 0;
    int r, c;
    for (r = 0; r < R; r++) {
        double s = 0;
        for (c = 0; c < ms[r].nCells; c++) s += a[r][c] * b[r][c];
        g = (r == 0) ? s : g + s;
    }
    return g;
}

static double gSumMag_levels(orc_matrix* ms, int R, double** a)
{
    double g = 0;
    int r, c;
    for (r = 0; r < R; r++) {
        double s = 0;
        for (c = 0; c < ms[r].nCells; c++) s += fabs(a[r][c]);
        g = (r == 0) ? s : g + s;
    }
    return g;
}

double orc_gSumProd(orc_world* w, double** a, double** b) { return gSumProd_levels(w->m, w->R, a, b); }
double orc_gSumMag(orc_world* w, double** a) { return gSumMag_levels(w->m, w->R, a); }

static double** alloc_fields(orc_matrix* ms, int R)
{
    double** f = (double**)malloc(sizeof(double*) * (size_t)R);
    int r;
    for (r = 0; r < R; r++) f[r] = (double*)calloc((size_t)(ms[r].nCells > 0 ? ms[r].nCells : 1), sizeof(double));
    return f;
}

static void free_fields(double** f, int R)
{
    int r;
    for (r = 0; r < R; r++) free(f[r]);
    free(f);
}

/* LM/lduMatrix/lduMatrixSolver.C:179-197
 *   sumA -> tmp; tmp *= gAverage(psi);
 *   return gSum(mag(Apsi - tmp) + mag(source - tmp)) + small_
 * gAverage (FieldFunctions.C:514-533): sum and count reduced together, then
 * sum/count. */
static double normFactor_levels(orc_matrix* ms, int R, double** psi, double** source, double** Apsi)
{
    double** tmp = alloc_fields(ms, R);
    double gs = 0, avg, nf = 0;
    long n = 0;
    int r, c;
    sumA_levels(ms, R, tmp);
    for (r = 0; r < R; r++) {
        double s = 0;
        for (c = 0; c < ms[r].nCells; c++) s += psi[r][c];
        gs = (r == 0) ? s : gs + s;
        n += ms[r].nCells;
    }
    avg = gs / (double)n;
    for (r = 0; r < R; r++) {
        double s = 0;
        for (c = 0; c < ms[r].nCells; c++) {
            double t = tmp[r][c] * avg;
            s += fabs(Apsi[r][c] - t) + fabs(source[r][c] - t);
        }
        nf = (r == 0) ? s : nf + s;
    }
    free_fields(tmp, R);
    return nf + ORC_SMALL;
}

double orc_normFactor(orc_world* w, double** psi, double** source, double** Apsi)
{
    return normFactor_levels(w->m, w->R, psi, source, Apsi);
}

/* ------------------------------------------------------------------------- */
/* preconditioners                                                            */
/* ------------------------------------------------------------------------- */

/* DICPreconditioner.C:57-84 (calcReciprocalD); DILU: DILUPreconditioner.C:57-85 */
static void calc_rD(const orc_matrix* m, double* rD, int dilu)
{
    int f, c;
    for (c = 0; c < m->nCells; c++) rD[c] = m->diag[c];
    if (dilu) {
        for (f = 0; f < m->nFaces; f++)
            rD[m->u[f]] -= m->upper[f] * m->lower[f] / rD[m->l[f]];
    } else {
        for (f = 0; f < m->nFaces; f++)
            rD[m->u[f]] -= m->upper[f] * m->upper[f] / rD[m->l[f]];
    }
    for (c = 0; c < m->nCells; c++) rD[c] = 1.0 / rD[c];
}

/* forward/backward face sweeps shared by DIC (DICPreconditioner.C:87-123) and
 * the DIC smoother (smoothers/DIC/DICSmoother.C:100-113) */
static void dic_sweeps(const orc_matrix* m, const double* rD, double* wA)
{
    int f;
    for (f = 0; f < m->nFaces; f++)
        wA[m->u[f]] -= rD[m->u[f]] * m->upper[f] * wA[m->l[f]];
    for (f = m->nFaces - 1; f >= 0; f--)
        wA[m->l[f]] -= rD[m->l[f]] * m->upper[f] * wA[m->u[f]];
}

/* DILUPreconditioner.C:88-135 (forward in losort order) */
static void dilu_sweeps(const orc_matrix* m, const double* rD, double* wA)
{
    int f;
    for (f = 0; f < m->nFaces; f++) {
        int sf = m->losort[f];
        wA[m->u[sf]] -= rD[m->u[sf]] * m->lower[sf] * wA[m->l[sf]];
    }
    for (f = m->nFaces - 1; f >= 0; f--)
        wA[m->l[f]] -= rD[m->l[f]] * m->upper[f] * wA[m->u[f]];
}

/* DILUPreconditioner.C:138-185 */
static void dilu_sweepsT(const orc_matrix* m, const double* rD, double* wT)
{
    int f;
    for (f = 0; f < m->nFaces; f++)
        wT[m->u[f]] -= rD[m->u[f]] * m->upper[f] * wT[m->l[f]];
    for (f = m->nFaces - 1; f >= 0; f--) {
        int sf = m->losort[f];
        wT[m->l[sf]] -= rD[m->l[sf]] * m->lower[sf] * wT[m->u[sf]];
    }
}

/* DILU smoother sweeps: smoothers/DILU/DILUSmoother.C:103-116 (plain face order) */
static void dilu_smoother_sweeps(const orc_matrix* m, const double* rD, double* rA)
{
    int f;
    for (f = 0; f < m->nFaces; f++)
        rA[m->u[f]] -= rD[m->u[f]] * m->lower[f] * rA[m->l[f]];
    for (f = m->nFaces - 1; f >= 0; f--)
        rA[m->l[f]] -= rD[m->l[f]] * m->upper[f] * rA[m->u[f]];
}

typedef struct {
    int kind;
    double** rD;       /* per region */
    double** rDuUpper; /* FDIC */
    double** rDlUpper;
} orc_precond;

static void precond_free(orc_precond* p, int R)
{
    int r;
    for (r = 0; r < R; r++) {
        if (p->rD) free(p->rD[r]);
        if (p->rDuUpper) free(p->rDuUpper[r]);
        if (p->rDlUpper) free(p->rDlUpper[r]);
    }
    free(p->rD); free(p->rDuUpper); free(p->rDlUpper);
    memset(p, 0, sizeof(*p));
}

static int precond_init(orc_precond* p, orc_matrix* ms, int R, int kind)
{
    int r, c, f;
    memset(p, 0, sizeof(*p));
    p->kind = kind;
    if (kind == ORC_PRECOND_NONE) return 0;
    p->rD = (double**)calloc((size_t)R, sizeof(double*));
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        p->rD[r] = (double*)malloc(sizeof(double) * (size_t)(m->nCells > 0 ? m->nCells : 1));
        switch (kind) {
        case ORC_PRECOND_DIAGONAL: /* diagonalPreconditioner.C:46-67 */
            for (c = 0; c < m->nCells; c++) p->rD[r][c] = 1.0 / m->diag[c];
            break;
        case ORC_PRECOND_DIC:
            calc_rD(m, p->rD[r], 0);
            break;
        case ORC_PRECOND_DILU:
            calc_rD(m, p->rD[r], 1);
            break;
        case ORC_PRECOND_FDIC: /* FDICPreconditioner.C:42-83 */
            if (!p->rDuUpper) {
                p->rDuUpper = (double**)calloc((size_t)R, sizeof(double*));
                p->rDlUpper = (double**)calloc((size_t)R, sizeof(double*));
            }
            for (c = 0; c < m->nCells; c++) p->rD[r][c] = m->diag[c];
            for (f = 0; f < m->nFaces; f++)
                p->rD[r][m->u[f]] -= (m->upper[f] * m->upper[f]) / p->rD[r][m->l[f]];
            for (c = 0; c < m->nCells; c++) p->rD[r][c] = 1.0 / p->rD[r][c];
            p->rDuUpper[r] = (double*)malloc(sizeof(double) * (size_t)(m->nFaces > 0 ? m->nFaces : 1));
            p->rDlUpper[r] = (double*)malloc(sizeof(double) * (size_t)(m->nFaces > 0 ? m->nFaces : 1));
            for (f = 0; f < m->nFaces; f++) {
                p->rDuUpper[r][f] = p->rD[r][m->u[f]] * m->upper[f];
                p->rDlUpper[r][f] = p->rD[r][m->l[f]] * m->upper[f];
            }
            break;
        default:
            return -1;
        }
    }
    return 0;
}

static void precond_apply(const orc_precond* p, orc_matrix* ms, int R,
                          double** wA, double** rA, int transpose)
{
    int r, c, f;
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        double* w = wA[r];
        const double* x = rA[r];
        switch (p->kind) {
        case ORC_PRECOND_NONE: /* noPreconditioner.C:58-74 */
            for (c = 0; c < m->nCells; c++) w[c] = x[c];
            break;
        case ORC_PRECOND_DIAGONAL: /* diagonalPreconditioner.C:70-87 */
            for (c = 0; c < m->nCells; c++) w[c] = p->rD[r][c] * x[c];
            break;
        case ORC_PRECOND_DIC: /* DICPreconditioner.C:87-123; symmetric: T == A */
            for (c = 0; c < m->nCells; c++) w[c] = p->rD[r][c] * x[c];
            dic_sweeps(m, p->rD[r], w);
            break;
        case ORC_PRECOND_FDIC: /* FDICPreconditioner.C:88-125 */
            for (c = 0; c < m->nCells; c++) w[c] = p->rD[r][c] * x[c];
            for (f = 0; f < m->nFaces; f++) w[m->u[f]] -= p->rDuUpper[r][f] * w[m->l[f]];
            for (f = m->nFaces - 1; f >= 0; f--) w[m->l[f]] -= p->rDlUpper[r][f] * w[m->u[f]];
            break;
        case ORC_PRECOND_DILU:
            for (c = 0; c < m->nCells; c++) w[c] = p->rD[r][c] * x[c];
            if (transpose) dilu_sweepsT(m, p->rD[r], w);
            else dilu_sweeps(m, p->rD[r], w);
            break;
        }
    }
}

int orc_precondition(orc_world* w, int precond, double** wA, double** rA, int transpose)
{
    orc_precond p;
    if (precond_init(&p, w->m, w->R, precond)) return -1;
    precond_apply(&p, w->m, w->R, wA, rA, transpose);
    precond_free(&p, w->R);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* smoothers                                                                  */
/* ------------------------------------------------------------------------- */

/* GaussSeidelSmoother.C:66-187 (sym == 0), symGaussSeidelSmoother.C:66-216
 * (sym == 1).  Boundary: bPrime = source, then the interface update with
 * NEGATED bouCoeffs (GaussSeidelSmoother.C:110-145): bPrime[fc] -= (-bou)*psiNbr,
 * using psi of all regions from BEFORE this sweep (Jacobi coupling). */
static void gs_rows(const orc_matrix* m, double* x, double* bP, int c0, int c1)
{
    int c, f;
    for (c = c0; c < c1; c++) {
        int fs = m->ownerStart[c], fe = m->ownerStart[c + 1];
        double psii = bP[c];
        for (f = fs; f < fe; f++) psii -= m->upper[f] * x[m->u[f]];
        psii /= m->diag[c];
        for (f = fs; f < fe; f++) bP[m->u[f]] -= m->lower[f] * psii;
        x[c] = psii;
    }
}

/* nonBlocking == 1: nonBlockingGaussSeidelSmoother.C:66-81,128-217.  The cells below
 * blockStart (the first cell on any interface) are swept BEFORE the interface update is
 * consumed, so a coupled cell receives the lower-side contributions of those cells first,
 * then the interface terms, then the rest: same operations as GaussSeidel, different
 * rounding order on the coupled cells.  No interface cell changes in the first phase, so
 * the neighbour values are still those from before the sweep. */
static void gs_sweeps(orc_matrix* ms, int R, double** psi, double** source, int nSweeps, int sym,
                      int nonBlocking)
{
    double** bPrime = alloc_fields(ms, R);
    int sweep, r, c, f, i, k;
    for (sweep = 0; sweep < nSweeps; sweep++) {
        for (r = 0; r < R; r++) {
            memcpy(bPrime[r], source[r], sizeof(double) * (size_t)ms[r].nCells);
        }
        if (nonBlocking) {
            for (r = 0; r < R; r++) {
                int blockStart = ms[r].nCells;
                for (i = 0; i < ms[r].nIf; i++)
                    for (k = 0; k < ms[r].ifs[i].n; k++)
                        if (ms[r].ifs[i].faceCells[k] < blockStart) blockStart = ms[r].ifs[i].faceCells[k];
                gs_rows(&ms[r], psi[r], bPrime[r], 0, blockStart);
            }
            for (r = 0; r < R; r++) update_interfaces(ms, R, r, bPrime[r], psi, 0, -1.0);
            for (r = 0; r < R; r++) {
                int blockStart = ms[r].nCells;
                for (i = 0; i < ms[r].nIf; i++)
                    for (k = 0; k < ms[r].ifs[i].n; k++)
                        if (ms[r].ifs[i].faceCells[k] < blockStart) blockStart = ms[r].ifs[i].faceCells[k];
                gs_rows(&ms[r], psi[r], bPrime[r], blockStart, ms[r].nCells);
            }
            continue;
        }
        for (r = 0; r < R; r++) update_interfaces(ms, R, r, bPrime[r], psi, 0, -1.0);
        /* all ranks sweep concurrently in the reference: interface values were
         * exchanged before any sweep started, so region order is irrelevant */
        for (r = 0; r < R; r++) {
            const orc_matrix* m = &ms[r];
            double* x = psi[r];
            double* bP = bPrime[r];
            gs_rows(m, x, bP, 0, m->nCells);
            if (sym) {
                for (c = m->nCells - 1; c >= 0; c--) {
                    int fs = m->ownerStart[c], fe = m->ownerStart[c + 1];
                    double psii = bP[c];
                    for (f = fs; f < fe; f++) psii -= m->upper[f] * x[m->u[f]];
                    psii /= m->diag[c];
                    for (f = fs; f < fe; f++) bP[m->u[f]] -= m->lower[f] * psii;
                    x[c] = psii;
                }
            }
        }
    }
    free_fields(bPrime, R);
}

/* multiColourGaussSeidel -- NOT one of the reference's smoothers: the north star's parallel form of
 * GaussSeidelSmoother.C:66-187.  The cells are coloured greedily in cell order (first colour no already
 * coloured neighbour has: include/ldu_b200.h ldu_colour_order); a sweep visits the colours in order and,
 * inside a colour, updates every cell from the CURRENT values of its neighbours (none has its colour):
 *     psi_c = ((bPrime_c - sum_{lower faces, ascending} lower_f psi_l) - sum_{upper faces, ascending} upper_f psi_u)/diag_c
 * with bPrime as in GaussSeidel (source + the Jacobi-coupled interface terms) and the row's operations in the
 * order the reference's loop performs them on a row (lower-side terms as their cells were visited, then the
 * upper side, then the division).  It IS the reference's Gauss-Seidel run on the mesh renumbered by colour
 * (renumberMesh-style, ldub200/renumber.py): same iterates up to the rounding of the row sums
 * (tests/test_multicolour_gs.py).  losort order = ascending face index of the faces whose upper cell is c. */
static int* greedy_colours(const orc_matrix* m, int* nColours)
{
    int n = m->nCells, c, f, k, q, nc = 0;
    int* colour = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* start = (int*)calloc((size_t)n + 1, sizeof(int));
    int* adj = (int*)malloc(sizeof(int) * (size_t)(2 * m->nFaces + 1));
    int* fill = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    char* mark = (char*)malloc((size_t)n + 1);
    for (f = 0; f < m->nFaces; f++) { start[m->l[f] + 1]++; start[m->u[f] + 1]++; }
    for (c = 0; c < n; c++) { start[c + 1] += start[c]; fill[c] = start[c]; colour[c] = -1; }
    for (f = 0; f < m->nFaces; f++) { adj[fill[m->l[f]]++] = m->u[f]; adj[fill[m->u[f]]++] = m->l[f]; }
    for (c = 0; c < n; c++) {
        for (q = 0; q <= nc; q++) mark[q] = 0;
        for (k = start[c]; k < start[c + 1]; k++) if (colour[adj[k]] >= 0) mark[colour[adj[k]]] = 1;
        q = 0;
        while (q < nc && mark[q]) q++;
        colour[c] = q;
        if (q == nc) nc++;
    }
    free(start); free(adj); free(fill); free(mark);
    *nColours = nc;
    return colour;
}

static void mcgs_sweeps(orc_matrix* ms, int R, double** psi, double** source, int nSweeps)
{
    double** bPrime = alloc_fields(ms, R);
    int** colour = (int**)malloc(sizeof(int*) * (size_t)R);
    int* nc = (int*)malloc(sizeof(int) * (size_t)R);
    int** losortStart = (int**)malloc(sizeof(int*) * (size_t)R);
    int sweep, r, c, f, k, q;
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        colour[r] = greedy_colours(m, &nc[r]);
        losortStart[r] = (int*)calloc((size_t)m->nCells + 1, sizeof(int));
        for (f = 0; f < m->nFaces; f++) losortStart[r][m->u[f] + 1]++;
        for (c = 0; c < m->nCells; c++) losortStart[r][c + 1] += losortStart[r][c];
    }
    for (sweep = 0; sweep < nSweeps; sweep++) {
        for (r = 0; r < R; r++) memcpy(bPrime[r], source[r], sizeof(double) * (size_t)ms[r].nCells);
        for (r = 0; r < R; r++) update_interfaces(ms, R, r, bPrime[r], psi, 0, -1.0);
        for (r = 0; r < R; r++) {
            const orc_matrix* m = &ms[r];
            double* x = psi[r];
            for (q = 0; q < nc[r]; q++) {
                for (c = 0; c < m->nCells; c++) {
                    double acc;
                    if (colour[r][c] != q) continue;
                    acc = bPrime[r][c];
                    for (k = losortStart[r][c]; k < losortStart[r][c + 1]; k++) {
                        f = m->losort[k];
                        acc -= m->lower[f] * x[m->l[f]];
                    }
                    for (f = m->ownerStart[c]; f < m->ownerStart[c + 1]; f++) acc -= m->upper[f] * x[m->u[f]];
                    x[c] = acc / m->diag[c];
                }
            }
        }
    }
    for (r = 0; r < R; r++) { free(colour[r]); free(losortStart[r]); }
    free(colour); free(nc); free(losortStart);
    free_fields(bPrime, R);
}

/* DICSmoother.C:67-116 / DILUSmoother.C:67-119 / FDICSmoother.C:98-146 */
static void dic_family_smooth(orc_matrix* ms, int R, double** psi, double** source,
                              int nSweeps, int kind)
{
    double** rA = alloc_fields(ms, R);
    orc_precond p;
    int sweep, r, c, f;
    precond_init(&p, ms, R,
                 kind == ORC_SMOOTHER_DILU ? ORC_PRECOND_DILU
               : kind == ORC_SMOOTHER_FDIC ? ORC_PRECOND_FDIC : ORC_PRECOND_DIC);
    for (sweep = 0; sweep < nSweeps; sweep++) {
        residual_levels(ms, R, rA, psi, source);
        for (r = 0; r < R; r++) {
            const orc_matrix* m = &ms[r];
            double* res = rA[r];
            for (c = 0; c < m->nCells; c++) res[c] *= p.rD[r][c];
            if (kind == ORC_SMOOTHER_DILU) {
                dilu_smoother_sweeps(m, p.rD[r], res);
            } else if (kind == ORC_SMOOTHER_FDIC) {
                for (f = 0; f < m->nFaces; f++) res[m->u[f]] -= p.rDuUpper[r][f] * res[m->l[f]];
                for (f = m->nFaces - 1; f >= 0; f--) res[m->l[f]] -= p.rDlUpper[r][f] * res[m->u[f]];
            } else {
                dic_sweeps(m, p.rD[r], res);
            }
            for (c = 0; c < m->nCells; c++) psi[r][c] += res[c];
        }
    }
    precond_free(&p, R);
    free_fields(rA, R);
}

static int smooth_levels(orc_matrix* ms, int R, int smoother, double** psi, double** source, int nSweeps)
{
    switch (smoother) {
    case ORC_SMOOTHER_GS:
        gs_sweeps(ms, R, psi, source, nSweeps, 0, 0);
        return 0;
    case ORC_SMOOTHER_NBGS:
        gs_sweeps(ms, R, psi, source, nSweeps, 0, 1);
        return 0;
    case ORC_SMOOTHER_SYMGS:
        gs_sweeps(ms, R, psi, source, nSweeps, 1, 0);
        return 0;
    case ORC_SMOOTHER_DIC:
    case ORC_SMOOTHER_DILU:
    case ORC_SMOOTHER_FDIC:
        dic_family_smooth(ms, R, psi, source, nSweeps, smoother);
        return 0;
    case ORC_SMOOTHER_DICGS: /* DICGaussSeidelSmoother.C:79-89: DIC then GS, nSweeps each */
        dic_family_smooth(ms, R, psi, source, nSweeps, ORC_SMOOTHER_DIC);
        gs_sweeps(ms, R, psi, source, nSweeps, 0, 0);
        return 0;
    case ORC_SMOOTHER_DILUGS:
        dic_family_smooth(ms, R, psi, source, nSweeps, ORC_SMOOTHER_DILU);
        gs_sweeps(ms, R, psi, source, nSweeps, 0, 0);
        return 0;
    case ORC_SMOOTHER_MCGS:
        mcgs_sweeps(ms, R, psi, source, nSweeps);
        return 0;
    }
    return -1;
}

int orc_smooth(orc_world* w, int smoother, double** psi, double** source, int nSweeps)
{
    return smooth_levels(w->m, w->R, smoother, psi, source, nSweeps);
}

/* ------------------------------------------------------------------------- */
/* SolverPerformance                                                          */
/* ------------------------------------------------------------------------- */

/* SolverPerformance.C:59-91 */
static int check_convergence(orc_perf* p, double tol, double relTol)
{
    if (p->finalResidual < tol
        || (relTol > ORC_SMALL && p->finalResidual < relTol * p->initialResidual))
        p->converged = 1;
    else
        p->converged = 0;
    return p->converged;
}

/* SolverPerformance.C:31-43 */
static int check_singularity(orc_perf* p, double x)
{
    p->singular = (x < ORC_VSMALL);
    return p->singular;
}

static void hist_push(double* hist, int cap, int* n, double v)
{
    if (hist && *n < cap) hist[*n] = v;
    (*n)++;
}

/* ------------------------------------------------------------------------- */
/* PCG (LM/solvers/PCG/PCG.C:65-182) and PBiCG (LM/solvers/PBiCG/PBiCG.C:65-198)*/
/* ------------------------------------------------------------------------- */

static int gamg_precondition(orc_world* w, const orc_controls* c, double** wA, double** rA);

static void krylov_solve(orc_world* w, orc_matrix* ms, int R, const orc_controls* ctl,
                         double** psi, double** source, orc_perf* perf,
                         double* hist, int histCap, int bicg)
{
    double** pA = alloc_fields(ms, R);
    double** wA = alloc_fields(ms, R);
    double** rA = alloc_fields(ms, R);
    double** pT = NULL; double** wT = NULL; double** rT = NULL;
    double wArA = ORC_GREAT, wArAold = wArA, normFactor;
    int r, c, nh = 0;

    memset(perf, 0, sizeof(*perf));
    amul_levels(ms, R, wA, psi);
    if (bicg) {
        pT = alloc_fields(ms, R); wT = alloc_fields(ms, R); rT = alloc_fields(ms, R);
        /* Tmul on this level set */
        {
            orc_world tmp = *w; tmp.m = ms; tmp.R = R;
            orc_tmul(&tmp, wT, psi);
        }
    }
    for (r = 0; r < R; r++)
        for (c = 0; c < ms[r].nCells; c++) {
            rA[r][c] = source[r][c] - wA[r][c];
            if (bicg) rT[r][c] = source[r][c] - wT[r][c];
        }
    normFactor = normFactor_levels(ms, R, psi, source, wA);
    perf->initialResidual = gSumMag_levels(ms, R, rA) / normFactor;
    perf->finalResidual = perf->initialResidual;
    hist_push(hist, histCap, &nh, perf->finalResidual);

    if (!check_convergence(perf, ctl->tolerance, ctl->relTol)) {
        orc_precond pre;
        int useGamg = (ctl->preconditioner == ORC_PRECOND_GAMG);
        if (!useGamg) precond_init(&pre, ms, R, ctl->preconditioner);
        /* GAMGPreconditioner is constructed here (PCG.C:111-115): no coarse level -> FatalError */
        if (useGamg && !w->hBuilt && orc_gamg_build(w, ctl)) w->failed = 1;
        else do {
            double wApA, alpha;
            wArAold = wArA;
            if (useGamg) {
                gamg_precondition(w, ctl, wA, rA);
            } else {
                precond_apply(&pre, ms, R, wA, rA, 0);
                if (bicg) precond_apply(&pre, ms, R, wT, rT, 1);
            }
            wArA = gSumProd_levels(ms, R, wA, bicg ? rT : rA);
            if (perf->nIterations == 0) {
                for (r = 0; r < R; r++)
                    for (c = 0; c < ms[r].nCells; c++) {
                        pA[r][c] = wA[r][c];
                        if (bicg) pT[r][c] = wT[r][c];
                    }
            } else {
                double beta = wArA / wArAold;
                for (r = 0; r < R; r++)
                    for (c = 0; c < ms[r].nCells; c++) {
                        pA[r][c] = wA[r][c] + beta * pA[r][c];
                        if (bicg) pT[r][c] = wT[r][c] + beta * pT[r][c];
                    }
            }
            amul_levels(ms, R, wA, pA);
            if (bicg) {
                orc_world tmp = *w; tmp.m = ms; tmp.R = R;
                orc_tmul(&tmp, wT, pT);
            }
            wApA = gSumProd_levels(ms, R, wA, bicg ? pT : pA);
            if (check_singularity(perf, fabs(wApA) / normFactor)) break;
            alpha = wArA / wApA;
            for (r = 0; r < R; r++)
                for (c = 0; c < ms[r].nCells; c++) {
                    psi[r][c] += alpha * pA[r][c];
                    rA[r][c] -= alpha * wA[r][c];
                    if (bicg) rT[r][c] -= alpha * wT[r][c];
                }
            perf->finalResidual = gSumMag_levels(ms, R, rA) / normFactor;
            hist_push(hist, histCap, &nh, perf->finalResidual);
        } while (perf->nIterations++ < ctl->maxIter
                 && !check_convergence(perf, ctl->tolerance, ctl->relTol));
        if (!useGamg) precond_free(&pre, R);
    }
    free_fields(pA, R); free_fields(wA, R); free_fields(rA, R);
    if (bicg) { free_fields(pT, R); free_fields(wT, R); free_fields(rT, R); }
}

/* ------------------------------------------------------------------------- */
/* smoothSolver (LM/solvers/smoothSolver/smoothSolver.C:77-180)                */
/* ------------------------------------------------------------------------- */

static void smooth_solve(orc_matrix* ms, int R, const orc_controls* ctl,
                         double** psi, double** source, orc_perf* perf,
                         double* hist, int histCap)
{
    int nh = 0, r, c;
    memset(perf, 0, sizeof(*perf));
    if (ctl->nSweeps < 0) {
        smooth_levels(ms, R, ctl->smoother, psi, source, -ctl->nSweeps);
        perf->nIterations -= ctl->nSweeps;
        return;
    }
    {
        double** Apsi = alloc_fields(ms, R);
        double** res = alloc_fields(ms, R);
        double normFactor;
        amul_levels(ms, R, Apsi, psi);
        normFactor = normFactor_levels(ms, R, psi, source, Apsi);
        for (r = 0; r < R; r++)
            for (c = 0; c < ms[r].nCells; c++) res[r][c] = source[r][c] - Apsi[r][c];
        perf->initialResidual = gSumMag_levels(ms, R, res) / normFactor;
        perf->finalResidual = perf->initialResidual;
        hist_push(hist, histCap, &nh, perf->finalResidual);
        if (!check_convergence(perf, ctl->tolerance, ctl->relTol)) {
            do {
                smooth_levels(ms, R, ctl->smoother, psi, source, ctl->nSweeps);
                residual_levels(ms, R, res, psi, source);
                perf->finalResidual = gSumMag_levels(ms, R, res) / normFactor;
                hist_push(hist, histCap, &nh, perf->finalResidual);
            } while ((perf->nIterations += ctl->nSweeps) < ctl->maxIter
                     && !check_convergence(perf, ctl->tolerance, ctl->relTol));
        }
        free_fields(Apsi, R); free_fields(res, R);
    }
}

/* diagonalSolver.C:62-81 */
static void diagonal_solve(orc_matrix* ms, int R, double** psi, double** source, orc_perf* perf)
{
    int r, c;
    memset(perf, 0, sizeof(*perf));
    for (r = 0; r < R; r++)
        for (c = 0; c < ms[r].nCells; c++) psi[r][c] = source[r][c] / ms[r].diag[c];
    perf->converged = 1;
}

/* ------------------------------------------------------------------------- */
/* GAMG agglomeration                                                         */
/* ------------------------------------------------------------------------- */

/* pairGAMGAgglomerate.C:31-198 — one level of greedy pairwise clustering */
static int* pair_agglomerate(int* nCoarseOut, int nFine, int nFaces,
                             const int* lowerAddr, const int* upperAddr,
                             const double* faceWeights)
{
    int* cellFaces = (int*)malloc(sizeof(int) * (size_t)(2 * nFaces + 1));
    int* cellFaceOffsets = (int*)malloc(sizeof(int) * (size_t)(nFine + 1));
    int* nNbrs = (int*)calloc((size_t)(nFine + 1), sizeof(int));
    int* coarseCellMap = (int*)malloc(sizeof(int) * (size_t)(nFine > 0 ? nFine : 1));
    int f, c, nCoarse = 0;

    for (f = 0; f < nFaces; f++) nNbrs[upperAddr[f]]++;
    for (f = 0; f < nFaces; f++) nNbrs[lowerAddr[f]]++;
    cellFaceOffsets[0] = 0;
    for (c = 0; c < nFine; c++) cellFaceOffsets[c + 1] = cellFaceOffsets[c] + nNbrs[c];
    for (c = 0; c < nFine; c++) nNbrs[c] = 0;
    for (f = 0; f < nFaces; f++) {
        cellFaces[cellFaceOffsets[upperAddr[f]] + nNbrs[upperAddr[f]]] = f;
        nNbrs[upperAddr[f]]++;
    }
    for (f = 0; f < nFaces; f++) {
        cellFaces[cellFaceOffsets[lowerAddr[f]] + nNbrs[lowerAddr[f]]] = f;
        nNbrs[lowerAddr[f]]++;
    }
    for (c = 0; c < nFine; c++) coarseCellMap[c] = -1;

    for (c = 0; c < nFine; c++) {
        if (coarseCellMap[c] < 0) {
            int matchFaceNo = -1, fo;
            double maxFaceWeight = -ORC_SCALAR_GREAT;
            for (fo = cellFaceOffsets[c]; fo < cellFaceOffsets[c + 1]; fo++) {
                int facei = cellFaces[fo];
                if (coarseCellMap[upperAddr[facei]] < 0
                    && coarseCellMap[lowerAddr[facei]] < 0
                    && faceWeights[facei] > maxFaceWeight) {
                    matchFaceNo = facei;
                    maxFaceWeight = faceWeights[facei];
                }
            }
            if (matchFaceNo >= 0) {
                coarseCellMap[upperAddr[matchFaceNo]] = nCoarse;
                coarseCellMap[lowerAddr[matchFaceNo]] = nCoarse;
                nCoarse++;
            } else {
                int clusterMatchFaceNo = -1;
                double clusterMaxFaceCoeff = -ORC_SCALAR_GREAT;
                for (fo = cellFaceOffsets[c]; fo < cellFaceOffsets[c + 1]; fo++) {
                    int facei = cellFaces[fo];
                    if (faceWeights[facei] > clusterMaxFaceCoeff) {
                        clusterMatchFaceNo = facei;
                        clusterMaxFaceCoeff = faceWeights[facei];
                    }
                }
                if (clusterMatchFaceNo >= 0) {
                    int a = coarseCellMap[upperAddr[clusterMatchFaceNo]];
                    int b = coarseCellMap[lowerAddr[clusterMatchFaceNo]];
                    coarseCellMap[c] = a > b ? a : b;
                }
            }
        }
    }
    for (c = 0; c < nFine; c++) {
        if (coarseCellMap[c] < 0) {
            coarseCellMap[c] = nCoarse;
            nCoarse++;
        }
    }
    /* reverse the map ordering (pairGAMGAgglomerate.C:186-195) */
    nCoarse--;
    for (c = 0; c < nFine; c++) coarseCellMap[c] = nCoarse - coarseCellMap[c];
    nCoarse++;

    free(cellFaces); free(cellFaceOffsets); free(nNbrs);
    *nCoarseOut = nCoarse;
    return coarseCellMap;
}

/* GAMGAgglomerateLduAddressing.C:31-190 — coarse owner/neighbour + face map */
static void agglomerate_addressing(int nCoarseCells, int nFineFaces,
                                   const int* lowerAddr, const int* upperAddr,
                                   const int* restrictMap,
                                   int** faceRestrictOut, int* nCoarseFacesOut,
                                   int** coarseOwnerOut, int** coarseNeighbourOut)
{
    int maxNnbrs = 10;
    int* cCellnFaces = (int*)calloc((size_t)(nCoarseCells > 0 ? nCoarseCells : 1), sizeof(int));
    int* cCellFaces = (int*)malloc(sizeof(int) * (size_t)maxNnbrs * (size_t)(nCoarseCells > 0 ? nCoarseCells : 1));
    int* faceRestrictAddr = (int*)malloc(sizeof(int) * (size_t)(nFineFaces > 0 ? nFineFaces : 1));
    int* initCoarseNeighb = (int*)malloc(sizeof(int) * (size_t)(nFineFaces > 0 ? nFineFaces : 1));
    int nCoarseFaces = 0, f, i, j, cci;
    int *coarseOwner, *coarseNeighbour, *coarseFaceMap, coarseFacei;

    for (f = 0; f < nFineFaces; f++) {
        int rmUpper = restrictMap[upperAddr[f]];
        int rmLower = restrictMap[lowerAddr[f]];
        if (rmUpper == rmLower) {
            faceRestrictAddr[f] = -(rmUpper + 1);
        } else {
            int cOwn = rmUpper, cNei = rmLower, nbrFound = 0;
            int* ccFaces;
            if (rmUpper > rmLower) { cOwn = rmLower; cNei = rmUpper; }
            ccFaces = &cCellFaces[(size_t)maxNnbrs * (size_t)cOwn];
            for (i = 0; i < cCellnFaces[cOwn]; i++) {
                if (initCoarseNeighb[ccFaces[i]] == cNei) {
                    nbrFound = 1;
                    faceRestrictAddr[f] = ccFaces[i];
                    break;
                }
            }
            if (!nbrFound) {
                if (cCellnFaces[cOwn] >= maxNnbrs) {
                    int oldMax = maxNnbrs;
                    maxNnbrs *= 2;
                    cCellFaces = (int*)realloc(cCellFaces, sizeof(int) * (size_t)maxNnbrs * (size_t)nCoarseCells);
                    for (i = nCoarseCells - 1; i >= 0; i--) {
                        int* oldp = &cCellFaces[(size_t)oldMax * (size_t)i];
                        int* newp = &cCellFaces[(size_t)maxNnbrs * (size_t)i];
                        for (j = cCellnFaces[i] - 1; j >= 0; j--) newp[j] = oldp[j];
                    }
                    ccFaces = &cCellFaces[(size_t)maxNnbrs * (size_t)cOwn];
                }
                ccFaces[cCellnFaces[cOwn]] = nCoarseFaces;
                initCoarseNeighb[nCoarseFaces] = cNei;
                faceRestrictAddr[f] = nCoarseFaces;
                cCellnFaces[cOwn]++;
                nCoarseFaces++;
            }
        }
    }
    coarseOwner = (int*)malloc(sizeof(int) * (size_t)(nCoarseFaces > 0 ? nCoarseFaces : 1));
    coarseNeighbour = (int*)malloc(sizeof(int) * (size_t)(nCoarseFaces > 0 ? nCoarseFaces : 1));
    coarseFaceMap = (int*)malloc(sizeof(int) * (size_t)(nCoarseFaces > 0 ? nCoarseFaces : 1));
    coarseFacei = 0;
    for (cci = 0; cci < nCoarseCells; cci++) {
        int* cFaces = &cCellFaces[(size_t)maxNnbrs * (size_t)cci];
        for (i = 0; i < cCellnFaces[cci]; i++) {
            coarseOwner[coarseFacei] = cci;
            coarseNeighbour[coarseFacei] = initCoarseNeighb[cFaces[i]];
            coarseFaceMap[cFaces[i]] = coarseFacei;
            coarseFacei++;
        }
    }
    for (f = 0; f < nFineFaces; f++)
        if (faceRestrictAddr[f] >= 0) faceRestrictAddr[f] = coarseFaceMap[faceRestrictAddr[f]];

    free(cCellnFaces); free(cCellFaces); free(initCoarseNeighb); free(coarseFaceMap);
    *faceRestrictOut = faceRestrictAddr;
    *nCoarseFacesOut = nCoarseFaces;
    *coarseOwnerOut = coarseOwner;
    *coarseNeighbourOut = coarseNeighbour;
}

/* coarse processor interface: processorGAMGInterface.C:47-126.  Coarse faces
 * are the distinct (master coarse cell, slave coarse cell) pairs in order of
 * first appearance; master = lower region index. */
static void agglomerate_interface(const orc_iface* fine, int myRegion,
                                  const int* localRestrict /* per fine iface face */,
                                  const int* nbrRestrict,
                                  orc_iface* coarse)
{
    int n = fine->n, i, j, nc = 0;
    int* fc = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* fra = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* pa = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* pb = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    /* simple hash on the pair to keep this O(n) */
    int hsize = 1; int* htab;
    while (hsize < 4 * n + 8) hsize <<= 1;
    htab = (int*)malloc(sizeof(int) * (size_t)hsize);
    for (i = 0; i < hsize; i++) htab[i] = -1;
    for (i = 0; i < n; i++) {
        int a, b; unsigned h;
        if (myRegion < fine->nbrRegion) { a = localRestrict[i]; b = nbrRestrict[i]; }
        else { a = nbrRestrict[i]; b = localRestrict[i]; }
        h = ((unsigned)a * 2654435761u) ^ ((unsigned)b * 40503u + 0x9e3779b9u);
        h &= (unsigned)(hsize - 1);
        j = -1;
        while (htab[h] >= 0) {
            int k = htab[h];
            if (pa[k] == a && pb[k] == b) { j = k; break; }
            h = (h + 1) & (unsigned)(hsize - 1);
        }
        if (j < 0) {
            pa[nc] = a; pb[nc] = b;
            fc[nc] = localRestrict[i];
            htab[h] = nc;
            fra[i] = nc;
            nc++;
        } else {
            fra[i] = j;
        }
    }
    free(htab); free(pa); free(pb);
    memset(coarse, 0, sizeof(*coarse));
    coarse->nbrRegion = fine->nbrRegion;
    coarse->nbrInterface = fine->nbrInterface;
    coarse->n = nc;
    coarse->faceCells = fc;
    coarse->faceRestrict = fra;
    coarse->nFineFaces = n;
    coarse->owned = 1;
}

/* GAMGInterface.C:61-75 agglomerateCoeffs */
static double* agglomerate_coeffs(const orc_iface* coarse, const double* fineCoeffs)
{
    double* c = (double*)calloc((size_t)(coarse->n > 0 ? coarse->n : 1), sizeof(double));
    int i;
    for (i = 0; i < coarse->nFineFaces; i++) c[coarse->faceRestrict[i]] += fineCoeffs[i];
    return c;
}

/* GAMGSolverAgglomerateMatrix.C:31-207 — coefficients of one coarse level */
static void agglomerate_matrix(const orc_matrix* fine, orc_matrix* coarse,
                               const int* restrictAddr, const int* faceRestrictAddr)
{
    int i, f, p;
    double* cdiag = (double*)calloc((size_t)(coarse->nCells > 0 ? coarse->nCells : 1), sizeof(double));
    double* cupper = (double*)calloc((size_t)(coarse->nFaces > 0 ? coarse->nFaces : 1), sizeof(double));
    double* clower = NULL;
    /* restrictField: GAMGAgglomerationTemplates.C:31-59 */
    for (i = 0; i < fine->nCells; i++) cdiag[restrictAddr[i]] += fine->diag[i];
    if (!fine->symmetric) {
        clower = (double*)calloc((size_t)(coarse->nFaces > 0 ? coarse->nFaces : 1), sizeof(double));
        for (f = 0; f < fine->nFaces; f++) {
            int cFace = faceRestrictAddr[f];
            if (cFace >= 0) {
                if (coarse->l[cFace] == restrictAddr[fine->l[f]]) {
                    cupper[cFace] += fine->upper[f];
                    clower[cFace] += fine->lower[f];
                } else {
                    cupper[cFace] += fine->lower[f];
                    clower[cFace] += fine->upper[f];
                }
            } else {
                cdiag[-1 - cFace] += fine->upper[f] + fine->lower[f];
            }
        }
    } else {
        for (f = 0; f < fine->nFaces; f++) {
            int cFace = faceRestrictAddr[f];
            if (cFace >= 0) cupper[cFace] += fine->upper[f];
            else cdiag[-1 - cFace] += 2 * fine->upper[f];
        }
    }
    coarse->own_diag = cdiag; coarse->own_upper = cupper; coarse->own_lower = clower;
    coarse->diag = cdiag; coarse->upper = cupper;
    coarse->symmetric = fine->symmetric;
    coarse->lower = clower ? clower : cupper;
    for (p = 0; p < coarse->nIf; p++) {
        free(coarse->ifs[p].bou); free(coarse->ifs[p].intc);
        coarse->ifs[p].bou = agglomerate_coeffs(&coarse->ifs[p], fine->ifs[p].bou);
        coarse->ifs[p].intc = agglomerate_coeffs(&coarse->ifs[p], fine->ifs[p].intc);
    }
}

static int imin(int a, int b) { return a < b ? a : b; }

/*
 * pairGAMGAgglomerate.C:201-292 (level loop), GAMGAgglomeration.C:53-61
 * (continueAgglomerating: AND over ranks), combineLevels
 * (pairGAMGAgglomerationCombineLevels.C:32-95), then GAMGSolver.C:86-89:
 * agglomerateMatrix for every level.
 */
int orc_gamg_build(orc_world* w, const orc_controls* ctl)
{
    const int R = w->R;
    int r, nPairLevels = 0, nCreated = 0, i, f;
    double** fw = (double**)calloc((size_t)R, sizeof(double*));
    int* fwOwned = (int*)calloc((size_t)R, sizeof(int));

    if (w->cached) {
        /* the agglomeration found on the mesh is used whatever this dictionary says about
         * agglomerator / mergeLevels / nCellsInCoarsestLevel; coefficients are agglomerated afresh */
        free(fw); free(fwOwned);
        nCreated = w->h[0].nLevels;
        for (i = 0; i < nCreated; i++)
            for (r = 0; r < R; r++) {
                orc_hierarchy* h = &w->h[r];
                const orc_matrix* fm = i ? h->level[i - 1] : &w->m[r];
                free(h->level[i]->own_diag); free(h->level[i]->own_upper); free(h->level[i]->own_lower);
                agglomerate_matrix(fm, h->level[i], h->restrictAddr[i], h->faceRestrictAddr[i]);
            }
        if (!ctl->cacheAgglomeration) w->cached = 0;   /* this solver deletes it when it is done */
        w->hBuilt = 1;
        return 0;
    }
    for (r = 0; r < R; r++) hierarchy_clear(&w->h[r]);
    w->hBuilt = 0;

    for (r = 0; r < R; r++) {
        const orc_matrix* m = &w->m[r];
        if (ctl->useFaceWeights && w->faceWeights[r]) {
            fw[r] = (double*)w->faceWeights[r];
        } else {
            /* algebraicPairGAMGAgglomeration.C:47-56: mag(matrix.upper()) */
            fw[r] = (double*)malloc(sizeof(double) * (size_t)(m->nFaces > 0 ? m->nFaces : 1));
            for (f = 0; f < m->nFaces; f++) fw[r][f] = fabs(m->upper[f]);
            fwOwned[r] = 1;
        }
    }

    while (nCreated < ORC_MAXLEVELS - 1) {
        int cont = 1;
        int** maps = (int**)calloc((size_t)R, sizeof(int*));
        int* nCoarse = (int*)calloc((size_t)R, sizeof(int));
        for (r = 0; r < R; r++) {
            const orc_matrix* fm = nCreated ? w->h[r].level[nCreated - 1] : &w->m[r];
            maps[r] = pair_agglomerate(&nCoarse[r], fm->nCells, fm->nFaces, fm->l, fm->u, fw[r]);
            if (!(nCoarse[r] >= ctl->nCellsInCoarsestLevel)) cont = 0;
        }
        if (!cont) {
            for (r = 0; r < R; r++) free(maps[r]);
            free(maps); free(nCoarse);
            break;
        }
        /* coarse addressing per region */
        for (r = 0; r < R; r++) {
            orc_hierarchy* h = &w->h[r];
            const orc_matrix* fm = nCreated ? h->level[nCreated - 1] : &w->m[r];
            orc_matrix* cm = (orc_matrix*)calloc(1, sizeof(orc_matrix));
            int nCoarseFaces; int *fra, *cown, *cnei;
            agglomerate_addressing(nCoarse[r], fm->nFaces, fm->l, fm->u, maps[r],
                                   &fra, &nCoarseFaces, &cown, &cnei);
            cm->nCells = nCoarse[r];
            cm->nFaces = nCoarseFaces;
            cm->own_l = cown; cm->own_u = cnei; cm->l = cown; cm->u = cnei;
            calc_addressing(cm);
            h->restrictAddr[nCreated] = maps[r];
            h->faceRestrictAddr[nCreated] = fra;
            h->nFine[nCreated] = fm->nCells;
            h->nFineFaces[nCreated] = fm->nFaces;
            h->nCoarse[nCreated] = nCoarse[r];
            h->level[nCreated] = cm;
            h->nLevels = nCreated + 1;
        }
        /* coarse interfaces (needs every region's restrict map) */
        for (r = 0; r < R; r++) {
            orc_hierarchy* h = &w->h[r];
            const orc_matrix* fm = nCreated ? h->level[nCreated - 1] : &w->m[r];
            orc_matrix* cm = h->level[nCreated];
            int p;
            cm->nIf = fm->nIf;
            cm->ifs = (orc_iface*)calloc((size_t)(fm->nIf > 0 ? fm->nIf : 1), sizeof(orc_iface));
            for (p = 0; p < fm->nIf; p++) {
                const orc_iface* it = &fm->ifs[p];
                const orc_matrix* nfm = nCreated ? w->h[it->nbrRegion].level[nCreated - 1] : &w->m[it->nbrRegion];
                const orc_iface* nb = &nfm->ifs[it->nbrInterface];
                int* loc = (int*)malloc(sizeof(int) * (size_t)(it->n > 0 ? it->n : 1));
                int* nbr = (int*)malloc(sizeof(int) * (size_t)(it->n > 0 ? it->n : 1));
                for (i = 0; i < it->n; i++) {
                    loc[i] = maps[r][it->faceCells[i]];
                    nbr[i] = maps[it->nbrRegion][nb->faceCells[i]];
                }
                agglomerate_interface(it, r, loc, nbr, &cm->ifs[p]);
                free(loc); free(nbr);
            }
        }
        /* restrict the face weights: GAMGAgglomerationTemplates.C:63-83 */
        for (r = 0; r < R; r++) {
            orc_hierarchy* h = &w->h[r];
            const orc_matrix* cm = h->level[nCreated];
            double* nfw = (double*)calloc((size_t)(cm->nFaces > 0 ? cm->nFaces : 1), sizeof(double));
            for (f = 0; f < h->nFineFaces[nCreated]; f++) {
                int cf = h->faceRestrictAddr[nCreated][f];
                if (cf >= 0) nfw[cf] += fw[r][f];
            }
            if (fwOwned[r]) free(fw[r]);
            fw[r] = nfw; fwOwned[r] = 1;
        }
        if (nPairLevels % ctl->mergeLevels) {
            /* combineLevels(curLevel = nCreated) */
            const int cur = nCreated, prev = nCreated - 1;
            for (r = 0; r < R; r++) {
                orc_hierarchy* h = &w->h[r];
                int* curRes = h->restrictAddr[cur];
                int* prevRes = h->restrictAddr[prev];
                int* curF = h->faceRestrictAddr[cur];
                int* prevF = h->faceRestrictAddr[prev];
                orc_matrix* pm = h->level[prev];
                orc_matrix* cm = h->level[cur];
                int p;
                h->nCoarse[prev] = h->nCoarse[cur];
                for (i = 0; i < h->nFineFaces[prev]; i++) {
                    if (prevF[i] >= 0) prevF[i] = curF[prevF[i]];
                    else prevF[i] = -curRes[-prevF[i] - 1] - 1;
                }
                for (i = 0; i < h->nFine[prev]; i++) prevRes[i] = curRes[prevRes[i]];
                /* GAMGInterface::combine (GAMGInterface.C:39-49) */
                for (p = 0; p < pm->nIf; p++) {
                    orc_iface* pi = &pm->ifs[p];
                    orc_iface* ci = &cm->ifs[p];
                    for (i = 0; i < pi->nFineFaces; i++)
                        pi->faceRestrict[i] = ci->faceRestrict[pi->faceRestrict[i]];
                    free(pi->faceCells);
                    pi->faceCells = ci->faceCells; ci->faceCells = NULL;
                    pi->n = ci->n;
                    /* keep pi->nFineFaces: still maps the finer level's faces */
                    ci->owned = 1;
                }
                /* previous level's addressing replaced by the coarser one */
                {
                    orc_iface* keepIfs = pm->ifs; int keepN = pm->nIf;
                    pm->ifs = NULL; pm->nIf = 0;
                    matrix_clear(pm);
                    *pm = *cm;
                    /* cm's interfaces are discarded, pm keeps the combined ones */
                    {
                        int q;
                        for (q = 0; q < cm->nIf; q++) {
                            free(cm->ifs[q].faceCells); free(cm->ifs[q].bou);
                            free(cm->ifs[q].intc); free(cm->ifs[q].faceRestrict);
                        }
                        free(cm->ifs);
                    }
                    pm->ifs = keepIfs; pm->nIf = keepN;
                    free(cm);
                }
                free(curRes); free(curF);
                h->restrictAddr[cur] = NULL; h->faceRestrictAddr[cur] = NULL;
                h->level[cur] = NULL;
                h->nLevels = cur;
            }
        } else {
            nCreated++;
        }
        nPairLevels++;
        free(maps); free(nCoarse);
    }
    for (r = 0; r < R; r++) if (fwOwned[r]) free(fw[r]);
    free(fw); free(fwOwned);

    for (r = 0; r < R; r++) w->h[r].nLevels = nCreated;
    if (nCreated == 0) return -1; /* GAMGSolver.C:108-126: "No coarse levels created" */

    /* coefficients: GAMGSolver.C:86-89 */
    for (i = 0; i < nCreated; i++)
        for (r = 0; r < R; r++) {
            orc_hierarchy* h = &w->h[r];
            const orc_matrix* fm = i ? h->level[i - 1] : &w->m[r];
            agglomerate_matrix(fm, h->level[i], h->restrictAddr[i], h->faceRestrictAddr[i]);
        }
    w->hBuilt = 1;
    w->cached = ctl->cacheAgglomeration ? 1 : 0;
    return 0;
}

int orc_gamg_nlevels(orc_world* w, int r) { return w->h[r].nLevels; }
int orc_gamg_level_ncells(orc_world* w, int r, int level) { return w->h[r].level[level]->nCells; }
int orc_gamg_level_nfaces(orc_world* w, int r, int level) { return w->h[r].level[level]->nFaces; }
int orc_gamg_level_nfine(orc_world* w, int r, int level) { return w->h[r].nFine[level]; }
const int* orc_gamg_restrict(orc_world* w, int r, int level) { return w->h[r].restrictAddr[level]; }
const int* orc_gamg_face_restrict(orc_world* w, int r, int level) { return w->h[r].faceRestrictAddr[level]; }
const int* orc_gamg_level_lower(orc_world* w, int r, int level) { return w->h[r].level[level]->l; }
const int* orc_gamg_level_upper(orc_world* w, int r, int level) { return w->h[r].level[level]->u; }
const double* orc_gamg_level_diag(orc_world* w, int r, int level) { return w->h[r].level[level]->diag; }
const double* orc_gamg_level_upperCoef(orc_world* w, int r, int level) { return w->h[r].level[level]->upper; }
const double* orc_gamg_level_lowerCoef(orc_world* w, int r, int level) { return w->h[r].level[level]->lower; }

/* ------------------------------------------------------------------------- */
/* GAMG solve                                                                 */
/* ------------------------------------------------------------------------- */

/* matrices of all regions at GAMG level i (0 = finest) gathered in one array */
static orc_matrix* level_set(orc_world* w, int lev)
{
    orc_matrix* ms = (orc_matrix*)malloc(sizeof(orc_matrix) * (size_t)w->R);
    int r;
    for (r = 0; r < w->R; r++) ms[r] = lev ? *w->h[r].level[lev - 1] : w->m[r];
    return ms;
}

/* GAMGSolverScale.C:31-75 */
static void gamg_scale(orc_matrix* ms, int R, double** field, double** Acf, double** source)
{
    double num = 0, den = 0, sf;
    int r, c;
    amul_levels(ms, R, Acf, field);
    for (r = 0; r < R; r++) {
        double n = 0, d = 0;
        for (c = 0; c < ms[r].nCells; c++) {
            n += source[r][c] * field[r][c];
            d += Acf[r][c] * field[r][c];
        }
        num = (r == 0) ? n : num + n;
        den = (r == 0) ? d : den + d;
    }
    /* stabilise(y, VSMALL): primitives/Scalar/Scalar.H — y<0 ? y-VSMALL : y+VSMALL */
    sf = num / (den < 0 ? den - ORC_VSMALL : den + ORC_VSMALL);
    for (r = 0; r < R; r++)
        for (c = 0; c < ms[r].nCells; c++)
            field[r][c] = sf * field[r][c] + (source[r][c] - sf * Acf[r][c]) / ms[r].diag[c];
}

/* GAMGSolverInterpolate.C:30-83 */
static void gamg_interpolate(orc_matrix* ms, int R, double** psi, double** Apsi)
{
    int r, c, f;
    for (r = 0; r < R; r++) {
        const orc_matrix* m = &ms[r];
        for (c = 0; c < m->nCells; c++) Apsi[r][c] = 0;
        for (f = 0; f < m->nFaces; f++) {
            Apsi[r][m->u[f]] += m->lower[f] * psi[r][m->l[f]];
            Apsi[r][m->l[f]] += m->upper[f] * psi[r][m->u[f]];
        }
    }
    for (r = 0; r < R; r++) update_interfaces(ms, R, r, Apsi[r], psi, 0, 1.0);
    for (r = 0; r < R; r++)
        for (c = 0; c < ms[r].nCells; c++) psi[r][c] = -Apsi[r][c] / ms[r].diag[c];
}

typedef struct {
    int nLev;                 /* coarse levels */
    orc_matrix** sets;        /* [nLev+1], 0 = finest */
    double*** corr;           /* coarseCorrFields[lev][r] */
    double*** src;            /* coarseSources[lev][r] */
} gamg_state;

static void gamg_state_init(orc_world* w, gamg_state* s)
{
    int i;
    s->nLev = w->h[0].nLevels;
    s->sets = (orc_matrix**)malloc(sizeof(orc_matrix*) * (size_t)(s->nLev + 1));
    s->corr = (double***)malloc(sizeof(double**) * (size_t)s->nLev);
    s->src = (double***)malloc(sizeof(double**) * (size_t)s->nLev);
    for (i = 0; i <= s->nLev; i++) s->sets[i] = level_set(w, i);
    for (i = 0; i < s->nLev; i++) {
        s->corr[i] = alloc_fields(s->sets[i + 1], w->R);
        s->src[i] = alloc_fields(s->sets[i + 1], w->R);
    }
}

static void gamg_state_free(orc_world* w, gamg_state* s)
{
    int i;
    for (i = 0; i < s->nLev; i++) {
        free_fields(s->corr[i], w->R);
        free_fields(s->src[i], w->R);
    }
    for (i = 0; i <= s->nLev; i++) free(s->sets[i]);
    free(s->sets); free(s->corr); free(s->src);
}

static void restrict_field(orc_world* w, int lev, double** cf, double** ff)
{
    int r, i;
    for (r = 0; r < w->R; r++) {
        const orc_hierarchy* h = &w->h[r];
        for (i = 0; i < h->nCoarse[lev]; i++) cf[r][i] = 0;
        for (i = 0; i < h->nFine[lev]; i++) cf[r][h->restrictAddr[lev][i]] += ff[r][i];
    }
}

static void prolong_field(orc_world* w, int lev, double** ff, double** cf)
{
    int r, i;
    for (r = 0; r < w->R; r++) {
        const orc_hierarchy* h = &w->h[r];
        for (i = 0; i < h->nFine[lev]; i++) ff[r][i] = cf[r][h->restrictAddr[lev][i]];
    }
}

/* GAMGSolverSolve.C:430-487 (ICCG = PCG+DIC, BICCG = PBiCG+DILU; ICCG.C:40-109) */
static void gamg_solve_coarsest(orc_world* w, gamg_state* s, const orc_controls* ctl)
{
    const int cl = s->nLev - 1;
    orc_controls cc = *ctl;
    orc_perf perf;
    int r, c, asym = !s->sets[cl + 1][0].symmetric;
    if (ctl->solver != ORC_SOLVER_GAMG) {
        /* GAMG used as a preconditioner reads tolerance/relTol from its own
         * sub-dictionary (GAMGPreconditioner.C:44-63 -> GAMGSolver ctor) */
        cc.tolerance = ctl->precTolerance;
        cc.relTol = ctl->precRelTol;
    }
    for (r = 0; r < w->R; r++)
        for (c = 0; c < s->sets[cl + 1][r].nCells; c++) s->corr[cl][r][c] = 0;
    cc.maxIter = 1000; /* ICCG.C:49-68 builds a dict with only tolerance/relTol */
    cc.preconditioner = asym ? ORC_PRECOND_DILU : ORC_PRECOND_DIC;
    krylov_solve(w, s->sets[cl + 1], w->R, &cc, s->corr[cl], s->src[cl], &perf, NULL, 0, asym);
}

/* GAMGSolverSolve.C:120-364 */
static void gamg_vcycle(orc_world* w, gamg_state* s, const orc_controls* ctl,
                        double** psi, double** source, double** Apsi,
                        double** finestCorrection, double** finestResidual)
{
    const int R = w->R;
    const int coarsestLevel = s->nLev - 1;
    const int scaleCorrection = ctl->scaleCorrection < 0 ? w->m[0].symmetric : ctl->scaleCorrection;
    int lev, r, c;

    restrict_field(w, 0, s->src[0], finestResidual);

    for (lev = 0; lev < coarsestLevel; lev++) {
        if (ctl->nPreSweeps) {
            for (r = 0; r < R; r++)
                for (c = 0; c < s->sets[lev + 1][r].nCells; c++) s->corr[lev][r][c] = 0.0;
            smooth_levels(s->sets[lev + 1], R, ctl->smoother, s->corr[lev], s->src[lev],
                          imin(ctl->nPreSweeps + ctl->preSweepsLevelMultiplier * lev, ctl->maxPreSweeps));
            /* ACf is a sub-field of Apsi */
            if (scaleCorrection && lev < coarsestLevel - 1)
                gamg_scale(s->sets[lev + 1], R, s->corr[lev], Apsi, s->src[lev]);
            amul_levels(s->sets[lev + 1], R, Apsi, s->corr[lev]);
            for (r = 0; r < R; r++)
                for (c = 0; c < s->sets[lev + 1][r].nCells; c++) s->src[lev][r][c] -= Apsi[r][c];
        }
        restrict_field(w, lev + 1, s->src[lev + 1], s->src[lev]);
    }

    gamg_solve_coarsest(w, s, ctl);

    for (lev = coarsestLevel - 1; lev >= 0; lev--) {
        /* preSmoothedCoarseCorrField lives in finestCorrection */
        if (ctl->nPreSweeps)
            for (r = 0; r < R; r++)
                memcpy(finestCorrection[r], s->corr[lev][r],
                       sizeof(double) * (size_t)s->sets[lev + 1][r].nCells);
        prolong_field(w, lev + 1, s->corr[lev], s->corr[lev + 1]);
        if (ctl->interpolateCorrection)
            gamg_interpolate(s->sets[lev + 1], R, s->corr[lev], Apsi);
        if (scaleCorrection && lev < coarsestLevel - 1)
            gamg_scale(s->sets[lev + 1], R, s->corr[lev], Apsi, s->src[lev]);
        if (ctl->nPreSweeps)
            for (r = 0; r < R; r++)
                for (c = 0; c < s->sets[lev + 1][r].nCells; c++)
                    s->corr[lev][r][c] += finestCorrection[r][c];
        smooth_levels(s->sets[lev + 1], R, ctl->smoother, s->corr[lev], s->src[lev],
                      imin(ctl->nPostSweeps + ctl->postSweepsLevelMultiplier * lev, ctl->maxPostSweeps));
    }

    prolong_field(w, 0, finestCorrection, s->corr[0]);
    if (ctl->interpolateCorrection)
        gamg_interpolate(s->sets[0], R, finestCorrection, Apsi);
    if (scaleCorrection)
        gamg_scale(s->sets[0], R, finestCorrection, Apsi, finestResidual);
    for (r = 0; r < R; r++)
        for (c = 0; c < w->m[r].nCells; c++) psi[r][c] += finestCorrection[r][c];
    smooth_levels(s->sets[0], R, ctl->smoother, psi, source, ctl->nFinestSweeps);
}

/* GAMGSolverSolve.C:34-117 */
static int gamg_solve(orc_world* w, const orc_controls* ctl, double** psi, double** source,
                      orc_perf* perf, double* hist, int histCap)
{
    const int R = w->R;
    double** Apsi = alloc_fields(w->m, R);
    double** finestCorrection = alloc_fields(w->m, R);
    double** finestResidual = alloc_fields(w->m, R);
    double normFactor;
    int r, c, nh = 0;
    memset(perf, 0, sizeof(*perf));
    if (orc_gamg_build(w, ctl)) { /* the reference rebuilds coefficients every solve */
        free_fields(Apsi, R); free_fields(finestCorrection, R); free_fields(finestResidual, R);
        return -1;
    }
    amul_levels(w->m, R, Apsi, psi);
    normFactor = normFactor_levels(w->m, R, psi, source, Apsi);
    for (r = 0; r < R; r++)
        for (c = 0; c < w->m[r].nCells; c++) finestResidual[r][c] = source[r][c] - Apsi[r][c];
    perf->initialResidual = gSumMag_levels(w->m, R, finestResidual) / normFactor;
    perf->finalResidual = perf->initialResidual;
    hist_push(hist, histCap, &nh, perf->finalResidual);
    if (!check_convergence(perf, ctl->tolerance, ctl->relTol)) {
        gamg_state s;
        gamg_state_init(w, &s);
        do {
            gamg_vcycle(w, &s, ctl, psi, source, Apsi, finestCorrection, finestResidual);
            amul_levels(w->m, R, Apsi, psi);
            for (r = 0; r < R; r++)
                for (c = 0; c < w->m[r].nCells; c++) {
                    finestResidual[r][c] = source[r][c];
                    finestResidual[r][c] -= Apsi[r][c];
                }
            perf->finalResidual = gSumMag_levels(w->m, R, finestResidual) / normFactor;
            hist_push(hist, histCap, &nh, perf->finalResidual);
        } while (++perf->nIterations < ctl->maxIter
                 && !check_convergence(perf, ctl->tolerance, ctl->relTol));
        gamg_state_free(w, &s);
    }
    free_fields(Apsi, R); free_fields(finestCorrection, R); free_fields(finestResidual, R);
    return 0;
}

/* GAMGPreconditioner.C:81-128 */
static int gamg_precondition(orc_world* w, const orc_controls* ctl, double** wA, double** rA)
{
    const int R = w->R;
    double** AwA = alloc_fields(w->m, R);
    double** finestCorrection = alloc_fields(w->m, R);
    double** finestResidual = alloc_fields(w->m, R);
    gamg_state s;
    int r, c, cycle;
    if (!w->hBuilt && orc_gamg_build(w, ctl)) return -1;
    gamg_state_init(w, &s);
    for (r = 0; r < R; r++)
        for (c = 0; c < w->m[r].nCells; c++) { wA[r][c] = 0.0; finestResidual[r][c] = rA[r][c]; }
    for (cycle = 0; cycle < ctl->nVcycles; cycle++) {
        gamg_vcycle(w, &s, ctl, wA, rA, AwA, finestCorrection, finestResidual);
        if (cycle < ctl->nVcycles - 1) {
            amul_levels(w->m, R, AwA, wA);
            for (r = 0; r < R; r++)
                for (c = 0; c < w->m[r].nCells; c++) {
                    finestResidual[r][c] = rA[r][c];
                    finestResidual[r][c] -= AwA[r][c];
                }
        }
    }
    gamg_state_free(w, &s);
    free_fields(AwA, R); free_fields(finestCorrection, R); free_fields(finestResidual, R);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* lduMatrix::solver::New dispatch (LM/lduMatrix/lduMatrixSolver.C:40-136)     */
/* ------------------------------------------------------------------------- */

int orc_solve(orc_world* w, const orc_controls* c, double** psi, double** source,
              orc_perf* perf, double* residualHistory, int histCap)
{
    w->hBuilt = 0;
    /* matrix.diagonal() -> diagonalSolver unconditionally (lduMatrixSolver.C:52-66) */
    if (w->diagonalOnly || c->solver == ORC_SOLVER_DIAGONAL) {
        diagonal_solve(w->m, w->R, psi, source, perf);
        return 0;
    }
    switch (c->solver) {
    case ORC_SOLVER_PCG:
        w->failed = 0;
        krylov_solve(w, w->m, w->R, c, psi, source, perf, residualHistory, histCap, 0);
        return w->failed ? -1 : 0;
    case ORC_SOLVER_PBICG:
        /* GAMGPreconditioner has no preconditionT: the reference stops with "Not implemented"
         * in the first iteration (lduMatrix.H:492-505, PBiCG.C:139) */
        if (c->preconditioner == ORC_PRECOND_GAMG) return -1;
        krylov_solve(w, w->m, w->R, c, psi, source, perf, residualHistory, histCap, 1);
        return 0;
    case ORC_SOLVER_SMOOTH:
        smooth_solve(w->m, w->R, c, psi, source, perf, residualHistory, histCap);
        return 0;
    case ORC_SOLVER_GAMG:
        return gamg_solve(w, c, psi, source, perf, residualHistory, histCap);
    }
    return -1;
}
