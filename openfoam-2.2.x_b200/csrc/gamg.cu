// GAMG: pairwise agglomeration (host, once per mesh), Galerkin-by-summation
// coarse matrices (device, every solve) and the V-cycle (device).
//
//   GAMGSolver ctor / solve / Vcycle     solvers/GAMG/GAMGSolver.C:44-127, GAMGSolverSolve.C:34-487
//   scale / interpolate                  solvers/GAMG/GAMGSolverScale.C:31-75, GAMGSolverInterpolate.C:30-83
//   agglomerateMatrix                    solvers/GAMG/GAMGSolverAgglomerateMatrix.C:31-207
//   pair agglomeration                   solvers/GAMG/GAMGAgglomerations/pairGAMGAgglomeration/pairGAMGAgglomerate.C:31-292
//   coarse addressing                    .../GAMGAgglomeration/GAMGAgglomerateLduAddressing.C:31-286
//   level merging                        .../pairGAMGAgglomeration/pairGAMGAgglomerationCombineLevels.C:32-95
//   restrict / prolong                   .../GAMGAgglomeration/GAMGAgglomerationTemplates.C:31-100
//   coarse processor interfaces          solvers/GAMG/interfaces/processorGAMGInterface/processorGAMGInterface.C:47-126
// (paths relative to /root/reference/src/OpenFOAM/matrices/lduMatrix/)
//
// The reference's scatter-add loops (restrictField, agglomerateMatrix) are
// turned into gathers over inverted maps whose lists are kept in ascending
// fine index, i.e. the order the scatter loop adds them in: coarse matrices and
// restricted fields are bit-identical to the reference's.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <utility>

#include "comm.cuh"
#include "epilogue.cuh"
#include "reduce.cuh"
#include "sweeps.h"
#include "gamg_host.h"

namespace ldu {

constexpr int kMaxLevels = 50;             // GAMGAgglomeration.C:74
constexpr double kVSmallG = 1.0e-300;

template <class T>
static int to_device(ldu_context* ctx, T** d, const std::vector<T>& h)
{
    *d = nullptr;
    LDU_CUDA(cudaMalloc((void**)d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    if (!h.empty())
        LDU_CUDA(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return LDU_OK;
}

static void free_level(GamgLevel* L)
{
    if (!L) return;
    cudaFree(L->d_restrict);
    cudaFree(L->d_cellStart);
    cudaFree(L->d_cellFine);
    cudaFree(L->d_faceStart);
    cudaFree(L->d_faceFine);
    cudaFree(L->d_intStart);
    cudaFree(L->d_intFine);
    cudaFree(L->d_ifStart);
    cudaFree(L->d_ifFine);
    cudaFree(L->d_corr);
    cudaFree(L->d_src);
    if (L->coarse) ldu_matrix_destroy(L->coarse);
    delete L;
}

void gamg_free(ldu_matrix* m)
{
    for (GamgLevel* L : m->levels) free_level(L);
    m->levels.clear();
    m->hierarchyValid = false;
}

// neighbour side's restrict map on every coupled face of `fine` (the reference
// ships it with initInternalFieldTransfer, GAMGAgglomerateLduAddressing.C:216-262)
static int exchange_restrict(ldu_matrix* fine, const std::vector<int>& cmap, std::vector<int>& nbrValues)
{
    nbrValues.assign(fine->nIfFaces, 0);
    if (!fine->nIfFaces) return LDU_OK;
    ldu_context* ctx = fine->ctx;
    std::vector<double> asDouble(cmap.begin(), cmap.end());
    double* d = work_vec(fine, W_TMP);
    LDU_CUDA(cudaMemcpyAsync(d, asDouble.data(), asDouble.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LDU_TRY(comm_halo_exchange(fine, d, false));
    std::vector<double> recv(fine->nIfFaces);
    LDU_CUDA(cudaMemcpyAsync(recv.data(), fine->d_recv, recv.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t k = 0; k < recv.size(); k++) nbrValues[k] = (int)recv[k];
    return LDU_OK;
}

// continueAgglomerating: AND over ranks (GAMGAgglomeration.C:53-61)
static int all_ranks_agree(ldu_context* ctx, bool mine, bool& all)
{
    all = mine;
    if (!ctx->comm.connected || ctx->comm.nRanks == 1) return LDU_OK;
    double v = mine ? 1.0 : 0.0;
    LDU_CUDA(cudaMemcpyAsync(ctx->d_red, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    LDU_TRY(comm_allreduce(ctx, ctx->d_red, 1));
    LDU_CUDA(cudaMemcpyAsync(&v, ctx->d_red, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    all = (v > ctx->comm.nRanks - 0.5);
    return LDU_OK;
}

struct HostLevel {  // one pairing step before merging
    std::vector<int> cmap, faceMap, cOwner, cNeighbour;
    int nCoarse = 0;
    std::vector<std::vector<int>> ifCells;       // coarse interface faceCells
    std::vector<std::vector<int>> ifRestrict;    // fine if-face -> coarse if-face
};

static int build_level_device(ldu_matrix* fine, const HostLevel& H, GamgLevel** out)
{
    ldu_context* ctx = fine->ctx;
    GamgLevel* L = new GamgLevel();
    L->nFine = fine->nCells;
    L->nFineFaces = fine->nFaces;
    L->nCoarse = H.nCoarse;
    L->nCoarseFaces = (int)H.cOwner.size();
    L->h_restrict = H.cmap;
    L->h_faceRestrict = H.faceMap;
    L->h_ifRestrict = H.ifRestrict;
    LDU_TRY(to_device(ctx, &L->d_restrict, H.cmap));
    // inverted maps, lists in ascending fine index
    std::vector<int> cellStart(H.nCoarse + 1, 0), cellFine(L->nFine);
    for (int i = 0; i < L->nFine; i++) cellStart[H.cmap[i] + 1]++;
    for (int c = 0; c < H.nCoarse; c++) cellStart[c + 1] += cellStart[c];
    {
        std::vector<int> fill(cellStart.begin(), cellStart.end() - 1);
        for (int i = 0; i < L->nFine; i++) cellFine[fill[H.cmap[i]]++] = i;
    }
    std::vector<int> faceStart(L->nCoarseFaces + 1, 0), intStart(H.nCoarse + 1, 0);
    for (int f = 0; f < L->nFineFaces; f++) {
        const int cf = H.faceMap[f];
        if (cf >= 0) faceStart[cf + 1]++;
        else intStart[-1 - cf + 1]++;
    }
    for (int c = 0; c < L->nCoarseFaces; c++) faceStart[c + 1] += faceStart[c];
    for (int c = 0; c < H.nCoarse; c++) intStart[c + 1] += intStart[c];
    std::vector<int> faceFine(faceStart[L->nCoarseFaces]), intFine(intStart[H.nCoarse]);
    {
        std::vector<int> ffill(faceStart.begin(), faceStart.end() - 1), ifill(intStart.begin(), intStart.end() - 1);
        for (int f = 0; f < L->nFineFaces; f++) {
            const int cf = H.faceMap[f];
            if (cf >= 0) {
                // orientation of the fine face relative to the coarse face
                // (GAMGSolverAgglomerateMatrix.C:151-166): flipped faces swap upper/lower
                const bool same = (H.cOwner[cf] == H.cmap[fine->h_l[f]]);
                faceFine[ffill[cf]++] = same ? f : (f | (int)0x80000000);
            } else {
                intFine[ifill[-1 - cf]++] = f;
            }
        }
    }
    LDU_TRY(to_device(ctx, &L->d_cellStart, cellStart));
    LDU_TRY(to_device(ctx, &L->d_cellFine, cellFine));
    LDU_TRY(to_device(ctx, &L->d_faceStart, faceStart));
    LDU_TRY(to_device(ctx, &L->d_faceFine, faceFine));
    LDU_TRY(to_device(ctx, &L->d_intStart, intStart));
    LDU_TRY(to_device(ctx, &L->d_intFine, intFine));

    // coarse matrix object (addressing + interfaces)
    std::vector<int> sizes, nbrRank, nbrIf;
    std::vector<const int*> cellsPtr;
    for (size_t p = 0; p < fine->ifs.size(); p++) {
        sizes.push_back((int)H.ifCells[p].size());
        cellsPtr.push_back(H.ifCells[p].data());
        nbrRank.push_back(fine->ifs[p].nbrRank);
        nbrIf.push_back(fine->ifs[p].nbrInterface);
    }
    LDU_TRY(ldu_matrix_create(ctx, H.nCoarse, L->nCoarseFaces, H.cOwner.data(), H.cNeighbour.data(),
                              (int)fine->ifs.size(), sizes.data(), cellsPtr.data(), nbrRank.data(),
                              nbrIf.data(), &L->coarse));
    L->coarse->isCoarse = true;
    // interface coefficient gather lists: coarse concatenated index -> fine concatenated indices
    if (fine->nIfFaces) {
        std::vector<int> ifStart(L->coarse->nIfFaces + 1, 0), ifFine(fine->nIfFaces);
        for (size_t p = 0; p < fine->ifs.size(); p++)
            for (int i = 0; i < fine->ifs[p].n; i++)
                ifStart[L->coarse->ifs[p].offset + H.ifRestrict[p][i] + 1]++;
        for (int k = 0; k < L->coarse->nIfFaces; k++) ifStart[k + 1] += ifStart[k];
        std::vector<int> fill(ifStart.begin(), ifStart.end() - 1);
        for (size_t p = 0; p < fine->ifs.size(); p++)
            for (int i = 0; i < fine->ifs[p].n; i++)
                ifFine[fill[L->coarse->ifs[p].offset + H.ifRestrict[p][i]]++] = fine->ifs[p].offset + i;
        LDU_TRY(to_device(ctx, &L->d_ifStart, ifStart));
        LDU_TRY(to_device(ctx, &L->d_ifFine, ifFine));
    }
    LDU_CUDA(cudaMalloc((void**)&L->d_corr, std::max(H.nCoarse, 1) * sizeof(double)));
    LDU_CUDA(cudaMalloc((void**)&L->d_src, std::max(H.nCoarse, 1) * sizeof(double)));
    LDU_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = L;
    return LDU_OK;
}

// a matrix-less carrier for the interfaces of the mesh currently being paired:
// lets the restrict-map exchange reuse the halo kernels
static int make_carrier(ldu_context* ctx, int nCells, const std::vector<std::vector<int>>& ifCells,
                        const ldu_matrix* top, ldu_matrix** out)
{
    std::vector<int> sizes, nbrRank, nbrIf;
    std::vector<const int*> ptrs;
    for (size_t p = 0; p < ifCells.size(); p++) {
        sizes.push_back((int)ifCells[p].size());
        ptrs.push_back(ifCells[p].data());
        nbrRank.push_back(top->ifs[p].nbrRank);
        nbrIf.push_back(top->ifs[p].nbrInterface);
    }
    return ldu_matrix_create(ctx, nCells, 0, nullptr, nullptr, (int)sizes.size(), sizes.data(), ptrs.data(),
                             nbrRank.data(), nbrIf.data(), out);
}

// pairGAMGAgglomeration::agglomerate(mesh, faceWeights): pairGAMGAgglomerate.C:201-292
int gamg_build(ldu_matrix* m, const ldu_controls* c)
{
    // a hierarchy handed over by the host (ldu_gamg_set_level: the reference's own GAMGAgglomeration)
    // is used as it is; the agglomeration keys of the controls are the host's business then
    if (m->externalHierarchy && m->hierarchyValid) return LDU_OK;
    if (m->hierarchyValid && c->cacheAgglomeration
        && m->hierarchyControls.nCellsInCoarsestLevel == c->nCellsInCoarsestLevel
        && m->hierarchyControls.mergeLevels == c->mergeLevels
        && m->hierarchyControls.useFaceWeights == c->useFaceWeights)
        return LDU_OK;
    gamg_free(m);
    ldu_context* ctx = m->ctx;
    if (c->mergeLevels < 1) {
        set_error("GAMG: mergeLevels must be >= 1");
        return LDU_EINVAL;
    }
    std::vector<double> w;
    if (c->useFaceWeights) {
        if ((int)m->h_faceWeights.size() != m->nFaces) {
            set_error("GAMG: faceAreaPair agglomeration needs ldu_matrix_set_face_weights");
            return LDU_EINVAL;
        }
        w = m->h_faceWeights;
    } else {  // algebraicPair: mag(upper) (algebraicPairGAMGAgglomeration.C:47-56)
        w.resize(m->nFaces);
        LDU_CUDA(cudaStreamSynchronize(ctx->stream));
        if (m->nFaces)
            LDU_CUDA(cudaMemcpy(w.data(), m->d_upper, m->nFaces * sizeof(double), cudaMemcpyDeviceToHost));
        for (double& x : w) x = std::fabs(x);
    }

    // mesh currently being paired
    int curN = m->nCells;
    std::vector<int> curL = m->h_l, curU = m->h_u;
    std::vector<std::vector<int>> curIfCells(m->ifs.size());
    if (m->nIfFaces) {
        std::vector<int> cells(m->nIfFaces);
        LDU_CUDA(cudaMemcpy(cells.data(), m->d_ifCells, cells.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t p = 0; p < m->ifs.size(); p++)
            curIfCells[p].assign(cells.begin() + m->ifs[p].offset,
                                 cells.begin() + m->ifs[p].offset + m->ifs[p].n);
    }

    std::vector<HostLevel> HL;   // created (merged) levels
    int nPairLevels = 0;
    while ((int)HL.size() < kMaxLevels - 1) {
        HostLevel H;
        H.cmap = pair_cluster(curN, curL, curU, w, H.nCoarse);
        bool cont;
        LDU_TRY(all_ranks_agree(ctx, H.nCoarse >= c->nCellsInCoarsestLevel, cont));
        if (!cont) break;
        coarse_addressing(H.nCoarse, curL, curU, H.cmap, H.faceMap, H.cOwner, H.cNeighbour);
        H.ifCells.resize(curIfCells.size());
        H.ifRestrict.resize(curIfCells.size());
        if (m->nIfFaces) {
            ldu_matrix* carrier = nullptr;
            LDU_TRY(make_carrier(ctx, curN, curIfCells, m, &carrier));
            std::vector<int> nbrVals;
            int rc = exchange_restrict(carrier, H.cmap, nbrVals);
            if (rc == LDU_OK) {
                for (size_t p = 0; p < curIfCells.size(); p++) {
                    std::vector<int> local(curIfCells[p].size()), nbr(curIfCells[p].size());
                    for (size_t i = 0; i < local.size(); i++) {
                        local[i] = H.cmap[curIfCells[p][i]];
                        nbr[i] = nbrVals[carrier->ifs[p].offset + i];
                    }
                    agglomerate_interface(ctx->comm.rank, m->ifs[p].nbrRank, local, nbr, H.ifCells[p],
                                          H.ifRestrict[p]);
                }
            }
            ldu_matrix_destroy(carrier);
            LDU_TRY(rc);
        }
        // restrict the face weights (GAMGAgglomerationTemplates.C:63-83)
        std::vector<double> cw(H.cOwner.size(), 0.0);
        for (size_t f = 0; f < H.faceMap.size(); f++)
            if (H.faceMap[f] >= 0) cw[H.faceMap[f]] += w[f];
        w.swap(cw);
        curN = H.nCoarse;
        curL = H.cOwner;
        curU = H.cNeighbour;
        curIfCells = H.ifCells;

        if (nPairLevels % c->mergeLevels) {
            // combineLevels (pairGAMGAgglomerationCombineLevels.C:32-95): fold this
            // pairing into the previous level
            HostLevel& P = HL.back();
            for (int& v : P.faceMap) v = (v >= 0) ? H.faceMap[v] : -H.cmap[-v - 1] - 1;
            for (int& v : P.cmap) v = H.cmap[v];
            for (size_t p = 0; p < P.ifRestrict.size(); p++)
                for (int& v : P.ifRestrict[p]) v = H.ifRestrict[p][v];
            P.ifCells = H.ifCells;
            P.cOwner = H.cOwner;
            P.cNeighbour = H.cNeighbour;
            P.nCoarse = H.nCoarse;
        } else {
            HL.push_back(std::move(H));
        }
        nPairLevels++;
    }
    if (HL.empty()) {
        set_error("GAMG: no coarse levels created, matrix too small or nCellsInCoarsestLevel too large "
                  "(GAMGSolver.C:108-126)");
        return LDU_EINVAL;
    }
    ldu_matrix* fine = m;
    for (const HostLevel& H : HL) {
        GamgLevel* L = nullptr;
        LDU_TRY(build_level_device(fine, H, &L));
        m->levels.push_back(L);
        fine = L->coarse;
    }
    m->hierarchyValid = true;
    m->hierarchyControls = *c;
    return LDU_OK;
}

// ---------------------------------------------------------------------------
// device: coefficient agglomeration (every solve)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) agg_diag_kernel(
    int nCoarse, const int* __restrict__ cellStart, const int* __restrict__ cellFine,
    const int* __restrict__ intStart, const int* __restrict__ intFine, const double* __restrict__ fDiag,
    const double* __restrict__ fUpper, const double* __restrict__ fLower, bool asym, double* __restrict__ cDiag)
{
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= nCoarse) return;
    double acc = 0.0;  // restrictField: cf = 0; cf[map[i]] += ff[i]
    for (int k = cellStart[c]; k < cellStart[c + 1]; k++) acc = __dadd_rn(acc, fDiag[cellFine[k]]);
    for (int k = intStart[c]; k < intStart[c + 1]; k++) {
        const int f = intFine[k];
        // sym: += 2*fineUpper ; asym: += fineUpper + fineLower  (GAMGSolverAgglomerateMatrix.C:175-177,201)
        acc = __dadd_rn(acc, asym ? __dadd_rn(fUpper[f], fLower[f]) : __dmul_rn(2.0, fUpper[f]));
    }
    cDiag[c] = acc;
}

__global__ void __launch_bounds__(kBlock) agg_faces_kernel(
    int nCoarseFaces, const int* __restrict__ faceStart, const int* __restrict__ faceFine,
    const double* __restrict__ fUpper, const double* __restrict__ fLower, bool asym,
    double* __restrict__ cUpper, double* __restrict__ cLower)
{
    const int cf = blockIdx.x * kBlock + threadIdx.x;
    if (cf >= nCoarseFaces) return;
    double up = 0.0, lo = 0.0;
    for (int k = faceStart[cf]; k < faceStart[cf + 1]; k++) {
        const int e = faceFine[k];
        const int f = e & 0x7fffffff;
        const bool flipped = e < 0;
        if (!asym) {
            up = __dadd_rn(up, fUpper[f]);
        } else if (!flipped) {
            up = __dadd_rn(up, fUpper[f]);
            lo = __dadd_rn(lo, fLower[f]);
        } else {
            up = __dadd_rn(up, fLower[f]);
            lo = __dadd_rn(lo, fUpper[f]);
        }
    }
    cUpper[cf] = up;
    if (asym) cLower[cf] = lo;
}

// GAMGInterface::agglomerateCoeffs (GAMGInterface.C:61-75)
__global__ void __launch_bounds__(kBlock) agg_iface_kernel(int nCoarseIf, const int* __restrict__ ifStart,
                                                            const int* __restrict__ ifFine,
                                                            const double* __restrict__ fBou,
                                                            const double* __restrict__ fInt,
                                                            double* __restrict__ cBou, double* __restrict__ cInt)
{
    const int k = blockIdx.x * kBlock + threadIdx.x;
    if (k >= nCoarseIf) return;
    double b = 0.0, in = 0.0;
    for (int e = ifStart[k]; e < ifStart[k + 1]; e++) {
        b = __dadd_rn(b, fBou[ifFine[e]]);
        in = __dadd_rn(in, fInt[ifFine[e]]);
    }
    cBou[k] = b;
    cInt[k] = in;
}

static int agglomerate_coefficients(ldu_matrix* m)
{
    cudaStream_t st = m->ctx->stream;
    ldu_matrix* fine = m;
    for (GamgLevel* L : m->levels) {
        ldu_matrix* cm = L->coarse;
        const bool asym = !fine->symmetric;
        if (asym && !cm->ownLower) {
            LDU_CUDA(cudaMalloc((void**)&cm->d_lower, std::max(cm->nFaces, 1) * sizeof(double)));
            cm->ownLower = true;
        } else if (!asym && cm->ownLower) {
            LDU_CUDA(cudaStreamSynchronize(st));
            cudaFree(cm->d_lower);
            cm->ownLower = false;
        }
        if (!asym) cm->d_lower = cm->d_upper;
        cm->symmetric = !asym;
        agg_diag_kernel<<<(L->nCoarse + kBlock - 1) / kBlock, kBlock, 0, st>>>(
            L->nCoarse, L->d_cellStart, L->d_cellFine, L->d_intStart, L->d_intFine, fine->d_diag, fine->d_upper,
            fine->d_lower, asym, cm->d_diag);
        count_launch();
        if (L->nCoarseFaces) {
            agg_faces_kernel<<<(L->nCoarseFaces + kBlock - 1) / kBlock, kBlock, 0, st>>>(
                L->nCoarseFaces, L->d_faceStart, L->d_faceFine, fine->d_upper, fine->d_lower, asym, cm->d_upper,
                cm->d_lower);
            count_launch();
        }
        if (cm->nIfFaces) {
            agg_iface_kernel<<<(cm->nIfFaces + kBlock - 1) / kBlock, kBlock, 0, st>>>(
                cm->nIfFaces, L->d_ifStart, L->d_ifFine, fine->d_bou, fine->d_int, cm->d_bou, cm->d_int);
            count_launch();
        }
        LDU_CUDA(cudaGetLastError());
        cm->haveCoeffs = true;
        fine = cm;
    }
    return LDU_OK;
}

// ---------------------------------------------------------------------------
// device: restrict / prolong / scale / interpolate
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) restrict_kernel(int nCoarse, const int* __restrict__ cellStart,
                                                           const int* __restrict__ cellFine,
                                                           const double* __restrict__ ff, double* __restrict__ cf)
{
    const int c = blockIdx.x * kBlock + threadIdx.x;
    if (c >= nCoarse) return;
    double acc = 0.0;
    for (int k = cellStart[c]; k < cellStart[c + 1]; k++) acc = __dadd_rn(acc, ff[cellFine[k]]);
    cf[c] = acc;
}

struct ProlongMap {
    double* ff;
    const double* cf;
    const int* map;
    __device__ void operator()(int i) const { ff[i] = cf[map[i]]; }
};

struct ScaleDotMap {  // GAMGSolverScale.C:51-58
    const double* source;
    const double* field;
    const double* Acf;
    __device__ void operator()(int i, double (&acc)[2]) const
    {
        acc[0] = __dadd_rn(acc[0], __dmul_rn(source[i], field[i]));
        acc[1] = __dadd_rn(acc[1], __dmul_rn(Acf[i], field[i]));
    }
};

struct EpiScale {  // sf = num/stabilise(den, VSMALL)  (GAMGSolverScale.C:60-62, Scalar.H:290-300)
    SolverScalars* dst;
    __device__ void operator()(SolverScalars*, const double* t) const
    {
        const double den = t[1] >= 0 ? __dadd_rn(t[1], kVSmallG) : __dsub_rn(t[1], kVSmallG);
        dst->scaleNum = t[0];
        dst->scaleDen = t[1];
        dst->sf = __ddiv_rn(t[0], den);
    }
};

struct ScaleApplyMap {  // field = sf*field + (source - sf*Acf)/D  (GAMGSolverScale.C:69-74)
    const SolverScalars* S;
    double* field;
    const double* source;
    const double* Acf;
    const double* diag;
    __device__ void operator()(int i) const
    {
        const double sf = S->sf;
        field[i] = __dadd_rn(__dmul_rn(sf, field[i]),
                             __ddiv_rn(__dsub_rn(source[i], __dmul_rn(sf, Acf[i])), diag[i]));
    }
};

struct SubMap {  // a -= b
    double* a;
    const double* b;
    __device__ void operator()(int i) const { a[i] = __dsub_rn(a[i], b[i]); }
};

struct AddMapG {
    double* a;
    const double* b;
    __device__ void operator()(int i) const { a[i] = __dadd_rn(a[i], b[i]); }
};

struct ZeroMap {
    double* a;
    __device__ void operator()(int i) const { a[i] = 0.0; }
};

struct CopyMapG {
    double* a;
    const double* b;
    __device__ void operator()(int i) const { a[i] = b[i]; }
};

struct ResidualFromMap {  // finestResidual = source; finestResidual -= Apsi ; sum|res|
    double* res;
    const double* source;
    const double* Apsi;
    __device__ void operator()(int i, double (&acc)[1]) const
    {
        const double r = __dsub_rn(source[i], Apsi[i]);
        res[i] = r;
        acc[0] = __dadd_rn(acc[0], fabs(r));
    }
};

struct NegDivMap {  // psi = -Apsi/diag (GAMGSolverInterpolate.C:78-82)
    double* psi;
    const double* Apsi;
    const double* diag;
    __device__ void operator()(int i) const { psi[i] = __ddiv_rn(-Apsi[i], diag[i]); }
};

static int restrict_field(ldu_matrix* any, GamgLevel* L, const double* ff, double* cf)
{
    restrict_kernel<<<(L->nCoarse + kBlock - 1) / kBlock, kBlock, 0, any->ctx->stream>>>(
        L->nCoarse, L->d_cellStart, L->d_cellFine, ff, cf);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

static int gamg_scale(ldu_matrix* A, double* field, double* Acf, const double* source)
{
    LDU_TRY(k_amul(A, Acf, field, false));
    LDU_TRY((launch_map_reduce<2, false>(A, A->nCells, ScaleDotMap{source, field, Acf}, EpiScale{A->d_scalars})));
    return launch_map<false>(A, A->nCells, ScaleApplyMap{A->d_scalars, field, source, Acf, A->d_diag});
}


static int gamg_interpolate(ldu_matrix* A, double* psi, double* Apsi)
{
    LDU_TRY(k_offdiag(A, Apsi, psi));
    return launch_map<false>(A, A->nCells, NegDivMap{psi, Apsi, A->d_diag});
}

// ---------------------------------------------------------------------------
// coarsest level: ICCG / BICCG (GAMGSolverSolve.C:430-487, ICCG.C:40-109)
// ---------------------------------------------------------------------------
// Single-rank case: the coarsest matrix has ~nCellsInCoarsestLevel cells, far
// below one warp's worth of work, so the whole PCG+DIC (or PBiCG+DILU) solve
// runs sequentially in ONE thread, in the reference's own loop order — exact
// same arithmetic as the reference including the dot products.
struct CoarsestArgs {
    int n, nf, asym, maxIter;
    const int* l;
    const int* u;
    const int* losort;
    const double* diag;
    const double* upper;
    const double* lower;
    const double* source;
    double* psi;
    double *pA, *wA, *rA, *rD, *pT, *wT, *rT;
    double tol, relTol;
};

__global__ void coarsest_solve_kernel(CoarsestArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = a.n, nf = a.nf;
    const bool bicg = a.asym != 0;
    double wArA = 1.0e+20, wArAold = wArA;
    for (int c = 0; c < n; c++) a.psi[c] = 0.0;
    // wA = A psi (psi == 0 still goes through the loops like the reference)
    for (int c = 0; c < n; c++) a.wA[c] = __dmul_rn(a.diag[c], a.psi[c]);
    for (int f = 0; f < nf; f++) {
        a.wA[a.u[f]] = __dadd_rn(a.wA[a.u[f]], __dmul_rn(a.lower[f], a.psi[a.l[f]]));
        a.wA[a.l[f]] = __dadd_rn(a.wA[a.l[f]], __dmul_rn(a.upper[f], a.psi[a.u[f]]));
    }
    if (bicg) {
        for (int c = 0; c < n; c++) a.wT[c] = __dmul_rn(a.diag[c], a.psi[c]);
        for (int f = 0; f < nf; f++) {
            a.wT[a.u[f]] = __dadd_rn(a.wT[a.u[f]], __dmul_rn(a.upper[f], a.psi[a.l[f]]));
            a.wT[a.l[f]] = __dadd_rn(a.wT[a.l[f]], __dmul_rn(a.lower[f], a.psi[a.u[f]]));
        }
    }
    for (int c = 0; c < n; c++) {
        a.rA[c] = __dsub_rn(a.source[c], a.wA[c]);
        if (bicg) a.rT[c] = __dsub_rn(a.source[c], a.wT[c]);
    }
    // normFactor (lduMatrixSolver.C:179-197), sumA into pA
    for (int c = 0; c < n; c++) a.pA[c] = a.diag[c];
    for (int f = 0; f < nf; f++) {
        a.pA[a.u[f]] = __dadd_rn(a.pA[a.u[f]], a.lower[f]);
        a.pA[a.l[f]] = __dadd_rn(a.pA[a.l[f]], a.upper[f]);
    }
    double sumPsi = 0;
    for (int c = 0; c < n; c++) sumPsi = __dadd_rn(sumPsi, a.psi[c]);
    const double avg = __ddiv_rn(sumPsi, (double)n);
    double nfac = 0, res = 0;
    for (int c = 0; c < n; c++) {
        const double t = __dmul_rn(a.pA[c], avg);
        nfac = __dadd_rn(nfac, __dadd_rn(fabs(__dsub_rn(a.wA[c], t)), fabs(__dsub_rn(a.source[c], t))));
    }
    nfac = __dadd_rn(nfac, 1.0e-20);
    for (int c = 0; c < n; c++) res = __dadd_rn(res, fabs(a.rA[c]));
    const double initial = __ddiv_rn(res, nfac);
    double final_ = initial;
    auto converged = [&]() {
        return final_ < a.tol || (a.relTol > 1.0e-20 && final_ < __dmul_rn(a.relTol, initial));
    };
    if (converged()) return;
    // DIC / DILU diagonal
    for (int c = 0; c < n; c++) a.rD[c] = a.diag[c];
    for (int f = 0; f < nf; f++)
        a.rD[a.u[f]] = __dsub_rn(a.rD[a.u[f]], __ddiv_rn(__dmul_rn(a.upper[f], a.lower[f]), a.rD[a.l[f]]));
    for (int c = 0; c < n; c++) a.rD[c] = __ddiv_rn(1.0, a.rD[c]);
    int nIter = 0;
    do {
        wArAold = wArA;
        for (int c = 0; c < n; c++) a.wA[c] = __dmul_rn(a.rD[c], a.rA[c]);
        if (!bicg) {  // DICPreconditioner.C:87-123
            for (int f = 0; f < nf; f++)
                a.wA[a.u[f]] = __dsub_rn(a.wA[a.u[f]], __dmul_rn(__dmul_rn(a.rD[a.u[f]], a.upper[f]), a.wA[a.l[f]]));
            for (int f = nf - 1; f >= 0; f--)
                a.wA[a.l[f]] = __dsub_rn(a.wA[a.l[f]], __dmul_rn(__dmul_rn(a.rD[a.l[f]], a.upper[f]), a.wA[a.u[f]]));
        } else {  // DILUPreconditioner.C:88-185
            for (int k = 0; k < nf; k++) {
                const int f = a.losort[k];
                a.wA[a.u[f]] = __dsub_rn(a.wA[a.u[f]], __dmul_rn(__dmul_rn(a.rD[a.u[f]], a.lower[f]), a.wA[a.l[f]]));
            }
            for (int f = nf - 1; f >= 0; f--)
                a.wA[a.l[f]] = __dsub_rn(a.wA[a.l[f]], __dmul_rn(__dmul_rn(a.rD[a.l[f]], a.upper[f]), a.wA[a.u[f]]));
            for (int c = 0; c < n; c++) a.wT[c] = __dmul_rn(a.rD[c], a.rT[c]);
            for (int f = 0; f < nf; f++)
                a.wT[a.u[f]] = __dsub_rn(a.wT[a.u[f]], __dmul_rn(__dmul_rn(a.rD[a.u[f]], a.upper[f]), a.wT[a.l[f]]));
            for (int k = nf - 1; k >= 0; k--) {
                const int f = a.losort[k];
                a.wT[a.l[f]] = __dsub_rn(a.wT[a.l[f]], __dmul_rn(__dmul_rn(a.rD[a.l[f]], a.lower[f]), a.wT[a.u[f]]));
            }
        }
        wArA = 0;
        for (int c = 0; c < n; c++) wArA = __dadd_rn(wArA, __dmul_rn(a.wA[c], bicg ? a.rT[c] : a.rA[c]));
        if (nIter == 0) {
            for (int c = 0; c < n; c++) {
                a.pA[c] = a.wA[c];
                if (bicg) a.pT[c] = a.wT[c];
            }
        } else {
            const double beta = __ddiv_rn(wArA, wArAold);
            for (int c = 0; c < n; c++) {
                a.pA[c] = __dadd_rn(a.wA[c], __dmul_rn(beta, a.pA[c]));
                if (bicg) a.pT[c] = __dadd_rn(a.wT[c], __dmul_rn(beta, a.pT[c]));
            }
        }
        for (int c = 0; c < n; c++) a.wA[c] = __dmul_rn(a.diag[c], a.pA[c]);
        for (int f = 0; f < nf; f++) {
            a.wA[a.u[f]] = __dadd_rn(a.wA[a.u[f]], __dmul_rn(a.lower[f], a.pA[a.l[f]]));
            a.wA[a.l[f]] = __dadd_rn(a.wA[a.l[f]], __dmul_rn(a.upper[f], a.pA[a.u[f]]));
        }
        if (bicg) {
            for (int c = 0; c < n; c++) a.wT[c] = __dmul_rn(a.diag[c], a.pT[c]);
            for (int f = 0; f < nf; f++) {
                a.wT[a.u[f]] = __dadd_rn(a.wT[a.u[f]], __dmul_rn(a.upper[f], a.pT[a.l[f]]));
                a.wT[a.l[f]] = __dadd_rn(a.wT[a.l[f]], __dmul_rn(a.lower[f], a.pT[a.u[f]]));
            }
        }
        double wApA = 0;
        for (int c = 0; c < n; c++) wApA = __dadd_rn(wApA, __dmul_rn(a.wA[c], bicg ? a.pT[c] : a.pA[c]));
        if (__ddiv_rn(fabs(wApA), nfac) < 1.0e-300) break;
        const double alpha = __ddiv_rn(wArA, wApA);
        for (int c = 0; c < n; c++) {
            a.psi[c] = __dadd_rn(a.psi[c], __dmul_rn(alpha, a.pA[c]));
            a.rA[c] = __dsub_rn(a.rA[c], __dmul_rn(alpha, a.wA[c]));
            if (bicg) a.rT[c] = __dsub_rn(a.rT[c], __dmul_rn(alpha, a.wT[c]));
        }
        res = 0;
        for (int c = 0; c < n; c++) res = __dadd_rn(res, fabs(a.rA[c]));
        final_ = __ddiv_rn(res, nfac);
    } while (nIter++ < a.maxIter && !converged());
}

// Coupled case (one region per GPU): the same solve by ONE WARP per rank, all ranks at once.  Lane 0 runs the
// region's loops in the reference's order as above; the three global sums of an iteration are the in-kernel
// all-reduce over the peer windows (comm_allreduce_warp: partial sums added in rank order, bit-identical on every
// rank), the processor-interface update of Amul / Tmul is a put + flag + wait on the same windows by the warp.
// No host round trip, no launch per operation: an iteration costs four NVLink round trips instead of ~15 launches
// and a host poll every fourth iteration (measured 2 ranks, 128^3: 2.3-3.6 ms per V-cycle in the generic solver).
// The kernels of all ranks must be running together; each is a single warp, so nothing can keep one from
// starting.  (PCG.C:65-182, PBiCG.C:65-190, lduMatrixUpdateMatrixInterfaces.C:30-266)
struct CoupledArgs {
    CoarsestArgs a;
    CommDev comm;
    const IfaceDev* ifs;
    int nIfs, nIfFaces;
    const int* ifCells;
    const double* bou;
    const double* intc;
    double* recv;
    SolverScalars* S;      // the finest matrix's scalars: commError / done on a time-out
};

// psi[faceCells] of every interface into the neighbours' windows, wait for theirs, gather into recv; false on time-out
// the same exchange with tagged words (every interface of every rank has at most kLLFaces faces and a slot below
// kLLIfs): value and flag arrive together, nobody waits for a fence
__device__ bool coupled_halo_ll(const CoupledArgs& A, const double* x)
{
    if (A.nIfs == 0) return true;
    const int lane = threadIdx.x & 31;
    const CommDev& c = A.comm;
    WindowHeader* me = win_hdr(c, c.rank);
    __syncwarp();      // x was written by lane 0
    // one exchange per neighbour rank (counted per pair, see WindowHeader::haloSent): the tag of interface k
    for (int k = 0; k < A.nIfs; k++) {
        const IfaceDev it = A.ifs[k];
        const unsigned long long epoch = me->haloSent[it.nbrRank] + 1ull;
        if (lane < it.n)
            ll_store_sys(&win_hdr(c, it.nbrRank)->haloLL[(int)(epoch & 1ull)][it.nbrInterface][lane],
                         x[A.ifCells[it.offset + lane]], (unsigned int)epoch);
    }
    bool ok = true;
    for (int k = 0; k < A.nIfs; k++) {
        const IfaceDev it = A.ifs[k];
        const unsigned long long epoch = me->haloSent[it.nbrRank] + 1ull;
        double v = 0.0;
        if (lane < it.n) {
            ok = ok && ll_wait_sys(&me->haloLL[(int)(epoch & 1ull)][k][lane], (unsigned int)epoch, c.timeoutCycles, v);
            A.recv[it.offset + lane] = v;
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (lane == 0) {
        for (int k = 0; k < A.nIfs; k++) {
            const int r = A.ifs[k].nbrRank;
            bool first = true;
            for (int j = 0; j < k; j++) first = first && A.ifs[j].nbrRank != r;
            if (first) me->haloSent[r] = me->haloSent[r] + 1ull;
        }
    }
    __syncwarp();
    return ok;
}

// the fall-back for interfaces with more than kLLFaces faces (a large nCellsInCoarsestLevel): window slots, system
// fence, epoch flags -- the protocol of halo_put_kernel / interface_wait_kernel run by the one warp
__device__ bool coupled_halo(const CoupledArgs& A, const double* x)
{
    if (A.nIfs == 0) return true;
    const int lane = threadIdx.x & 31;
    const CommDev& c = A.comm;
    WindowHeader* me = win_hdr(c, c.rank);
    __syncwarp();      // x was written by lane 0
    for (int k = 0; k < A.nIfs; k++) {
        const IfaceDev it = A.ifs[k];
        const int par = (int)((me->haloSent[it.nbrRank] + 1ull) & 1ull);
        double* dst = win_halo(c, it.nbrRank, par, it.nbrInterface);
        for (int i = lane; i < it.n; i += 32) dst[i] = x[A.ifCells[it.offset + i]];
    }
    __threadfence_system();
    __syncwarp();
    bool ok = true;
    if (lane == 0) {
        for (int k = 0; k < A.nIfs; k++) {
            const int r = A.ifs[k].nbrRank;
            bool first = true;
            for (int j = 0; j < k; j++) first = first && A.ifs[j].nbrRank != r;
            if (!first) continue;
            const unsigned long long epoch = me->haloSent[r] + 1ull;
            me->haloSent[r] = epoch;
            st_release_sys(&win_hdr(c, r)->haloSeq[(int)(epoch & 1ull)][c.rank], epoch);
        }
        for (int k = 0; k < A.nIfs && ok; k++) {
            const unsigned long long epoch = me->haloSent[A.ifs[k].nbrRank];
            ok = wait_epoch(&me->haloSeq[(int)(epoch & 1ull)][A.ifs[k].nbrRank], epoch, c.timeoutCycles);
        }
    }
    ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
    __syncwarp();      // lane 0 has advanced haloSent
    if (!ok) return false;
    for (int k = 0; k < A.nIfs; k++) {
        const IfaceDev it = A.ifs[k];
        const int par = (int)(me->haloSent[it.nbrRank] & 1ull);
        const double* src = win_halo(c, c.rank, par, k);
        for (int i = lane; i < it.n; i += 32) A.recv[it.offset + i] = ld_volatile_f64(src + i);
    }
    __syncwarp();
    return true;
}

__global__ void __launch_bounds__(32, 1) coarsest_coupled_kernel(CoupledArgs A)
{
    const CoarsestArgs& a = A.a;
    const int lane = threadIdx.x;
    const bool lead = lane == 0;
    const int n = a.n, nf = a.nf;
    const bool bicg = a.asym != 0;
    auto fail = [&]() {
        if (lead) {
            A.S->commError = 1;
            A.S->done = 1;
        }
    };
    // y = A x (transpose: T x) of the region, then the coupled rows; x has been exchanged into recv
    auto amul_local = [&](double* y, const double* x, bool transpose) {
        const double* lo = transpose ? a.upper : a.lower;
        const double* up = transpose ? a.lower : a.upper;
        const double* cf = transpose ? A.intc : A.bou;
        for (int c = 0; c < n; c++) y[c] = __dmul_rn(a.diag[c], x[c]);
        for (int f = 0; f < nf; f++) {
            y[a.u[f]] = __dadd_rn(y[a.u[f]], __dmul_rn(lo[f], x[a.l[f]]));
            y[a.l[f]] = __dadd_rn(y[a.l[f]], __dmul_rn(up[f], x[a.u[f]]));
        }
        for (int k = 0; k < A.nIfFaces; k++) {
            const int c = A.ifCells[k];
            y[c] = __dsub_rn(y[c], __dmul_rn(cf[k], A.recv[k]));
        }
    };
    // do the interfaces of EVERY rank fit the tagged halo words?  (all ranks must take the same path)
    bool llHalo;
    {
        double big[1] = {0.0};
        if (lead)
            for (int k = 0; k < A.nIfs; k++)
                if (A.ifs[k].n > kLLFaces || A.ifs[k].nbrInterface >= kLLIfs || k >= kLLIfs) big[0] = 1.0;
        if (!comm_allreduce_warp_ll<1>(A.comm, big, A.S)) return;
        llHalo = big[0] == 0.0;
    }
    auto halo = [&](const double* x) { return llHalo ? coupled_halo_ll(A, x) : coupled_halo(A, x); };
    double wArA = 1.0e+20, wArAold = wArA;
    if (lead)
        for (int c = 0; c < n; c++) a.psi[c] = 0.0;
    if (!halo(a.psi)) return fail();
    if (lead) {
        amul_local(a.wA, a.psi, false);
        if (bicg) amul_local(a.wT, a.psi, true);
        for (int c = 0; c < n; c++) {
            a.rA[c] = __dsub_rn(a.source[c], a.wA[c]);
            if (bicg) a.rT[c] = __dsub_rn(a.source[c], a.wT[c]);
        }
        // sumA into pA (lduMatrixATmul.C:156-205)
        for (int c = 0; c < n; c++) a.pA[c] = a.diag[c];
        for (int f = 0; f < nf; f++) {
            a.pA[a.u[f]] = __dadd_rn(a.pA[a.u[f]], a.lower[f]);
            a.pA[a.l[f]] = __dadd_rn(a.pA[a.l[f]], a.upper[f]);
        }
        for (int k = 0; k < A.nIfFaces; k++) a.pA[A.ifCells[k]] = __dsub_rn(a.pA[A.ifCells[k]], A.bou[k]);
    }
    double t2[2] = {0.0, 0.0};
    if (lead) {
        for (int c = 0; c < n; c++) t2[0] = __dadd_rn(t2[0], a.psi[c]);
        t2[1] = (double)n;
    }
    if (!comm_allreduce_warp_ll<2>(A.comm, t2, A.S)) return;
    const double avg = __ddiv_rn(t2[0], t2[1]);
    t2[0] = t2[1] = 0.0;
    if (lead) {
        for (int c = 0; c < n; c++) {
            const double t = __dmul_rn(a.pA[c], avg);
            t2[0] = __dadd_rn(t2[0], __dadd_rn(fabs(__dsub_rn(a.wA[c], t)), fabs(__dsub_rn(a.source[c], t))));
        }
        for (int c = 0; c < n; c++) t2[1] = __dadd_rn(t2[1], fabs(a.rA[c]));
    }
    if (!comm_allreduce_warp_ll<2>(A.comm, t2, A.S)) return;
    const double nfac = __dadd_rn(t2[0], 1.0e-20);
    const double initial = __ddiv_rn(t2[1], nfac);
    double final_ = initial;
    auto converged = [&]() {
        return final_ < a.tol || (a.relTol > 1.0e-20 && final_ < __dmul_rn(a.relTol, initial));
    };
    if (converged()) return;
    if (lead) {   // DIC / DILU diagonal of the region (block-local, as in the reference)
        for (int c = 0; c < n; c++) a.rD[c] = a.diag[c];
        for (int f = 0; f < nf; f++)
            a.rD[a.u[f]] = __dsub_rn(a.rD[a.u[f]], __ddiv_rn(__dmul_rn(a.upper[f], a.lower[f]), a.rD[a.l[f]]));
        for (int c = 0; c < n; c++) a.rD[c] = __ddiv_rn(1.0, a.rD[c]);
    }
    int nIter = 0;
    do {
        wArAold = wArA;
        double t1[1] = {0.0};
        if (lead) {
            for (int c = 0; c < n; c++) a.wA[c] = __dmul_rn(a.rD[c], a.rA[c]);
            if (!bicg) {
                for (int f = 0; f < nf; f++)
                    a.wA[a.u[f]] = __dsub_rn(a.wA[a.u[f]], __dmul_rn(__dmul_rn(a.rD[a.u[f]], a.upper[f]), a.wA[a.l[f]]));
                for (int f = nf - 1; f >= 0; f--)
                    a.wA[a.l[f]] = __dsub_rn(a.wA[a.l[f]], __dmul_rn(__dmul_rn(a.rD[a.l[f]], a.upper[f]), a.wA[a.u[f]]));
            } else {
                for (int k = 0; k < nf; k++) {
                    const int f = a.losort[k];
                    a.wA[a.u[f]] = __dsub_rn(a.wA[a.u[f]], __dmul_rn(__dmul_rn(a.rD[a.u[f]], a.lower[f]), a.wA[a.l[f]]));
                }
                for (int f = nf - 1; f >= 0; f--)
                    a.wA[a.l[f]] = __dsub_rn(a.wA[a.l[f]], __dmul_rn(__dmul_rn(a.rD[a.l[f]], a.upper[f]), a.wA[a.u[f]]));
                for (int c = 0; c < n; c++) a.wT[c] = __dmul_rn(a.rD[c], a.rT[c]);
                for (int f = 0; f < nf; f++)
                    a.wT[a.u[f]] = __dsub_rn(a.wT[a.u[f]], __dmul_rn(__dmul_rn(a.rD[a.u[f]], a.upper[f]), a.wT[a.l[f]]));
                for (int k = nf - 1; k >= 0; k--) {
                    const int f = a.losort[k];
                    a.wT[a.l[f]] = __dsub_rn(a.wT[a.l[f]], __dmul_rn(__dmul_rn(a.rD[a.l[f]], a.lower[f]), a.wT[a.u[f]]));
                }
            }
            for (int c = 0; c < n; c++) t1[0] = __dadd_rn(t1[0], __dmul_rn(a.wA[c], bicg ? a.rT[c] : a.rA[c]));
        }
        if (!comm_allreduce_warp_ll<1>(A.comm, t1, A.S)) return;
        wArA = t1[0];
        if (lead) {
            if (nIter == 0) {
                for (int c = 0; c < n; c++) {
                    a.pA[c] = a.wA[c];
                    if (bicg) a.pT[c] = a.wT[c];
                }
            } else {
                const double beta = __ddiv_rn(wArA, wArAold);
                for (int c = 0; c < n; c++) {
                    a.pA[c] = __dadd_rn(a.wA[c], __dmul_rn(beta, a.pA[c]));
                    if (bicg) a.pT[c] = __dadd_rn(a.wT[c], __dmul_rn(beta, a.pT[c]));
                }
            }
        }
        if (!halo(a.pA)) return fail();
        if (lead) amul_local(a.wA, a.pA, false);
        if (bicg) {
            if (!halo(a.pT)) return fail();
            if (lead) amul_local(a.wT, a.pT, true);
        }
        t1[0] = 0.0;
        if (lead)
            for (int c = 0; c < n; c++) t1[0] = __dadd_rn(t1[0], __dmul_rn(a.wA[c], bicg ? a.pT[c] : a.pA[c]));
        if (!comm_allreduce_warp_ll<1>(A.comm, t1, A.S)) return;
        const double wApA = t1[0];
        if (__ddiv_rn(fabs(wApA), nfac) < 1.0e-300) break;
        const double alpha = __ddiv_rn(wArA, wApA);
        t1[0] = 0.0;
        if (lead) {
            for (int c = 0; c < n; c++) {
                a.psi[c] = __dadd_rn(a.psi[c], __dmul_rn(alpha, a.pA[c]));
                a.rA[c] = __dsub_rn(a.rA[c], __dmul_rn(alpha, a.wA[c]));
                if (bicg) a.rT[c] = __dsub_rn(a.rT[c], __dmul_rn(alpha, a.wT[c]));
            }
            for (int c = 0; c < n; c++) t1[0] = __dadd_rn(t1[0], fabs(a.rA[c]));
        }
        if (!comm_allreduce_warp_ll<1>(A.comm, t1, A.S)) return;
        final_ = __ddiv_rn(t1[0], nfac);
    } while (nIter++ < a.maxIter && !converged());
}

static int solve_coarsest(ldu_matrix* top, const ldu_controls* c, bool asPrecond)
{
    GamgLevel* L = top->levels.back();
    ldu_matrix* cm = L->coarse;
    const double tol = asPrecond ? c->precTolerance : c->tolerance;
    const double relTol = asPrecond ? c->precRelTol : c->relTol;
    // LDU_GAMG_GENERIC_COARSEST=1 sends every coarsest level to the generic solver (tests)
    const char* generic = getenv("LDU_GAMG_GENERIC_COARSEST");
    const bool small = cm->nCells <= 8192 && !(generic && generic[0] == '1');
    const bool coupled = cm->ctx->comm.connected && cm->ctx->comm.nRanks > 1 && !cm->ctx->comm.selfOnly;
    if (small && (coupled || cm->nIfFaces == 0)) {
        CoarsestArgs a;
        a.n = cm->nCells;
        a.nf = cm->nFaces;
        a.asym = cm->symmetric ? 0 : 1;
        a.maxIter = 1000;  // ICCG's dictionary carries tolerance/relTol only (ICCG.C:44-54)
        a.l = cm->d_l;
        a.u = cm->d_u;
        a.losort = cm->d_losort;
        a.diag = cm->d_diag;
        a.upper = cm->d_upper;
        a.lower = cm->d_lower;
        a.source = L->d_src;
        a.psi = L->d_corr;
        a.pA = work_vec(cm, W_PA);
        a.wA = work_vec(cm, W_WA);
        a.rA = work_vec(cm, W_RA);
        a.rD = work_vec(cm, W_RD);
        a.pT = work_vec(cm, W_PT);
        a.wT = work_vec(cm, W_WT);
        a.rT = work_vec(cm, W_RT);
        a.tol = tol;
        a.relTol = relTol;
        if (coupled) {
            CoupledArgs A;
            A.a = a;
            A.comm = comm_dev(cm->ctx);
            LDU_TRY(comm_halo_table(cm, &A.ifs));
            A.nIfs = (int)cm->ifs.size();
            A.nIfFaces = cm->nIfFaces;
            A.ifCells = cm->d_ifCells;
            A.bou = cm->d_bou;
            A.intc = cm->d_int;
            A.recv = cm->d_recv;
            A.S = top->d_scalars;
            coarsest_coupled_kernel<<<1, 32, 0, cm->ctx->stream>>>(A);
        } else {
            coarsest_solve_kernel<<<1, 32, 0, cm->ctx->stream>>>(a);
        }
        count_launch();
        LDU_CUDA(cudaGetLastError());
        return LDU_OK;
    }
    // generic (multi-rank or large coarsest level): the device-resident Krylov solver
    ldu_controls cc = *c;
    cc.solver = cm->symmetric ? LDU_SOLVER_PCG : LDU_SOLVER_PBICG;
    cc.preconditioner = cm->symmetric ? LDU_PRECOND_DIC : LDU_PRECOND_DILU;
    cc.maxIter = 1000;
    cc.tolerance = tol;
    cc.relTol = relTol;
    cc.checkInterval = 4;
    LDU_TRY(launch_map<false>(cm, cm->nCells, ZeroMap{L->d_corr}));
    ldu_solver_performance perf;
    int rc = solve_device(cm, &cc, L->d_corr, L->d_src, &perf);
    // leave the coarse matrix's guard flag clear for later kernels on it
    LDU_CUDA(cudaMemsetAsync(cm->d_scalars, 0, sizeof(SolverScalars), cm->ctx->stream));
    return rc;
}

// ---------------------------------------------------------------------------
// V-cycle (GAMGSolverSolve.C:120-364)
// ---------------------------------------------------------------------------
struct VcycleState {
    std::vector<Smoother> smoothers;  // [nLevels+1], 0 = finest
    bool ready = false;
};

static ldu_matrix* level_matrix(ldu_matrix* m, int lev) { return lev == 0 ? m : m->levels[lev - 1]->coarse; }

static int vcycle_init(ldu_matrix* m, const ldu_controls* c, VcycleState& vs)
{
    const int nLev = (int)m->levels.size();
    vs.smoothers.resize(nLev + 1);
    for (int i = 0; i <= nLev; i++) LDU_TRY(smoother_setup(level_matrix(m, i), c->smoother, vs.smoothers[i]));
    vs.ready = true;
    return LDU_OK;
}

static void vcycle_release(VcycleState& vs)
{
    for (Smoother& s : vs.smoothers) smoother_release(s);
    vs.smoothers.clear();
}

static double g_cycleTimes[4] = {0, 0, 0, 0};   // debug: us in restriction / coarsest solve / prolongation+smoothing, cycles

static int vcycle(ldu_matrix* m, const ldu_controls* c, VcycleState& vs, double* psi, const double* source,
                  double* Apsi, double* finestCorrection, double* finestResidual, bool asPrecond)
{
    const int nLev = (int)m->levels.size();
    const int coarsestLevel = nLev - 1;
    const bool scaleCorrection = c->scaleCorrection < 0 ? m->symmetric : (c->scaleCorrection != 0);
    std::vector<GamgLevel*>& Ls = m->levels;

    const double t0 = getenv("LDU_GAMG_TIMING") ? std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0.0;
    LDU_TRY(restrict_field(m, Ls[0], finestResidual, Ls[0]->d_src));
    for (int lev = 0; lev < coarsestLevel; lev++) {
        ldu_matrix* A = Ls[lev]->coarse;
        if (c->nPreSweeps) {
            LDU_TRY(launch_map<false>(A, A->nCells, ZeroMap{Ls[lev]->d_corr}));
            LDU_TRY(smoother_apply(A, vs.smoothers[lev + 1], Ls[lev]->d_corr, Ls[lev]->d_src,
                                   std::min(c->nPreSweeps + c->preSweepsLevelMultiplier * lev, c->maxPreSweeps)));
            if (scaleCorrection && lev < coarsestLevel - 1)
                LDU_TRY(gamg_scale(A, Ls[lev]->d_corr, Apsi, Ls[lev]->d_src));
            LDU_TRY(k_amul(A, Apsi, Ls[lev]->d_corr, false));
            LDU_TRY(launch_map<false>(A, A->nCells, SubMap{Ls[lev]->d_src, Apsi}));
        }
        LDU_TRY(restrict_field(m, Ls[lev + 1], Ls[lev]->d_src, Ls[lev + 1]->d_src));
    }

    // debug (LDU_GAMG_TIMING=1): host wall time of the three parts of a cycle, with a stream synchronize between them
    static const bool timing = getenv("LDU_GAMG_TIMING") != nullptr;
    auto now = [&]() {
        cudaStreamSynchronize(m->ctx->stream);
        return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    const double t1 = timing ? now() : 0.0;
    LDU_TRY(solve_coarsest(m, c, asPrecond));
    const double t2 = timing ? now() : 0.0;

    for (int lev = coarsestLevel - 1; lev >= 0; lev--) {
        ldu_matrix* A = Ls[lev]->coarse;
        if (c->nPreSweeps)  // preSmoothedCoarseCorrField lives in finestCorrection
            LDU_TRY(launch_map<false>(A, A->nCells, CopyMapG{finestCorrection, Ls[lev]->d_corr}));
        LDU_TRY(launch_map<false>(A, A->nCells, ProlongMap{Ls[lev]->d_corr, Ls[lev + 1]->d_corr, Ls[lev + 1]->d_restrict}));
        if (c->interpolateCorrection) LDU_TRY(gamg_interpolate(A, Ls[lev]->d_corr, Apsi));
        if (scaleCorrection && lev < coarsestLevel - 1)
            LDU_TRY(gamg_scale(A, Ls[lev]->d_corr, Apsi, Ls[lev]->d_src));
        if (c->nPreSweeps) LDU_TRY(launch_map<false>(A, A->nCells, AddMapG{Ls[lev]->d_corr, finestCorrection}));
        LDU_TRY(smoother_apply(A, vs.smoothers[lev + 1], Ls[lev]->d_corr, Ls[lev]->d_src,
                               std::min(c->nPostSweeps + c->postSweepsLevelMultiplier * lev, c->maxPostSweeps)));
    }

    LDU_TRY(launch_map<false>(m, m->nCells, ProlongMap{finestCorrection, Ls[0]->d_corr, Ls[0]->d_restrict}));
    if (c->interpolateCorrection) LDU_TRY(gamg_interpolate(m, finestCorrection, Apsi));
    if (scaleCorrection) LDU_TRY(gamg_scale(m, finestCorrection, Apsi, finestResidual));
    LDU_TRY(launch_map<false>(m, m->nCells, AddMapG{psi, finestCorrection}));
    const int rcLast = smoother_apply(m, vs.smoothers[0], psi, source, c->nFinestSweeps);
    if (timing) {
        const double t3 = now();
        g_cycleTimes[0] += t1 - t0;
        g_cycleTimes[1] += t2 - t1;
        g_cycleTimes[2] += t3 - t2;
        g_cycleTimes[3] += 1.0;
    }
    return rcLast;
}

static int prepare_hierarchy(ldu_matrix* m, const ldu_controls* c)
{
    LDU_TRY(gamg_build(m, c));
    LDU_TRY(agglomerate_coefficients(m));
    // scalars of every level start clear (guards read them)
    for (GamgLevel* L : m->levels) {
        L->coarse->referenceOrderSums = m->referenceOrderSums;
        LDU_TRY(ensure_scalars(L->coarse));
        LDU_CUDA(cudaMemsetAsync(L->coarse->d_scalars, 0, sizeof(SolverScalars), m->ctx->stream));
    }
    return LDU_OK;
}

// GAMGSolver::solve (GAMGSolverSolve.C:34-117)
int gamg_solve(ldu_matrix* m, const ldu_controls* c, double* psi, const double* source,
               ldu_solver_performance* perf)
{
    (void)perf;
    const int n = m->nCells;
    LDU_TRY(prepare_hierarchy(m, c));
    double* Apsi = work_vec(m, W_APSI);
    double* finestCorrection = work_vec(m, W_CORR);
    double* finestResidual = work_vec(m, W_RES);
    LDU_TRY(init_scalars(m, c));
    LDU_TRY(solve_prologue(m, psi, source, Apsi, finestResidual, finestCorrection));
    SolverScalars hs;
    LDU_TRY(read_scalars(m, &hs));
    if (hs.done) return LDU_OK;
    // the flag must be clear while the V-cycle's unguarded and guarded kernels run
    VcycleState vs;
    LDU_TRY(vcycle_init(m, c, vs));
    auto one_cycle = [&]() -> int {
        LDU_TRY(vcycle(m, c, vs, psi, source, Apsi, finestCorrection, finestResidual, false));
        LDU_TRY(k_amul(m, Apsi, psi, false));
        return launch_map_reduce<1, false>(m, n, ResidualFromMap{finestResidual, source, Apsi}, EpiResidual<false>{1});
    };
    // A V-cycle is a couple of hundred launches, most of them on levels so small that the kernel is shorter than
    // the launch call: from the second cycle on the whole cycle (V-cycle + residual) is replayed as ONE CUDA graph.
    // The first cycle runs eagerly (it allocates the lazily created work fields, schedules and colourings, which
    // a capture must not do).  Single-region solves whose coarsest level is solved by the one-kernel solver only
    // (the coupled coarsest solver polls the host); anything the capture refuses falls back to eager launches.
    const char* graphEnv = getenv("LDU_GAMG_GRAPH");
    GamgLevel* Lc = m->levels.back();
    // ... and only with the multi-colour smoother: the lexicographic sweeps (dataflow and box kernels) tag their
    // words with an epoch the HOST increments per launch, which a replayed graph would freeze
    // Several regions: the halo and all-reduce epochs live in the exchange windows and are advanced by the kernels
    // themselves, and the coupled coarsest level is one kernel too, so the cycle of every rank replays as a graph
    // just the same (the ranks' kernels meet through the windows, not through the host).
    const bool coupled = m->ctx->comm.connected && m->ctx->comm.nRanks > 1 && !m->ctx->comm.selfOnly;
    bool useGraph = !(graphEnv && graphEnv[0] == '0') && (coupled || (m->ctx->comm.nRanks == 1 && m->nIfFaces == 0))
                    && c->smoother == LDU_SMOOTHER_MCGS && Lc->coarse->nCells <= 8192
                    && !getenv("LDU_GAMG_GENERIC_COARSEST") && !getenv("LDU_GAMG_TIMING");
    cudaGraphExec_t exec = nullptr;
    cudaStream_t st = m->ctx->stream;
    int rc = LDU_OK;
    for (int cycle = 0;; cycle++) {
        if (useGraph && cycle >= 1) {
            if (!exec) {
                cudaGraph_t graph = nullptr;
                const long long launchesBefore = g_launches;
                bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
                if (ok) {
                    const int crc = one_cycle();
                    const cudaError_t e = cudaStreamEndCapture(st, &graph);
                    ok = crc == LDU_OK && e == cudaSuccess && graph != nullptr;
                    m->graphLaunches = (int)(g_launches - launchesBefore);
                    g_launches = launchesBefore;        // nothing has run yet
                }
                if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
                if (graph) cudaGraphDestroy(graph);
                if (!ok) {
                    cudaGetLastError();
                    exec = nullptr;
                    useGraph = false;
                }
            }
            if (exec) {
                if (cudaGraphLaunch(exec, st) != cudaSuccess) {
                    rc = cuda_fail(cudaGetLastError(), "cudaGraphLaunch", __FILE__, __LINE__);
                    break;
                }
                g_launches += m->graphLaunches;
                rc = read_scalars(m, &hs);
                if (rc != LDU_OK || hs.done) break;
                continue;
            }
        }
        rc = one_cycle();
        if (rc != LDU_OK) break;
        rc = read_scalars(m, &hs);
        if (rc != LDU_OK || hs.done) break;
    }
    if (exec) cudaGraphExecDestroy(exec);
    vcycle_release(vs);
    if (getenv("LDU_GAMG_TIMING") && g_cycleTimes[3] > 0) {
        fprintf(stderr, "[ldu gamg rank %d] %d cycles: restrict %.0f us, coarsest %.0f us, prolong+smooth %.0f us per cycle\n",
                m->ctx->comm.rank, (int)g_cycleTimes[3], g_cycleTimes[0] / g_cycleTimes[3], g_cycleTimes[1] / g_cycleTimes[3],
                g_cycleTimes[2] / g_cycleTimes[3]);
        g_cycleTimes[0] = g_cycleTimes[1] = g_cycleTimes[2] = g_cycleTimes[3] = 0;
    }
    return rc;
}

// GAMGPreconditioner::precondition (GAMGPreconditioner.C:81-128)
int gamg_precondition(ldu_matrix* m, const ldu_controls* c, double* wA, const double* rA)
{
    const int n = m->nCells;
    // hierarchy + coarse coefficients are built once per outer solve: the
    // preconditioner object lives as long as the Krylov loop (PCG.C:117-121)
    if (!m->precondHierarchyReady) {
        LDU_TRY(prepare_hierarchy(m, c));
        m->precondHierarchyReady = true;
    }
    double* AwA = work_vec(m, W_APSI);
    double* finestCorrection = work_vec(m, W_CORR);
    double* finestResidual = work_vec(m, W_RES);
    VcycleState vs;
    LDU_TRY(vcycle_init(m, c, vs));
    int rc = launch_map<false>(m, n, ZeroMap{wA});
    if (rc == LDU_OK) rc = launch_map<false>(m, n, CopyMapG{finestResidual, rA});
    for (int cycle = 0; cycle < c->nVcycles && rc == LDU_OK; cycle++) {
        rc = vcycle(m, c, vs, wA, rA, AwA, finestCorrection, finestResidual, true);
        if (rc == LDU_OK && cycle < c->nVcycles - 1) {
            rc = k_amul(m, AwA, wA, false);
            if (rc == LDU_OK) rc = launch_map<false>(m, n, CopyMapG{finestResidual, rA});
            if (rc == LDU_OK) rc = launch_map<false>(m, n, SubMap{finestResidual, AwA});
        }
    }
    vcycle_release(vs);
    return rc;
}

}  // namespace ldu

using namespace ldu;

extern "C" {

int ldu_gamg_build(ldu_matrix* m, const ldu_controls* controls)
{
    if (!m || !controls) return LDU_EINVAL;
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    LDU_TRY(gamg_build(m, controls));
    if (m->haveCoeffs) LDU_TRY(agglomerate_coefficients(m));
    LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return LDU_OK;
}

int ldu_gamg_begin_levels(ldu_matrix* m)
{
    if (!m) return LDU_EINVAL;
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
    gamg_free(m);
    m->externalHierarchy = true;
    return LDU_OK;
}

int ldu_gamg_set_level(ldu_matrix* m, int level, int nFine, const int* restrictAddr, int nFineFaces,
                       const int* faceRestrictAddr, int nCoarse, int nCoarseFaces, const int* coarseLower,
                       const int* coarseUpper, const int* coarseIfSizes, const int* const* coarseIfFaceCells,
                       const int* const* ifRestrictAddr)
{
    if (!m || !m->externalHierarchy || m->hierarchyValid) {
        set_error("ldu_gamg_set_level: call ldu_gamg_begin_levels first");
        return LDU_EINVAL;
    }
    if (level != (int)m->levels.size() || level >= kMaxLevels - 1) {
        set_error("ldu_gamg_set_level: levels must be given in order, starting at 0");
        return LDU_EINVAL;
    }
    ldu_matrix* fine = level == 0 ? m : m->levels[level - 1]->coarse;
    if (nFine != fine->nCells || nFineFaces != fine->nFaces || nCoarse < 1 || nCoarseFaces < 0 || !restrictAddr
        || (nFineFaces && !faceRestrictAddr) || (nCoarseFaces && (!coarseLower || !coarseUpper))) {
        set_error("ldu_gamg_set_level: sizes do not match the level above");
        return LDU_EINVAL;
    }
    LDU_CUDA(cudaSetDevice(m->ctx->device));
    HostLevel H;
    H.nCoarse = nCoarse;
    H.cmap.assign(restrictAddr, restrictAddr + nFine);
    H.faceMap.assign(faceRestrictAddr, faceRestrictAddr + nFineFaces);
    H.cOwner.assign(coarseLower, coarseLower + nCoarseFaces);
    H.cNeighbour.assign(coarseUpper, coarseUpper + nCoarseFaces);
    for (int v : H.cmap)
        if (v < 0 || v >= nCoarse) {
            set_error("ldu_gamg_set_level: restrictAddr out of range");
            return LDU_EINVAL;
        }
    for (int v : H.faceMap)
        if (v >= nCoarseFaces || -1 - v >= nCoarse) {
            set_error("ldu_gamg_set_level: faceRestrictAddr out of range");
            return LDU_EINVAL;
        }
    const size_t nIf = fine->ifs.size();
    H.ifCells.resize(nIf);
    H.ifRestrict.resize(nIf);
    if (nIf && (!coarseIfSizes || !coarseIfFaceCells || !ifRestrictAddr)) {
        set_error("ldu_gamg_set_level: the matrix has coupled patches, their coarse interfaces are missing");
        return LDU_EINVAL;
    }
    for (size_t p = 0; p < nIf; p++) {
        H.ifCells[p].assign(coarseIfFaceCells[p], coarseIfFaceCells[p] + coarseIfSizes[p]);
        H.ifRestrict[p].assign(ifRestrictAddr[p], ifRestrictAddr[p] + fine->ifs[p].n);
        for (int v : H.ifRestrict[p])
            if (v < 0 || v >= coarseIfSizes[p]) {
                set_error("ldu_gamg_set_level: interface restrict addressing out of range");
                return LDU_EINVAL;
            }
    }
    GamgLevel* L = nullptr;
    LDU_TRY(build_level_device(fine, H, &L));
    m->levels.push_back(L);
    return LDU_OK;
}

int ldu_gamg_end_levels(ldu_matrix* m)
{
    if (!m || !m->externalHierarchy) return LDU_EINVAL;
    if (m->levels.empty()) {
        set_error("GAMG: no coarse levels created, matrix too small or nCellsInCoarsestLevel too large "
                  "(GAMGSolver.C:108-126)");
        return LDU_EINVAL;
    }
    m->hierarchyValid = true;
    return LDU_OK;
}

int ldu_gamg_internal_levels(ldu_matrix* m)
{
    if (!m) return LDU_EINVAL;
    if (m->externalHierarchy) {
        LDU_CUDA(cudaSetDevice(m->ctx->device));
        LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
        gamg_free(m);
        m->externalHierarchy = false;
    }
    return LDU_OK;
}

int ldu_gamg_nlevels(ldu_matrix* m) { return m ? (int)m->levels.size() : 0; }

int ldu_gamg_level_sizes(ldu_matrix* m, int level, int* nFine, int* nCoarse, int* nCoarseFaces)
{
    if (!m || level < 0 || level >= (int)m->levels.size()) return LDU_EINVAL;
    GamgLevel* L = m->levels[level];
    if (nFine) *nFine = L->nFine;
    if (nCoarse) *nCoarse = L->nCoarse;
    if (nCoarseFaces) *nCoarseFaces = L->nCoarseFaces;
    return LDU_OK;
}

int ldu_gamg_level_restrict(ldu_matrix* m, int level, int* restrictAddr)
{
    if (!m || level < 0 || level >= (int)m->levels.size() || !restrictAddr) return LDU_EINVAL;
    GamgLevel* L = m->levels[level];
    memcpy(restrictAddr, L->h_restrict.data(), L->h_restrict.size() * sizeof(int));
    return LDU_OK;
}

int ldu_gamg_level_coeffs(ldu_matrix* m, int level, double* diag, double* upper, double* lower)
{
    if (!m || level < 0 || level >= (int)m->levels.size()) return LDU_EINVAL;
    ldu_matrix* cm = m->levels[level]->coarse;
    LDU_CUDA(cudaStreamSynchronize(m->ctx->stream));
    if (diag) LDU_CUDA(cudaMemcpy(diag, cm->d_diag, cm->nCells * sizeof(double), cudaMemcpyDeviceToHost));
    if (upper && cm->nFaces)
        LDU_CUDA(cudaMemcpy(upper, cm->d_upper, cm->nFaces * sizeof(double), cudaMemcpyDeviceToHost));
    if (lower && cm->nFaces)
        LDU_CUDA(cudaMemcpy(lower, cm->d_lower, cm->nFaces * sizeof(double), cudaMemcpyDeviceToHost));
    return LDU_OK;
}

}  // extern "C"
