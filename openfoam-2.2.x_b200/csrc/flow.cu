// Dataflow (launch-free, barrier-free) triangular sweeps.
//
// The order-dependent loops of the reference — DIC/DILU/FDIC forward/backward
// substitution (DICPreconditioner.C:57-123, DILUPreconditioner.C:57-185,
// FDICPreconditioner.C:42-125) and the Gauss-Seidel sweeps
// (GaussSeidelSmoother.C:151-176, symGaussSeidelSmoother.C:151-205) — run as ONE
// persistent kernel per sweep instead of one launch per dependency level
// (646 levels per sweep on the 216^3 box):
//
//   * rows are laid out in dependency-level order (the schedule of sweeps.cu) in
//     chunks of 32 that never straddle a level; warps claim chunks in that order
//     from an atomic counter, so everything a row waits for has already been
//     claimed by a warp that is running: no deadlock, no grid-wide barrier;
//   * every finished row publishes {value, epoch} as ONE 16-byte word (two 8-byte
//     halves, each carrying the epoch, like NCCL's LL protocol): a consumer polls
//     that word in L2 and gets the value and the "ready" flag in the same load —
//     no fences, one L2 round trip per dependency level;
//   * a row consumes its dependencies strictly in the reference's face order with
//     unfused multiply/subtract, so results stay BIT-IDENTICAL to the sequential
//     reference loops.
//
// (A first version claimed chunks in plain row order and passed in-chunk values
// through shared memory; on the 10M-cell box its in-flight window of ~300k rows
// covered only 6.5 k-planes and throttled the wavefront pipeline to 18 ms per
// sweep — slower than launch-per-level.  Level order removes the window effect.)
#include <algorithm>
#include <cstdlib>

#include "reduce.cuh"
#include "sweeps.h"

namespace ldu {

#ifndef LDU_FLOW_SLEEP
#define LDU_FLOW_SLEEP 0
#endif
constexpr bool kFlowSleep = LDU_FLOW_SLEEP != 0;
constexpr long long kFlowTimeout = 4000000000ll;  // ~2 s without progress: bail out loudly

struct LLWord {  // 16 bytes, 16-byte aligned
    unsigned int lo, f0, hi, f1;
};

__device__ __forceinline__ void ll_store(LLWord* p, double v, unsigned int epoch)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(epoch),
                 "r"((unsigned int)(b >> 32)), "r"(epoch)
                 : "memory");
}

// spin until row `col` of this sweep is published; false on timeout
__device__ __forceinline__ bool ll_wait(const LLWord* p, unsigned int epoch, double& v, SolverScalars* S)
{
    unsigned int lo, f0, hi, f1;
    long long t0 = 0;
    for (int spin = 0;; spin++) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1)
                     : "l"(p)
                     : "memory");
        if (f0 == epoch && f1 == epoch) break;
        if (kFlowSleep) __nanosleep(spin < 4 ? 20 : 100);
        if ((spin & 1023) == 1023) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > kFlowTimeout) {
                S->commError = 2;
                S->done = 1;
                v = 0.0;
                return false;
            }
        }
    }
    v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    return true;
}

// Wait for up to kJoint published words TOGETHER: every round requests all that are still missing
// at once, so the wait costs one L2 round trip after the last of them lands.  (Waiting for them one
// after the other costs a round trip each: on a box the lower neighbours of a row all sit in the
// level before and arrive together — measured 0.62 us per level on a 1-D chain, 1.5 on a 2-D
// sheet, 1.6-2.3 on 3-D boxes with sequential waits.)
constexpr int kJoint = 4;

__device__ __forceinline__ bool ll_wait_joint(const LLWord* ll, const int (&col)[8], int n, unsigned int epoch,
                                              double (&v)[kJoint], SolverScalars* S)
{
    unsigned int lo[kJoint], f0[kJoint], hi[kJoint], f1[kJoint];
    bool ready[kJoint];
#pragma unroll
    for (int j = 0; j < kJoint; j++) ready[j] = j >= n;
    long long t0 = 0;
    for (int spin = 0;; spin++) {
#pragma unroll
        for (int j = 0; j < kJoint; j++)
            if (!ready[j])
                asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(lo[j]), "=r"(f0[j]), "=r"(hi[j]), "=r"(f1[j])
                             : "l"(ll + col[j])
                             : "memory");
        bool all = true;
#pragma unroll
        for (int j = 0; j < kJoint; j++) {
            if (!ready[j]) {
                if (f0[j] == epoch && f1[j] == epoch) {
                    ready[j] = true;
                    v[j] = __longlong_as_double((long long)(((unsigned long long)hi[j] << 32) | lo[j]));
                } else all = false;
            }
        }
        if (all) return true;
        if ((spin & 1023) == 1023) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > kFlowTimeout) {
                S->commError = 2;
                S->done = 1;
                return false;
            }
        }
    }
}

enum { FLOW_DIC = 0, FLOW_DIC_PRE = 1, FLOW_RD = 2, FLOW_GS = 3, FLOW_GS_STORE = 4 };
enum { FLOWB_DIC = 0, FLOWB_DIC_PRE = 1, FLOWB_GS = 2 };

struct FlowArgs {
    SolverScalars* S;
    bool guarded;
    int nChunks;
    unsigned int epoch;
    const int* chunkRows;    // [nChunks*32] row or -1, level order
    LLWord* ll;              // [nCells] published values of the running sweep
    const int* losortStart;
    const int* losort;
    const int* lowerCol;
    const int* ownerStart;
    const int* u;
    const double* rD;
    const double* coef;      // coefficient of the awaited entries
    const double* coef2;     // RD: lower[] ; GS: upper[] (non-awaited part)
    const double* diag;
    const double* r;         // DIC init source / GS bPrime / reverse GS bLower
    double* bLower;          // GS_STORE: bPrime after the lower part
    double* w;               // the field being produced (w, rD or psi)
    bool init;
};

template <int MODE>
__global__ void __launch_bounds__(kBlock) flow_fwd_kernel(FlowArgs a)
{
    if (a.guarded && a.S->done) return;
    const int lane = threadIdx.x & 31;
    // chunks are dealt round-robin to the warps of the (co-resident, cooperative)
    // grid: warp w handles chunks w, w+W, w+2W, ... in level order, so whatever a
    // row waits for belongs to a chunk some running warp reaches first
    const int nWarps = gridDim.x * (kBlock / 32);
    for (int chunk = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); chunk < a.nChunks; chunk += nWarps) {
        const int row = a.chunkRows[chunk * 32 + lane];
        // static data of the row goes to registers BEFORE the gate, so that only
        // the dependency values themselves sit on the critical path
        constexpr int kPre = 8;
        int pcol[kPre];
        double pc[kPre], pc2[kPre];
        int kBeg = 0, kEnd = 0;
        double acc = 0.0, rDc = 0.0;
        constexpr bool kGS = (MODE == FLOW_GS || MODE == FLOW_GS_STORE);
        constexpr int kPreU = kGS ? 6 : 1;
        double uc[kPreU], uv[kPreU], dg = 1.0;
        int fU0 = 0, fU1 = 0;
        if (row >= 0) {
            kBeg = a.losortStart[row];
            kEnd = a.losortStart[row + 1];
            if (MODE == FLOW_DIC || MODE == FLOW_DIC_PRE) {
                rDc = a.rD[row];
                acc = a.init ? __dmul_rn(rDc, a.r[row]) : a.w[row];
            } else if (MODE == FLOW_RD) {
                acc = a.diag[row];
            } else {
                acc = a.r[row];
            }
#pragma unroll
            for (int j = 0; j < kPre; j++) {
                if (kBeg + j < kEnd) {
                    const int f = a.losort[kBeg + j];
                    pcol[j] = a.lowerCol[kBeg + j];
                    const double c = a.coef[f];
                    pc[j] = (MODE == FLOW_DIC) ? __dmul_rn(rDc, c) : c;
                    pc2[j] = (MODE == FLOW_RD) ? a.coef2[f] : 0.0;
                }
            }
            if (kGS) {
                // Gauss-Seidel: the upper part uses values of the PREVIOUS sweep — the rows above
                // all await this row, so they cannot have been finalised yet and their old values
                // can be fetched now, off the critical path (only the arithmetic stays in order)
                fU0 = a.ownerStart[row];
                fU1 = a.ownerStart[row + 1];
                dg = a.diag[row];
#pragma unroll
                for (int j = 0; j < kPreU; j++) {
                    if (fU0 + j < fU1) {
                        uc[j] = a.coef2[fU0 + j];
                        uv[j] = a.w[a.u[fU0 + j]];
                    }
                }
            }
        }
        if (row >= 0) {
            double vj[kJoint];
            const int nAwait = kEnd - kBeg;
            bool okj = ll_wait_joint(a.ll, pcol, nAwait < kJoint ? nAwait : kJoint, a.epoch, vj, a.S);
#pragma unroll
            for (int j = 0; j < kPre; j++) {
                if (okj && kBeg + j < kEnd) {
                    double v;
                    if (j < kJoint) v = vj[j];
                    else if (!ll_wait(a.ll + pcol[j], a.epoch, v, a.S)) break;
                    if (MODE == FLOW_RD) acc = __dsub_rn(acc, __ddiv_rn(__dmul_rn(pc[j], pc2[j]), v));
                    else acc = __dsub_rn(acc, __dmul_rn(pc[j], v));
                }
            }
            for (int k = kBeg + kPre; k < kEnd; k++) {   // rows with more than kPre lower faces
                const int f = a.losort[k];
                const double c = a.coef[f];
                double v;
                if (!ll_wait(a.ll + a.lowerCol[k], a.epoch, v, a.S)) break;
                if (MODE == FLOW_DIC) acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rDc, c), v));
                else if (MODE == FLOW_RD) acc = __dsub_rn(acc, __ddiv_rn(__dmul_rn(c, a.coef2[f]), v));
                else acc = __dsub_rn(acc, __dmul_rn(c, v));
            }
            if (MODE == FLOW_GS || MODE == FLOW_GS_STORE) {
                if (MODE == FLOW_GS_STORE) a.bLower[row] = acc;
                // upper part: values of the previous sweep (rows above cannot have
                // been finalised yet: they all await this row)
#pragma unroll
                for (int j = 0; j < kPreU; j++)
                    if (fU0 + j < fU1) acc = __dsub_rn(acc, __dmul_rn(uc[j], uv[j]));
                for (int f = fU0 + kPreU; f < fU1; f++)   // cells that own more than kPreU faces
                    acc = __dsub_rn(acc, __dmul_rn(a.coef2[f], a.w[a.u[f]]));
                acc = __ddiv_rn(acc, dg);
            }
            a.w[row] = acc;
            ll_store(a.ll + row, acc, a.epoch);
        }
        __syncwarp();
    }
}

// backward sweeps: a row awaits its upper neighbours.  DIC/DILU/FDIC consume the
// upper faces in DESCENDING face order (DICPreconditioner.C:118-121), the reverse
// Gauss-Seidel sweep in ascending order (symGaussSeidelSmoother.C:190-196).
template <int MODE>
__global__ void __launch_bounds__(kBlock) flow_bwd_kernel(FlowArgs a)
{
    if (a.guarded && a.S->done) return;
    const int lane = threadIdx.x & 31;
    // chunks are dealt round-robin to the warps of the (co-resident, cooperative)
    // grid: warp w handles chunks w, w+W, w+2W, ... in level order, so whatever a
    // row waits for belongs to a chunk some running warp reaches first
    const int nWarps = gridDim.x * (kBlock / 32);
    for (int chunk = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); chunk < a.nChunks; chunk += nWarps) {
        const int row = a.chunkRows[chunk * 32 + lane];
        constexpr int kPre = 8;
        int pcol[kPre];
        double pc[kPre];
        int f0 = 0, fEnd = 0;
        double acc = 0.0, rDc = 0.0, dg = 1.0;
        if (row >= 0) {
            f0 = a.ownerStart[row];
            fEnd = a.ownerStart[row + 1];
            if (MODE == FLOWB_GS) {
                acc = a.r[row];
                dg = a.diag[row];   // before the gate: only the awaited values sit on the critical path
            } else {
                rDc = a.rD[row];
                acc = a.w[row];
            }
#pragma unroll
            for (int j = 0; j < kPre; j++) {
                // j-th entry in consumption order: ascending faces for the reverse
                // Gauss-Seidel sweep, descending for DIC/DILU/FDIC
                const int f = (MODE == FLOWB_GS) ? f0 + j : fEnd - 1 - j;
                if (f >= f0 && f < fEnd) {
                    pcol[j] = a.u[f];
                    const double c = a.coef[f];
                    pc[j] = (MODE == FLOWB_DIC) ? __dmul_rn(rDc, c) : c;
                }
            }
        }
        if (row >= 0) {
            const int deg = fEnd - f0;
            double vj[kJoint];
            bool okj = ll_wait_joint(a.ll, pcol, deg < kJoint ? deg : kJoint, a.epoch, vj, a.S);
#pragma unroll
            for (int j = 0; j < kPre; j++) {
                if (okj && j < deg) {
                    double v;
                    if (j < kJoint) v = vj[j];
                    else if (!ll_wait(a.ll + pcol[j], a.epoch, v, a.S)) break;
                    acc = __dsub_rn(acc, __dmul_rn(pc[j], v));
                }
            }
            for (int j = kPre; j < deg; j++) {
                const int f = (MODE == FLOWB_GS) ? f0 + j : fEnd - 1 - j;
                double v;
                if (!ll_wait(a.ll + a.u[f], a.epoch, v, a.S)) break;
                if (MODE == FLOWB_DIC) acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(rDc, a.coef[f]), v));
                else acc = __dsub_rn(acc, __dmul_rn(a.coef[f], v));
            }
            if (MODE == FLOWB_GS) acc = __ddiv_rn(acc, dg);
            a.w[row] = acc;
            ll_store(a.ll + row, acc, a.epoch);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool flow_enabled()
{
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("LDU_SWEEPS");  // "levels" selects the launch-per-level path
        mode = (e && std::string(e) == "levels") ? 0 : 1;
    }
    return mode == 1;
}

// level-ordered chunk table: each level starts a new chunk
struct FlowDir {   // device tables of one sweep direction (owned by the matrix)
    int* chunkRows = nullptr;
    int nChunks = 0;
};

static int build_dir(ldu_matrix* m, const Schedule& s, FlowDir& d)
{
    std::vector<int> rows(std::max(m->nCells, 1));
    if (m->nCells)
        LDU_CUDA(cudaMemcpy(rows.data(), s.d_rows, m->nCells * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<int> table;
    table.reserve((size_t)m->nCells + 32 * (size_t)s.nLevels);
    for (int L = 0; L < s.nLevels; L++) {   // each level starts a new chunk
        for (int i = s.levelStart[L]; i < s.levelStart[L + 1]; i++) table.push_back(rows[i]);
        while (table.size() % 32) table.push_back(-1);
    }
    d.nChunks = (int)(table.size() / 32);
    LDU_CUDA(cudaMalloc((void**)&d.chunkRows, std::max<size_t>(table.size(), 1) * sizeof(int)));
    if (!table.empty())
        LDU_CUDA(cudaMemcpy(d.chunkRows, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice));
    return LDU_OK;
}

void flow_free(ldu_matrix* m)
{
    for (int i = 0; i < 2; i++) {
        FlowDir* d = reinterpret_cast<FlowDir*>(m->flowDir[i]);
        if (!d) continue;
        cudaFree(d->chunkRows);
        delete d;
        m->flowDir[i] = nullptr;
    }
    cudaFree(m->d_ll);
    m->d_ll = nullptr;
}

static int flow_prepare(ldu_matrix* m, FlowArgs& a, bool guarded, bool backward)
{
    if (!m->d_ll) {
        LDU_TRY(build_schedules(m));
        FlowDir* f = new FlowDir();
        FlowDir* b = new FlowDir();
        m->flowDir[0] = f;
        m->flowDir[1] = b;
        LDU_TRY(build_dir(m, m->fwd, *f));
        LDU_TRY(build_dir(m, m->bwd, *b));
        LDU_CUDA(cudaMalloc((void**)&m->d_ll, std::max(m->nCells, 1) * sizeof(LLWord)));
        LDU_CUDA(cudaMemsetAsync(m->d_ll, 0, std::max(m->nCells, 1) * sizeof(LLWord), m->ctx->stream));
        m->flowEpoch = 0;
    }
    const FlowDir* d = reinterpret_cast<const FlowDir*>(m->flowDir[backward ? 1 : 0]);
    a = FlowArgs();
    a.S = m->d_scalars;
    a.guarded = guarded;
    a.nChunks = d->nChunks;
    a.chunkRows = d->chunkRows;
    a.epoch = (unsigned int)(++m->flowEpoch);
    a.ll = reinterpret_cast<LLWord*>(m->d_ll);
    a.losortStart = m->d_losortStart;
    a.losort = m->d_losort;
    a.lowerCol = m->d_lowerCol;
    a.ownerStart = m->d_ownerStart;
    a.u = m->d_u;
    a.diag = m->d_diag;
    return LDU_OK;
}

// All warps of a sweep must be co-resident (they wait on each other): the grid is
// capped by the occupancy of the kernel and launched cooperatively, which makes
// the driver refuse a grid that cannot be resident instead of deadlocking.
template <class K>
static int flow_launch(ldu_matrix* m, K kernel, FlowArgs& a)
{
    static int perSm = -1;
    if (perSm < 0) {
        const char* e = getenv("LDU_FLOW_BLOCKS");
        perSm = e ? atoi(e) : 2;   // measured on B200 (216^3 DIC-PCG): 1 -> 250, 2 -> 265, 4 -> 218, 8 -> 202 it/s
        if (perSm < 1) perSm = 1;
    }
    int maxPerSm = 0;
    LDU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxPerSm, kernel, kBlock, 0));
    if (maxPerSm < 1) {
        set_error("dataflow sweep kernel cannot be resident");
        return LDU_ECUDA;
    }
    const int blocksWanted = (a.nChunks + kBlock / 32 - 1) / (kBlock / 32);
    const int grid = std::max(1, std::min(blocksWanted, m->ctx->smCount * std::min(perSm, maxPerSm)));
    void* params[] = {&a};
    LDU_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(kBlock), params, 0, m->ctx->stream));
    count_launch();
    return LDU_OK;
}

#define FLOW_LAUNCH(kernel, a) LDU_TRY(flow_launch(m, kernel, a))

int flow_forward(ldu_matrix* m, const double* rD, const double* coef, bool pre, const double* r, double* w,
                 bool init)
{
    if (m->nCells <= 0) return LDU_OK;
    FlowArgs a;
    LDU_TRY(flow_prepare(m, a, true, false));
    a.rD = rD;
    a.coef = coef;
    a.r = r;
    a.w = w;
    a.init = init;
    if (pre) FLOW_LAUNCH(flow_fwd_kernel<FLOW_DIC_PRE>, a);
    else FLOW_LAUNCH(flow_fwd_kernel<FLOW_DIC>, a);
    return LDU_OK;
}

int flow_backward(ldu_matrix* m, const double* rD, const double* coef, bool pre, double* w)
{
    if (m->nCells <= 0) return LDU_OK;
    FlowArgs a;
    LDU_TRY(flow_prepare(m, a, true, true));
    a.rD = rD;
    a.coef = coef;
    a.w = w;
    if (pre) FLOW_LAUNCH(flow_bwd_kernel<FLOWB_DIC_PRE>, a);
    else FLOW_LAUNCH(flow_bwd_kernel<FLOWB_DIC>, a);
    return LDU_OK;
}

int flow_rD(ldu_matrix* m, double* rD, const double* upper, const double* lower)
{
    if (m->nCells <= 0) return LDU_OK;
    FlowArgs a;
    LDU_TRY(flow_prepare(m, a, false, false));
    a.coef = upper;
    a.coef2 = lower;
    a.w = rD;
    FLOW_LAUNCH(flow_fwd_kernel<FLOW_RD>, a);
    return LDU_OK;
}

int flow_gs(ldu_matrix* m, const double* bPrime, double* bLower, double* psi, bool sym)
{
    if (m->nCells <= 0) return LDU_OK;
    FlowArgs a;
    LDU_TRY(flow_prepare(m, a, true, false));
    a.coef = m->d_lower;
    a.coef2 = m->d_upper;
    a.r = bPrime;
    a.bLower = bLower;
    a.w = psi;
    if (sym) {
        FLOW_LAUNCH(flow_fwd_kernel<FLOW_GS_STORE>, a);
        FlowArgs b;
        LDU_TRY(flow_prepare(m, b, true, true));
        b.coef = m->d_upper;
        b.r = bLower;
        b.w = psi;
        FLOW_LAUNCH(flow_bwd_kernel<FLOWB_GS>, b);
    } else {
        FLOW_LAUNCH(flow_fwd_kernel<FLOW_GS>, a);
    }
    return LDU_OK;
}

}  // namespace ldu
