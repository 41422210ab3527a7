// Line-pipelined triangular sweeps for structured hex boxes.
//
// blockMesh boxes (the icoFoam cavity family: the PCG configurations of
// BASELINE.json) number cells lexicographically, c = (k*ny + j)*nx + i, and list
// the faces of a cell in the order +i, +j, +k.  ldu_matrix_create detects that
// pattern from the addressing alone (detect_box).  For such matrices the DIC /
// DILU / FDIC substitutions — whose dependency DAG is the i+j+k hyperplane
// order, 3n-2 levels deep — are run as a pipelined wavefront:
//
//   * one warp owns a tile of 32 consecutive x-lines of one k-plane; lane l
//     walks line j0+l along i, skewed by l steps, so that
//       - the i-neighbour is the lane's own previous result (a register),
//       - the j-neighbour is the previous result of lane l-1 (one warp shuffle),
//       - only the k-neighbour (same lane of the tile one plane below) and lane
//         0's j-neighbour (lane 31 of the previous tile) come from other warps;
//   * those two are read from a {value, epoch} publication buffer polled in L2
//     (same 16-byte words as flow.cu), so a plane trails the plane below it by a
//     few steps instead of a kernel launch or a grid barrier;
//   * all per-row operands the sweep streams (coefficients of the three lower
//     and three upper faces, rD, diag) are kept in a *skewed tile layout*
//     [tile][step][lane], written once per coefficient update, so every step of
//     a warp is one coalesced 256-byte row per operand.
//
// Arithmetic per row is the reference's: terms in ascending face order for the
// forward sweeps (k-, j-, i-neighbour), descending for the backward ones
// (k+, j+, i+), unfused multiply/subtract — results are BIT-IDENTICAL to
// DICPreconditioner.C:87-123 / DILUPreconditioner.C:88-185 and to the generic
// dataflow path.  Matrices that are not boxes never come here.
#include <algorithm>
#include <cstdlib>

#include "reduce.cuh"
#include "sweeps.h"

namespace ldu {

constexpr int kSWarps = 4;  // warps per CTA of the sweep kernels
constexpr long long kStencilTimeout = 4000000000ll;

struct LLW {
    unsigned int lo, f0, hi, f1;
};

__device__ __forceinline__ void llw_store(LLW* p, double v, unsigned int epoch)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(epoch),
                 "r"((unsigned int)(b >> 32)), "r"(epoch)
                 : "memory");
}

__device__ __forceinline__ double llw_wait(const LLW* p, unsigned int epoch, SolverScalars* S)
{
    unsigned int lo, f0, hi, f1;
    long long t0 = 0;
    for (int spin = 0;; spin++) {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1)
                     : "l"(p)
                     : "memory");
        if (f0 == epoch && f1 == epoch) break;
        if ((spin & 1023) == 1023) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > kStencilTimeout) {
                S->commError = 2;
                S->done = 1;
                return 0.0;
            }
        }
    }
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

struct BoxDev {
    int nx, ny, nz;
    int nJ;      // tiles per plane = ceil(ny/32)
    int steps;   // nx + 31 skewed steps per tile
    int nTiles;  // nz*nJ
};

__host__ __device__ inline long long tile_pos(const BoxDev& b, int i, int j, int k)
{
    const int J = j >> 5, l = j & 31;
    return ((long long)(k * b.nJ + J) * b.steps + (i + l)) * 32 + l;
}

// ---------------------------------------------------------------------------
// coefficient view in skewed tile layout (once per ldu_matrix_set_coeffs)
// ---------------------------------------------------------------------------
struct StencilView {
    // coefficient multiplying the k-/j-/i- neighbour (faces where this cell is the
    // upper cell) and the i+/j+/k+ neighbour (faces it owns); "U" = upper[], "L" = lower[]
    double* UL[3];  // upper[f] of the lower faces   (order k, j, i)
    double* LL[3];  // lower[f] of the lower faces   (aliases UL when symmetric)
    double* UU[3];  // upper[f] of the upper faces   (order i, j, k)
    double* LU[3];  // lower[f] of the upper faces   (aliases UU when symmetric)
    double* diag;
};

__global__ void __launch_bounds__(kBlock) stencil_build_kernel(BoxDev b, const int* __restrict__ ownerStart,
                                                               const double* __restrict__ diag,
                                                               const double* __restrict__ upper,
                                                               const double* __restrict__ lower, bool asym,
                                                               StencilView v)
{
    const int n = b.nx * b.ny * b.nz;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock) {
        const int i = c % b.nx, j = (c / b.nx) % b.ny, k = c / (b.nx * b.ny);
        const long long p = tile_pos(b, i, j, k);
        const int hasI = i < b.nx - 1, hasJ = j < b.ny - 1, hasK = k < b.nz - 1;
        const int os = ownerStart[c];
        const int fI = os, fJ = os + hasI, fK = os + hasI + hasJ;
        v.diag[p] = diag[c];
        v.UU[0][p] = hasI ? upper[fI] : 0.0;
        v.UU[1][p] = hasJ ? upper[fJ] : 0.0;
        v.UU[2][p] = hasK ? upper[fK] : 0.0;
        // lower faces: the +k / +j / +i face of the cell below / behind / to the left
        // (those cells have the same i, j flags where it matters)
        const int gK = k > 0 ? ownerStart[c - b.nx * b.ny] + hasI + hasJ : -1;
        const int gJ = j > 0 ? ownerStart[c - b.nx] + hasI : -1;
        const int gI = i > 0 ? ownerStart[c - 1] : -1;
        v.UL[0][p] = gK >= 0 ? upper[gK] : 0.0;
        v.UL[1][p] = gJ >= 0 ? upper[gJ] : 0.0;
        v.UL[2][p] = gI >= 0 ? upper[gI] : 0.0;
        if (asym) {
            v.LU[0][p] = hasI ? lower[fI] : 0.0;
            v.LU[1][p] = hasJ ? lower[fJ] : 0.0;
            v.LU[2][p] = hasK ? lower[fK] : 0.0;
            v.LL[0][p] = gK >= 0 ? lower[gK] : 0.0;
            v.LL[1][p] = gJ >= 0 ? lower[gJ] : 0.0;
            v.LL[2][p] = gI >= 0 ? lower[gI] : 0.0;
        }
    }
}

__global__ void __launch_bounds__(kBlock) stencil_scatter_kernel(BoxDev b, const double* __restrict__ src,
                                                                 double* __restrict__ dst)
{
    const int n = b.nx * b.ny * b.nz;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock) {
        const int i = c % b.nx, j = (c / b.nx) % b.ny, k = c / (b.nx * b.ny);
        dst[tile_pos(b, i, j, k)] = src[c];
    }
}

// ---------------------------------------------------------------------------
// sweeps
// ---------------------------------------------------------------------------
struct StencilArgs {
    SolverScalars* S;
    bool guarded;
    BoxDev b;
    unsigned int epoch;
    LLW* ll;            // [nTiles*steps*32]
    const double* rDt;  // tile layout
    const double* c0;   // coefficient of the first awaited neighbour in consumption order
    const double* c1;
    const double* c2;
    const double* r;    // natural layout (forward init source)
    double* w;          // natural layout, in/out
    bool init;
};

// One sweep kernel for both directions.
//   forward  (BWD = false): w[c] = rD*r[c] - (rD*ck)*w[c-nx*ny] - (rD*cj)*w[c-nx] - (rD*ci)*w[c-1]
//   backward (BWD = true) : w[c] = w[c]    - (rD*ck)*w[c+nx*ny] - (rD*cj)*w[c+nx] - (rD*ci)*w[c+1]
// Tiles are taken in dependency order (ascending / descending), the rows of the
// skewed layout are walked up / down, and the j-neighbour comes from the lane
// below / above.  The operands of the next group of kU steps — coefficient rows,
// source values and the published {value, epoch} words of the k-neighbours — are
// loaded while the current group is computed; a prefetched word whose epoch is
// not there yet is polled again when it is needed (warp-uniform slow path).
//
// One warp per tile makes the loop issue- and latency-bound (ncu on the first
// version: 190 instructions per step, 3.4 cycles each), so the body is branch-free:
// every lane executes every step, operand rows are loaded unconditionally (the tile
// arrays are zero-padded by kPadRows rows at both ends), and whether a cell exists
// only predicates the subtractions, the stores and the update of `prev`.
constexpr int kU = 4;
constexpr int kPadRows = 2 * kU;

struct StepOps {
    double rD, ck, cj, ci, src;
    unsigned int klo, kf0, khi, kf1;   // k-neighbour word (all lanes)
    unsigned int jlo, jf0, jhi, jf1;   // cross-tile j-neighbour word (edge lane only)
};

__device__ __forceinline__ void ll_peek(const LLW* p, unsigned int& lo, unsigned int& f0, unsigned int& hi,
                                        unsigned int& f1)
{
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1)
                 : "l"(p)
                 : "memory");
}

template <bool BWD, bool INIT, bool HASK>
__device__ __forceinline__ void stencil_tile(const StencilArgs& a, int T, int k, int J, int lane)
{
    const int nx = a.b.nx, ny = a.b.ny, nJ = a.b.nJ, steps = a.b.steps;
    const unsigned int epoch = a.epoch;
    const long long tileStride = (long long)steps * 32;
    const int j = J * 32 + lane;
    const bool lineValid = j < ny;
    const bool hasJ = lineValid && (BWD ? (j < ny - 1) : (j > 0));
    const bool edgeJ = hasJ && (BWD ? (lane == 31) : (lane == 0));   // j-neighbour lives in another tile
    // element of layout row tf is ptr[tf*32]
    const long long base = (long long)T * tileStride + lane;
    const double* pRD = a.rDt + base;
    const double* pC0 = a.c0 + base;
    const double* pC1 = a.c1 + base;
    const double* pC2 = a.c2 + base;
    LLW* pSelf = a.ll + base;
    const LLW* pK = a.ll + base + (BWD ? (long long)nJ : -(long long)nJ) * tileStride;
    // forward: lane 31 of the previous tile, row tf+31; backward: lane 0 of the next tile, row tf-31
    const LLW* pJ = edgeJ ? (BWD ? a.ll + (long long)(T + 1) * tileStride - 31 * 32
                                 : a.ll + (long long)(T - 1) * tileStride + 31 * 32 + 31)
                          : pSelf;   // any valid address: the word is ignored
    // natural layout: cell i = tf - lane of line j
    const long long rowBase = ((long long)k * ny + j) * nx - lane;
    const double* pSrc = (INIT ? a.r : a.w) + rowBase;
    double* pW = a.w + rowBase;
    const int tLo = lineValid ? lane : 0x7fffffff;   // the cell exists for tLo <= tf < tHi
    const int tHi = lineValid ? lane + nx : -1;
    const int tI = BWD ? lane + nx - 1 : lane;        // row without an i-neighbour

    auto load = [&](int tf, StepOps& o) {
        const int e = tf * 32;
        o.rD = pRD[e];
        o.ck = pC0[e];
        o.cj = pC1[e];
        o.ci = pC2[e];
        o.src = (tf >= tLo && tf < tHi) ? pSrc[tf] : 0.0;
        if (HASK) ll_peek(pK + e, o.klo, o.kf0, o.khi, o.kf1);
        ll_peek(pJ + e, o.jlo, o.jf0, o.jhi, o.jf1);
    };

    double prev = 0.0;
    auto step = [&](int tf, const StepOps& o) {
        const bool active = tf >= tLo && tf < tHi;
        const int e = tf * 32;
        double vj = BWD ? __shfl_down_sync(0xffffffffu, prev, 1) : __shfl_up_sync(0xffffffffu, prev, 1);
        double vk = HASK ? __hiloint2double((int)o.khi, (int)o.klo) : 0.0;
        if (edgeJ) vj = __hiloint2double((int)o.jhi, (int)o.jlo);
        // prefetched words that had not been published yet: poll (rare once the
        // pipeline is primed; warp-uniform branch)
        const bool lateK = HASK && active && (o.kf0 != epoch || o.kf1 != epoch);
        const bool lateJ = edgeJ && active && (o.jf0 != epoch || o.jf1 != epoch);
        if (__any_sync(0xffffffffu, lateK || lateJ)) {
            if (lateK) vk = llw_wait(pK + e, epoch, a.S);
            if (lateJ) vj = llw_wait(pJ + e, epoch, a.S);
        }
        double acc = INIT ? __dmul_rn(o.rD, o.src) : o.src;
        if (HASK) acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(o.rD, o.ck), vk));
        const double tj = __dmul_rn(__dmul_rn(o.rD, o.cj), vj);
        if (hasJ) acc = __dsub_rn(acc, tj);
        const double ti = __dmul_rn(__dmul_rn(o.rD, o.ci), prev);
        if (tf != tI) acc = __dsub_rn(acc, ti);
        if (active) {
            pW[tf] = acc;
            llw_store(pSelf + e, acc, epoch);
            prev = acc;
        }
    };

    // layout rows are visited 0,1,2,... (forward) or steps-1, steps-2, ... (backward);
    // the last group may run up to kU-1 rows past the tile (padding rows, no cell)
    const int t0 = BWD ? steps - 1 : 0, dt = BWD ? -1 : 1;
    const int nGroups = (steps + kU - 1) / kU;
    StepOps cur[kU], nxt[kU];
#pragma unroll
    for (int q = 0; q < kU; q++) load(t0 + dt * q, cur[q]);
    for (int g = 0; g < nGroups; g++) {
        const int tg = t0 + dt * g * kU;
#pragma unroll
        for (int q = 0; q < kU; q++) load(tg + dt * (kU + q), nxt[q]);
#pragma unroll
        for (int q = 0; q < kU; q++) step(tg + dt * q, cur[q]);
#pragma unroll
        for (int q = 0; q < kU; q++) cur[q] = nxt[q];
    }
}

template <bool BWD, bool INIT>
__global__ void __launch_bounds__(kSWarps * 32) stencil_sweep_kernel(StencilArgs a)
{
    if (a.guarded && a.S->done) return;
    const int lane = threadIdx.x & 31;
    const int nWarps = gridDim.x * kSWarps;
    for (int Tr = blockIdx.x * kSWarps + (threadIdx.x >> 5); Tr < a.b.nTiles; Tr += nWarps) {
        const int T = BWD ? a.b.nTiles - 1 - Tr : Tr;
        const int k = T / a.b.nJ, J = T - k * a.b.nJ;
        const bool hasK = BWD ? (k < a.b.nz - 1) : (k > 0);   // warp-uniform
        if (hasK) stencil_tile<BWD, INIT, true>(a, T, k, J, lane);
        else stencil_tile<BWD, INIT, false>(a, T, k, J, lane);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct StencilState {
    BoxDev b;
    long long padded = 0;
    StencilView v;
    bool asym = false;
    long long viewGen = -1;      // m->coefGen the view was built for
    LLW* ll = nullptr;
    unsigned int epoch = 0;
    // rD in tile layout, keyed by (source pointer, sweep generation); two slots
    double* rDt[2] = {nullptr, nullptr};
    const double* rDsrc[2] = {nullptr, nullptr};
    long long rDgen[2] = {-1, -1};
    int rDnext = 0;
};

// tile arrays carry kPadRows zero rows in front of tile 0 and behind the last tile
static int alloc_padded(void** user, size_t elems, size_t elemBytes, cudaStream_t st)
{
    const size_t pad = (size_t)kPadRows * 32;
    unsigned char* raw = nullptr;
    LDU_CUDA(cudaMalloc((void**)&raw, (elems + 2 * pad) * elemBytes));
    LDU_CUDA(cudaMemsetAsync(raw, 0, (elems + 2 * pad) * elemBytes, st));
    *user = raw + pad * elemBytes;
    return LDU_OK;
}

static void free_padded(void* user, size_t elemBytes)
{
    if (user) cudaFree((unsigned char*)user - (size_t)kPadRows * 32 * elemBytes);
}

void stencil_free(ldu_matrix* m)
{
    StencilState* s = reinterpret_cast<StencilState*>(m->stencil);
    if (!s) return;
    for (int d = 0; d < 3; d++) {
        free_padded(s->v.UL[d], sizeof(double));
        free_padded(s->v.UU[d], sizeof(double));
        if (s->asym) {
            free_padded(s->v.LL[d], sizeof(double));
            free_padded(s->v.LU[d], sizeof(double));
        }
    }
    free_padded(s->v.diag, sizeof(double));
    free_padded(s->ll, sizeof(LLW));
    free_padded(s->rDt[0], sizeof(double));
    free_padded(s->rDt[1], sizeof(double));
    delete s;
    m->stencil = nullptr;
}

static int stencil_state(ldu_matrix* m, StencilState** out)
{
    StencilState* s = reinterpret_cast<StencilState*>(m->stencil);
    const bool asym = !m->symmetric;
    if (s && s->asym != asym) {  // symmetry changed between solves: rebuild storage
        stencil_free(m);
        s = nullptr;
    }
    if (!s) {
        s = new StencilState();
        m->stencil = s;
        s->b.nx = m->box[0];
        s->b.ny = m->box[1];
        s->b.nz = m->box[2];
        s->b.nJ = (s->b.ny + 31) / 32;
        s->b.steps = s->b.nx + 31;
        s->b.nTiles = s->b.nz * s->b.nJ;
        s->padded = (long long)s->b.nTiles * s->b.steps * 32;
        s->asym = asym;
        cudaStream_t st = m->ctx->stream;
        const size_t ne = (size_t)s->padded;
        for (int d = 0; d < 3; d++) {
            LDU_TRY(alloc_padded((void**)&s->v.UL[d], ne, sizeof(double), st));
            LDU_TRY(alloc_padded((void**)&s->v.UU[d], ne, sizeof(double), st));
            if (asym) {
                LDU_TRY(alloc_padded((void**)&s->v.LL[d], ne, sizeof(double), st));
                LDU_TRY(alloc_padded((void**)&s->v.LU[d], ne, sizeof(double), st));
            } else {
                s->v.LL[d] = s->v.UL[d];
                s->v.LU[d] = s->v.UU[d];
            }
        }
        LDU_TRY(alloc_padded((void**)&s->v.diag, ne, sizeof(double), st));
        LDU_TRY(alloc_padded((void**)&s->ll, ne, sizeof(LLW), st));
        LDU_TRY(alloc_padded((void**)&s->rDt[0], ne, sizeof(double), st));
        LDU_TRY(alloc_padded((void**)&s->rDt[1], ne, sizeof(double), st));
    }
    if (s->viewGen != m->coefGen) {
        const int grid = grid_for(m->ctx, m->nCells);
        stencil_build_kernel<<<grid, kBlock, 0, m->ctx->stream>>>(s->b, m->d_ownerStart, m->d_diag, m->d_upper,
                                                                  m->d_lower, asym, s->v);
        count_launch();
        LDU_CUDA(cudaGetLastError());
        s->viewGen = m->coefGen;
    }
    *out = s;
    return LDU_OK;
}

static int stencil_rD(ldu_matrix* m, StencilState* s, const double* rD, const double** out)
{
    for (int q = 0; q < 2; q++)
        if (s->rDsrc[q] == rD && s->rDgen[q] == m->sweepGen) {
            *out = s->rDt[q];
            return LDU_OK;
        }
    const int q = s->rDnext;
    s->rDnext ^= 1;
    stencil_scatter_kernel<<<grid_for(m->ctx, m->nCells), kBlock, 0, m->ctx->stream>>>(s->b, rD, s->rDt[q]);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    s->rDsrc[q] = rD;
    s->rDgen[q] = m->sweepGen;
    *out = s->rDt[q];
    return LDU_OK;
}

template <class K>
static int stencil_launch(ldu_matrix* m, K kernel, StencilArgs& a)
{
    int maxPerSm = 0;
    LDU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxPerSm, kernel, kSWarps * 32, 0));
    if (maxPerSm < 1) {
        set_error("stencil sweep kernel cannot be resident");
        return LDU_ECUDA;
    }
    const int blocksWanted = (a.b.nTiles + kSWarps - 1) / kSWarps;
    const int grid = std::max(1, std::min(blocksWanted, m->ctx->smCount * maxPerSm));
    void* params[] = {&a};
    LDU_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(kSWarps * 32), params, 0,
                                         m->ctx->stream));
    count_launch();
    return LDU_OK;
}

// which: coefficient array of the matrix the generic caller passed (upper or lower)
int stencil_forward(ldu_matrix* m, const double* rD, const double* coef, const double* r, double* w, bool init)
{
    StencilState* s;
    LDU_TRY(stencil_state(m, &s));
    StencilArgs a;
    a.S = m->d_scalars;
    a.guarded = true;
    a.b = s->b;
    a.epoch = ++s->epoch;
    a.ll = s->ll;
    LDU_TRY(stencil_rD(m, s, rD, &a.rDt));
    double* const* c = (coef == m->d_upper) ? s->v.UL : s->v.LL;   // coefficient on the lower faces
    a.c0 = c[0];
    a.c1 = c[1];
    a.c2 = c[2];
    a.r = r;
    a.w = w;
    a.init = init;
    return init ? stencil_launch(m, stencil_sweep_kernel<false, true>, a)
                : stencil_launch(m, stencil_sweep_kernel<false, false>, a);
}

int stencil_backward(ldu_matrix* m, const double* rD, const double* coef, double* w)
{
    StencilState* s;
    LDU_TRY(stencil_state(m, &s));
    StencilArgs a;
    a.S = m->d_scalars;
    a.guarded = true;
    a.b = s->b;
    a.epoch = ++s->epoch;
    a.ll = s->ll;
    LDU_TRY(stencil_rD(m, s, rD, &a.rDt));
    double* const* c = (coef == m->d_upper) ? s->v.UU : s->v.LU;   // coefficient on the upper faces
    a.c0 = c[2];  // descending face order: +k, +j, +i
    a.c1 = c[1];
    a.c2 = c[0];
    a.r = nullptr;
    a.w = w;
    a.init = false;
    return stencil_launch(m, stencil_sweep_kernel<true, false>, a);
}

}  // namespace ldu
