// lduMatrix::Amul / Tmul / sumA / residual as cell-row kernels.
//
// The reference walks faces and scatter-adds into both cells of each face
// (matrices/lduMatrix/lduMatrix/lduMatrixATmul.C:34-92).  Here each cell row is
// gathered by one thread from the two CSR views of the LDU addressing:
//   lower part of row c: k in [losortStart[c], losortStart[c+1]) -> face losort[k],
//                        column l[face], coefficient lower[face]
//   upper part of row c: f in [ownerStart[c], ownerStart[c+1])   -> column u[f],
//                        coefficient upper[f]
// and accumulated in exactly the order the reference's face loop reaches that
// row (diag, lower faces ascending, upper faces ascending) with separate
// multiply and add (no FMA), so Apsi is BIT-IDENTICAL to the reference.
// Symmetric matrices keep a single coefficient array (lower aliases upper), so
// the HBM traffic stays at the LDU minimum: each coefficient is fetched from
// DRAM once and served to its second row from L2.
#include <algorithm>
#include <cstdlib>

#include "epilogue.cuh"
#include "comm.cuh"
#include "reduce.cuh"

namespace ldu {

struct RowView {
    const int* __restrict__ ownerStart;
    const int* __restrict__ losortStart;
    const int* __restrict__ losort;
    const int* __restrict__ lowerCol;
    const int* __restrict__ lowerPacked;   // (column << 5 | face position), see ldu_internal.h
    const int* __restrict__ u;
    const double* __restrict__ diag;
    const double* __restrict__ lowerCoef;  // coefficient applied to the lower-part entries
    const double* __restrict__ upperCoef;  // coefficient applied to the upper-part entries
};

static RowView row_view(const ldu_matrix* m, bool transpose)
{
    RowView v;
    v.ownerStart = m->d_ownerStart;
    v.losortStart = m->d_losortStart;
    v.losort = m->d_losort;
    v.lowerCol = m->d_lowerCol;
    v.lowerPacked = m->d_lowerPacked;
    v.u = m->d_u;
    v.diag = m->d_diag;
    // Amul: Apsi[u] += lower*psi[l]; Apsi[l] += upper*psi[u]
    // Tmul: Tpsi[u] += upper*psi[l]; Tpsi[l] += lower*psi[u]   (lduMatrixATmul.C:134-138)
    v.lowerCoef = transpose ? m->d_upper : m->d_lower;
    v.upperCoef = transpose ? m->d_lower : m->d_upper;
    return v;
}

// MODE 0: y = A x   1: y = b - A x (residual)   2: y = rowsum(A) (sumA)   3: y = (A - diag) x
//      4: y = H(x) = -(A - diag) x (lduMatrixTemplates.C:33-65)   5: y = H1 = -rowsum(A - diag) (lduMatrixATmul.C:298-327)
// One term of a row in the reference's order; `valid` predicates it (rows are processed in
// batches of kBatch entries whose loads are all issued before the first use, so that a thread
// has ~2*kBatch gathers in flight instead of one dependent chain per face).
template <int MODE>
__device__ __forceinline__ double row_term(double acc, double a, double xv, bool valid)
{
    double r;
    if (MODE == 0 || MODE == 3) r = __dadd_rn(acc, __dmul_rn(a, xv));
    else if (MODE == 1 || MODE == 4) r = __dsub_rn(acc, __dmul_rn(a, xv));
    else if (MODE == 5) r = __dsub_rn(acc, a);
    else r = __dadd_rn(acc, a);
    return valid ? r : acc;
}

constexpr int kBatch = 4;

template <int MODE, bool PACKED>
__device__ __forceinline__ double row_apply(const RowView& v, int c, const double* __restrict__ x,
                                            const double* __restrict__ b)
{
    const int k0 = v.losortStart[c], k1 = v.losortStart[c + 1];
    const int f0 = v.ownerStart[c], f1 = v.ownerStart[c + 1];
    double acc;
    if (MODE == 0) acc = __dmul_rn(v.diag[c], x[c]);
    else if (MODE == 1) acc = __dsub_rn(b[c], __dmul_rn(v.diag[c], x[c]));
    else if (MODE == 2) acc = v.diag[c];
    else acc = 0.0;
    constexpr bool kNoX = (MODE == 2 || MODE == 5);
    for (int k = k0; k < k1; k += kBatch) {
        int col[kBatch], face[kBatch];
        double a[kBatch], xv[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; j++) {
            const int kk = min(k + j, k1 - 1);   // clamped: a valid entry, its term is dropped below
            if (PACKED) {   // 4 bytes per lower entry from DRAM; ownerStart[col] is an L2 hit
                const int w = v.lowerPacked[kk];
                col[j] = w >> 5;
                face[j] = w & 31;
            } else {
                col[j] = v.lowerCol[kk];
                face[j] = v.losort[kk];
            }
        }
#pragma unroll
        for (int j = 0; j < kBatch; j++) {
            if (PACKED) face[j] += v.ownerStart[col[j]];
            if (!kNoX) xv[j] = x[col[j]];
            else xv[j] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < kBatch; j++) a[j] = v.lowerCoef[face[j]];
#pragma unroll
        for (int j = 0; j < kBatch; j++) acc = row_term<MODE>(acc, a[j], xv[j], k + j < k1);
    }
    for (int f = f0; f < f1; f += kBatch) {
        int col[kBatch];
        double a[kBatch], xv[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; j++) {
            const int ff = min(f + j, f1 - 1);
            a[j] = v.upperCoef[ff];
            col[j] = v.u[ff];
        }
#pragma unroll
        for (int j = 0; j < kBatch; j++) xv[j] = (!kNoX) ? x[col[j]] : 0.0;
#pragma unroll
        for (int j = 0; j < kBatch; j++) acc = row_term<MODE>(acc, a[j], xv[j], f + j < f1);
    }
    return acc;
}

template <int MODE, bool PACKED, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) row_kernel(int n, RowView v, double* __restrict__ y,
                                                      const double* __restrict__ x,
                                                      const double* __restrict__ b,
                                                      const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock)
        y[c] = row_apply<MODE, PACKED>(v, c, x, b);
}

// ---------------------------------------------------------------------------
// Row kernel for blockMesh boxes.  detect_box (context.cu) has established that cell
// c = (k*ny + j)*nx + i owns exactly its +i, +j, +k faces, in that order: columns and face
// indices follow from (i, j, k) alone, so the kernel reads NO addressing array — only diag, x,
// the coefficients (each from DRAM once: the three faces a cell owns stream coalesced, the three
// faces below / behind / left of it were streamed by those neighbours and hit L2) and writes y.
// DRAM traffic 24 N + 8 F instead of the LDU-minimal 24 N + 16 F the roofline counts.  Same
// per-row order as row_kernel (diag; k-, j-, i- neighbour; +i, +j, +k neighbour), unfused: the
// results are bit-identical.
// ---------------------------------------------------------------------------
struct BoxView {
    int nx, ny, nz;
    unsigned int mulNx, mulNy;   // ceil(2^32 / nx), ceil(2^32 / ny): c / nx = umulhi(c, mulNx) for c < 2^31 ... checked on the host
    const double* __restrict__ diag;
    const double* __restrict__ lowerCoef;
    const double* __restrict__ upperCoef;
};

// number of faces owned by the cells before c = (i, j, k): every cell owns 3 minus one for each
// of i == nx-1, j == ny-1, k == nz-1
__device__ __forceinline__ int box_owner_start(const BoxView& v, int c, int i, int j, int k)
{
    const int jk = k * v.ny + j;                      // cells before c with i == nx-1: one per finished line
    int start = 3 * c - jk;
    start -= k * v.nx + (j == v.ny - 1 ? i : 0);      // ... with j == ny-1
    start -= (k == v.nz - 1) ? c - k * v.nx * v.ny : 0;   // ... with k == nz-1
    return start;
}

template <int MODE, int MINB>   // 0: y = A x   1: y = b - A x
__global__ void __launch_bounds__(kBlock, MINB) box_row_kernel(int n, BoxView v, double* __restrict__ y,
                                                             const double* __restrict__ x,
                                                             const double* __restrict__ b,
                                                             const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    const int nx = v.nx, ny = v.ny, nz = v.nz, nxy = nx * ny;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock) {
        const int jk = (int)__umulhi((unsigned int)c, v.mulNx);
        const int i = c - jk * nx;
        const int k = (int)__umulhi((unsigned int)jk, v.mulNy);
        const int j = jk - k * ny;
        const bool hasI = i < nx - 1, hasJ = j < ny - 1, hasK = k < nz - 1;
        const int os = box_owner_start(v, c, i, j, k);
        // faces where c is the upper cell: the +k / +j / +i face of the cell below / behind / left
        // (those cells have the same flags where it matters)
        const int fk = k > 0 ? box_owner_start(v, c - nxy, i, j, k - 1) + hasI + hasJ : 0;
        const int fj = j > 0 ? box_owner_start(v, c - nx, i, j - 1, k) + hasI : 0;
        const int fi = i > 0 ? box_owner_start(v, c - 1, i - 1, j, k) : 0;
        // all loads first
        const double d = v.diag[c], xc = x[c];
        const double bk = k > 0 ? v.lowerCoef[fk] : 0.0, xk = k > 0 ? x[c - nxy] : 0.0;
        const double bj = j > 0 ? v.lowerCoef[fj] : 0.0, xj = j > 0 ? x[c - nx] : 0.0;
        const double bi = i > 0 ? v.lowerCoef[fi] : 0.0, xi = i > 0 ? x[c - 1] : 0.0;
        const double ai = hasI ? v.upperCoef[os] : 0.0, xI = hasI ? x[c + 1] : 0.0;
        const double aj = hasJ ? v.upperCoef[os + hasI] : 0.0, xJ = hasJ ? x[c + nx] : 0.0;
        const double ak = hasK ? v.upperCoef[os + hasI + hasJ] : 0.0, xK = hasK ? x[c + nxy] : 0.0;
        double acc = MODE == 0 ? __dmul_rn(d, xc) : __dsub_rn(b[c], __dmul_rn(d, xc));
        acc = row_term<MODE>(acc, bk, xk, k > 0);
        acc = row_term<MODE>(acc, bj, xj, j > 0);
        acc = row_term<MODE>(acc, bi, xi, i > 0);
        acc = row_term<MODE>(acc, ai, xI, hasI);
        acc = row_term<MODE>(acc, aj, xJ, hasJ);
        acc = row_term<MODE>(acc, ak, xK, hasK);
        y[c] = acc;
    }
}

// Amul on a box fused with <A x, x> (PCG.C:153-155) and the alpha epilogue: the products y[c]*x[c] are
// summed per thread in row order, then through the fixed-shape reduction of reduce.cuh
template <class Epi>
__global__ void __launch_bounds__(kBlock, 6) box_amul_dot_kernel(int n, BoxView v, double* __restrict__ y,
                                                                  const double* __restrict__ x, Epi epi, ReduceCtx rc)
{
    if (rc.S->done) return;
    const int nx = v.nx, ny = v.ny, nz = v.nz, nxy = nx * ny;
    double dot[1] = {0.0};
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock) {
        const int jk = (int)__umulhi((unsigned int)c, v.mulNx);
        const int i = c - jk * nx;
        const int k = (int)__umulhi((unsigned int)jk, v.mulNy);
        const int j = jk - k * ny;
        const bool hasI = i < nx - 1, hasJ = j < ny - 1, hasK = k < nz - 1;
        const int os = box_owner_start(v, c, i, j, k);
        const int fk = k > 0 ? box_owner_start(v, c - nxy, i, j, k - 1) + hasI + hasJ : 0;
        const int fj = j > 0 ? box_owner_start(v, c - nx, i, j - 1, k) + hasI : 0;
        const int fi = i > 0 ? box_owner_start(v, c - 1, i - 1, j, k) : 0;
        const double d = v.diag[c], xc = x[c];
        const double bk = k > 0 ? v.lowerCoef[fk] : 0.0, xk = k > 0 ? x[c - nxy] : 0.0;
        const double bj = j > 0 ? v.lowerCoef[fj] : 0.0, xj = j > 0 ? x[c - nx] : 0.0;
        const double bi = i > 0 ? v.lowerCoef[fi] : 0.0, xi = i > 0 ? x[c - 1] : 0.0;
        const double ai = hasI ? v.upperCoef[os] : 0.0, xI = hasI ? x[c + 1] : 0.0;
        const double aj = hasJ ? v.upperCoef[os + hasI] : 0.0, xJ = hasJ ? x[c + nx] : 0.0;
        const double ak = hasK ? v.upperCoef[os + hasI + hasJ] : 0.0, xK = hasK ? x[c + nxy] : 0.0;
        double acc = __dmul_rn(d, xc);
        acc = row_term<0>(acc, bk, xk, k > 0);
        acc = row_term<0>(acc, bj, xj, j > 0);
        acc = row_term<0>(acc, bi, xi, i > 0);
        acc = row_term<0>(acc, ai, xI, hasI);
        acc = row_term<0>(acc, aj, xJ, hasJ);
        acc = row_term<0>(acc, ak, xK, hasK);
        y[c] = acc;
        dot[0] = __dadd_rn(dot[0], __dmul_rn(acc, xc));
    }
    reduce_tail<1>(dot, rc, epi);
}

// exact c / d by multiply-high for every 0 <= c < limit?  (checked on the host, once per matrix)
static bool magic_div_ok(unsigned int d, unsigned int mul, unsigned int limit)
{
    // it suffices to test the multiples of d and their predecessors
    for (unsigned long long q = 0; q * d < limit; q++) {
        const unsigned long long lo = q * d, hi = std::min<unsigned long long>(lo + d - 1, limit - 1);
        if ((unsigned int)((lo * mul) >> 32) != q || (unsigned int)((hi * mul) >> 32) != q) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------
// TMA-staged row kernel.  Persistent CTAs walk blocks of kRowBlock consecutive rows.  The
// streams of a block that consecutive rows read at a stride - the upper coefficients and
// columns of the faces the rows own (one contiguous face range), the packed lower entries
// (one contiguous range) and the two row-pointer slices - are fetched by cp.async.bulk
// (TMA) into shared memory, double buffered, completion on an mbarrier per stage: the DRAM
// side sees long sequential bursts issued a whole block ahead instead of 8-byte loads at a
// 24-byte stride, and the per-row loops read shared memory.  diag / x / y are read and
// written directly (already coalesced); x[col] and the lower coefficients are gathers that
// hit L1/L2.  Same per-row operation order as row_kernel: bit-identical results.
// ---------------------------------------------------------------------------
struct StagedView {
    RowView v;
    const int4* blocks;   // {f0a, nFa, k0a, nKa} per row block
    int nBlocks, faceCap, lowerCap;
};

__device__ __forceinline__ void mbar_init(unsigned int bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity)
{
    unsigned int done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

constexpr int kStagedThreads = 256;

__host__ __device__ inline unsigned int staged_stage_bytes(int faceCap, int lowerCap)
{
    // coef[faceCap] f64 | u[faceCap] i32 | packed[lowerCap] i32 | ownerStart, losortStart [kRowBlock + 4] i32
    return (unsigned int)faceCap * 12u + (unsigned int)lowerCap * 4u + 2u * (kRowBlock + 4) * 4u;
}

template <int MODE>
__global__ void __launch_bounds__(kStagedThreads) row_staged_kernel(int n, StagedView sv, double* __restrict__ y,
                                                                     const double* __restrict__ x,
                                                                     const double* __restrict__ b,
                                                                     const SolverScalars* __restrict__ guard)
{
    extern __shared__ __align__(128) unsigned char stage_mem[];
    __shared__ unsigned long long full[2];
    if (guard && guard->done) return;
    const RowView& v = sv.v;
    const int tid = threadIdx.x;
    const unsigned int stageBytes = staged_stage_bytes(sv.faceCap, sv.lowerCap);
    const unsigned int smem0 = (unsigned int)__cvta_generic_to_shared(stage_mem);
    const unsigned int bar0 = (unsigned int)__cvta_generic_to_shared(&full[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned int offU = (unsigned int)sv.faceCap * 8u;
    const unsigned int offP = offU + (unsigned int)sv.faceCap * 4u;
    const unsigned int offOs = offP + (unsigned int)sv.lowerCap * 4u;
    const unsigned int offLs = offOs + (kRowBlock + 4) * 4u;

    auto issue = [&](int blk, int s, const int4& d) {   // thread 0
        const unsigned int dst = smem0 + (unsigned int)s * stageBytes, bar = bar0 + 8u * (unsigned int)s;
        const unsigned int rowBytes = (kRowBlock + 4) * 4u;
        mbar_expect_tx(bar, (unsigned int)d.y * 12u + (unsigned int)d.w * 4u + 2u * rowBytes);
        if (d.y) {
            bulk_g2s(dst, v.upperCoef + d.x, (unsigned int)d.y * 8u, bar);
            bulk_g2s(dst + offU, v.u + d.x, (unsigned int)d.y * 4u, bar);
        }
        if (d.w) bulk_g2s(dst + offP, v.lowerPacked + d.z, (unsigned int)d.w * 4u, bar);
        bulk_g2s(dst + offOs, v.ownerStart + (size_t)blk * kRowBlock, rowBytes, bar);
        bulk_g2s(dst + offLs, v.losortStart + (size_t)blk * kRowBlock, rowBytes, bar);
    };

    int blk = blockIdx.x;
    int4 dCur = make_int4(0, 0, 0, 0), dNext = make_int4(0, 0, 0, 0);
    if (blk < sv.nBlocks) dCur = sv.blocks[blk];
    if (tid == 0 && blk < sv.nBlocks) issue(blk, 0, dCur);
    if (blk + (int)gridDim.x < sv.nBlocks) dNext = sv.blocks[blk + gridDim.x];
    for (int it = 0; blk < sv.nBlocks; blk += gridDim.x, it++) {
        const int s = it & 1;
        const int nxt = blk + gridDim.x;
        if (tid == 0 && nxt < sv.nBlocks) issue(nxt, s ^ 1, dNext);   // stage s^1 was released by the barrier below
        const int4 d = dCur;
        dCur = dNext;
        if (nxt + (int)gridDim.x < sv.nBlocks) dNext = sv.blocks[nxt + gridDim.x];
        const int r0 = blk * kRowBlock;
        // the coalesced streams go straight to registers while the staged ones land
        double dg[kRowBlock / kStagedThreads], xc[kRowBlock / kStagedThreads], bc[kRowBlock / kStagedThreads];
#pragma unroll
        for (int q = 0; q < kRowBlock / kStagedThreads; q++) {
            const int c = r0 + tid + q * kStagedThreads;
            dg[q] = xc[q] = bc[q] = 0.0;
            if (c < n) {
                if (MODE <= 2) dg[q] = v.diag[c];
                if (MODE != 2 && MODE != 5) xc[q] = x[c];
                if (MODE == 1) bc[q] = b[c];
            }
        }
        mbar_wait(bar0 + 8u * (unsigned int)s, (unsigned int)(it >> 1) & 1u);
        const unsigned char* st = stage_mem + (size_t)s * stageBytes;
        const double* coefS = reinterpret_cast<const double*>(st);
        const int* uS = reinterpret_cast<const int*>(st + offU);
        const int* pkS = reinterpret_cast<const int*>(st + offP);
        const int* osS = reinterpret_cast<const int*>(st + offOs);
        const int* lsS = reinterpret_cast<const int*>(st + offLs);
#pragma unroll
        for (int q = 0; q < kRowBlock / kStagedThreads; q++) {
            const int lr = tid + q * kStagedThreads;
            const int c = r0 + lr;
            if (c < n) {
                double acc;
                if (MODE == 0) acc = __dmul_rn(dg[q], xc[q]);
                else if (MODE == 1) acc = __dsub_rn(bc[q], __dmul_rn(dg[q], xc[q]));
                else if (MODE == 2) acc = dg[q];
                else acc = 0.0;
                const int k0 = lsS[lr] - d.z, k1 = lsS[lr + 1] - d.z;
                const int f0 = osS[lr] - d.x, f1 = osS[lr + 1] - d.x;
                for (int k = k0; k < k1; k += kBatch) {
                    int col[kBatch], face[kBatch];
                    double a[kBatch], xv[kBatch];
#pragma unroll
                    for (int j = 0; j < kBatch; j++) {
                        const int w = pkS[min(k + j, k1 - 1)];
                        col[j] = w >> 5;
                        face[j] = w & 31;
                    }
#pragma unroll
                    for (int j = 0; j < kBatch; j++) {
                        face[j] += __ldg(v.ownerStart + col[j]);
                        xv[j] = (MODE != 2 && MODE != 5) ? __ldg(x + col[j]) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < kBatch; j++) a[j] = __ldg(v.lowerCoef + face[j]);
#pragma unroll
                    for (int j = 0; j < kBatch; j++) acc = row_term<MODE>(acc, a[j], xv[j], k + j < k1);
                }
                for (int f = f0; f < f1; f += kBatch) {
                    double xv[kBatch];
#pragma unroll
                    for (int j = 0; j < kBatch; j++) xv[j] = (MODE != 2 && MODE != 5) ? __ldg(x + uS[min(f + j, f1 - 1)]) : 0.0;
#pragma unroll
                    for (int j = 0; j < kBatch; j++)
                        acc = row_term<MODE>(acc, coefS[min(f + j, f1 - 1)], xv[j], f + j < f1);
                }
                y[c] = acc;
            }
        }
        __syncthreads();   // everybody is done with stage s: it may be refilled
    }
}

// Interface contribution, one thread per boundary cell, entries in reference
// order:  result[cell] -= (sign*coeff[k]) * psiNbr[k]
// (processorFvPatchScalarField.C:92-144; sign = -1 restates the negated
// coefficients residual() and Gauss-Seidel use, lduMatrixATmul.C:236-244).
// MODE 2 (sumA): result[cell] -= coeff[k]   (lduMatrixATmul.C:187-198).
template <int MODE>
__global__ void __launch_bounds__(kBlock) interface_kernel(int nBRows, const int* __restrict__ bRowCell,
                                                            const int* __restrict__ bRowStart,
                                                            const int* __restrict__ bEntry,
                                                            const double* __restrict__ coeff,
                                                            const double* __restrict__ recv, double sign,
                                                            double* __restrict__ result,
                                                            const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= nBRows) return;
    const int c = bRowCell[r];
    double acc = result[c];
    for (int e = bRowStart[r]; e < bRowStart[r + 1]; e++) {
        const int k = bEntry[e];
        if (MODE == 2) acc = __dsub_rn(acc, coeff[k]);
        else acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(sign, coeff[k]), recv[k]));
    }
    result[c] = acc;
}

// updateMatrixInterfaces in ONE kernel: every block waits for the neighbours' halos of the exchange this rank has
// just sent its own for (what halo_recv_kernel does), then the coupled rows read the neighbour values straight from
// the window slots (what halo_recv_kernel copied into the receive buffer for interface_kernel<0>): one launch and one
// pass over the halo less per Amul / residual / Gauss-Seidel boundary update.  Same terms in the same order.
constexpr int kFusedIfs = 32;
__global__ void __launch_bounds__(kBlock) interface_wait_kernel(CommDev c, const IfaceDev* __restrict__ ifs, int nIfs,
                                                                 int nBRows, const int* __restrict__ bRowCell,
                                                                 const int* __restrict__ bRowStart,
                                                                 const int* __restrict__ bEntry,
                                                                 const double* __restrict__ coeff, double sign,
                                                                 double* __restrict__ result, SolverScalars* S,
                                                                 bool guarded)
{
    if (guarded && S->done) return;
    WindowHeader* me = win_hdr(c, c.rank);
    __shared__ bool ok;
    __shared__ const double* slot[kFusedIfs];     // window slot of interface q, shifted by its offset
    __shared__ int first[kFusedIfs + 1];
    if (threadIdx.x == 0) {
        ok = true;
        for (int k = 0; k < nIfs && ok; k++) {
            const unsigned long long epoch = me->haloSent[ifs[k].nbrRank];
            ok = wait_epoch(&me->haloSeq[(int)(epoch & 1ull)][ifs[k].nbrRank], epoch, c.timeoutCycles);
        }
        if (!ok) {
            S->commError = 1;
            S->done = 1;
        }
    }
    if (threadIdx.x < nIfs) {
        const IfaceDev it = ifs[threadIdx.x];
        const int par = (int)(me->haloSent[it.nbrRank] & 1ull);
        slot[threadIdx.x] = win_halo(c, c.rank, par, threadIdx.x) - it.offset;
        first[threadIdx.x] = it.offset;
        if (threadIdx.x == nIfs - 1) first[nIfs] = it.offset + it.n;
    }
    __syncthreads();
    if (!ok) return;
    const int r = blockIdx.x * kBlock + threadIdx.x;
    if (r >= nBRows) return;
    const int cell = bRowCell[r];
    double acc = result[cell];
    for (int e = bRowStart[r]; e < bRowStart[r + 1]; e++) {
        const int k = bEntry[e];
        int q = 0;
        while (q + 1 < nIfs && k >= first[q + 1]) q++;
        acc = __dsub_rn(acc, __dmul_rn(__dmul_rn(sign, coeff[k]), ld_volatile_f64(slot[q] + k)));
    }
    result[cell] = acc;
}

template <int MODE>
static int launch_rows(ldu_matrix* m, const RowView& v, double* y, const double* x, const double* b,
                       bool guarded)
{
    const int n = m->nCells;
    if (n <= 0) return LDU_OK;
    // one row per thread, grid sized in whole waves of the SM count
    long long blocks = ((long long)n + kBlock - 1) / kBlock;
    const long long cap = (long long)m->ctx->smCount * 16;
    if (blocks > cap) blocks = cap;
    // read per launch: bench.py times both row kernels in one process
    const char* boxEnv = getenv("LDU_AMUL_BOX");
    const bool boxOff = boxEnv && boxEnv[0] == '0';
    if ((MODE == 0 || MODE == 1) && m->box[0] > 0 && !boxOff && m->boxDivOk >= 0) {
        if (m->boxDivOk == 0) {   // first use: are the multiply-high divisions exact for this box?
            const unsigned int mulNx = (unsigned int)((0x100000000ull + m->box[0] - 1) / m->box[0]);
            const unsigned int mulNy = (unsigned int)((0x100000000ull + m->box[1] - 1) / m->box[1]);
            const bool ok = m->box[0] > 1 && m->box[1] > 1 && magic_div_ok(m->box[0], mulNx, (unsigned int)n)
                            && magic_div_ok(m->box[1], mulNy, (unsigned int)(n / m->box[0]) + 1u);
            m->boxDivOk = ok ? 1 : -1;
        }
        if (m->boxDivOk == 1) {
            BoxView bv;
            bv.nx = m->box[0];
            bv.ny = m->box[1];
            bv.nz = m->box[2];
            bv.mulNx = (unsigned int)((0x100000000ull + bv.nx - 1) / bv.nx);
            bv.mulNy = (unsigned int)((0x100000000ull + bv.ny - 1) / bv.ny);
            bv.diag = v.diag;
            bv.lowerCoef = v.lowerCoef;
            bv.upperCoef = v.upperCoef;
            constexpr int BM = (MODE == 0 || MODE == 1) ? MODE : 0;
            static const int minb = getenv("LDU_AMUL_BOX_MINB") ? atoi(getenv("LDU_AMUL_BOX_MINB")) : 6;
            const SolverScalars* g = guarded ? m->d_scalars : nullptr;
            if (minb >= 8) box_row_kernel<BM, 8><<<(int)blocks, kBlock, 0, m->ctx->stream>>>(n, bv, y, x, b, g);
            else if (minb >= 6) box_row_kernel<BM, 6><<<(int)blocks, kBlock, 0, m->ctx->stream>>>(n, bv, y, x, b, g);
            else box_row_kernel<BM, 4><<<(int)blocks, kBlock, 0, m->ctx->stream>>>(n, bv, y, x, b, g);
            count_launch();
            LDU_CUDA(cudaGetLastError());
            return LDU_OK;
        }
    }
    static const bool packedOff = getenv("LDU_AMUL_PACKED") && getenv("LDU_AMUL_PACKED")[0] == '0';
    // measured on B200 (216^3): the staged kernel 199 us, the batched row kernel 161 us -> opt-in
    static const bool stagedOff = !(getenv("LDU_AMUL_STAGED") && getenv("LDU_AMUL_STAGED")[0] == '1');
    if (m->d_rowBlocks && !stagedOff && !packedOff && n >= 8 * kRowBlock) {
        const unsigned int smem = 2u * staged_stage_bytes(m->rowFaceCap, m->rowLowerCap);
        if (smem <= 200u * 1024u) {
            static bool attr[6] = {false, false, false, false, false, false};
            if (!attr[MODE]) {
                LDU_CUDA(cudaFuncSetAttribute(row_staged_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              200 * 1024));
                attr[MODE] = true;
            }
            int perSm = 0;
            LDU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, row_staged_kernel<MODE>, kStagedThreads, smem));
            if (perSm >= 1) {
                StagedView sv;
                sv.v = v;
                sv.blocks = reinterpret_cast<const int4*>(m->d_rowBlocks);
                sv.nBlocks = m->nRowBlocks;
                sv.faceCap = m->rowFaceCap;
                sv.lowerCap = m->rowLowerCap;
                const int grid = std::min(m->nRowBlocks, m->ctx->smCount * perSm);
                row_staged_kernel<MODE><<<grid, kStagedThreads, smem, m->ctx->stream>>>(n, sv, y, x, b,
                                                                                       guarded ? m->d_scalars : nullptr);
                count_launch();
                LDU_CUDA(cudaGetLastError());
                return LDU_OK;
            }
        }
    }
    static const int minb = getenv("LDU_AMUL_MINB") ? atoi(getenv("LDU_AMUL_MINB")) : 8;
    const SolverScalars* g = guarded ? m->d_scalars : nullptr;
    cudaStream_t st = m->ctx->stream;
    if (v.lowerPacked && !packedOff) {
        if (minb >= 8) row_kernel<MODE, true, 8><<<(int)blocks, kBlock, 0, st>>>(n, v, y, x, b, g);
        else row_kernel<MODE, true, 1><<<(int)blocks, kBlock, 0, st>>>(n, v, y, x, b, g);
    } else {
        if (minb >= 8) row_kernel<MODE, false, 8><<<(int)blocks, kBlock, 0, st>>>(n, v, y, x, b, g);
        else row_kernel<MODE, false, 1><<<(int)blocks, kBlock, 0, st>>>(n, v, y, x, b, g);
    }
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

// updateMatrixInterfaces: wait for the halos started by comm_halo_put, then add
// the coupled contribution to the boundary rows
static int k_interfaces_finish(ldu_matrix* m, double* result, int whichCoeffs, double sign, bool guarded)
{
    if (!m->nIfFaces) return LDU_OK;
    const double* coeff = whichCoeffs ? m->d_int : m->d_bou;
    const int grid = (m->nBRows + kBlock - 1) / kBlock;
    static const bool fusedOff = getenv("LDU_IF_FUSED") && getenv("LDU_IF_FUSED")[0] == '0';
    if (!fusedOff && (int)m->ifs.size() <= kFusedIfs) {
        const IfaceDev* tab = nullptr;
        LDU_TRY(comm_halo_table(m, &tab));
        interface_wait_kernel<<<grid, kBlock, 0, m->ctx->stream>>>(comm_dev(m->ctx), tab, (int)m->ifs.size(), m->nBRows,
                                                                  m->d_bRowCell, m->d_bRowStart, m->d_bEntry, coeff, sign,
                                                                  result, m->d_scalars, guarded);
        count_launch();
        LDU_CUDA(cudaGetLastError());
        return LDU_OK;
    }
    LDU_TRY(comm_halo_recv(m, guarded));
    interface_kernel<0><<<grid, kBlock, 0, m->ctx->stream>>>(m->nBRows, m->d_bRowCell, m->d_bRowStart,
                                                            m->d_bEntry, coeff, m->d_recv, sign, result,
                                                            guarded ? m->d_scalars : nullptr);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

int k_interfaces(ldu_matrix* m, double* result, const double* psi, int whichCoeffs, double sign,
                 bool guarded)
{
    LDU_TRY(comm_halo_put(m, psi, guarded));
    return k_interfaces_finish(m, result, whichCoeffs, sign, guarded);
}

// halo puts first, interior rows while the halos fly over NVLink, coupled rows last
// (the reference's initMatrixInterfaces / face loop / updateMatrixInterfaces order)
int k_amul(ldu_matrix* m, double* Apsi, const double* psi, bool transpose, bool guarded)
{
    LDU_TRY(comm_halo_put(m, psi, guarded));
    LDU_TRY(launch_rows<0>(m, row_view(m, transpose), Apsi, psi, nullptr, guarded));
    // Amul uses interfaceBouCoeffs, Tmul interfaceIntCoeffs (lduMatrixATmul.C:57-64,118-125)
    return k_interfaces_finish(m, Apsi, transpose ? 1 : 0, 1.0, guarded);
}

int k_residual(ldu_matrix* m, double* rA, const double* psi, const double* source, bool guarded)
{
    LDU_TRY(comm_halo_put(m, psi, guarded));
    LDU_TRY(launch_rows<1>(m, row_view(m, false), rA, psi, source, guarded));
    return k_interfaces_finish(m, rA, 0, -1.0, guarded);
}

int k_offdiag(ldu_matrix* m, double* y, const double* x)
{
    LDU_TRY(comm_halo_put(m, x, false));
    LDU_TRY(launch_rows<3>(m, row_view(m, false), y, x, nullptr, false));
    return k_interfaces_finish(m, y, 0, 1.0, false);
}

// lduMatrix::H / H1: internal faces only (fvMatrix::H adds the boundary part itself)
int k_H(ldu_matrix* m, double* Hpsi, const double* psi)
{
    return launch_rows<4>(m, row_view(m, false), Hpsi, psi, nullptr, false);
}

int k_H1(ldu_matrix* m, double* H1)
{
    return launch_rows<5>(m, row_view(m, false), H1, nullptr, nullptr, false);
}

// lduMatrix::faceH (lduMatrixTemplates.C:79-113): upper*psi[u] - lower*psi[l] per face
__global__ void __launch_bounds__(kBlock) faceH_kernel(int nFaces, const int* __restrict__ l, const int* __restrict__ u,
                                                        const double* __restrict__ lower,
                                                        const double* __restrict__ upper,
                                                        const double* __restrict__ psi, double* __restrict__ out)
{
    for (int f = blockIdx.x * kBlock + threadIdx.x; f < nFaces; f += gridDim.x * kBlock)
        out[f] = __dsub_rn(__dmul_rn(upper[f], psi[u[f]]), __dmul_rn(lower[f], psi[l[f]]));
}

int k_faceH(ldu_matrix* m, double* faceHpsi, const double* psi)
{
    if (m->nFaces <= 0) return LDU_OK;
    faceH_kernel<<<grid_for(m->ctx, m->nFaces), kBlock, 0, m->ctx->stream>>>(m->nFaces, m->d_l, m->d_u, m->d_lower,
                                                                            m->d_upper, psi, faceHpsi);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

static bool box_view(ldu_matrix* m, const RowView& v, BoxView& bv)
{
    const char* boxEnv = getenv("LDU_AMUL_BOX");
    if (m->box[0] <= 0 || (boxEnv && boxEnv[0] == '0') || m->boxDivOk < 0) return false;
    const int n = m->nCells;
    const unsigned int mulNx = (unsigned int)((0x100000000ull + m->box[0] - 1) / m->box[0]);
    const unsigned int mulNy = (unsigned int)((0x100000000ull + m->box[1] - 1) / m->box[1]);
    if (m->boxDivOk == 0) {   // first use: are the multiply-high divisions exact for this box?
        const bool ok = m->box[0] > 1 && m->box[1] > 1 && magic_div_ok(m->box[0], mulNx, (unsigned int)n)
                        && magic_div_ok(m->box[1], mulNy, (unsigned int)(n / m->box[0]) + 1u);
        m->boxDivOk = ok ? 1 : -1;
    }
    if (m->boxDivOk != 1) return false;
    bv.nx = m->box[0];
    bv.ny = m->box[1];
    bv.nz = m->box[2];
    bv.mulNx = mulNx;
    bv.mulNy = mulNy;
    bv.diag = v.diag;
    bv.lowerCoef = v.lowerCoef;
    bv.upperCoef = v.upperCoef;
    return true;
}

bool k_amul_dot_available(ldu_matrix* m)
{
    BoxView bv;
    return m->nIfFaces == 0 && m->ctx->comm.nRanks == 1 && m->nCells > 0 && box_view(m, row_view(m, false), bv);
}

int k_amul_dot(ldu_matrix* m, double* Apsi, const double* psi)
{
    BoxView bv;
    if (!box_view(m, row_view(m, false), bv)) return LDU_EINVAL;
    box_amul_dot_kernel<<<grid_for(m->ctx, m->nCells), kBlock, 0, m->ctx->stream>>>(m->nCells, bv, Apsi, psi, EpiWApA(),
                                                                               make_rc(m));
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

int k_sumA(ldu_matrix* m, double* sumA)
{
    LDU_TRY(launch_rows<2>(m, row_view(m, false), sumA, nullptr, nullptr, false));
    if (m->nIfFaces) {
        const int grid = (m->nBRows + kBlock - 1) / kBlock;
        interface_kernel<2><<<grid, kBlock, 0, m->ctx->stream>>>(m->nBRows, m->d_bRowCell, m->d_bRowStart,
                                                                m->d_bEntry, m->d_bou, nullptr, 1.0, sumA, nullptr);
        count_launch();
        LDU_CUDA(cudaGetLastError());
    }
    return LDU_OK;
}

}  // namespace ldu
