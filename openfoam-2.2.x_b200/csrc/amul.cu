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
#include <cstdlib>

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
template <int MODE, bool PACKED>
__device__ __forceinline__ double row_apply(const RowView& v, int c, const double* __restrict__ x,
                                            const double* __restrict__ b)
{
    double acc;
    if (MODE == 0) acc = __dmul_rn(v.diag[c], x[c]);
    else if (MODE == 1) acc = __dsub_rn(b[c], __dmul_rn(v.diag[c], x[c]));
    else if (MODE == 2) acc = v.diag[c];
    else acc = 0.0;
    const int k0 = v.losortStart[c], k1 = v.losortStart[c + 1];
    for (int k = k0; k < k1; k++) {
        int col, face;
        if (PACKED) {   // 4 bytes per lower entry from DRAM; ownerStart[col] is an L2 hit
            const int w = v.lowerPacked[k];
            col = w >> 5;
            face = v.ownerStart[col] + (w & 31);
        } else {
            col = v.lowerCol[k];
            face = v.losort[k];
        }
        const double a = v.lowerCoef[face];
        if (MODE == 0 || MODE == 3) acc = __dadd_rn(acc, __dmul_rn(a, x[col]));
        else if (MODE == 1) acc = __dsub_rn(acc, __dmul_rn(a, x[col]));
        else acc = __dadd_rn(acc, a);
    }
    const int f0 = v.ownerStart[c], f1 = v.ownerStart[c + 1];
    for (int f = f0; f < f1; f++) {
        const double a = v.upperCoef[f];
        if (MODE == 0 || MODE == 3) acc = __dadd_rn(acc, __dmul_rn(a, x[v.u[f]]));
        else if (MODE == 1) acc = __dsub_rn(acc, __dmul_rn(a, x[v.u[f]]));
        else acc = __dadd_rn(acc, a);
    }
    return acc;
}

template <int MODE, bool PACKED>
__global__ void __launch_bounds__(kBlock) row_kernel(int n, RowView v, double* __restrict__ y,
                                                      const double* __restrict__ x,
                                                      const double* __restrict__ b,
                                                      const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock)
        y[c] = row_apply<MODE, PACKED>(v, c, x, b);
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
    static const bool packedOff = getenv("LDU_AMUL_PACKED") && getenv("LDU_AMUL_PACKED")[0] == '0';
    if (v.lowerPacked && !packedOff)
        row_kernel<MODE, true><<<(int)blocks, kBlock, 0, m->ctx->stream>>>(n, v, y, x, b, guarded ? m->d_scalars : nullptr);
    else
        row_kernel<MODE, false><<<(int)blocks, kBlock, 0, m->ctx->stream>>>(n, v, y, x, b, guarded ? m->d_scalars : nullptr);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

// updateMatrixInterfaces: wait for the halos started by comm_halo_put, then add
// the coupled contribution to the boundary rows
static int k_interfaces_finish(ldu_matrix* m, double* result, int whichCoeffs, double sign, bool guarded)
{
    if (!m->nIfFaces) return LDU_OK;
    LDU_TRY(comm_halo_recv(m, guarded));
    const double* coeff = whichCoeffs ? m->d_int : m->d_bou;
    const int grid = (m->nBRows + kBlock - 1) / kBlock;
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
