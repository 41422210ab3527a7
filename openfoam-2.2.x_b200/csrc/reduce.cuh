// Deterministic fused map + reduce + scalar epilogue.
//
// One launch does:   element-wise work over n cells (Map)  ->  NRED sums
//   -> block partials in HBM -> the LAST block to finish (ticket election)
//   sums the partials in a fixed order, all-reduces them across GPUs through
//   the peer-mapped exchange window (comm.cuh) and runs the scalar epilogue
//   (Epi) that the reference runs on the host between its loops (alpha, beta,
//   residual norm, convergence test: solvers/PCG/PCG.C:123-178).
// The summation tree depends only on (n, grid, block), never on timing, so
// every run and every rank count gives reproducible sums.
#pragma once

#include "comm.cuh"
#include "ldu_internal.h"

namespace ldu {

constexpr int kBlock = 256;

struct ReduceCtx {
    SolverScalars* S;
    double* partials;      // [gridDim.x * NRED]
    unsigned int* ticket;
    double* red;           // [kMaxRed] totals of this launch (also readable by the host)
    CommDev comm;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// fixed-shape block sum; result valid in thread 0
template <int NRED>
__device__ __forceinline__ void block_sum(double (&acc)[NRED], double (*smem)[kBlock / 32])
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NRED; k++) {
        double v = warp_sum(acc[k]);
        if (lane == 0) smem[k][wid] = v;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NRED; k++) {
            double v = (lane < kBlock / 32) ? smem[k][lane] : 0.0;
            v = warp_sum(v);
            acc[k] = v;
        }
    }
    __syncthreads();
}

// Tail of every reducing kernel: publish the block partials, elect the last
// block, finish the sums there, (multi-GPU) exchange them, run the epilogue.
template <int NRED, class Epi>
__device__ __forceinline__ void reduce_tail(double (&acc)[NRED], const ReduceCtx& rc, Epi epi)
{
    __shared__ double smem[NRED][kBlock / 32];
    __shared__ bool isLast;
    block_sum<NRED>(acc, smem);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NRED; k++) rc.partials[(size_t)blockIdx.x * NRED + k] = acc[k];
        __threadfence();
        unsigned int t = atomicAdd(rc.ticket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    double tot[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) tot[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kBlock) {
#pragma unroll
        for (int k = 0; k < NRED; k++)
            tot[k] = __dadd_rn(tot[k], __ldcg(&rc.partials[(size_t)b * NRED + k]));
    }
    block_sum<NRED>(tot, smem);
    if (threadIdx.x < 32) {      // the first warp, converged (block_sum ends with a barrier); totals valid in lane 0
        if (threadIdx.x == 0) *rc.ticket = 0u;  // re-arm for the next launch on this stream
        if (rc.comm.nRanks > 1) {
            // one or two sums travel as tagged words (one one-way NVLink latency), more through the mailboxes
            if constexpr (NRED <= 2) {
                if (rc.comm.llRed) comm_allreduce_warp_ll<NRED>(rc.comm, tot, rc.S);
                else comm_allreduce_warp<NRED>(rc.comm, tot, rc.S);
            } else {
                comm_allreduce_warp<NRED>(rc.comm, tot, rc.S);
            }
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < NRED; k++) rc.red[k] = tot[k];
            epi(rc.S, tot);
        }
    }
}

// Generic grid-stride map-reduce.  Map: void operator()(int i, double (&acc)[NRED]).
// guard == true: the launch is a no-op once the solve has finished on device.
template <int NRED, bool GUARD, class Map, class Epi>
__global__ void __launch_bounds__(kBlock) map_reduce_kernel(int n, Map map, Epi epi, ReduceCtx rc)
{
    if (GUARD && rc.S->done) return;
    double acc[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) acc[k] = 0.0;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) map(i, acc);
    reduce_tail<NRED>(acc, rc, epi);
}

// Reference-order variant ("referenceOrderSums"): ONE block; the element work is
// still done by 256 threads in parallel, but the NRED sums are accumulated by a
// single thread strictly left to right, i.e. in the order of the reference's
// sumProd / sumMag loops (fields/Fields/Field/FieldFunctions.C:363-385,422-434).
// Every reduced scalar, hence every iterate of every solver, is then
// BIT-IDENTICAL to the reference.  O(n) dependent adds: a verification mode for
// small and medium systems, not the production path.
template <int NRED, bool GUARD, class Map, class Epi>
__global__ void __launch_bounds__(kBlock) map_reduce_seq_kernel(int n, Map map, Epi epi, ReduceCtx rc)
{
    if (GUARD && rc.S->done) return;
    __shared__ double terms[NRED][kBlock];
    double tot[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) tot[k] = 0.0;
    for (int base = 0; base < n; base += kBlock) {
        const int i = base + threadIdx.x;
        double t[NRED];
#pragma unroll
        for (int k = 0; k < NRED; k++) t[k] = 0.0;
        if (i < n) map(i, t);   // 0 + term == term
#pragma unroll
        for (int k = 0; k < NRED; k++) terms[k][threadIdx.x] = t[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            const int m = min(kBlock, n - base);
            for (int j = 0; j < m; j++) {
#pragma unroll
                for (int k = 0; k < NRED; k++) tot[k] = __dadd_rn(tot[k], terms[k][j]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (rc.comm.nRanks > 1) comm_allreduce_dev<NRED>(rc.comm, tot, rc.S);
#pragma unroll
        for (int k = 0; k < NRED; k++) rc.red[k] = tot[k];
        epi(rc.S, tot);
    }
}

// Plain map (no reduction).
template <bool GUARD, class Map>
__global__ void __launch_bounds__(kBlock) map_kernel(int n, Map map, const SolverScalars* S)
{
    if (GUARD && S->done) return;
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) map(i);
}

struct NoEpi {
    __device__ void operator()(SolverScalars*, const double*) const {}
};

inline int grid_for(const ldu_context* ctx, int n)
{
    // persistent-style grid: a multiple of the SM count, capped by the work
    long long want = ((long long)n + kBlock - 1) / kBlock;
    long long cap = (long long)ctx->smCount * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

inline ReduceCtx make_rc(ldu_matrix* m)
{
    ReduceCtx rc;
    rc.S = m->d_scalars;
    rc.partials = m->ctx->d_partials;
    rc.ticket = m->ctx->d_ticket;
    rc.red = m->ctx->d_red;
    rc.comm = comm_dev(m->ctx);
    return rc;
}

template <int NRED, bool GUARD, class Map, class Epi>
inline int launch_map_reduce(ldu_matrix* m, int n, Map map, Epi epi)
{
    if (m->referenceOrderSums) {
        map_reduce_seq_kernel<NRED, GUARD, Map, Epi><<<1, kBlock, 0, m->ctx->stream>>>(n, map, epi, make_rc(m));
        count_launch();
        LDU_CUDA(cudaGetLastError());
        return LDU_OK;
    }
    const int grid = grid_for(m->ctx, n);
    map_reduce_kernel<NRED, GUARD, Map, Epi><<<grid, kBlock, 0, m->ctx->stream>>>(n, map, epi, make_rc(m));
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

template <bool GUARD, class Map>
inline int launch_map(ldu_matrix* m, int n, Map map)
{
    if (n <= 0) return LDU_OK;
    const int grid = grid_for(m->ctx, n);
    map_kernel<GUARD, Map><<<grid, kBlock, 0, m->ctx->stream>>>(n, map, m->d_scalars);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

}  // namespace ldu
