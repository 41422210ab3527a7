// Device side of the multi-GPU exchange: peers write straight into each
// other's exchange window over NVLink (CUDA IPC mapped), no host involvement.
//
// Replaces, for the solver hot path, the reference's MPI calls:
//   reduce(scalar, sumOp) / reduce(vector2D, sumOp)   src/Pstream/mpi/UPstream.C:174-204
//   processor-patch halo Irecv/Isend + wait            src/Pstream/mpi/UIPread.C:280-300,
//                                                      UOPwrite.C:97-110, UPstream.C:257-341
//
// Window layout (identical on every rank):
//   WindowHeader | halo[2][maxInterfaces][slotStride] doubles
// Every message carries a monotonically increasing epoch; two parity halves make
// a slot reusable as soon as the NEXT exchange has completed.
#pragma once

#include "ldu_internal.h"

namespace ldu {

struct WindowHeader {
    unsigned long long redEpoch;                    // local counters (owner writes)
    unsigned long long haloEpoch;
    unsigned long long redSeq[2][kMaxRanks];        // written by peers
    double redVal[2][kMaxRanks][kRedSlots];
    unsigned long long haloSeq[2][kMaxRanks];       // written by peers: "my halos of epoch e landed"
    unsigned int haloTicket;
    unsigned int pad;
};

struct CommDev {
    int rank;
    int nRanks;
    unsigned char* const* peer;   // device array [nRanks] of mapped window bases
    int maxInterfaces;
    long long slotStride;         // doubles per interface slot
    long long timeoutCycles;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double ld_volatile_f64(const double* p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ WindowHeader* win_hdr(const CommDev& c, int r)
{
    return reinterpret_cast<WindowHeader*>(c.peer[r]);
}

__device__ __forceinline__ double* win_halo(const CommDev& c, int r, int parity, int iface)
{
    double* base = reinterpret_cast<double*>(c.peer[r] + sizeof(WindowHeader));
    return base + ((long long)parity * c.maxInterfaces + iface) * c.slotStride;
}

// spin until *flag >= epoch; false on timeout
__device__ __forceinline__ bool wait_epoch(const unsigned long long* flag, unsigned long long epoch,
                                           long long timeoutCycles)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch) {
        if (clock64() - t0 > timeoutCycles) return false;
        __nanosleep(20);
    }
    return true;
}

// All-reduce (sum) of NRED doubles across ranks, called by ONE thread per rank
// (the last block of a reducing kernel).  Every rank sums the contributions in
// rank order, so all ranks obtain the bit-identical result.
template <int NRED>
__device__ __forceinline__ void comm_allreduce_dev(const CommDev& c, double (&v)[NRED], SolverScalars* S)
{
    WindowHeader* me = win_hdr(c, c.rank);
    const unsigned long long epoch = me->redEpoch + 1;
    me->redEpoch = epoch;
    const int par = (int)(epoch & 1ull);
    for (int r = 0; r < c.nRanks; r++) {
        WindowHeader* w = win_hdr(c, r);
#pragma unroll
        for (int k = 0; k < NRED; k++) w->redVal[par][c.rank][k] = v[k];
        __threadfence_system();
        st_release_sys(&w->redSeq[par][c.rank], epoch);
    }
    double tot[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) tot[k] = 0.0;
    for (int r = 0; r < c.nRanks; r++) {
        if (!wait_epoch(&me->redSeq[par][r], epoch, c.timeoutCycles)) {
            if (S) { S->commError = 1; S->done = 1; }
            return;
        }
#pragma unroll
        for (int k = 0; k < NRED; k++) {
            const double x = ld_volatile_f64(&me->redVal[par][r][k]);
            tot[k] = (r == 0) ? x : __dadd_rn(tot[k], x);
        }
    }
#pragma unroll
    for (int k = 0; k < NRED; k++) v[k] = tot[k];
}

// The same all-reduce done by the 32 lanes of ONE converged warp (the first warp of the last block of a reducing
// kernel): lane r talks to rank r -- the nRanks stores + system fences + releases, and then the nRanks waits, run
// side by side instead of one after the other (at 8 ranks the serial version cost ~3x the NVLink round trip per
// reduction, three reductions per PCG iteration).  v is taken from lane 0; the sum is formed in rank order by
// every lane alike, so the result is bit-identical to comm_allreduce_dev's and identical on all ranks.
template <int NRED>
__device__ __forceinline__ void comm_allreduce_warp(const CommDev& c, double (&v)[NRED], SolverScalars* S)
{
    const int lane = threadIdx.x & 31;
    WindowHeader* me = win_hdr(c, c.rank);
    unsigned long long epoch = 0;
    if (lane == 0) {
        epoch = me->redEpoch + 1;
        me->redEpoch = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    const int par = (int)(epoch & 1ull);
    double mine[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) mine[k] = __shfl_sync(0xffffffffu, v[k], 0);
    bool ok = true;
    double x[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) x[k] = 0.0;
    if (lane < c.nRanks) {
        WindowHeader* w = win_hdr(c, lane);
#pragma unroll
        for (int k = 0; k < NRED; k++) w->redVal[par][c.rank][k] = mine[k];
        __threadfence_system();
        st_release_sys(&w->redSeq[par][c.rank], epoch);
        ok = wait_epoch(&me->redSeq[par][lane], epoch, c.timeoutCycles);
        if (ok) {
#pragma unroll
            for (int k = 0; k < NRED; k++) x[k] = ld_volatile_f64(&me->redVal[par][lane][k]);
        }
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (S && lane == 0) { S->commError = 1; S->done = 1; }
        return;
    }
    double tot[NRED];
#pragma unroll
    for (int k = 0; k < NRED; k++) tot[k] = __shfl_sync(0xffffffffu, x[k], 0);
    for (int r = 1; r < c.nRanks; r++) {
#pragma unroll
        for (int k = 0; k < NRED; k++) tot[k] = __dadd_rn(tot[k], __shfl_sync(0xffffffffu, x[k], r));
    }
#pragma unroll
    for (int k = 0; k < NRED; k++) v[k] = tot[k];
}

CommDev comm_dev(const ldu_context* ctx);

}  // namespace ldu
