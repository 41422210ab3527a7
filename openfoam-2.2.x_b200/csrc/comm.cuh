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
// Every message carries a monotonically increasing epoch (halos: one counter per pair of ranks, reductions: one
// per rank, all ranks take part in every reduction); two parity halves make a slot reusable as soon as the NEXT
// exchange has completed.
#pragma once

#include "ldu_internal.h"

namespace ldu {

// A double and its message tag in one 16-byte word, each 8-byte half carrying the tag (NCCL-LL style): the value
// IS the flag, so a message costs one one-way NVLink latency -- no system fence, no separate release store.  Used by
// the coupled coarsest-level kernel (gamg.cu), whose iterations are nothing but such messages.
struct alignas(16) LLMsg {
    unsigned int lo, t0, hi, t1;
};
constexpr int kLLIfs = 16;      // interfaces per rank served by the LL halo words
constexpr int kLLFaces = 32;    // faces per interface

struct WindowHeader {
    unsigned long long redEpoch;                    // local counters (owner writes)
    unsigned long long haloSent[kMaxRanks];         // halo exchanges done with rank r, counted per PAIR of ranks: a rank
                                                    // whose matrix has no face coupled to r does not take part and
                                                    // does not fall out of step (both ends of a pair count alike)
    unsigned long long redSeq[2][kMaxRanks];        // written by peers
    double redVal[2][kMaxRanks][kRedSlots];
    unsigned long long haloSeq[2][kMaxRanks];       // written by peers: "my halos of epoch e landed"
    unsigned int haloTicket;
    unsigned int pad;
    LLMsg redLL[2][kMaxRanks][2];                  // written by peers: all-reduce partials, tagged with the epoch
    LLMsg haloLL[2][kLLIfs][kLLFaces];             // written by peers: halo values of small interfaces, tagged
};

__device__ __forceinline__ void ll_store_sys(LLMsg* p, double v, unsigned int tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(tag),
                 "r"((unsigned int)(b >> 32)), "r"(tag)
                 : "memory");
}

// poll a word of this rank's own window until both halves carry `tag`; false on time-out
__device__ __forceinline__ bool ll_wait_sys(const LLMsg* p, unsigned int tag, long long timeoutCycles, double& v)
{
    const long long t0 = clock64();
    for (;;) {
        unsigned int lo, a, hi, b;
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(a), "=r"(hi), "=r"(b) : "l"(p) : "memory");
        if (a == tag && b == tag) {
            v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
            return true;
        }
        if (clock64() - t0 > timeoutCycles) return false;
    }
}

struct CommDev {
    int rank;
    int nRanks;
    unsigned char* const* peer;   // device array [nRanks] of mapped window bases
    int maxInterfaces;
    long long slotStride;         // doubles per interface slot
    long long timeoutCycles;
    int llRed;                    // all-reduces of one or two sums use the tagged words (LDU_RED_LL=0: the mailboxes)
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
// Returns false (on every lane) after a time-out.
template <int NRED>
__device__ __forceinline__ bool comm_allreduce_warp(const CommDev& c, double (&v)[NRED], SolverScalars* S)
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
        return false;
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
    return true;
}

// comm_allreduce_warp with tagged words (NRED <= 2): lane r sends the partials to rank r and polls rank r's words
// in its own window; same rank-order sum, same epoch counter (the two flavours may be mixed: a rank is never more
// than one reduction ahead of another, so the two parity halves of either buffer are enough).
template <int NRED>
__device__ __forceinline__ bool comm_allreduce_warp_ll(const CommDev& c, double (&v)[NRED], SolverScalars* S)
{
    static_assert(NRED <= 2, "redLL holds two words per rank");
    const int lane = threadIdx.x & 31;
    WindowHeader* me = win_hdr(c, c.rank);
    unsigned long long epoch = 0;
    if (lane == 0) {
        epoch = me->redEpoch + 1;
        me->redEpoch = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    const int par = (int)(epoch & 1ull);
    const unsigned int tag = (unsigned int)epoch;
    double x[NRED];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NRED; k++) x[k] = __shfl_sync(0xffffffffu, v[k], 0);
    if (lane < c.nRanks) {
        WindowHeader* w = win_hdr(c, lane);
#pragma unroll
        for (int k = 0; k < NRED; k++) ll_store_sys(&w->redLL[par][c.rank][k], x[k], tag);
#pragma unroll
        for (int k = 0; k < NRED; k++) ok = ok && ll_wait_sys(&me->redLL[par][lane][k], tag, c.timeoutCycles, x[k]);
    }
    if (!__all_sync(0xffffffffu, ok)) {
        if (S && lane == 0) { S->commError = 1; S->done = 1; }
        return false;
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
    return true;
}

CommDev comm_dev(const ldu_context* ctx);

// one coupled interface of a matrix: faces [offset, offset+n) of the concatenated interface arrays go to slot
// nbrInterface of rank nbrRank's window
struct IfaceDev {
    int offset, n, nbrRank, nbrInterface;
};
int comm_halo_table(ldu_matrix* m, const IfaceDev** tab);

}  // namespace ldu
