// Plane-stacked pipelined triangular sweeps for structured hex boxes (second generation).
//
// Same arithmetic and the same skewed tile layout as stencil.cu (one warp owns a
// tile of 32 x-lines of one k-plane, lane l walks line j0+l along i skewed by l
// steps, i-neighbour = own register, j-neighbour = one shuffle), but the three
// things that bounded the first version are removed:
//
//   * the k-neighbour hop.  W consecutive k-planes of the same 32-line column are
//     stacked in ONE CTA (W compute warps); plane k hands its results to plane k+1
//     through a shared-memory ring of {value, tag} words (a plane trails the one
//     below it by a shared-memory latency, ~10^2 cycles) instead of a word polled
//     in L2 (~10^3 cycles).  Only every W-th plane crosses CTAs.
//   * natural-layout traffic in the sweep.  The vector being substituted lives in
//     tile layout for both sweeps: every operand and the result is one coalesced
//     256-byte row per step.  pack/unpack kernels transpose 32x32 blocks through
//     shared memory between the natural cell order and the tile layout.
//   * memory latency on the compute warps.  A helper warp per CTA (a) polls the
//     {value, epoch} words published by other CTAs in global memory (the plane
//     below the stack, the last line of the previous column) and forwards them into
//     shared-memory rings, (b) issues cp.async.bulk.prefetch.L2 for the operand
//     rows ~48 steps ahead of every compute warp.  The compute warps only see
//     coalesced loads that hit L2, shared memory and registers.
//
// The coefficient operands are stored premultiplied, P = rD[c]*coef[f], the product the
// reference forms first (wA[u] -= rD[u]*upper[f]*wA[l], DICPreconditioner.C:108-121),
// so each term is one multiply and one subtract and results stay BIT-IDENTICAL to the
// reference and to the generic dataflow path.
//
// CTAs claim their (k-group, column) tile from an atomic ticket in dependency order, so
// a CTA only ever waits on CTAs that are already running: no co-residency requirement.
#include <algorithm>
#include <cstdlib>

#include "reduce.cuh"
#include "sweeps.h"

namespace ldu {

namespace {

constexpr int kRing = 16;          // ring depth (steps) of every shared-memory hand-off
constexpr int kD = 8;              // operand rows in flight per plane: register FIFO depth = unroll factor
constexpr int kPad2 = 16;          // zero rows in front of / behind the tile arrays
constexpr int kC = 4;              // steps per operand chunk (one TMA bulk copy per operand)
constexpr int kNS = 3;             // operand chunks in flight per plane
constexpr int kA = 2;              // ticks between a word entering shared memory and its use
constexpr int kF = 4;              // ticks a global word is requested ahead of being forwarded
constexpr int kPfChunk = 16;       // rows per L2 bulk prefetch (16 * 256 B = 4 KB)
constexpr int kPfAhead = 48;       // rows the prefetch runs ahead of a compute warp
constexpr long long kTimeout2 = 4000000000ll;

struct LLW {
    unsigned int lo, f0, hi, f1;
};

struct Box2 {
    int nx, ny, nz;
    int nJ;      // columns of 32 lines per plane
    int steps;   // nx + 31 skewed steps per tile
    int nTiles;  // nz * nJ
    int nKg;     // k-groups of W planes
};

__device__ __forceinline__ void g_store(LLW* p, double v, unsigned int tag)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(tag),
                 "r"((unsigned int)(b >> 32)), "r"(tag)
                 : "memory");
}

__device__ __forceinline__ void g_peek(const LLW* p, LLW& w)
{
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w.lo), "=r"(w.f0), "=r"(w.hi), "=r"(w.f1)
                 : "l"(p)
                 : "memory");
}

__device__ __forceinline__ bool ok(const LLW& w, unsigned int tag) { return w.f0 == tag && w.f1 == tag; }
__device__ __forceinline__ double val(const LLW& w) { return __hiloint2double((int)w.hi, (int)w.lo); }

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void l2_prefetch(const void* p, unsigned int bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct S2Args {
    SolverScalars* S;
    int guarded;
    Box2 b;
    unsigned int epoch;
    const double* pk;   // premultiplied coefficient of the k-, j-, i-neighbour (tile layout)
    const double* pj;
    const double* pi;
    double* Y;          // the vector being substituted, tile layout, in place
    LLW* gK;            // [nKg][nJ][steps][32]  last plane of a stack, for the next stack
    LLW* gJ;            // [nTiles][steps]       edge line of a tile, for the next column
    unsigned int* ticket;  // [0] claimed, [1] finished
    int prefetch;
    unsigned long long* trace;   // debug (LDU_S2_TRACE): per CTA {ticket, start, first plane done, last plane done, helper loops, helper done}
};

struct Shared2 {
    unsigned long long bar;  // the tick mbarrier
    volatile int abort;
    int ticket;
    int finished;
};

// warp-uniform abort / timeout test for the spin loops (called convergently)
__device__ __forceinline__ bool give_up(Shared2* sh, long long& tstart)
{
    bool bad = sh->abort != 0;
    if (tstart == 0) tstart = clock64();
    else if (clock64() - tstart > kTimeout2) bad = true;
    return __any_sync(0xffffffffu, bad);
}

__device__ __forceinline__ double lds64(unsigned int addr)
{
    double v;
    asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ void sts64(unsigned int addr, double v)
{
    asm volatile("st.volatile.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// The planes of a stack and the helper warp tick together through one mbarrier in shared
// memory (one arrive per warp and tick): arriving and waiting are separate operations, so
// a warp can arrive as soon as its result is handed off and do the rest of its step while
// the others catch up.
__device__ __forceinline__ void mbar_init(unsigned int bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned int bar)
{
    unsigned long long state;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(state) : "r"(bar) : "memory");
    (void)state;
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

__device__ __forceinline__ void cp_async8(unsigned int dst, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct StepOps2 {
    double pk, pj, pi, src;
};

// Shared memory of one CTA:
//   kx [W][2][32]     plane p's result of its step n, slot n&1: read by plane p+1 one tick later
//   hk [kRing][32]    helper -> first plane: results of the stack below, slot n % kRing
//   hj [W][kRing]     helper -> edge lanes: last line of the previous column, slot n % kRing
template <int W>
struct Smem2 {
    double kx[W][2][32];
    double hk[kRing][32];
    double hj[W][kRing];
    double ops[W][kNS][4][kC][32];       // operand ring of every plane: pk, pj, pi, src rows of kNS chunks
    unsigned long long full[W][kNS];     // mbarrier per ring slot: the chunk's bytes have landed
    Shared2 sh;
};

// One stack of up to W planes of one column.  All warps of the CTA tick together through
// barrier 1: in interval T (between barrier T and barrier T+1) plane p executes its step
// n = T - p, reading what plane p-1 wrote in interval T-1.  The helper warp is part of the
// barrier: the words of other CTAs that interval T reads were put into shared memory kA
// ticks earlier, from {value, epoch} words requested from global memory kF ticks before
// that, so in steady state nobody polls: every warp only waits in the hardware barrier.
//
// Critical path of a tick: barrier release -> LDS of the k-neighbour -> DMUL, 3 x DADD
// (the reference's operation order) -> STS -> barrier arrive.  Everything else (operand
// loads kD steps ahead into a rotating register FIFO, the j/i terms, stores to global
// memory, publication for other CTAs) is issued after the arrive, in the barrier's shadow.
template <int W, bool BWD>
__global__ void __launch_bounds__((W + 1) * 32) sweep2_kernel(S2Args a)
{
    extern __shared__ uint4 smem_raw[];
    Smem2<W>* sm = reinterpret_cast<Smem2<W>*>(smem_raw);
    Shared2* sh = &sm->sh;

    if (a.guarded && a.S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        sh->ticket = (int)atomicAdd(&a.ticket[0], 1u);
        sh->abort = 0;
        sh->finished = 0;
    }
    {
        double* q0 = &sm->hk[0][0];   // zero rows: what a plane without a k-neighbour reads
        for (int q = threadIdx.x; q < kRing * 32 + W * kRing; q += (W + 1) * 32) q0[q] = 0.0;
    }
    __syncthreads();

    const int nx = a.b.nx, ny = a.b.ny, nz = a.b.nz, nJ = a.b.nJ, steps = a.b.steps;
    const int nCta = a.b.nKg * nJ;
    const int tr = BWD ? nCta - 1 - sh->ticket : sh->ticket;
    const int kg = tr / nJ, J = tr - kg * nJ;
    const int k0 = kg * W;
    const int Wg = min(W, nz - k0);                               // planes in this stack
    const bool prevGroup = BWD ? (kg + 1 < a.b.nKg) : (kg > 0);   // a stack feeds this one
    const bool nextGroup = BWD ? (kg > 0) : (kg + 1 < a.b.nKg);   // this one feeds a stack
    const bool jIn = BWD ? (J + 1 < nJ) : (J > 0);                // a column feeds this one
    const bool jOut = BWD ? (J > 0) : (J + 1 < nJ);
    const unsigned int epoch = a.epoch;
    const long long tileStride = (long long)steps * 32;
    const int t0 = BWD ? steps - 1 : 0, dt = BWD ? -1 : 1;
    // The planes of a stack and the helper tick together through named barrier 1 (measured on
    // B200, tests/micro/sync_latency.cu: bar.sync + LDS + 4 dependent FP64 ops + STS = 105
    // cycles per tick for 8 warps, 142 for 16; an mbarrier costs 150-370).  Everything of a
    // step that does not feed the hand-off is issued after the barrier instruction.
    const int nThreads = (Wg + 1) * 32;
    auto tick = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(nThreads) : "memory"); };
    unsigned long long* trace = a.trace ? a.trace + 8ull * (unsigned int)sh->ticket : nullptr;
    if (trace && threadIdx.x == 0) {
        trace[0] = ((unsigned long long)kg << 32) | (unsigned int)J;
        trace[1] = gtimer();
    }

    if (warp < W) {
        // ------------------------------------------------------------------ compute warp
        const int p = warp;
        if (p < Wg) {
            const int k = BWD ? k0 + Wg - 1 - p : k0 + p;
            const int T = k * nJ + J;
            const int j = J * 32 + lane;
            const bool lineValid = j < ny;
            const bool hasJ = lineValid && (BWD ? (j < ny - 1) : (j > 0));
            const bool edgeIn = jIn && hasJ && (BWD ? (lane == 31) : (lane == 0));
            const bool edgeOut = lineValid && jOut && (BWD ? (lane == 0) : (lane == 31));
            const bool hasConsumer = p + 1 < Wg;
            const bool pubK = (p == Wg - 1) && nextGroup;
            const long long base = (long long)T * tileStride + lane;
            const double* pPk = a.pk + base;
            const double* pPj = a.pj + base;
            const double* pPi = a.pi + base;
            double* pY = a.Y + base;
            // k-neighbour values: the plane below in this stack (2 slots) or the helper's ring
            const unsigned int kIn = (unsigned int)__cvta_generic_to_shared(p > 0 ? &sm->kx[p - 1][0][lane] : &sm->hk[0][lane]);
            const unsigned int kMask = p > 0 ? 1u : (unsigned int)(kRing - 1);
            const unsigned int kOut = (unsigned int)__cvta_generic_to_shared(&sm->kx[p][0][lane]);
            const unsigned int jInA = (unsigned int)__cvta_generic_to_shared(&sm->hj[p][0]);
            LLW* gKself = a.gK + ((long long)(kg * nJ + J) * steps) * 32 + lane;
            LLW* gJself = a.gJ + (long long)T * steps;
            const unsigned int nxEff = lineValid ? (unsigned int)nx : 0u;   // cell i = t - lane exists iff i < nxEff
            const int iFirst = BWD ? nx - 1 : 0;                            // cell without an i-neighbour

            // Operand stream: TMA.  The rows of a chunk of kC steps are contiguous in the tile layout
            // (kC * 256 bytes per operand), so one lane fetches them with four cp.async.bulk copies
            // into a ring of kNS chunks in shared memory, completion on an mbarrier per ring slot.
            // (Register-destined loads all hang on the warp's few counting scoreboards - waiting for
            // the oldest row waits for the youngest too; per-lane cp.async costs the LSU 8 cycles per
            // instruction and warp, measured as 35 cycles per warp and tick.)
            const unsigned int opsA = (unsigned int)__cvta_generic_to_shared(&sm->ops[p][0][0][0][0]);
            const unsigned int fullA = (unsigned int)__cvta_generic_to_shared(&sm->full[p][0]);
            if (lane == 0)
                for (int q = 0; q < kNS; q++) mbar_init(fullA + 8u * q, 1);
            __syncwarp();
            const int nChunks = (steps + kC - 1) / kC;
            auto issue = [&](int c, unsigned int slot) {   // lane 0: rows of steps kC*c .. kC*c + kC-1
                const int n0 = c * kC;
                const int rowLo = BWD ? t0 - n0 - (kC - 1) : n0;      // lowest row of the chunk
                const long long e = (long long)rowLo * 32 - lane;      // p* pointers carry + lane
                const unsigned int d = opsA + slot * (4u * kC * 256u);
                const unsigned int bar = fullA + 8u * slot;
                mbar_expect_tx(bar, 4u * kC * 256u);
                bulk_g2s(d, pPk + e, kC * 256u, bar);
                bulk_g2s(d + kC * 256u, pPj + e, kC * 256u, bar);
                bulk_g2s(d + 2u * kC * 256u, pPi + e, kC * 256u, bar);
                bulk_g2s(d + 3u * kC * 256u, pY + e, kC * 256u, bar);
            };
            auto fetch = [&](unsigned int slot, int q, StepOps2& o) {   // step q of the chunk in `slot`
                const unsigned int r = (unsigned int)(BWD ? kC - 1 - q : q);
                const unsigned int d = opsA + slot * (4u * kC * 256u) + r * 256u + (unsigned int)lane * 8u;
                o.pk = lds64(d);
                o.pj = lds64(d + kC * 256u);
                o.pi = lds64(d + 2u * kC * 256u);
                o.src = lds64(d + 3u * kC * 256u);
            };
            if (lane == 0)
                for (int c = 0; c < kNS && c < nChunks; c++) issue(c, (unsigned int)c);
            for (int q = 0; q <= p; q++) tick();   // plane p starts in interval p

            StepOps2 o;   // operands of the coming step
            mbar_wait(fullA, 0u);
            fetch(0u, 0, o);
            double prev = 0.0;
            // j- and i-terms of the coming step: known before the barrier opens.  Step 0 only has
            // the edge lane's cell (i == iFirst): no i-term, a j-term if a column feeds this one.
            double ti = 0.0;
            double tj = edgeIn ? __dmul_rn(o.pj, lds64(jInA)) : 0.0;
            unsigned int slot = 0, useCount = 0;   // ring slot of the current chunk, times the ring wrapped
            for (int c = 0; c < nChunks; c++) {
                const int tb = t0 + dt * c * kC;
                const unsigned int nextSlot = slot + 1 == kNS ? 0u : slot + 1;
                const unsigned int nextUse = slot + 1 == kNS ? useCount + 1 : useCount;
#pragma unroll
                for (int q = 0; q < kC; q++) {
                    const unsigned int n = (unsigned int)(c * kC + q);
                    if ((int)n < steps) {
                        const int t = tb + dt * q;
                        const int i = t - lane;
                        const bool active = (unsigned int)i < nxEff;
                        // ---- critical path: k-neighbour -> result -> hand-off -> barrier
                        const double vk = lds64(kIn + (n & kMask) * 256u);
                        // without a k-neighbour pk == +0 and vk == +0: src stays as it is
                        double acc = __dsub_rn(o.src, __dmul_rn(o.pk, vk));
                        acc = __dsub_rn(acc, tj);    // tj, ti == +0 where the neighbour does not exist
                        acc = __dsub_rn(acc, ti);
                        if (hasConsumer) sts64(kOut + (n & 1u) * 256u, acc);
                        if (trace && lane == 0 && sh->ticket == a.prefetch && n + p + 1 >= 110u && n + p + 1 < 138u)
                            a.trace[8ull * nCta + (unsigned long long)p * 28ull + (n + p + 1 - 110u)] = (unsigned long long)clock64();
                        tick();
                        // ---- behind the barrier instruction: nothing here feeds another warp this tick
                        if (active) {
                            pY[t * 32] = acc;
                            prev = acc;
                        }
                        if (pubK) g_store(gKself + t * 32, acc, epoch);
                        if (edgeOut && active) g_store(gJself + t, acc, epoch);
                        if (q + 1 < kC) {
                            fetch(slot, q + 1, o);
                        } else {
                            // every lane has its operands of this chunk in registers: refill the slot,
                            // then move on to the next chunk (its bytes landed long ago)
                            __syncwarp();
                            if (lane == 0 && c + kNS < nChunks) issue(c + kNS, slot);
                            if (c + 1 < nChunks) {
                                mbar_wait(fullA + 8u * nextSlot, nextUse & 1u);
                                fetch(nextSlot, 0, o);
                            }
                        }
                        {   // terms of step n + 1
                            double vj = BWD ? __shfl_down_sync(0xffffffffu, prev, 1) : __shfl_up_sync(0xffffffffu, prev, 1);
                            const double vje = lds64(jInA + ((n + 1u) & (unsigned int)(kRing - 1)) * 8u);
                            if (edgeIn) vj = vje;
                            tj = hasJ ? __dmul_rn(o.pj, vj) : 0.0;
                            ti = (i + dt != iFirst) ? __dmul_rn(o.pi, prev) : 0.0;
                        }
                    }
                }
                slot = nextSlot;
                useCount = nextUse;
            }
            for (int q = p + 1; q < Wg; q++) tick();
            if (trace && lane == 0 && p == 0) trace[2] = gtimer();
            if (trace && lane == 0 && p == Wg - 1) trace[3] = gtimer();
        }
    } else {
        // ------------------------------------------------------------------ helper warp
        // k-row r (needed by plane 0 in interval r) and j-row r of plane q (needed in interval
        // r + q) enter shared memory kA ticks early; their global words are requested kF
        // ticks before that.  A word that is not there yet when its turn comes is polled.
        const int kgProd = BWD ? kg + 1 : kg - 1;
        const LLW* gKrow = a.gK + ((long long)((prevGroup ? kgProd : kg) * nJ + J) * steps) * 32 + lane;
        const int kq = BWD ? k0 + Wg - 1 - lane : k0 + lane;           // plane of stack position `lane`
        const int Tq = (lane < Wg ? kq : k0) * nJ + J;
        const LLW* gJrow = a.gJ + (long long)(jIn ? (BWD ? Tq + 1 : Tq - 1) : Tq) * steps;
        const unsigned int hkA = (unsigned int)__cvta_generic_to_shared(&sm->hk[0][lane]);
        const unsigned int hjA = (unsigned int)__cvta_generic_to_shared(&sm->hj[lane < Wg ? lane : 0][0]);
        const int nkEnd = prevGroup ? steps : 0;
        const int njEnd = (jIn && lane < Wg) ? nx : 0;
        const int jShift = BWD ? -31 : 31;
        const long long pbase = (long long)Tq * tileStride;
        const bool pfK = (lane > 0) || prevGroup;
        bool dead = false;
        unsigned long long loops = 0;
        LLW wk[kF], wj[kF];
#pragma unroll
        for (int q = 0; q < kF; q++) wk[q].f0 = wk[q].f1 = wj[q].f0 = wj[q].f1 = 0u;

        // forward k-row rk and this lane's j-row rj (if they exist) from the words in hand,
        // polling global memory for any that has not been published yet
        auto forward = [&](int rk, int rj, LLW& k_, LLW& j_) {
            const bool doK = rk >= 0 && rk < nkEnd;          // warp-uniform
            const bool doJ = rj >= 0 && rj < njEnd;          // per lane
            long long tstart = 0;
            for (int spin = 0; !dead; spin++) {
                const bool bad = (doK && !ok(k_, epoch)) || (doJ && !ok(j_, epoch));
                if (!__any_sync(0xffffffffu, bad)) break;
                loops++;
                if (doK) g_peek(gKrow + (t0 + dt * rk) * 32, k_);
                if (doJ) g_peek(gJrow + (t0 + dt * rj + jShift), j_);
                if ((spin & 63) == 63 && give_up(sh, tstart)) dead = true;
            }
            if (doK) sts64(hkA + (unsigned int)(rk & (kRing - 1)) * 256u, val(k_));
            if (doJ) sts64(hjA + (unsigned int)(rj & (kRing - 1)) * 8u, val(j_));
        };
        auto request = [&](int rk, int rj, LLW& k_, LLW& j_) {
            if (rk >= 0 && rk < nkEnd) g_peek(gKrow + (t0 + dt * rk) * 32, k_);
            if (rj >= 0 && rj < njEnd) g_peek(gJrow + (t0 + dt * rj + jShift), j_);
        };

        // rows of the first kA intervals, then the requests of the next kF
        for (int r = 0; r < kA; r++) {
            LLW k_, j_;
            k_.f0 = k_.f1 = j_.f0 = j_.f1 = 0u;
            forward(r, r - lane, k_, j_);
        }
#pragma unroll
        for (int q = 0; q < kF; q++) request(kA + q, kA + q - lane, wk[q], wj[q]);

        const int total = steps + Wg;
        for (int Tb = 0; Tb < total; Tb += kF) {
#pragma unroll
            for (int q = 0; q < kF; q++) {
                const int T = Tb + q;
                if (T < total) {
                    // interval T-1 is running and reads rows T-1 (k) and T-1-lane (j)
                    forward(T + kA, T + kA - lane, wk[q], wj[q]);
                    request(T + kA + kF, T + kA + kF - lane, wk[q], wj[q]);
                    if (false && lane < Wg && ((T - lane) & (kPfChunk - 1)) == 0) {
                        // operand rows kPfAhead steps ahead of plane `lane` into L2
                        const int r0 = T - lane + kPfAhead;
                        const int nrows = min(kPfChunk, steps - r0);
                        if (r0 >= 0 && nrows > 0) {
                            const int tlo = BWD ? steps - r0 - nrows : r0;
                            const long long off = pbase + (long long)tlo * 32;
                            const unsigned int bytes = (unsigned int)nrows * 256u;
                            if (pfK) l2_prefetch(a.pk + off, bytes);
                            l2_prefetch(a.pj + off, bytes);
                            l2_prefetch(a.pi + off, bytes);
                            l2_prefetch(a.Y + off, bytes);
                        }
                    }
                    if (trace && lane == 0 && sh->ticket == a.prefetch && T >= 110 && T < 138)
                        a.trace[8ull * nCta + (unsigned long long)W * 28ull + (T - 110)] = (unsigned long long)clock64();
                    tick();
                }
            }
        }
        if (dead) {
            sh->abort = 1;
            a.S->commError = 2;
            a.S->done = 1;
        }
        if (trace && lane == 0) {
            trace[4] = loops;
            trace[5] = gtimer();
        }
    }

    // last warp out of the last CTA re-arms the ticket for the next launch
    __syncwarp();
    if (lane == 0) {
        const int f = atomicAdd(&sh->finished, 1);
        if (f == W) {
            __threadfence();
            const unsigned int done = atomicAdd(&a.ticket[1], 1u);
            if (done == (unsigned int)nCta - 1u) {
                a.ticket[0] = 0u;
                a.ticket[1] = 0u;
                __threadfence();
            }
        }
    }
}

// ---------------------------------------------------------------------------
// natural cell order <-> tile layout, 32 steps x 32 lines per CTA through shared memory
// ---------------------------------------------------------------------------
template <bool MUL>
__global__ void __launch_bounds__(256) pack2_kernel(Box2 b, int nBlk, const double* __restrict__ src,
                                                     const double* __restrict__ rD, double* __restrict__ Y,
                                                     const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    __shared__ double s[32][33];
    const int T = blockIdx.x / nBlk, tb = (blockIdx.x - T * nBlk) * 32;
    const int k = T / b.nJ, J = T - k * b.nJ;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = wid; r < 32; r += 8) {
        const int j = J * 32 + r, i = tb - r + lane;
        double v = 0.0;
        if (j < b.ny && i >= 0 && i < b.nx) {
            const long long c = ((long long)k * b.ny + j) * b.nx + i;
            v = MUL ? __dmul_rn(rD[c], src[c]) : src[c];
        }
        s[r][lane] = v;
    }
    __syncthreads();
    for (int x = wid; x < 32; x += 8) {
        const int t = tb + x;
        if (t < b.steps) Y[((long long)T * b.steps + t) * 32 + lane] = s[lane][x];
    }
}

__global__ void __launch_bounds__(256) unpack2_kernel(Box2 b, int nBlk, const double* __restrict__ Y,
                                                       double* __restrict__ dst,
                                                       const SolverScalars* __restrict__ guard)
{
    if (guard && guard->done) return;
    __shared__ double s[32][33];
    const int T = blockIdx.x / nBlk, tb = (blockIdx.x - T * nBlk) * 32;
    const int k = T / b.nJ, J = T - k * b.nJ;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int x = wid; x < 32; x += 8) {
        const int t = tb + x;
        s[lane][x] = (t < b.steps) ? Y[((long long)T * b.steps + t) * 32 + lane] : 0.0;
    }
    __syncthreads();
    for (int r = wid; r < 32; r += 8) {
        const int j = J * 32 + r, i = tb - r + lane;
        if (j < b.ny && i >= 0 && i < b.nx) dst[((long long)k * b.ny + j) * b.nx + i] = s[r][lane];
    }
}

// premultiplied coefficients in tile layout: F* for the forward sweep (lower faces of a
// cell, neighbour order k-, j-, i-), B* for the backward one (upper faces, k+, j+, i+)
struct Products {
    double* F[3];
    double* B[3];
};

__host__ __device__ inline long long tile_pos2(const Box2& b, int i, int j, int k)
{
    const int J = j >> 5, l = j & 31;
    return ((long long)(k * b.nJ + J) * b.steps + (i + l)) * 32 + l;
}

__global__ void __launch_bounds__(kBlock) products2_kernel(Box2 b, const int* __restrict__ ownerStart,
                                                           const double* __restrict__ rD,
                                                           const double* __restrict__ coefF,
                                                           const double* __restrict__ coefB, Products P)
{
    const int n = b.nx * b.ny * b.nz;
    for (int c = blockIdx.x * kBlock + threadIdx.x; c < n; c += gridDim.x * kBlock) {
        const int i = c % b.nx, j = (c / b.nx) % b.ny, k = c / (b.nx * b.ny);
        const long long p = tile_pos2(b, i, j, k);
        const int hasI = i < b.nx - 1, hasJ = j < b.ny - 1, hasK = k < b.nz - 1;
        const int os = ownerStart[c];
        const double r = rD[c];
        // faces this cell owns, in the order +i, +j, +k
        P.B[2][p] = hasI ? __dmul_rn(r, coefB[os]) : 0.0;
        P.B[1][p] = hasJ ? __dmul_rn(r, coefB[os + hasI]) : 0.0;
        P.B[0][p] = hasK ? __dmul_rn(r, coefB[os + hasI + hasJ]) : 0.0;
        // faces where it is the upper cell: the +k / +j / +i face of the cell below /
        // behind / to the left (those cells have the same i, j flags where it matters)
        P.F[0][p] = k > 0 ? __dmul_rn(r, coefF[ownerStart[c - b.nx * b.ny] + hasI + hasJ]) : 0.0;
        P.F[1][p] = j > 0 ? __dmul_rn(r, coefF[ownerStart[c - b.nx] + hasI]) : 0.0;
        P.F[2][p] = i > 0 ? __dmul_rn(r, coefF[ownerStart[c - 1]]) : 0.0;
    }
}

struct ProductSlot {
    Products P;
    const double* rD = nullptr;
    const double* coefF = nullptr;
    const double* coefB = nullptr;
    long long sweepGen = -1, coefGen = -1;
    long long lastUse = 0;
    bool allocated = false;
};

struct State2 {
    Box2 b;
    int W = 16;
    long long padded = 0;
    double* Y = nullptr;
    LLW* gK = nullptr;
    LLW* gJ = nullptr;
    unsigned int* ticket = nullptr;
    unsigned int epoch = 0;
    ProductSlot slot[2];
    long long useClock = 0;
    bool attrSet = false;
    unsigned long long* trace = nullptr;
};

int alloc_padded2(void** user, size_t elems, size_t elemBytes, cudaStream_t st)
{
    const size_t pad = (size_t)kPad2 * 32;
    unsigned char* raw = nullptr;
    LDU_CUDA(cudaMalloc((void**)&raw, (elems + 2 * pad) * elemBytes));
    LDU_CUDA(cudaMemsetAsync(raw, 0, (elems + 2 * pad) * elemBytes, st));
    *user = raw + pad * elemBytes;
    return LDU_OK;
}

void free_padded2(void* user, size_t elemBytes)
{
    if (user) cudaFree((unsigned char*)user - (size_t)kPad2 * 32 * elemBytes);
}

size_t smem_bytes(int W)
{
    return ((size_t)W * 2 * 32 + (size_t)kRing * 32 + (size_t)W * kRing + (size_t)W * kNS * 4 * kC * 32 + (size_t)W * kNS) * sizeof(double)
           + sizeof(Shared2) + 64;
}

int pick_W(int nz)
{
    const char* e = getenv("LDU_STENCIL_W");
    if (e) {
        const int w = atoi(e);
        if (w == 4 || w == 8 || w == 15 || w == 16) return w;
    }
    return nz >= 32 ? 15 : (nz >= 6 ? 8 : 4);
}

int state2(ldu_matrix* m, State2** out)
{
    State2* s = reinterpret_cast<State2*>(m->stencil2);
    if (!s) {
        s = new State2();
        m->stencil2 = s;
        Box2& b = s->b;
        b.nx = m->box[0];
        b.ny = m->box[1];
        b.nz = m->box[2];
        b.nJ = (b.ny + 31) / 32;
        b.steps = b.nx + 31;
        b.nTiles = b.nz * b.nJ;
        s->W = pick_W(b.nz);
        b.nKg = (b.nz + s->W - 1) / s->W;
        s->padded = (long long)b.nTiles * b.steps * 32;
        cudaStream_t st = m->ctx->stream;
        LDU_TRY(alloc_padded2((void**)&s->Y, (size_t)s->padded, sizeof(double), st));
        const size_t nK = (size_t)b.nKg * b.nJ * b.steps * 32, nJw = (size_t)b.nTiles * b.steps;
        LDU_CUDA(cudaMalloc((void**)&s->gK, nK * sizeof(LLW)));
        LDU_CUDA(cudaMemsetAsync(s->gK, 0, nK * sizeof(LLW), st));
        LDU_CUDA(cudaMalloc((void**)&s->gJ, nJw * sizeof(LLW)));
        LDU_CUDA(cudaMemsetAsync(s->gJ, 0, nJw * sizeof(LLW), st));
        LDU_CUDA(cudaMalloc((void**)&s->ticket, 2 * sizeof(unsigned int)));
        LDU_CUDA(cudaMemsetAsync(s->ticket, 0, 2 * sizeof(unsigned int), st));
    }
    *out = s;
    return LDU_OK;
}

int products_for(ldu_matrix* m, State2* s, const double* rD, const double* coefF, const double* coefB,
                 Products* out)
{
    s->useClock++;
    for (int q = 0; q < 2; q++) {
        ProductSlot& ps = s->slot[q];
        if (ps.allocated && ps.rD == rD && ps.coefF == coefF && ps.coefB == coefB && ps.sweepGen == m->sweepGen
            && ps.coefGen == m->coefGen) {
            ps.lastUse = s->useClock;
            *out = ps.P;
            return LDU_OK;
        }
    }
    ProductSlot& ps = s->slot[0].lastUse <= s->slot[1].lastUse ? s->slot[0] : s->slot[1];
    cudaStream_t st = m->ctx->stream;
    if (!ps.allocated) {
        for (int d = 0; d < 3; d++) {
            LDU_TRY(alloc_padded2((void**)&ps.P.F[d], (size_t)s->padded, sizeof(double), st));
            LDU_TRY(alloc_padded2((void**)&ps.P.B[d], (size_t)s->padded, sizeof(double), st));
        }
        ps.allocated = true;
    }
    products2_kernel<<<grid_for(m->ctx, m->nCells), kBlock, 0, st>>>(s->b, m->d_ownerStart, rD, coefF, coefB, ps.P);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    ps.rD = rD;
    ps.coefF = coefF;
    ps.coefB = coefB;
    ps.sweepGen = m->sweepGen;
    ps.coefGen = m->coefGen;
    ps.lastUse = s->useClock;
    *out = ps.P;
    return LDU_OK;
}

template <int W>
int launch_sweeps(ldu_matrix* m, State2* s, S2Args& a, const Products& P)
{
    const size_t smem = smem_bytes(W);
    if (!s->attrSet) {
        LDU_CUDA(cudaFuncSetAttribute(sweep2_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDU_CUDA(cudaFuncSetAttribute(sweep2_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->attrSet = true;
    }
    const int grid = s->b.nKg * s->b.nJ;
    cudaStream_t st = m->ctx->stream;
    a.epoch = ++s->epoch;
    a.pk = P.F[0];
    a.pj = P.F[1];
    a.pi = P.F[2];
    const char* tracePath = getenv("LDU_S2_TRACE");
    if (tracePath && !s->trace) {
        LDU_CUDA(cudaMalloc((void**)&s->trace, ((size_t)grid * 8 + 17 * 28) * sizeof(unsigned long long)));
    }
    a.trace = tracePath ? s->trace : nullptr;
    if (a.trace) LDU_CUDA(cudaMemsetAsync(a.trace, 0, ((size_t)grid * 8 + 17 * 28) * sizeof(unsigned long long), st));
    sweep2_kernel<W, false><<<grid, (W + 1) * 32, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    if (a.trace) {   // debug only: dump the forward sweep's per-CTA timeline
        std::vector<unsigned long long> h((size_t)grid * 8 + 17 * 28);
        LDU_CUDA(cudaMemcpyAsync(h.data(), a.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        LDU_CUDA(cudaStreamSynchronize(st));
        if (FILE* f = fopen(tracePath, "w")) {
            for (int c = 0; c < grid; c++)
                fprintf(f, "%d %d %d %llu %llu %llu %llu %llu %llu %llu %llu\n", c, (int)(h[8 * c] >> 32),
                        (int)(h[8 * c] & 0xffffffffu), h[8 * c + 1], h[8 * c + 2], h[8 * c + 3], h[8 * c + 4], h[8 * c + 5],
                        h[8 * c + 6] & ((1ull << 40) - 1), h[8 * c + 6] >> 40, h[8 * c + 7]);
            for (int c = 0; c < 17; c++) {
                fprintf(f, "T%d", c);
                for (int n = 0; n < 28; n++) fprintf(f, " %llu", h[(size_t)grid * 8 + 28 * c + n]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        a.trace = nullptr;
    }
    a.epoch = ++s->epoch;
    a.pk = P.B[0];
    a.pj = P.B[1];
    a.pi = P.B[2];
    sweep2_kernel<W, true><<<grid, (W + 1) * 32, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

}  // namespace

int stencil_version(const ldu_matrix* m)
{
    if (m->box[0] <= 0 || !flow_enabled()) return 0;
    const char* e = getenv("LDU_STENCIL");   // 0: generic dataflow sweeps, 1: stencil.cu, 2 (default): stencil2.cu
    const int ver = e ? atoi(e) : 2;
    return (ver < 0 || ver > 2) ? 2 : ver;
}

void stencil2_free(ldu_matrix* m)
{
    State2* s = reinterpret_cast<State2*>(m->stencil2);
    if (!s) return;
    free_padded2(s->Y, sizeof(double));
    cudaFree(s->gK);
    cudaFree(s->gJ);
    cudaFree(s->ticket);
    cudaFree(s->trace);
    for (int q = 0; q < 2; q++)
        if (s->slot[q].allocated)
            for (int d = 0; d < 3; d++) {
                free_padded2(s->slot[q].P.F[d], sizeof(double));
                free_padded2(s->slot[q].P.B[d], sizeof(double));
            }
    delete s;
    m->stencil2 = nullptr;
}

// w = backward(forward(init ? rD*r : w)): both substitutions of a DIC / DILU / FDIC
// application.  coefF multiplies the lower faces in the forward sweep, coefB the upper
// faces in the backward sweep (DIC: upper/upper, DILU: lower/upper, DILU^T: upper/lower).
int stencil2_apply(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, const double* r,
                   double* w, bool init)
{
    State2* s;
    LDU_TRY(state2(m, &s));
    Products P;
    LDU_TRY(products_for(m, s, rD, coefF, coefB, &P));
    cudaStream_t st = m->ctx->stream;
    const int nBlk = (s->b.steps + 31) / 32;
    const int gridT = s->b.nTiles * nBlk;
    if (init) pack2_kernel<true><<<gridT, 256, 0, st>>>(s->b, nBlk, r, rD, s->Y, m->d_scalars);
    else pack2_kernel<false><<<gridT, 256, 0, st>>>(s->b, nBlk, w, nullptr, s->Y, m->d_scalars);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    S2Args a;
    a.S = m->d_scalars;
    a.guarded = 1;
    a.b = s->b;
    a.Y = s->Y;
    a.gK = s->gK;
    a.gJ = s->gJ;
    a.ticket = s->ticket;
    { const char* e = getenv("LDU_S2_PREFETCH"); a.prefetch = e ? atoi(e) : 0; }
    if (s->W == 16) LDU_TRY(launch_sweeps<16>(m, s, a, P));
    else if (s->W == 15) LDU_TRY(launch_sweeps<15>(m, s, a, P));
    else if (s->W == 8) LDU_TRY(launch_sweeps<8>(m, s, a, P));
    else LDU_TRY(launch_sweeps<4>(m, s, a, P));
    unpack2_kernel<<<gridT, 256, 0, st>>>(s->b, nBlk, s->Y, w, m->d_scalars);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

}  // namespace ldu
