// Plane-stacked pipelined triangular sweeps for structured hex boxes (second generation).
//
// Same arithmetic and the same skewed tile layout as stencil.cu (one warp owns a tile of 32
// x-lines of one k-plane, lane l walks line j0+l along i skewed by l steps, i-neighbour = own
// register, j-neighbour = one shuffle), but what bounded the first version is removed:
//
//   * the k-neighbour hop.  W consecutive k-planes of the same 32-line column are stacked in
//     ONE CTA (W compute warps) that tick together through a named barrier: plane k hands
//     its result to plane k+1 through shared memory, one tick later, instead of through a
//     {value, epoch} word polled in L2 (~10^3 cycles per plane in the first version).  Only
//     every W-th plane crosses CTAs.
//   * natural-layout traffic in the sweep.  The vector being substituted lives in tile layout
//     for both sweeps: every operand and the result is one coalesced 256-byte row per step.
//     pack/unpack kernels transpose 32x32 blocks through shared memory between the natural
//     cell order and the tile layout.
//   * global-memory latency on the compute warps.  Operand rows stream through a cp.async ring
//     in shared memory; a helper warp per CTA polls the words other CTAs publish in global
//     memory (the plane below the stack, the last line of the previous column), four rows per
//     round trip, and forwards them into tagged rings in shared memory.
//
// The coefficient operands are stored premultiplied, P = rD[c]*coef[f], the product the
// reference forms first (wA[u] -= rD[u]*upper[f]*wA[l], DICPreconditioner.C:108-121), so each
// term is one multiply and one subtract and results stay BIT-IDENTICAL to the reference and
// to the generic dataflow path.
//
// CTAs claim their (stack, column) tile from an atomic ticket in dependency order, so a CTA
// only ever waits on CTAs that are already running: no co-residency requirement.
//
// What was measured on B200 while getting here (216^3, microseconds per sweep incl. its share
// of pack/unpack; tests/perf_sweeps.py, tests/micro/sync_latency.cu):
//   first generation (one word in L2 per cell)                               1263
//   stacks + per-step {value, tag} polling between planes                     595   (54 % of all
//       instructions were polls; every plane catches up with the one below and then polls)
//   stacks ticking through bar.sync, operands double-buffered in registers    528   (the planes
//       reach their refill burst in different ticks, so every tick has a straggler)
//   the same through an mbarrier (split arrive/wait)                       640-760   (an mbarrier
//       tick costs 150-370 cycles against 40-60 for bar.sync)
//   per-step register FIFO of operands                                        790   (loads share
//       the warp's counting scoreboards: waiting for the oldest waits for the youngest)
//   TMA bulk chunks + mbarrier per ring slot                                  585   (chunk
//       boundaries of the planes fall into different ticks again)
//   cp.async ring, uniform steps, asynchronous helper (this file)             452   (W = 6)
//   + operands as 32-byte structs (2 x 16-byte cp.async, 2 x LDS.128), loop not unrolled,
//     role-specialised loops, out-of-line slow paths: 35 instead of 90 instructions per
//     step, the first CTA runs at 0.20 us per step instead of 0.28 ...          515   (the hops
//       between CTAs, not the step, set the pace: ~3.7 us per stack, ~17 us per column)
//   + thread-block clusters of 8 stacks handing over through distributed shared
//     memory (st.shared::cluster into the next CTA's ring)                  474-493   (two CTAs
//       per SM and cluster placement slow the first CTA to 0.39 us per step)
//   + helper forwarding the j-words of all planes with one load (lane = plane x row), 6 k-rows
//     per round trip, optionally split into a k-helper and a j-helper warp     472-477   (column
//       hops 13 instead of 17-25 us, stack hops 5.5 instead of 3.7 us: no net gain)
// A tick still costs ~470 cycles (60-70 dependent instructions per step and plane); the
// barrier + hand-off chain alone is 105.
// Round 2 added two more generations below (sweep3_kernel: one register-stacked warp per group, 590-610 us;
// sweep4_kernel: chains of eight such warps per CTA, 344 us -- the default); their headers and DESIGN.md 4
// have the measurements.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "epilogue.cuh"
#include "reduce.cuh"
#include "sweeps.h"

namespace ldu {

namespace {

constexpr int kRing = 16;          // ring depth (steps) of every shared-memory hand-off
constexpr int kD = 8;              // operand rows in flight per plane: register FIFO depth = unroll factor
constexpr int kPad2 = kD + 2;      // zero rows in front of / behind the tile arrays
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
    int W3;      // 0: tile layout [tile][step][lane]; > 0: stacked layout of the register-stacked sweeps, see tile_row
    int ticks;   // stacked layout: steps + W3 / blk - 1 ticks per (stack, column) group
    int blk;     // stacked layout: planes per block (1: every plane lags one tick behind the plane below; b: the b
                 // planes of a block run the SAME step in a tick, blocks lag one tick), see tile_row
};

// Row (32 doubles) of step t of the tile (plane k, column J).
//   tile layout    : [k][J][t]
//   stacked layout : [kg = k / W][J][tick = t + p][p = k % W] -- the W rows one warp of sweep3_kernel
//     needs in one tick (plane p of the stack lags p steps behind plane 0) are adjacent: 256*W bytes per
//     operand and tick, one contiguous stream per (stack, column) group in both sweep directions.
//     Blocked (blk = planes per warp of sweep4_kernel<.., BLK = true>): tick = t + p / blk.  The planes of a warp
//     run the same step in a tick, one after the other (the k-neighbour of the second plane is the value the warp
//     has just computed): a stack of 16 planes spans 8 ticks instead of 16, at the price of a dependent chain of
//     blk cells inside a tick.
__device__ __forceinline__ long long tile_row(const Box2& b, int k, int J, int t)
{
    if (b.W3 == 0) return (long long)(k * b.nJ + J) * b.steps + t;
    const int kg = k / b.W3, p = k - kg * b.W3;
    return ((long long)(kg * b.nJ + J) * b.ticks + (t + p / b.blk)) * b.W3 + p;
}

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
    unsigned long long* trace;   // debug (LDU_S2_TRACE): per CTA {ticket, start, first plane done, last plane done, helper loops, helper done}
    int dbg;                     // debug (LDU_S3_DBG): 1 = operands by direct loads (no cp.async ring), 2 = block-wide barrier
                                 // around the ring refill, 4 = fence before the words are published
};

struct Shared2 {
    volatile int prog[32];   // steps completed by plane p, published every 4 steps
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

__device__ __forceinline__ void s_peek_a(unsigned int addr, LLW& w)
{
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(w.lo), "=r"(w.f0), "=r"(w.hi), "=r"(w.f1)
                 : "r"(addr)
                 : "memory");
}

__device__ __forceinline__ void s_store_a(unsigned int addr, unsigned int lo, unsigned int hi, unsigned int tag)
{
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(lo), "r"(tag), "r"(hi), "r"(tag)
                 : "memory");
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
//   kx  [W][2][32]        plane p's result of its step n, slot n&1: read by plane p+1 one tick later
//   hk  [kRing][32] LLW   helper -> first plane: results of the stack below, {value, tag = n + 1}
//   hj  [W][kRing]  LLW   helper -> edge lanes: last line of the previous column, {value, tag = n + 1}
//   ops [W][kD][4][32]    operand ring of every plane: pk, pj, pi, src rows of the next kD steps
//   prog[W]               steps completed by plane p (flow control of the helper's rings)
template <int W>
struct Smem2 {
    double kx[W][2][32];
    LLW hk[kRing][32];
    LLW hj[W][kRing];
    double ops[W][kD][4][32];
    Shared2 sh;
};

// One stack of up to W planes of one column.  The planes tick together through named barrier 1:
// in interval T plane p executes its step n = T - p, reading what plane p-1 wrote in interval
// T-1 (measured on B200, tests/micro/sync_latency.cu: bar.sync + LDS + 4 dependent FP64 ops +
// STS = 105 cycles per tick for 8 warps; an mbarrier costs 150-370).  Every step of every plane
// costs the same (a straggler delays all planes every tick): the operand rows of the next kD
// steps stream through a cp.async ring, one commit group per step, so no step has a refill
// burst and nothing waits on register scoreboards shared with younger loads.
// The helper warp is not part of the barrier.  It polls the {value, epoch} words other CTAs
// publish in global memory (the plane below the stack, the last line of the previous column),
// four rows per round trip, and forwards them into tagged rings in shared memory, as far
// ahead as the ring allows; only the first plane and the edge lanes ever look at a tag.
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
    if (threadIdx.x < 32) sh->prog[threadIdx.x] = 0;
    {
        // tags 0 everywhere: never equal to n + 1
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        uint4* q0 = reinterpret_cast<uint4*>(&sm->hk[0][0]);
        for (int q = threadIdx.x; q < kRing * 32 + W * kRing; q += (W + 1) * 32) q0[q] = z;
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
    const int nThreads = Wg * 32;
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
            const bool fromHelper = (p == 0) && prevGroup;       // warp-uniform
            const bool hasConsumer = p + 1 < Wg;
            const bool pubK = (p == Wg - 1) && nextGroup;
            const long long base = (long long)T * tileStride + lane;
            const double* pPk = a.pk + base;
            const double* pPj = a.pj + base;
            const double* pPi = a.pi + base;
            double* pY = a.Y + base;
            const unsigned int kIn = (unsigned int)__cvta_generic_to_shared(&sm->kx[p > 0 ? p - 1 : 0][0][lane]);
            const unsigned int kOut = (unsigned int)__cvta_generic_to_shared(&sm->kx[p][0][lane]);
            const unsigned int hkA = (unsigned int)__cvta_generic_to_shared(&sm->hk[0][lane]);
            const unsigned int hjA = (unsigned int)__cvta_generic_to_shared(&sm->hj[p][0]);
            const unsigned int opsA = (unsigned int)__cvta_generic_to_shared(&sm->ops[p][0][0][lane]);
            LLW* gKself = a.gK + ((long long)(kg * nJ + J) * steps) * 32 + lane;
            LLW* gJself = a.gJ + (long long)T * steps;
            const unsigned int nxEff = lineValid ? (unsigned int)nx : 0u;   // cell i = t - lane exists iff i < nxEff
            const int iFirst = BWD ? nx - 1 : 0;                            // cell without an i-neighbour

            auto issue = [&](int n_, unsigned int slot) {
                const long long e = (long long)(t0 + dt * n_) * 32;
                const unsigned int d = opsA + slot * 1024u;
                cp_async8(d, pPk + e);
                cp_async8(d + 256u, pPj + e);
                cp_async8(d + 512u, pPi + e);
                cp_async8(d + 768u, pY + e);
                cp_async_commit();
            };
            auto fetch = [&](unsigned int slot, StepOps2& o) {
                const unsigned int d = opsA + slot * 1024u;
                o.pk = lds64(d);
                o.pj = lds64(d + 256u);
                o.pi = lds64(d + 512u);
                o.src = lds64(d + 768u);
            };
            bool dead = false;
            // value of a word the helper forwards; polls shared memory while it is not there yet
            auto helper_word = [&](unsigned int addr, unsigned int tag, bool need) -> double {
                LLW w;
                s_peek_a(addr, w);
                if (__any_sync(0xffffffffu, need && !ok(w, tag))) {
                    long long tstart = 0;
                    for (int spin = 0; !dead; spin++) {
                        s_peek_a(addr, w);
                        if (!__any_sync(0xffffffffu, need && !ok(w, tag))) break;
                        if ((spin & 1023) == 1023 && give_up(sh, tstart)) dead = true;
                    }
                }
                return val(w);
            };
#pragma unroll
            for (int q = 0; q < kD; q++) issue(q, (unsigned int)q);
            for (int q = 0; q < p; q++) tick();   // plane p starts in interval p

            StepOps2 o;   // operands of the coming step
            cp_async_wait<kD - 2>();   // rows of steps 0 and 1: the loop fetches one step ahead
            fetch(0u, o);
            double prev = 0.0;
            // k-neighbour value and j- / i-terms of the coming step.  Step 0 only has the edge
            // lane's cell (i == iFirst): no i-term, a j-term if a column feeds this one.
            double vk = 0.0;   // no k-neighbour: pk == +0 too and src stays as it is
            if (p > 0) vk = lds64(kIn);
            else if (fromHelper) vk = helper_word(hkA, 1u, true);
            double ti = 0.0;
            double tj = 0.0;
            if (jIn) {
                const double vje = helper_word(hjA, 1u, edgeIn);
                if (edgeIn) tj = __dmul_rn(o.pj, vje);
            }
            for (int nb = 0; nb < steps; nb += kD) {
                const int tb = t0 + dt * nb;
#pragma unroll
                for (int q = 0; q < kD; q++) {
                    const unsigned int n = (unsigned int)(nb + q);
                    if ((int)n < steps) {
                        const int t = tb + dt * q;
                        const int i = t - lane;
                        const bool active = (unsigned int)i < nxEff;
                        // ---- result -> hand-off -> barrier
                        double acc = __dsub_rn(o.src, __dmul_rn(o.pk, vk));
                        acc = __dsub_rn(acc, tj);    // tj, ti == +0 where the neighbour does not exist
                        acc = __dsub_rn(acc, ti);
                        if (hasConsumer) sts64(kOut + (n & 1u) * 256u, acc);
                        tick();
                        // ---- first what the next result waits for longest: the k-neighbour's value
                        if (p > 0) vk = lds64(kIn + ((n + 1u) & 1u) * 256u);
                        else if (fromHelper) vk = helper_word(hkA + ((n + 1u) & (unsigned int)(kRing - 1)) * 512u, n + 2u, (int)n + 1 < steps);
                        fetch((unsigned int)((q + 1) % kD), o);        // landed: see the wait below
                        if (active) prev = acc;
                        {   // terms of step n + 1
                            double vj = BWD ? __shfl_down_sync(0xffffffffu, prev, 1) : __shfl_up_sync(0xffffffffu, prev, 1);
                            if (jIn) {
                                const bool need = edgeIn && (unsigned int)(i + dt) < nxEff;
                                const double vje = helper_word(hjA + ((n + 1u) & (unsigned int)(kRing - 1)) * 16u, n + 2u, need);
                                if (edgeIn) vj = vje;
                            }
                            tj = hasJ ? __dmul_rn(o.pj, vj) : 0.0;
                            ti = (i + dt != iFirst) ? __dmul_rn(o.pi, prev) : 0.0;
                        }
                        // ---- then what nobody in this CTA waits for
                        if (active) pY[t * 32] = acc;
                        if (pubK) g_store(gKself + t * 32, acc, epoch);     // warp-uniform branch
                        if (jOut) {
                            if (edgeOut && active) g_store(gJself + t, acc, epoch);
                        }
                        issue((int)n + kD, (unsigned int)q);          // this step's slot is free again
                        cp_async_wait<kD - 2>();                       // the rows of step n + 2 have landed
                        if (((q + 1) & 3) == 0 && lane == 0) sh->prog[p] = (int)n + 1;
                    }
                }
            }
            cp_async_wait<0>();
            for (int q = p + 1; q < Wg; q++) tick();
            if (dead) {
                sh->abort = 1;
                a.S->commError = 2;
                a.S->done = 1;
            }
            __syncwarp();
            if (lane == 0) sh->prog[p] = steps;
            if (trace && lane == 0 && p == 0) trace[2] = gtimer();
            if (trace && lane == 0 && p == Wg - 1) trace[3] = gtimer();
        }
    } else {
        // ------------------------------------------------------------------ helper warp
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
        int nk = 0, nj = 0;     // rows forwarded so far (nj: for the plane of this lane)
        long long tstart = 0;
        unsigned long long loops = 0;
        for (int spin = 0;; loops++) {
            const int pr = (lane < Wg) ? sh->prog[lane] : 0x7fffffff;
            const int minProg = __reduce_min_sync(0xffffffffu, pr);
            const int p0 = __shfl_sync(0xffffffffu, pr, 0);
            if (__any_sync(0xffffffffu, sh->abort != 0)) break;
            bool did = false;
            // a ring slot is free again once its reader has finished the step kRing before;
            // prog is published every 4 steps
            const int capK = min(nkEnd, p0 + kRing - 1), capJ = min(njEnd, pr + kRing - 1);
            LLW wk[4], wj[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (nk + q < capK) g_peek(gKrow + (t0 + dt * (nk + q)) * 32, wk[q]);
                if (nj + q < capJ) g_peek(gJrow + (t0 + dt * (nj + q) + jShift), wj[q]);
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {   // k-rows: all lanes of a row must be there
                if (nk >= capK || !__all_sync(0xffffffffu, ok(wk[q], epoch))) break;
                s_store_a(hkA + (unsigned int)(nk & (kRing - 1)) * 512u, wk[q].lo, wk[q].hi, (unsigned int)nk + 1u);
                nk++;
                did = true;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {   // j-words: every lane for its own plane
                if (nj >= capJ || !ok(wj[q], epoch)) break;
                s_store_a(hjA + (unsigned int)(nj & (kRing - 1)) * 16u, wj[q].lo, wj[q].hi, (unsigned int)nj + 1u);
                nj++;
                did = true;
            }
            did = __any_sync(0xffffffffu, did);
            const bool jDone = __all_sync(0xffffffffu, nj >= njEnd);
            if (nk >= nkEnd && jDone) break;    // everything forwarded: the planes finish alone
            if (did) {
                spin = 0;
                tstart = 0;
            } else {
                spin++;
                if ((spin & 63) == 63 && give_up(sh, tstart)) {
                    sh->abort = 1;
                    a.S->commError = 2;
                    a.S->done = 1;
                    break;
                }
            }
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
// Third generation: REGISTER-STACKED planes (LDU_STENCIL=3, opt-in).
//
// What bounded sweep2_kernel was the tick: W warps of a CTA meeting at a named barrier once per
// step and handing a 256-byte row to the next plane through shared memory (~470 cycles per tick
// measured, ~1200 with the CTA-to-CTA hops).  Here ONE WARP owns the whole stack: W consecutive
// k-planes of a 32-line column live in the registers of the same lanes (res[p] = result of plane p
// in the previous tick), so
//   * the k-neighbour of plane p is the register res[p-1] of the same lane (no shared memory, no barrier),
//   * the j-neighbour is one shuffle, the i-neighbour the lane's own register,
//   * the W planes of a tick are W independent dependency chains (mul, sub, mul, sub, mul, sub): the
//     FP64 pipe stays busy without any other warp,
//   * nothing synchronises inside the tick.
// Operands arrive through a cp.async ring from the STACKED layout (tile_row): the 4 x W rows of a tick
// are 4 contiguous blocks of 256 W bytes.  Only the faces of a (stack, column) group cross CTAs, as
// {value, epoch} words in L2:
//   gK[group][step][lane]      the last plane of the stack below (above, backward sweep)
//   gJ[group][i + p][p]        the edge line of the previous (next) column, one 16 W-byte row per tick
// A helper warp (the CTA's second warp, on another SM sub-partition) polls them, several rows per round
// trip, and forwards them into tagged rings in shared memory: the compute warp never has a global load in
// flight (loads kept in registers across ticks share the warp's counting scoreboards with the LDS of the
// tick: the first version of this kernel paid one L2 round trip per tick for its look-ahead polls).
// Groups are claimed from an atomic ticket in dependency order (no co-residency requirement).
// Arithmetic and its order are those of sweep2_kernel: bit-identical to the reference.
// ---------------------------------------------------------------------------
constexpr int kRing3 = 16;          // ticks of forwarded face words kept in shared memory
constexpr int kOpRingBytes = 65536; // operand ring: 64 KB / (4 W 256 B) ticks deep

template <int W>
struct Smem3 {
    LLW hk[kRing3][32];     // helper -> compute: k-face row of loop tick s in slot s % kRing3, tag s + 1
    LLW hj[kRing3][W];      // helper -> compute: j-face row
    volatile int prog;      // ticks completed by the compute warp (flow control of the rings)
    volatile int abort;
    int ticket;
    int pad;
};

__device__ __forceinline__ void cp_async16(unsigned int dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int W, bool BWD>
__global__ void __launch_bounds__(64, 1) sweep3_kernel(S2Args a)
{
    constexpr int kD3 = kOpRingBytes / (4 * W * 256);           // operand ring depth in ticks
    constexpr unsigned int kOpBytes = W * 256u, kSlotBytes = 4u * kOpBytes;
    extern __shared__ uint4 smem_raw[];
    Smem3<W>* sm = reinterpret_cast<Smem3<W>*>(reinterpret_cast<unsigned char*>(smem_raw) + kOpRingBytes);
    if (a.guarded && a.S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        sm->ticket = (int)atomicAdd(&a.ticket[0], 1u);
        sm->prog = 0;
        sm->abort = 0;
    }
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);      // tags 0: never equal to s + 1
        uint4* q0 = reinterpret_cast<uint4*>(&sm->hk[0][0]);
        for (int q = threadIdx.x; q < kRing3 * (32 + W); q += 64) q0[q] = z;
    }
    __syncthreads();
    const int tk = sm->ticket;

    const int nx = a.b.nx, nJ = a.b.nJ, steps = a.b.steps, ticks = a.b.ticks, nKg = a.b.nKg;
    const int nCta = nKg * nJ;
    const int tr = BWD ? nCta - 1 - tk : tk;
    const int kg = tr / nJ, J = tr - kg * nJ;
    const int g = kg * nJ + J;
    const int Wg = min(W, a.b.nz - kg * W);                    // planes of this stack that exist
    const bool kIn = BWD ? (kg + 1 < nKg) : (kg > 0);
    const bool kOut = BWD ? (kg > 0) : (kg + 1 < nKg);
    const bool jIn = BWD ? (J + 1 < nJ) : (J > 0);
    const bool jOut = BWD ? (J > 0) : (J + 1 < nJ);
    const unsigned int epoch = a.epoch;
    const int nRho = nx + W - 1;                               // rows of a group's gJ block
    const LLW* gKin = a.gK + ((long long)(BWD ? g + nJ : g - nJ) * steps) * 32 + lane;
    const LLW* gJin = a.gJ + ((long long)(BWD ? g + 1 : g - 1) * nRho) * W + (lane < W ? lane : 0);
    // loop tick s -> layout tick; rows of the face words loop tick s needs (negative: none)
    auto sigma_of = [&](int s_) { return BWD ? ticks - 1 - s_ : s_; };
    auto k_row = [&](int s_) -> int {
        const int sg = sigma_of(s_);
        const int t = BWD ? sg - (W - 1) : sg;
        return (kIn && t >= 0 && t < steps) ? t : -1;
    };
    auto j_row = [&](int s_) -> int {
        const int sg = sigma_of(s_);
        const int rho = BWD ? sg - 31 : sg;
        return (jIn && rho >= 0 && rho < nRho) ? rho : -1;
    };
    unsigned long long* trace = a.trace ? a.trace + 8ull * (unsigned int)tk : nullptr;

    if (warp == 1) {
        // ------------------------------------------------------------------ helper warp
        if (!kIn && !jIn) return;
        const unsigned int hkA = (unsigned int)__cvta_generic_to_shared(&sm->hk[0][lane]);
        const unsigned int hjA = (unsigned int)__cvta_generic_to_shared(&sm->hj[0][lane < W ? lane : 0]);
        int sk = kIn ? 0 : ticks, sj = jIn ? 0 : ticks;     // next loop tick to forward, per ring
        long long tstart = 0;
        for (int spin = 0; sk < ticks || sj < ticks;) {
            if (sm->abort) break;
            const int cap = min(ticks, sm->prog + kRing3 - 1);   // slot s % kRing3 is free once tick s - kRing3 is done
            // skip the ticks that need no word
            while (sk < cap && k_row(sk) < 0) sk++;
            while (sj < cap && j_row(sj) < 0) sj++;
            LLW wk[4], wj[4];
            int rk[4], rj[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                rk[q] = (sk + q < cap) ? k_row(sk + q) : -1;
                rj[q] = (sj + q < cap) ? j_row(sj + q) : -1;
                if (rk[q] >= 0) g_peek(gKin + (long long)rk[q] * 32, wk[q]);
                if (rj[q] >= 0 && lane < W) g_peek(gJin + (long long)rj[q] * W, wj[q]);
            }
            bool did = false;
#pragma unroll
            for (int q = 0; q < 4; q++) {       // rows become valid in order
                if (rk[q] < 0 || !__all_sync(0xffffffffu, ok(wk[q], epoch))) break;
                s_store_a(hkA + (unsigned int)(sk & (kRing3 - 1)) * 512u, wk[q].lo, wk[q].hi, (unsigned int)sk + 1u);
                sk++;
                did = true;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (rj[q] < 0 || !__all_sync(0xffffffffu, lane >= W || ok(wj[q], epoch))) break;
                if (lane < W) s_store_a(hjA + (unsigned int)(sj & (kRing3 - 1)) * (W * 16u), wj[q].lo, wj[q].hi, (unsigned int)sj + 1u);
                sj++;
                did = true;
            }
            if (did) {
                spin = 0;
                tstart = 0;
            } else if ((++spin & 63) == 63) {
                if (tstart == 0) tstart = clock64();
                else if (clock64() - tstart > kTimeout2) {
                    sm->abort = 1;
                    a.S->commError = 2;
                    a.S->done = 1;
                    break;
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warp
    const long long elemBase = (long long)g * ticks * W * 32;  // first element of the group's stream
    const double* pOp[4] = {a.pk + elemBase, a.pj + elemBase, a.pi + elemBase, a.Y + elemBase};
    double* pY = a.Y + elemBase + lane;
    LLW* gKout = a.gK + ((long long)g * steps) * 32 + lane;
    LLW* gJout = a.gJ + ((long long)g * nRho) * W;
    const unsigned int ring = (unsigned int)__cvta_generic_to_shared(smem_raw);
    const double* ringP = reinterpret_cast<const double*>(smem_raw) + lane;
    const unsigned int hkA = (unsigned int)__cvta_generic_to_shared(&sm->hk[0][lane]);
    const unsigned int hjA = (unsigned int)__cvta_generic_to_shared(&sm->hj[0][0]);
    const int edgeLane = BWD ? 31 : 0, pubLane = BWD ? 0 : 31;

    auto issue = [&](int s_, unsigned int slot) {
        if (s_ < ticks) {
            const long long e = (long long)sigma_of(s_) * (W * 32);
            const unsigned int d = ring + slot * kSlotBytes + (unsigned int)lane * 16u;
#pragma unroll
            for (int o = 0; o < 4; o++) {
                const double* src = pOp[o] + e + lane * 2;
#pragma unroll
                for (int c = 0; c < W / 2; c++) cp_async16(d + o * kOpBytes + c * 512u, src + c * 64);
            }
        }
        cp_async_commit();
    };
    long long trWaits = 0, trWaitCyc = 0, trOpCyc = 0;
    const long long trStartC = clock64();
    if (trace && lane == 0) {
        trace[0] = ((unsigned long long)kg << 32) | (unsigned int)J;
        trace[1] = gtimer();
    }
    bool dead = false;

    double res[W];
#pragma unroll
    for (int p = 0; p < W; p++) res[p] = 0.0;
#pragma unroll
    for (int q = 0; q < kD3; q++) issue(q, (unsigned int)q);

    unsigned int slot = 0;
#pragma unroll 1
    for (int s_ = 0; s_ < ticks; s_++) {
        const int sg = sigma_of(s_);
        const unsigned int hslot = (unsigned int)(s_ & (kRing3 - 1)), tag = (unsigned int)s_ + 1u;
        // ---- faces of the group, forwarded by the helper
        double vkin = 0.0;
        double ve[W];
#pragma unroll
        for (int p = 0; p < W; p++) ve[p] = 0.0;
        if (k_row(s_) >= 0) {
            LLW w;
            s_peek_a(hkA + hslot * 512u, w);
            if (!__all_sync(0xffffffffu, ok(w, tag))) {
                const long long c0 = clock64();
                while (!dead) {
                    s_peek_a(hkA + hslot * 512u, w);
                    if (__all_sync(0xffffffffu, ok(w, tag))) break;
                    if (sm->abort || clock64() - c0 > kTimeout2) dead = true;
                    dead = __any_sync(0xffffffffu, dead);
                }
                if (trace) { trWaits++; trWaitCyc += clock64() - c0; }
            }
            vkin = val(w);
        }
        if (j_row(s_) >= 0) {
            LLW w[W];
            bool good = true;
#pragma unroll
            for (int p = 0; p < W; p++) {
                s_peek_a(hjA + hslot * (W * 16u) + p * 16u, w[p]);
                good = good && ok(w[p], tag);
            }
            if (!__all_sync(0xffffffffu, good)) {     // every lane reads the same words; the vote keeps the warp together
                const long long c0 = clock64();
                while (!dead) {
                    good = true;
#pragma unroll
                    for (int p = 0; p < W; p++) {
                        s_peek_a(hjA + hslot * (W * 16u) + p * 16u, w[p]);
                        good = good && ok(w[p], tag);
                    }
                    if (__all_sync(0xffffffffu, good)) break;
                    if (sm->abort || clock64() - c0 > kTimeout2) dead = true;
                    dead = __any_sync(0xffffffffu, dead);
                }
                if (trace) { trWaits += 1ll << 32; trWaitCyc += clock64() - c0; }
            }
#pragma unroll
            for (int p = 0; p < W; p++) ve[p] = val(w[p]);
        }
        // ---- operands of this tick have landed
        {
            const long long c0 = trace ? clock64() : 0;
            cp_async_wait<kD3 - 1>();
            __syncwarp();
            if (trace) trOpCyc += clock64() - c0;
        }
        const double* o = ringP + slot * (kSlotBytes / 8);
        double opk[W], opj[W], opi[W], src[W], vj[W], nr[W];
        if (a.dbg & 1) {
            const long long e = (long long)sg * (W * 32) + lane;
#pragma unroll
            for (int p = 0; p < W; p++) {
                opk[p] = __ldcg(pOp[0] + e + p * 32);
                opj[p] = __ldcg(pOp[1] + e + p * 32);
                opi[p] = __ldcg(pOp[2] + e + p * 32);
                src[p] = __ldcg(pOp[3] + e + p * 32);
            }
        } else {
#pragma unroll
        for (int p = 0; p < W; p++) {
            opk[p] = o[p * 32];
            opj[p] = o[W * 32 + p * 32];
            opi[p] = o[2 * W * 32 + p * 32];
            src[p] = o[3 * W * 32 + p * 32];
        }
        }
#pragma unroll
        for (int p = 0; p < W; p++) {
            vj[p] = BWD ? __shfl_down_sync(0xffffffffu, res[p], 1) : __shfl_up_sync(0xffffffffu, res[p], 1);
            if (jIn && lane == edgeLane) vj[p] = ve[p];
        }
        // ---- W planes, each one step: res[p] <- src - pk*vk - pj*vj - pi*res[p], all from the old values
#pragma unroll
        for (int p = 0; p < W; p++) {
            const double vk = BWD ? (p < W - 1 ? res[p < W - 1 ? p + 1 : p] : vkin) : (p > 0 ? res[p > 0 ? p - 1 : p] : vkin);
            double acc = __dsub_rn(src[p], __dmul_rn(opk[p], vk));
            acc = __dsub_rn(acc, __dmul_rn(opj[p], vj[p]));
            nr[p] = __dsub_rn(acc, __dmul_rn(opi[p], res[p]));
        }
#pragma unroll
        for (int p = 0; p < W; p++) res[p] = nr[p];
        // ---- results: the faces other groups wait for first, then the vector itself
        if (a.dbg & 4) __threadfence();
        if (kOut) {
            const int t = BWD ? sg : sg - (W - 1);
            if (t >= 0 && t < steps) g_store(gKout + (long long)t * 32, BWD ? res[0] : res[W - 1], epoch);
        }
        if (jOut) {
            const int rho = BWD ? sg : sg - 31;
            if (rho >= 0 && rho < nRho && lane == pubLane) {
#pragma unroll
                for (int p = 0; p < W; p++) g_store(gJout + (long long)rho * W + p, res[p], epoch);
            }
        }
#pragma unroll
        for (int p = 0; p < W; p++) {
            const int t = sg - p;
            if (t >= 0 && t < steps && p < Wg) pY[((long long)sg * W + p) * 32] = res[p];
        }
        // ---- refill this tick's ring slot
        __syncwarp();
        issue(s_ + kD3, slot);
        slot = slot + 1u == (unsigned int)kD3 ? 0u : slot + 1u;
        if (lane == 0) sm->prog = s_ + 1;
    }
    cp_async_wait<0>();
    if (dead && lane == 0) {
        sm->abort = 1;
        a.S->commError = 2;
        a.S->done = 1;
    }
    if (trace && lane == 0) {
        trace[2] = gtimer();
        trace[3] = (unsigned long long)trWaits;
        trace[4] = (unsigned long long)trWaitCyc;
        trace[5] = (unsigned long long)trOpCyc;
        trace[6] = (unsigned long long)(clock64() - trStartC);
    }
    // the last CTA re-arms the ticket for the next launch
    if (lane == 0) {
        __threadfence();
        const unsigned int doneCtas = atomicAdd(&a.ticket[1], 1u);
        if (doneCtas == (unsigned int)nCta - 1u) {
            a.ticket[0] = 0u;
            a.ticket[1] = 0u;
            __threadfence();
        }
    }
}


// ---------------------------------------------------------------------------
// Fourth generation: CHAINED register-stacked warps (LDU_STENCIL=4).
//
// What the trace of sweep3_kernel showed: a lone warp with 8 planes in its registers needs 0.57 us per tick (it
// has nobody to hide its LDS / SHFL / STG / LDGSTS latencies behind), and every stack-to-stack hop through L2 costs
// 3 us.  Here a CTA is a CHAIN of M such warps with W planes each (W = 2): warp q hands the result of its last
// plane to warp q+1 through a ring of tagged 16-byte words in shared memory ({value, tick+1}; the consumer polls
// its word, the producer throttles on the consumer's progress counter), so
//   * a tick of a warp is ~90 instructions instead of ~350,
//   * the M warps of a CTA run on different SM sub-partitions as a dataflow pipeline: no barrier,
//   * a CTA covers M*W planes: only every (M*W)-th plane crosses CTAs through L2.
// Layout, operand rings (cp.async, per warp), the helper warp that forwards the group's faces from L2 and the
// arithmetic are those of sweep3_kernel with W3 = M*W planes per stack.
// ---------------------------------------------------------------------------
constexpr int kRing4 = 16;     // ticks of forwarded j-face words
// ticks of warp-to-warp hand-off words (also the ring of the stack's k-face words): 16, 8 for chains of 16 warps
template <int M> struct Hand4 { static constexpr int value = M > 8 ? 8 : 16; };
constexpr int kD4 = 8;         // operand ring depth (ticks) of every warp

template <int W, int M>
struct Smem4 {
    double ops[M][kD4][4][W][32];   // operand rings
    LLW kx[M][Hand4<M>::value][32];          // kx[q]: input words of pipeline warp q (q = 0: from the helper)
    LLW hj[kRing4][M * W];          // helper -> all warps: j-face row of loop tick s in slot s % kRing4, tag s + 1
    volatile int prog[M];           // ticks completed by pipeline warp q
    volatile int abort;
    int ticket;
};

template <int W, int M, bool BWD, bool BLK, bool DIV = false>
__global__ void __launch_bounds__((M + 2) * 32, 1) sweep4_kernel(S2Args a)
{
    constexpr int MW = M * W;
    constexpr int kHand4 = Hand4<M>::value;
    extern __shared__ uint4 smem_raw[];
    Smem4<W, M>* sm = reinterpret_cast<Smem4<W, M>*>(smem_raw);
    if (a.guarded && a.S->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        sm->ticket = (int)atomicAdd(&a.ticket[0], 1u);
        sm->abort = 0;
    }
    if (threadIdx.x < M) sm->prog[threadIdx.x] = 0;
    {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);      // tags 0: never a valid tag
        uint4* q0 = reinterpret_cast<uint4*>(&sm->kx[0][0][0]);
        for (int q = threadIdx.x; q < M * kHand4 * 32 + kRing4 * MW; q += (M + 2) * 32) q0[q] = z;
    }
    __syncthreads();
    const int tk = sm->ticket;

    const int nx = a.b.nx, nJ = a.b.nJ, steps = a.b.steps, ticks = a.b.ticks, nKg = a.b.nKg;
    const int nCta = nKg * nJ;
    const int tr = BWD ? nCta - 1 - tk : tk;
    const int kg = tr / nJ, J = tr - kg * nJ;
    const int g = kg * nJ + J;
    const int Wg = min(MW, a.b.nz - kg * MW);                  // planes of this stack that exist
    const bool kIn = BWD ? (kg + 1 < nKg) : (kg > 0);
    const bool kOut = BWD ? (kg > 0) : (kg + 1 < nKg);
    const bool jIn = BWD ? (J + 1 < nJ) : (J > 0);
    const bool jOut = BWD ? (J > 0) : (J + 1 < nJ);
    const unsigned int epoch = a.epoch;
    constexpr int kSpan = BLK ? M : MW;                        // ticks a stack spans: one per plane, blocked one per warp
    const int nRho = nx + kSpan - 1;                           // rows of a group's gJ block
    auto sigma_of = [&](int s_) { return BWD ? ticks - 1 - s_ : s_; };
    auto k_row = [&](int s_) -> int {
        const int sg = sigma_of(s_);
        const int t = BWD ? sg - (kSpan - 1) : sg;
        return (kIn && t >= 0 && t < steps) ? t : -1;
    };
    auto j_row = [&](int s_) -> int {
        const int sg = sigma_of(s_);
        const int rho = BWD ? sg - 31 : sg;
        return (jIn && rho >= 0 && rho < nRho) ? rho : -1;
    };

    if (warp >= M) {
        // ------------------------------------------------------------------ helper warps
        // warp M forwards the stack's k-face rows (32 words per tick, for chain warp 0), warp M+1 the column's
        // j-face rows (M*W words per tick, for all warps): up to kPoll rows per L2 round trip each, so that a
        // consumer that runs a few ticks behind its producer is never paced by the polling
        constexpr int kPoll = 8;
        const bool forK = (warp == M);
        if (forK ? !kIn : !jIn) return;
        if (forK) {
            // ---- k-face rows: all 32 words of a row are published by one warp in one instruction
            const LLW* gIn = a.gK + ((long long)(BWD ? g + nJ : g - nJ) * steps) * 32 + lane;
            const unsigned int outA = (unsigned int)__cvta_generic_to_shared(&sm->kx[0][0][lane]);
            int sn = 0;                                          // next loop tick to forward
            long long tstart = 0;
            for (int spin = 0; sn < ticks;) {
                if (sm->abort) break;
                const int cap = min(ticks, sm->prog[0] + kHand4 - 1);
                while (sn < cap && k_row(sn) < 0) sn++;
                LLW w[kPoll];
                int r[kPoll];
#pragma unroll
                for (int i = 0; i < kPoll; i++) {
                    r[i] = (sn + i < cap) ? k_row(sn + i) : -1;
                    if (r[i] >= 0) g_peek(gIn + (long long)r[i] * 32, w[i]);
                }
                bool did = false;
#pragma unroll
                for (int i = 0; i < kPoll; i++) {       // rows become valid in order
                    if (r[i] < 0 || !__all_sync(0xffffffffu, ok(w[i], epoch))) break;
                    s_store_a(outA + (unsigned int)(sn & (kHand4 - 1)) * 512u, w[i].lo, w[i].hi, (unsigned int)sn + 1u);
                    sn++;
                    did = true;
                }
                if (did) {
                    spin = 0;
                    tstart = 0;
                } else if ((++spin & 63) == 63) {
                    if (tstart == 0) tstart = clock64();
                    else if (clock64() - tstart > kTimeout2) {
                        sm->abort = 1;
                        a.S->commError = 2;
                        a.S->done = 1;
                        break;
                    }
                }
            }
            return;
        }
        // ---- j-face words: lane p forwards the words of plane p on its own.  The warps of the producing chain
        // publish their words of a row at different times (warp q runs q ticks behind warp 0) and the warps of
        // this chain need them at different times: a row-wise hand-over would make warp 0 wait for warp M-1
        {
            const int p = lane < MW ? lane : 0;
            const bool mine = lane < MW;
            const LLW* gIn = a.gJ + ((long long)(BWD ? g + 1 : g - 1) * nRho) * MW + p;
            const unsigned int outA = (unsigned int)__cvta_generic_to_shared(&sm->hj[0][p]);
            const int qOfPlane = BWD ? M - 1 - p / W : p / W;       // chain position of the warp that consumes plane p
            int sn = mine ? 0 : ticks;
            long long tstart = 0;
            for (int spin = 0;;) {
                if (__all_sync(0xffffffffu, sn >= ticks) || sm->abort) break;
                const int cap = min(ticks, sm->prog[qOfPlane] + kRing4 - 1);
                while (sn < cap && j_row(sn) < 0) sn++;
                LLW w[kPoll];
                int r[kPoll];
#pragma unroll
                for (int i = 0; i < kPoll; i++) {
                    r[i] = (sn + i < cap) ? j_row(sn + i) : -1;
                    if (r[i] >= 0) g_peek(gIn + (long long)r[i] * MW, w[i]);
                }
                bool did = false;
#pragma unroll
                for (int i = 0; i < kPoll; i++) {       // the words of a plane become valid in order
                    if (r[i] < 0 || !ok(w[i], epoch)) break;
                    s_store_a(outA + (unsigned int)(sn & (kRing4 - 1)) * (MW * 16u), w[i].lo, w[i].hi, (unsigned int)sn + 1u);
                    sn++;
                    did = true;
                }
                if (__any_sync(0xffffffffu, did)) {
                    spin = 0;
                    tstart = 0;
                } else if ((++spin & 63) == 63) {
                    if (tstart == 0) tstart = clock64();
                    else if (clock64() - tstart > kTimeout2) {
                        sm->abort = 1;
                        a.S->commError = 2;
                        a.S->done = 1;
                        break;
                    }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    // Everything a tick needs is a running pointer or a precomputed range of loop ticks: the tick is a chain of
    // dependent instructions of ONE warp, every instruction saved is 5-10 cycles of the sweep's critical path.
    const int q = BWD ? M - 1 - warp : warp;       // position in the chain: q = 0 takes the stack's k-face
    const int p0 = warp * W;                       // first plane (layout index) of this warp
    const int edgeLane = BWD ? 31 : 0, pubLane = BWD ? 0 : 31;
    const bool last = (q == M - 1);
    constexpr unsigned int kOpBytes = W * 256u, kSlotBytes = 4u * kOpBytes;
    const int dSig = BWD ? -1 : 1;
    const int sig0 = BWD ? ticks - 1 : 0;
    const long long tickStride = (long long)dSig * (MW * 32);   // elements per tick in the stream
    const long long elemBase = (long long)g * ticks * MW * 32 + (long long)p0 * 32 + (long long)sig0 * (MW * 32);
    // operand streams of the NEXT tick to prefetch (advance by tickStride per tick)
    const double* nPk = a.pk + elemBase + lane * 2;
    const double* nPj = a.pj + elemBase + lane * 2;
    const double* nPi = a.pi + elemBase + lane * 2;
    const double* nY = a.Y + elemBase + lane * 2;
    double* pY = a.Y + elemBase + lane;                         // result rows of the current tick
    const unsigned int ring = (unsigned int)__cvta_generic_to_shared(&sm->ops[warp][0][0][0][0]) + (unsigned int)lane * 16u;
    const double* ringP = &sm->ops[warp][0][0][0][lane];
    const unsigned int inA = (unsigned int)__cvta_generic_to_shared(&sm->kx[q][0][lane]);
    const unsigned int outA = (unsigned int)__cvta_generic_to_shared(&sm->kx[q + 1 < M ? q + 1 : q][0][lane]);
    const unsigned int hjA = (unsigned int)__cvta_generic_to_shared(&sm->hj[0][p0]);
    // loop-tick ranges (forward and backward alike: both run s = 0 .. ticks-1)
    //   k input from the helper (q == 0): rows exist for s in [0, steps)       (k_row)
    //   j input: s in [0, jEnd)                                               (j_row)
    //   k output (last warp): step t = s - (MW-1) (fwd) / sigma (bwd) in [0, steps)
    //   j output: rho = sigma - 31 (fwd) / sigma (bwd) in [0, nRho)
    //   Y row of plane p: t = sigma - (p0+p) in [0, steps)
    const int kInEnd = (q == 0) ? (kIn ? steps : 0) : ticks;    // q > 0: the word of warp q-1 (ticks 1..)
    const int jInEnd = jIn ? (BWD ? ticks - 31 : nRho) : 0;
    int yLo[W], yHi[W];
#pragma unroll
    for (int p = 0; p < W; p++) {
        // forward: sigma = s, valid for s in [p0+p, p0+p+steps); backward: sigma = ticks-1-s
        const int off = BLK ? warp : p0 + p;                    // tick offset of the plane in the layout
        const int lo = off, hi = off + steps;                   // sigma range
        yLo[p] = BWD ? ticks - hi : lo;
        yHi[p] = (p0 + p < Wg) ? (BWD ? ticks - lo : hi) : yLo[p];
    }
    // k output of the last warp: forward t = s - (MW-1) >= 0, < steps  -> s in [MW-1, MW-1+steps);
    // backward t = sigma = ticks-1-s < steps -> s in [ticks-steps, ticks)
    const int kOutLo = (last && kOut) ? (BWD ? ticks - steps : kSpan - 1) : 0;
    const int kOutHi = (last && kOut) ? (BWD ? ticks : kSpan - 1 + steps) : 0;
    LLW* gKout = a.gK + ((long long)g * steps) * 32 + lane + (long long)(BWD ? steps - 1 : 0) * 32 - (long long)kOutLo * dSig * 32;
    // j output: forward rho = s - 31 in [0, nRho) -> s in [31, 31+nRho); backward rho = ticks-1-s in [0, nRho)
    const int jOutLo = jOut ? (BWD ? ticks - nRho : 31) : 0;
    const int jOutHi = jOut ? (BWD ? ticks : 31 + nRho) : 0;
    LLW* gJout = a.gJ + ((long long)g * nRho) * MW + p0 + (long long)(BWD ? nRho - 1 : 0) * MW - (long long)jOutLo * dSig * MW;

    unsigned long long* trace = a.trace ? a.trace + 8ull * ((unsigned int)tk * M + q) : nullptr;
    const long long trStartC = clock64();
    if (trace && lane == 0) {
        trace[0] = ((unsigned long long)kg << 40) | ((unsigned long long)J << 20) | (unsigned int)q;
        trace[1] = gtimer();
    }
    bool dead = false;
    unsigned int spins = 0;                  // the time-out / abort test runs every 256th poll: a spinning warp shares
                                             // its sub-partition's issue slots with a working one
    auto spin_fail = [&](long long c0) {     // warp-uniform time-out / abort test of a spin loop
        bool bad = sm->abort != 0 || clock64() - c0 > kTimeout2;
        return __any_sync(0xffffffffu, bad);
    };
    auto issue_next = [&](unsigned int slot) {
        const unsigned int d = ring + slot * kSlotBytes;
        if (W == 1) {      // a row is 256 B: 16 lanes x 16 B
            if (lane < 16) {
                cp_async16(d, nPk);
                cp_async16(d + kOpBytes, nPj);
                cp_async16(d + 2u * kOpBytes, nPi);
                cp_async16(d + 3u * kOpBytes, nY);
            }
        }
#pragma unroll
        for (int c = 0; c < W / 2; c++) {
            cp_async16(d + c * 512u, nPk + c * 64);
            cp_async16(d + kOpBytes + c * 512u, nPj + c * 64);
            cp_async16(d + 2u * kOpBytes + c * 512u, nPi + c * 64);
            cp_async16(d + 3u * kOpBytes + c * 512u, nY + c * 64);
        }
        nPk += tickStride;
        nPj += tickStride;
        nPi += tickStride;
        nY += tickStride;
    };

    double res[W];
#pragma unroll
    for (int p = 0; p < W; p++) res[p] = 0.0;
    int nIssued = 0;
#pragma unroll
    for (int d_ = 0; d_ < kD4; d_++) {
        if (nIssued < ticks) { issue_next((unsigned int)d_); nIssued++; }
        cp_async_commit();
    }

    // The tick loop, compiled once per ROLE of the warp (what feeds it, what it feeds, whether its group has
    // j-faces): a lone warp pays 15-20 cycles for every branch, also for a uniform, loop-invariant one
    // (measured: the loop with everything but its control flow removed cost 500 of 820 cycles per tick).
    //   RIN  0: no k input, 1: from the helper (stack below), 2: from warp q-1
    //   ROUT 0: nothing, 1: hand-off to warp q+1, 2: the stack's k-face to L2
    //   STEADY: the middle segment of the loop, where every range test below is true
    // The operands of a tick are read from the ring at the END of the previous tick (software pipelining: their LDS
    // latency hides behind the next tick's face-word poll, and nothing but the poll stands between the start of a
    // tick and its hand-off): cur* hold the operands of the coming tick, `slot` is the ring slot they came from + 1
    double opk[W], opj[W], opi[W], src[W];
    unsigned int slot = 0;
    auto load_operands = [&]() {      // ring slot `slot` -> registers; refill it with tick `nIssued`; advance
        cp_async_wait<kD4 - 1>();
        __syncwarp();
        const double* o = ringP + slot * (kSlotBytes / 8);
#pragma unroll
        for (int p = 0; p < W; p++) {
            opk[p] = o[p * 32];
            opj[p] = o[W * 32 + p * 32];
            opi[p] = o[2 * W * 32 + p * 32];
            src[p] = o[3 * W * 32 + p * 32];
        }
        __syncwarp();
    };
    load_operands();
    if (nIssued < ticks) { issue_next(slot); nIssued++; }
    cp_async_commit();
    slot = slot + 1u == (unsigned int)kD4 ? 0u : slot + 1u;
    auto run = [&](auto RIN_, auto ROUT_, auto JI_, auto JO_, auto STEADY_, int sBegin, int sEnd) {
        constexpr int RIN = decltype(RIN_)::value, ROUT = decltype(ROUT_)::value;
        constexpr bool JI = decltype(JI_)::value, JO = decltype(JO_)::value, STEADY = decltype(STEADY_)::value;
#pragma unroll 1
        for (int s_ = sBegin; s_ < sEnd; s_++) {
            // ---- k input: from the helper (row of this tick, tag s+1) or from warp q-1 (its tick s-1, tag s)
            double vkin = 0.0;
            if (RIN == 1 ? (STEADY || s_ < kInEnd) : (RIN == 2 && (STEADY || s_ > 0))) {
                const unsigned int slotIn = (unsigned int)((RIN == 1 ? s_ : s_ - 1) & (kHand4 - 1));
                const unsigned int tag = (unsigned int)(RIN == 1 ? s_ + 1 : s_);
                LLW w;
                s_peek_a(inA + slotIn * 512u, w);
                if (!__all_sync(0xffffffffu, ok(w, tag))) {
                    const long long c0 = clock64();
                    while (!dead) {
                        s_peek_a(inA + slotIn * 512u, w);
                        if (__all_sync(0xffffffffu, ok(w, tag))) break;
                        if ((++spins & 255u) == 0u) dead = spin_fail(c0);
                    }
                }
                vkin = val(w);
            }
            // ---- j-face words of this warp's planes
            double ve[W];
#pragma unroll
            for (int p = 0; p < W; p++) ve[p] = 0.0;
            if (JI && (STEADY || s_ < jInEnd)) {
                const unsigned int hs = hjA + (unsigned int)(s_ & (kRing4 - 1)) * (MW * 16u), tag = (unsigned int)s_ + 1u;
                LLW w[W];
                bool good = true;
#pragma unroll
                for (int p = 0; p < W; p++) {
                    s_peek_a(hs + p * 16u, w[p]);
                    good = good && ok(w[p], tag);
                }
                if (!__all_sync(0xffffffffu, good)) {
                    const long long c0 = clock64();
                    while (!dead) {
                        good = true;
#pragma unroll
                        for (int p = 0; p < W; p++) {
                            s_peek_a(hs + p * 16u, w[p]);
                            good = good && ok(w[p], tag);
                        }
                        if (__all_sync(0xffffffffu, good)) break;
                        if ((++spins & 255u) == 0u) dead = spin_fail(c0);
                    }
                }
#pragma unroll
                for (int p = 0; p < W; p++) ve[p] = val(w[p]);
            }
            double vj[W], nr[W];
#pragma unroll
            for (int p = 0; p < W; p++) {
                vj[p] = BWD ? __shfl_down_sync(0xffffffffu, res[p], 1) : __shfl_up_sync(0xffffffffu, res[p], 1);
                if (JI) vj[p] = (lane == edgeLane) ? ve[p] : vj[p];
            }
            if (BLK) {
                // the planes of the warp one after the other, from the plane next to the k input: the k-neighbour
                // of the others is the value just computed
#pragma unroll
                for (int pp = 0; pp < W; pp++) {
                    const int p = BWD ? W - 1 - pp : pp;
                    const double vk = pp == 0 ? vkin : nr[BWD ? (p + 1 < W ? p + 1 : p) : (p > 0 ? p - 1 : p)];
                    if (DIV) {
                        // the recurrence of the DIC / DILU diagonal, rD[u] -= upper*lower / rD[l] in face order
                        // (DICPreconditioner.C:66-84): the operands are the products upper*lower, the values the
                        // diagonal itself.  A neighbour that does not exist has operand 0 and value 0: divide by 1.
                        double acc = __dsub_rn(src[p], __ddiv_rn(opk[p], vk == 0.0 ? 1.0 : vk));
                        acc = __dsub_rn(acc, __ddiv_rn(opj[p], vj[p] == 0.0 ? 1.0 : vj[p]));
                        nr[p] = __dsub_rn(acc, __ddiv_rn(opi[p], res[p] == 0.0 ? 1.0 : res[p]));
                    } else {
                        double acc = __dsub_rn(src[p], __dmul_rn(opk[p], vk));
                        acc = __dsub_rn(acc, __dmul_rn(opj[p], vj[p]));
                        nr[p] = __dsub_rn(acc, __dmul_rn(opi[p], res[p]));
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < W; p++) {
                    const double vk = BWD ? (p < W - 1 ? res[p < W - 1 ? p + 1 : p] : vkin) : (p > 0 ? res[p > 0 ? p - 1 : p] : vkin);
                    double acc = __dsub_rn(src[p], __dmul_rn(opk[p], vk));
                    acc = __dsub_rn(acc, __dmul_rn(opj[p], vj[p]));
                    nr[p] = __dsub_rn(acc, __dmul_rn(opi[p], res[p]));
                }
            }
#pragma unroll
            for (int p = 0; p < W; p++) res[p] = nr[p];
            // ---- hand-off to the next warp of the chain.  Its ring holds kHand4 ticks; every 4th tick make sure
            // the consumer has freed the next four slots (the word of tick X is read by the consumer in ITS tick
            // X+1: overwriting the slots of ticks s-kHand4 .. s-kHand4+3 needs the consumer to have completed tick s-kHand4+4)
            if (ROUT == 1) {
                if ((s_ & 3) == 0 && s_ >= kHand4 - 4 && __any_sync(0xffffffffu, sm->prog[q + 1] < s_ - (kHand4 - 5))) {
                    const long long c0 = clock64();
                    while (!dead) {     // warp votes keep the lanes together (volatile reads may differ per lane)
                        if (__all_sync(0xffffffffu, sm->prog[q + 1] >= s_ - (kHand4 - 5))) break;
                        if ((++spins & 255u) == 0u) dead = spin_fail(c0);
                    }
                }
                const double v = BWD ? res[0] : res[W - 1];
                const unsigned long long bb = (unsigned long long)__double_as_longlong(v);
                s_store_a(outA + (unsigned int)(s_ & (kHand4 - 1)) * 512u, (unsigned int)bb, (unsigned int)(bb >> 32), (unsigned int)s_ + 1u);
            } else if (ROUT == 2) {
                if (STEADY || (s_ >= kOutLo && s_ < kOutHi)) g_store(gKout + (long long)s_ * dSig * 32, BWD ? res[0] : res[W - 1], epoch);
            }
            if (JO) {
                if ((STEADY || (s_ >= jOutLo && s_ < jOutHi)) && lane == pubLane) {
#pragma unroll
                    for (int p = 0; p < W; p++) g_store(gJout + (long long)s_ * dSig * MW + p, res[p], epoch);
                }
            }
#pragma unroll
            for (int p = 0; p < W; p++)
                if (STEADY || (s_ >= yLo[p] && s_ < yHi[p])) pY[p * 32] = res[p];
            pY += tickStride;
            if (lane == 0) sm->prog[q] = s_ + 1;
            // ---- operands of the next tick into registers, its ring slot refilled
            if (STEADY || s_ + 1 < ticks) load_operands();
            if (STEADY || nIssued < ticks) { issue_next(slot); nIssued++; }
            cp_async_commit();
            slot = slot + 1u == (unsigned int)kD4 ? 0u : slot + 1u;
        }
    };
    {
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        using T = std::true_type;
        using F = std::false_type;
        const int rin = (q == 0) ? (kIn ? 1 : 0) : 2;
        const int rout = !last ? 1 : (kOut ? 2 : 0);
        // the steady segment: all planes of the warp inside their step range, all faces inside their row ranges,
        // the prefetch never past the end
        int sLo = 1, sHi = ticks - kD4 - 1;
#pragma unroll
        for (int p = 0; p < W; p++) {
            sLo = max(sLo, yLo[p]);
            sHi = min(sHi, yHi[p]);
        }
        if (rin == 1) sHi = min(sHi, kInEnd);
        if (jIn) sHi = min(sHi, jInEnd);
        if (rout == 2) { sLo = max(sLo, kOutLo); sHi = min(sHi, kOutHi); }
        if (jOut) { sLo = max(sLo, jOutLo); sHi = min(sHi, jOutHi); }
        if (sHi < sLo) sHi = sLo;
        auto three = [&](auto RIN_, auto ROUT_, auto JI_, auto JO_) {
            run(RIN_, ROUT_, JI_, JO_, F(), 0, sLo);
            run(RIN_, ROUT_, JI_, JO_, T(), sLo, sHi);
            run(RIN_, ROUT_, JI_, JO_, F(), sHi, ticks);
        };
        auto byJ = [&](auto RIN_, auto ROUT_) {
            if (jIn) { if (jOut) three(RIN_, ROUT_, T(), T()); else three(RIN_, ROUT_, T(), F()); }
            else { if (jOut) three(RIN_, ROUT_, F(), T()); else three(RIN_, ROUT_, F(), F()); }
        };
        // q == 0 is never the last warp (M >= 2): 5 roles
        if (rin == 0) byJ(I0(), I1());
        else if (rin == 1) byJ(I1(), I1());
        else if (rout == 1) byJ(I2(), I1());
        else if (rout == 2) byJ(I2(), I2());
        else byJ(I2(), I0());
    }
    cp_async_wait<0>();
    if (dead && lane == 0) {
        sm->abort = 1;
        a.S->commError = 2;
        a.S->done = 1;
    }
    if (trace && lane == 0) {
        trace[2] = gtimer();
        trace[7] = (unsigned long long)(clock64() - trStartC);
    }
    // the last warp of the chain of the last CTA re-arms the ticket for the next launch
    if (last && lane == 0) {
        __threadfence();
        const unsigned int doneCtas = atomicAdd(&a.ticket[1], 1u);
        if (doneCtas == (unsigned int)nCta - 1u) {
            a.ticket[0] = 0u;
            a.ticket[1] = 0u;
            __threadfence();
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
        if (t < b.steps) Y[tile_row(b, k, J, t) * 32 + lane] = s[lane][x];
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
        s[lane][x] = (t < b.steps) ? Y[tile_row(b, k, J, t) * 32 + lane] : 0.0;
    }
    __syncthreads();
    for (int r = wid; r < 32; r += 8) {
        const int j = J * 32 + r, i = tb - r + lane;
        if (j < b.ny && i >= 0 && i < b.nx) dst[((long long)k * b.ny + j) * b.nx + i] = s[r][lane];
    }
}

// unpack fused with a dot product: dst = Y (natural order), sum(dst * other) -> epilogue.
// Persistent over the tile blocks so that the partial sums fit the reduction scratch.
template <class Epi>
__global__ void __launch_bounds__(kBlock) unpack2_dot_kernel(Box2 b, int nBlk, int nTileBlocks,
                                                             const double* __restrict__ Y, double* __restrict__ dst,
                                                             const double* __restrict__ other, Epi epi, ReduceCtx rc)
{
    if (rc.S->done) return;
    __shared__ double s[32][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double acc[1] = {0.0};
    for (int tbk = blockIdx.x; tbk < nTileBlocks; tbk += gridDim.x) {
        const int T = tbk / nBlk, tb = (tbk - T * nBlk) * 32;
        const int k = T / b.nJ, J = T - k * b.nJ;
        for (int x = wid; x < 32; x += 8) {
            const int t = tb + x;
            s[lane][x] = (t < b.steps) ? Y[tile_row(b, k, J, t) * 32 + lane] : 0.0;
        }
        __syncthreads();
        for (int r = wid; r < 32; r += 8) {
            const int j = J * 32 + r, i = tb - r + lane;
            if (j < b.ny && i >= 0 && i < b.nx) {
                const long long c = ((long long)k * b.ny + j) * b.nx + i;
                const double v = s[r][lane];
                dst[c] = v;
                acc[0] = __dadd_rn(acc[0], __dmul_rn(v, other[c]));
            }
        }
        __syncthreads();
    }
    reduce_tail<1>(acc, rc, epi);
}

// PCG.C:166-172 in tile-block order: psi += alpha pA; rA -= alpha wA; sum |rA| -> epilogue; and
// Y = rD * rA in tile layout for the next preconditioner application
template <class Epi>
__global__ void __launch_bounds__(kBlock) xr_pack2_kernel(Box2 b, int nBlk, int nTileBlocks, double* __restrict__ psi,
                                                          double* __restrict__ rA, const double* __restrict__ pA,
                                                          const double* __restrict__ wA, const double* __restrict__ rD,
                                                          double* __restrict__ Y, Epi epi, ReduceCtx rc)
{
    if (rc.S->done) return;
    __shared__ double s[32][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double alpha = rc.S->alpha;
    double acc[1] = {0.0};
    for (int tbk = blockIdx.x; tbk < nTileBlocks; tbk += gridDim.x) {
        const int T = tbk / nBlk, tb = (tbk - T * nBlk) * 32;
        const int k = T / b.nJ, J = T - k * b.nJ;
        for (int r = wid; r < 32; r += 8) {
            const int j = J * 32 + r, i = tb - r + lane;
            double v = 0.0;
            if (j < b.ny && i >= 0 && i < b.nx) {
                const long long c = ((long long)k * b.ny + j) * b.nx + i;
                psi[c] = __dadd_rn(psi[c], __dmul_rn(alpha, pA[c]));
                const double rr = __dsub_rn(rA[c], __dmul_rn(alpha, wA[c]));
                rA[c] = rr;
                acc[0] = __dadd_rn(acc[0], fabs(rr));
                v = __dmul_rn(rD[c], rr);
            }
            s[r][lane] = v;
        }
        __syncthreads();
        for (int x = wid; x < 32; x += 8) {
            const int t = tb + x;
            if (t < b.steps) Y[tile_row(b, k, J, t) * 32 + lane] = s[lane][x];
        }
        __syncthreads();
    }
    reduce_tail<1>(acc, rc, epi);
}

// premultiplied coefficients in tile layout: F* for the forward sweep (lower faces of a
// cell, neighbour order k-, j-, i-), B* for the backward one (upper faces, k+, j+, i+)
struct Products {
    double* F[3];
    double* B[3];
};

// Same CTA shape as pack2_kernel: one (k, 32-line column) tile x 32 steps per block, read in natural
// cell order (lanes along i: coalesced), written in tile order (lanes along the line index:
// coalesced) through shared memory.  HALF 0: the backward products B[0..2], HALF 1: the forward F[0..2].
// (One thread per cell writing straight to its tile position spreads every warp's stores over 32
// cache lines: 961 us on 216^3 for what is 0.9 GB of traffic.)
template <int HALF>
__global__ void __launch_bounds__(256) products2_kernel(Box2 b, int nBlk, const int* __restrict__ ownerStart,
                                                         const double* __restrict__ rD,
                                                         const double* __restrict__ coefF,
                                                         const double* __restrict__ coefB, Products P)
{
    __shared__ double s[3][32][33];
    const int T = blockIdx.x / nBlk, tb = (blockIdx.x - T * nBlk) * 32;
    const int k = T / b.nJ, J = T - k * b.nJ;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = wid; r < 32; r += 8) {
        const int j = J * 32 + r, i = tb - r + lane;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (j < b.ny && i >= 0 && i < b.nx) {
            const int c = (k * b.ny + j) * b.nx + i;
            const int hasI = i < b.nx - 1, hasJ = j < b.ny - 1, hasK = k < b.nz - 1;
            const double rr = HALF == 2 ? 0.0 : rD[c];
            if (HALF == 0) {
                // faces this cell owns, in the order +i, +j, +k
                const int os = ownerStart[c];
                v2 = hasI ? __dmul_rn(rr, coefB[os]) : 0.0;
                v1 = hasJ ? __dmul_rn(rr, coefB[os + hasI]) : 0.0;
                v0 = hasK ? __dmul_rn(rr, coefB[os + hasI + hasJ]) : 0.0;
            } else if (HALF == 1) {
                // faces where it is the upper cell: the +k / +j / +i face of the cell below /
                // behind / to the left (those cells have the same i, j flags where it matters)
                v0 = k > 0 ? __dmul_rn(rr, coefF[ownerStart[c - b.nx * b.ny] + hasI + hasJ]) : 0.0;
                v1 = j > 0 ? __dmul_rn(rr, coefF[ownerStart[c - b.nx] + hasI]) : 0.0;
                v2 = i > 0 ? __dmul_rn(rr, coefF[ownerStart[c - 1]]) : 0.0;
            } else {
                // the same faces for the recurrence of the diagonal: upper*lower of the face (rD is not read)
                const int fk = k > 0 ? ownerStart[c - b.nx * b.ny] + hasI + hasJ : 0;
                const int fj = j > 0 ? ownerStart[c - b.nx] + hasI : 0;
                const int fi = i > 0 ? ownerStart[c - 1] : 0;
                v0 = k > 0 ? __dmul_rn(coefF[fk], coefB[fk]) : 0.0;
                v1 = j > 0 ? __dmul_rn(coefF[fj], coefB[fj]) : 0.0;
                v2 = i > 0 ? __dmul_rn(coefF[fi], coefB[fi]) : 0.0;
            }
        }
        s[0][r][lane] = v0;
        s[1][r][lane] = v1;
        s[2][r][lane] = v2;
    }
    __syncthreads();
    double* const o0 = HALF == 0 ? P.B[0] : P.F[0];     // HALF 2 writes the forward arrays too
    double* const o1 = HALF == 0 ? P.B[1] : P.F[1];
    double* const o2 = HALF == 0 ? P.B[2] : P.F[2];
    for (int x = wid; x < 32; x += 8) {
        const int t = tb + x;
        if (t < b.steps) {
            const long long p = tile_row(b, k, J, t) * 32 + lane;
            o0[p] = s[0][lane][x];
            o1[p] = s[1][lane][x];
            o2[p] = s[2][lane][x];
        }
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
    int gen = 2, M4 = 0;
    int W = 16;
    long long padded = 0;
    double* Y = nullptr;
    LLW* gK = nullptr;
    LLW* gJ = nullptr;
    unsigned int* ticket = nullptr;
    unsigned int epoch = 0;
    ProductSlot slot[2];
    long long useClock = 0;
    bool attrSet = false, attrSetRD = false;
    const double* yReadyFor = nullptr;   // Y already holds rD * (this vector), written by xr_pack2_kernel
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
    return (size_t)W * 2 * 32 * sizeof(double) + ((size_t)kRing * 32 + (size_t)W * kRing) * sizeof(LLW)
           + (size_t)W * kD * 4 * 32 * sizeof(double) + sizeof(Shared2) + 64;
}

int pick_W(int nz)
{
    const char* e = getenv("LDU_STENCIL_W");
    if (e) {
        const int w = atoi(e);
        if (w == 2 || w == 3 || w == 4 || w == 6 || w == 8 || w == 15 || w == 16) return w;
    }
    return nz >= 12 ? 6 : 4;
}

// 4 (default): chained register-stacked warps (sweep4_kernel); 2: plane-stacked CTAs (sweep2_kernel);
// 3: one register-stacked warp per group (sweep3_kernel).  Measured on B200, us per sweep on 216^3 / 108^3:
// generation 2: 449 / 183, generation 3: 590-610 / 260-290, generation 4 (8 warps x 2 planes): 371 / 174
int box_generation()
{
    const char* e = getenv("LDU_STENCIL");
    const int v = e ? atoi(e) : 4;
    return (v == 2 || v == 3) ? v : 4;
}

// chained warps (sweep4_kernel): warps per CTA; W = 2 planes per warp
int pick_M4(int nz)
{
    const char* e = getenv("LDU_STENCIL_M");
    if (e) {
        const int m = atoi(e);
        if (m == 2 || m == 4 || m == 8) return m;
    }
    return nz >= 32 ? 8 : nz >= 8 ? 4 : 2;
}

// planes per warp of the register-stacked sweeps: the chain of stack-to-stack hops costs
// nz/W * (W ticks + hop latency), a tick costs ~W * 12 ns: W = 8 is the minimum for 0.5-1 us hops
int pick_W3(int nz)
{
    const char* e = getenv("LDU_STENCIL_W");
    if (e) {
        const int w = atoi(e);
        if (w == 4 || w == 8 || w == 12 || w == 16) return w;
    }
    return nz >= 16 ? 8 : 4;
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
        const int gen = box_generation();
        const bool v3 = gen >= 3;
        s->gen = gen;
        s->M4 = gen == 4 ? pick_M4(b.nz) : 0;
        s->W = gen == 4 ? 2 * s->M4 : v3 ? pick_W3(b.nz) : pick_W(b.nz);
        b.nKg = (b.nz + s->W - 1) / s->W;
        // the two planes of a chain warp run the same step in a tick (blocked layout, see tile_row); LDU_STENCIL_BLK=1:
        // every plane one tick behind the plane below, as in the third generation
        b.blk = 1;
        if (gen == 4) {
            // the unblocked layout is kept (and tested) for chains of 8 warps only
            const char* be = getenv("LDU_STENCIL_BLK");
            b.blk = (be && atoi(be) == 1 && s->M4 == 8) ? 1 : 2;
        }
        b.W3 = v3 ? s->W : 0;
        b.ticks = b.steps + s->W / b.blk - 1;
        s->padded = v3 ? (long long)b.nKg * b.nJ * b.ticks * s->W * 32 : (long long)b.nTiles * b.steps * 32;
        cudaStream_t st = m->ctx->stream;
        LDU_TRY(alloc_padded2((void**)&s->Y, (size_t)s->padded, sizeof(double), st));
        const size_t nK = (size_t)b.nKg * b.nJ * b.steps * 32;
        const size_t nJw = v3 ? (size_t)b.nKg * b.nJ * (b.nx + s->W - 1) * s->W : (size_t)b.nTiles * b.steps;
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
    // same arrays, new values (the next solve): recompute in place; otherwise the least recently used slot
    int pick = s->slot[0].lastUse <= s->slot[1].lastUse ? 0 : 1;
    for (int q = 0; q < 2; q++)
        if (s->slot[q].allocated && s->slot[q].rD == rD && s->slot[q].coefF == coefF && s->slot[q].coefB == coefB) pick = q;
    ProductSlot& ps = s->slot[pick];
    cudaStream_t st = m->ctx->stream;
    if (!ps.allocated) {
        for (int d = 0; d < 3; d++) {
            LDU_TRY(alloc_padded2((void**)&ps.P.F[d], (size_t)s->padded, sizeof(double), st));
            LDU_TRY(alloc_padded2((void**)&ps.P.B[d], (size_t)s->padded, sizeof(double), st));
        }
        ps.allocated = true;
    }
    {
        const int nBlk = (s->b.steps + 31) / 32;
        const int grid = s->b.nTiles * nBlk;
        products2_kernel<0><<<grid, 256, 0, st>>>(s->b, nBlk, m->d_ownerStart, rD, coefF, coefB, ps.P);
        count_launch();
        products2_kernel<1><<<grid, 256, 0, st>>>(s->b, nBlk, m->d_ownerStart, rD, coefF, coefB, ps.P);
        count_launch();
    }
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
        LDU_CUDA(cudaMalloc((void**)&s->trace, ((size_t)grid * 8) * sizeof(unsigned long long)));
    }
    a.trace = tracePath ? s->trace : nullptr;
    if (a.trace) LDU_CUDA(cudaMemsetAsync(a.trace, 0, ((size_t)grid * 8) * sizeof(unsigned long long), st));
    sweep2_kernel<W, false><<<grid, (W + 1) * 32, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    if (a.trace) {   // debug only: dump the forward sweep's per-CTA timeline
        std::vector<unsigned long long> h((size_t)grid * 8);
        LDU_CUDA(cudaMemcpyAsync(h.data(), a.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        LDU_CUDA(cudaStreamSynchronize(st));
        if (FILE* f = fopen(tracePath, "w")) {
            for (int c = 0; c < grid; c++)
                fprintf(f, "%d %d %d %llu %llu %llu %llu %llu %llu %llu %llu\n", c, (int)(h[8 * c] >> 32),
                        (int)(h[8 * c] & 0xffffffffu), h[8 * c + 1], h[8 * c + 2], h[8 * c + 3], h[8 * c + 4], h[8 * c + 5],
                        h[8 * c + 6] & ((1ull << 40) - 1), h[8 * c + 6] >> 40, h[8 * c + 7]);
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

template <int W, int M, bool BLK>
int launch_sweeps4(ldu_matrix* m, State2* s, S2Args& a, const Products& P)
{
    const size_t smem = sizeof(Smem4<W, M>);
    if (!s->attrSet) {
        LDU_CUDA(cudaFuncSetAttribute(sweep4_kernel<W, M, false, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDU_CUDA(cudaFuncSetAttribute(sweep4_kernel<W, M, true, BLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->attrSet = true;
    }
    const int grid = s->b.nKg * s->b.nJ;
    cudaStream_t st = m->ctx->stream;
    const char* tracePath = getenv("LDU_S2_TRACE");
    const size_t nTrace = (size_t)grid * M * 8;
    if (tracePath && !s->trace) LDU_CUDA(cudaMalloc((void**)&s->trace, nTrace * sizeof(unsigned long long)));
    a.trace = tracePath ? s->trace : nullptr;
    if (a.trace) LDU_CUDA(cudaMemsetAsync(a.trace, 0, nTrace * sizeof(unsigned long long), st));
    a.dbg = getenv("LDU_S3_DBG") ? atoi(getenv("LDU_S3_DBG")) : 0;
    const char* only = getenv("LDU_S3_ONLY");     // debug: "fwd" / "bwd" runs one of the two sweeps
    a.epoch = ++s->epoch;
    a.pk = P.F[0];
    a.pj = P.F[1];
    a.pi = P.F[2];
    if (!(only && only[0] == 'b')) sweep4_kernel<W, M, false, BLK><<<grid, (M + 2) * 32, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    if (a.trace) {   // debug only: per-warp timeline of the forward sweep
        // columns: ticket kg J q start_ns end_ns wordsCycles mathCycles operandWaitCycles handoffStoreCycles totalCycles
        std::vector<unsigned long long> h(nTrace);
        LDU_CUDA(cudaMemcpyAsync(h.data(), a.trace, nTrace * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        LDU_CUDA(cudaStreamSynchronize(st));
        if (FILE* f = fopen(tracePath, "w")) {
            for (size_t c = 0; c < (size_t)grid * M; c++)
                fprintf(f, "%d %d %d %d %llu %llu %llu %llu %llu %llu %llu\n", (int)(c / M), (int)(h[8 * c] >> 40),
                        (int)((h[8 * c] >> 20) & 0xfffff), (int)(h[8 * c] & 0xfffff), h[8 * c + 1], h[8 * c + 2],
                        h[8 * c + 3], h[8 * c + 4], h[8 * c + 5], h[8 * c + 6], h[8 * c + 7]);
            fclose(f);
        }
        a.trace = nullptr;
    }
    a.epoch = ++s->epoch;
    a.pk = P.B[0];
    a.pj = P.B[1];
    a.pi = P.B[2];
    if (!(only && only[0] == 'f')) sweep4_kernel<W, M, true, BLK><<<grid, (M + 2) * 32, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

template <int W>
int launch_sweeps3(ldu_matrix* m, State2* s, S2Args& a, const Products& P)
{
    const size_t smem = (size_t)kOpRingBytes + sizeof(Smem3<W>);
    if (!s->attrSet) {
        LDU_CUDA(cudaFuncSetAttribute(sweep3_kernel<W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LDU_CUDA(cudaFuncSetAttribute(sweep3_kernel<W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->attrSet = true;
    }
    const int grid = s->b.nKg * s->b.nJ;
    cudaStream_t st = m->ctx->stream;
    const char* tracePath = getenv("LDU_S2_TRACE");
    if (tracePath && !s->trace) LDU_CUDA(cudaMalloc((void**)&s->trace, ((size_t)grid * 8) * sizeof(unsigned long long)));
    a.trace = tracePath ? s->trace : nullptr;
    a.dbg = getenv("LDU_S3_DBG") ? atoi(getenv("LDU_S3_DBG")) : 0;
    if (a.trace) LDU_CUDA(cudaMemsetAsync(a.trace, 0, ((size_t)grid * 8) * sizeof(unsigned long long), st));
    a.epoch = ++s->epoch;
    a.pk = P.F[0];
    a.pj = P.F[1];
    a.pi = P.F[2];
    const char* only = getenv("LDU_S3_ONLY");     // debug: "fwd" / "bwd" runs one of the two sweeps
    if (!(only && only[0] == 'b')) sweep3_kernel<W, false><<<grid, 64, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    if (a.trace) {   // debug only: per-group timeline of the forward sweep
        // columns: ticket kg J start_ns end_ns kWaits jWaits waitCycles operandWaitCycles totalCycles
        std::vector<unsigned long long> h((size_t)grid * 8);
        LDU_CUDA(cudaMemcpyAsync(h.data(), a.trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        LDU_CUDA(cudaStreamSynchronize(st));
        if (FILE* f = fopen(tracePath, "w")) {
            for (int c = 0; c < grid; c++)
                fprintf(f, "%d %d %d %llu %llu %llu %llu %llu %llu %llu\n", c, (int)(h[8 * c] >> 32),
                        (int)(h[8 * c] & 0xffffffffu), h[8 * c + 1], h[8 * c + 2], h[8 * c + 3] & 0xffffffffull,
                        h[8 * c + 3] >> 32, h[8 * c + 4], h[8 * c + 5], h[8 * c + 6]);
            fclose(f);
        }
        a.trace = nullptr;
    }
    a.epoch = ++s->epoch;
    a.pk = P.B[0];
    a.pj = P.B[1];
    a.pi = P.B[2];
    if (!(only && only[0] == 'f')) sweep3_kernel<W, true><<<grid, 64, smem, st>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

}  // namespace

int stencil_version(const ldu_matrix* m)
{
    if (m->box[0] <= 0 || !flow_enabled()) return 0;
    // 0: generic dataflow sweeps, 1: stencil.cu, 2 / 3 / 4 (default): the three kernel generations of
    // stencil2.cu (box_generation), which share every entry point
    const char* e = getenv("LDU_STENCIL");
    const int ver = e ? atoi(e) : 4;
    return (ver < 0 || ver >= 2) ? 2 : ver;
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
static int apply_core(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, const double* r,
                      double* w, bool init, const double* dotWith)
{
    State2* s;
    LDU_TRY(state2(m, &s));
    Products P;
    LDU_TRY(products_for(m, s, rD, coefF, coefB, &P));
    cudaStream_t st = m->ctx->stream;
    // the last CTA of a sweep re-arms the ticket itself; a sweep that was aborted half-way (exchange
    // time-out) would leave it armed wrongly for the next application
    LDU_CUDA(cudaMemsetAsync(s->ticket, 0, 2 * sizeof(unsigned int), st));
    const int nBlk = (s->b.steps + 31) / 32;
    const int gridT = s->b.nTiles * nBlk;
    if (init && s->yReadyFor == r && r != nullptr) {
        // xr_pack2_kernel has already left rD * r in the tile layout
    } else {
        if (init) pack2_kernel<true><<<gridT, 256, 0, st>>>(s->b, nBlk, r, rD, s->Y, m->d_scalars);
        else pack2_kernel<false><<<gridT, 256, 0, st>>>(s->b, nBlk, w, nullptr, s->Y, m->d_scalars);
        count_launch();
        LDU_CUDA(cudaGetLastError());
    }
    s->yReadyFor = nullptr;
    S2Args a;
    a.dbg = 0;
    a.S = m->d_scalars;
    a.guarded = 1;
    a.b = s->b;
    a.Y = s->Y;
    a.gK = s->gK;
    a.gJ = s->gJ;
    a.ticket = s->ticket;
    if (s->gen == 4 && s->M4 == 8 && s->b.blk == 2) LDU_TRY((launch_sweeps4<2, 8, true>(m, s, a, P)));
    else if (s->gen == 4 && s->M4 == 4) LDU_TRY((launch_sweeps4<2, 4, true>(m, s, a, P)));
    else if (s->gen == 4 && s->M4 == 2) LDU_TRY((launch_sweeps4<2, 2, true>(m, s, a, P)));
    else if (s->gen == 4) LDU_TRY((launch_sweeps4<2, 8, false>(m, s, a, P)));
    else if (s->b.W3 == 16) LDU_TRY(launch_sweeps3<16>(m, s, a, P));
    else if (s->b.W3 == 12) LDU_TRY(launch_sweeps3<12>(m, s, a, P));
    else if (s->b.W3 == 8) LDU_TRY(launch_sweeps3<8>(m, s, a, P));
    else if (s->b.W3 == 4) LDU_TRY(launch_sweeps3<4>(m, s, a, P));
    else if (s->W == 16) LDU_TRY(launch_sweeps<16>(m, s, a, P));
    else if (s->W == 15) LDU_TRY(launch_sweeps<15>(m, s, a, P));
    else if (s->W == 8) LDU_TRY(launch_sweeps<8>(m, s, a, P));
    else if (s->W == 6) LDU_TRY(launch_sweeps<6>(m, s, a, P));
    else if (s->W == 3) LDU_TRY(launch_sweeps<3>(m, s, a, P));
    else if (s->W == 2) LDU_TRY(launch_sweeps<2>(m, s, a, P));
    else LDU_TRY(launch_sweeps<4>(m, s, a, P));
    if (dotWith) {
        unpack2_dot_kernel<<<grid_for(m->ctx, m->nCells), kBlock, 0, st>>>(s->b, nBlk, gridT, s->Y, w, dotWith, EpiWArA(),
                                                                          make_rc(m));
    } else {
        unpack2_kernel<<<gridT, 256, 0, st>>>(s->b, nBlk, s->Y, w, m->d_scalars);
    }
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

namespace {
template <int W, int M>
int launch_rD4(ldu_matrix* m, State2* s, S2Args& a)
{
    const size_t smem = sizeof(Smem4<W, M>);
    if (!s->attrSetRD) {
        LDU_CUDA(cudaFuncSetAttribute(sweep4_kernel<W, M, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        s->attrSetRD = true;
    }
    a.epoch = ++s->epoch;
    sweep4_kernel<W, M, false, true, true><<<s->b.nKg * s->b.nJ, (M + 2) * 32, smem, m->ctx->stream>>>(a);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}
}  // namespace

// is the recurrence of the DIC / DILU diagonal available on the box path?  (fourth generation, blocked layout)
bool stencil2_rD_available(ldu_matrix* m)
{
    if (stencil_version(m) != 2) return false;
    static const bool off = getenv("LDU_STENCIL_RD") && getenv("LDU_STENCIL_RD")[0] == '0';
    if (off) return false;
    State2* s;
    if (state2(m, &s) != LDU_OK) return false;
    return s->gen == 4 && s->b.blk == 2;
}

// rD = diag; rD[u] -= upper[f]*lower[f] / rD[l] for the faces in order (DICPreconditioner.C:66-84 with lower = upper,
// DILUPreconditioner.C:66-85): one forward sweep of the chained warps with a division in place of the product.
// The reciprocal is left to the caller, as in calc_reciprocal_D.
int stencil2_rD(ldu_matrix* m, double* rD, const double* upper, const double* lower)
{
    State2* s;
    LDU_TRY(state2(m, &s));
    cudaStream_t st = m->ctx->stream;
    // the products go into the forward arrays of a product slot; the sweeps' own products are computed after rD
    // (they need it), so the slot is simply marked stale
    ProductSlot& ps = s->slot[s->slot[0].lastUse <= s->slot[1].lastUse ? 0 : 1];
    if (!ps.allocated) {
        for (int d = 0; d < 3; d++) {
            LDU_TRY(alloc_padded2((void**)&ps.P.F[d], (size_t)s->padded, sizeof(double), st));
            LDU_TRY(alloc_padded2((void**)&ps.P.B[d], (size_t)s->padded, sizeof(double), st));
        }
        ps.allocated = true;
    }
    ps.sweepGen = -1;
    ps.rD = nullptr;
    const int nBlk = (s->b.steps + 31) / 32;
    const int gridT = s->b.nTiles * nBlk;
    products2_kernel<2><<<gridT, 256, 0, st>>>(s->b, nBlk, m->d_ownerStart, nullptr, upper, lower, ps.P);
    count_launch();
    pack2_kernel<false><<<gridT, 256, 0, st>>>(s->b, nBlk, m->d_diag, nullptr, s->Y, nullptr);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    LDU_CUDA(cudaMemsetAsync(s->ticket, 0, 2 * sizeof(unsigned int), st));
    s->yReadyFor = nullptr;
    S2Args a;
    a.dbg = 0;
    a.S = m->d_scalars;
    a.guarded = 0;
    a.b = s->b;
    a.Y = s->Y;
    a.gK = s->gK;
    a.gJ = s->gJ;
    a.ticket = s->ticket;
    a.trace = nullptr;
    a.pk = ps.P.F[0];
    a.pj = ps.P.F[1];
    a.pi = ps.P.F[2];
    if (s->M4 == 8) LDU_TRY((launch_rD4<2, 8>(m, s, a)));
    else if (s->M4 == 4) LDU_TRY((launch_rD4<2, 4>(m, s, a)));
    else LDU_TRY((launch_rD4<2, 2>(m, s, a)));
    unpack2_kernel<<<gridT, 256, 0, st>>>(s->b, nBlk, s->Y, rD, nullptr);
    count_launch();
    LDU_CUDA(cudaGetLastError());
    return LDU_OK;
}

int stencil2_apply(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, const double* r,
                   double* w, bool init)
{
    return apply_core(m, rD, coefF, coefB, r, w, init, nullptr);
}

int stencil2_apply_dot(ldu_matrix* m, const double* rD, const double* coefF, const double* coefB, const double* r,
                       double* w, const double* dotWith)
{
    return apply_core(m, rD, coefF, coefB, r, w, true, dotWith);
}

int stencil2_xr_pack(ldu_matrix* m, const double* rD, double* psi, double* rA, const double* pA, const double* wA)
{
    State2* s;
    LDU_TRY(state2(m, &s));
    const int nBlk = (s->b.steps + 31) / 32;
    const int nTileBlocks = s->b.nTiles * nBlk;
    xr_pack2_kernel<<<grid_for(m->ctx, m->nCells), kBlock, 0, m->ctx->stream>>>(s->b, nBlk, nTileBlocks, psi, rA, pA, wA, rD,
                                                                               s->Y, EpiResidual<true>{1}, make_rc(m));
    count_launch();
    LDU_CUDA(cudaGetLastError());
    s->yReadyFor = rA;
    return LDU_OK;
}

void stencil2_invalidate(ldu_matrix* m)
{
    State2* s = reinterpret_cast<State2*>(m->stencil2);
    if (s) s->yReadyFor = nullptr;
}

}  // namespace ldu
