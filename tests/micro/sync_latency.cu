// Micro-benchmark (not a test): cost of one "tick" of W warps of one CTA synchronising through
// different primitives on sm_100a, with and without a short dependent FP64 chain per tick.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_latency sync_latency.cu && ./sync_latency
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(unsigned bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    unsigned long long st;
    asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}

// MODE 0 bar.sync named   1 __syncthreads   2 mbarrier (all lanes wait)   3 mbarrier (lane 0 waits + syncwarp)
//      4 mbarrier test_wait spin (lane 0)  5 dataflow chain through volatile shared flags (per-hop latency)
//      6 per-warp flag array: every warp publishes its tick, every warp polls all others (all-to-all flags)
template <int MODE, bool WORK>
__global__ void k(int nTicks, long long* out, double* sink)
{
    __shared__ unsigned long long bar;
    __shared__ volatile int flag[32];
    __shared__ volatile double val[32][2][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) mbar_init(b, W);
    if (threadIdx.x < 32) flag[threadIdx.x] = 0;
    val[w][0][lane] = 1.0;
    val[w][1][lane] = 1.0;
    __syncthreads();
    unsigned parity = 0;
    double acc = 1.0 + lane;
    const long long c0 = clock64();
    for (int t = 1; t <= nTicks; t++) {
        if (WORK) {
            // read the neighbour warp's value of the previous tick, 4 dependent FP64 ops, publish
            const double v = val[(w + W - 1) % W][(t - 1) & 1][lane];
            acc = acc - 0.5 * v;
            acc = acc - 0.25;
            acc = acc - 0.125;
            val[w][t & 1][lane] = acc;
        }
        if (MODE == 0) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
        else if (MODE == 1) __syncthreads();
        else if (MODE == 2) {
            __syncwarp();
            if (lane == 0) mbar_arrive(b);
            mbar_wait(b, parity);
            parity ^= 1;
        } else if (MODE == 3) {
            __syncwarp();
            if (lane == 0) { mbar_arrive(b); mbar_wait(b, parity); }
            __syncwarp();
            parity ^= 1;
        } else if (MODE == 4) {
            __syncwarp();
            if (lane == 0) { mbar_arrive(b); while (!mbar_test(b, parity)) {} }
            __syncwarp();
            parity ^= 1;
        } else if (MODE == 5) {
            // chain: warp w waits for warp w-1 to have finished tick t (warp 0 free-runs)
            if (w > 0) while (flag[w - 1] < t) {}
            __syncwarp();
            if (lane == 0) flag[w] = t;
        } else if (MODE == 6) {
            __syncwarp();
            if (lane == 0) flag[w] = t;
            if (lane < W) while (flag[lane] < t) {}
            __syncwarp();
        }
    }
    const long long c1 = clock64();
    if (threadIdx.x == blockDim.x - 1) out[0] = c1 - c0;
    if (acc == 12345.678) sink[0] = acc;
}

template <int MODE, bool WORK>
void run(const char* name, long long* d_out, double* d_sink)
{
    const int n = 2000;
    printf("%-44s", name);
    for (int W : {2, 4, 8, 9, 16}) {
        k<MODE, WORK><<<1, W * 32>>>(n, d_out, d_sink);
        k<MODE, WORK><<<1, W * 32>>>(n, d_out, d_sink);
        long long h = 0;
        cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("  W=%-2d %6.0f", W, (double)h / n);
    }
    printf("   cycles/tick\n");
}

int main()
{
    long long* d_out;
    double* d_sink;
    cudaMalloc(&d_out, 8);
    cudaMalloc(&d_sink, 8);
    run<0, false>("bar.sync named", d_out, d_sink);
    run<1, false>("__syncthreads", d_out, d_sink);
    run<2, false>("mbarrier, all lanes try_wait", d_out, d_sink);
    run<3, false>("mbarrier, lane 0 try_wait + syncwarp", d_out, d_sink);
    run<4, false>("mbarrier, lane 0 test_wait spin", d_out, d_sink);
    run<5, false>("flag chain w-1 -> w (per tick, pipelined)", d_out, d_sink);
    run<6, false>("all-to-all flags", d_out, d_sink);
    run<0, true>("bar.sync + LDS/4xDADD/STS", d_out, d_sink);
    run<1, true>("__syncthreads + work", d_out, d_sink);
    run<2, true>("mbarrier all lanes + work", d_out, d_sink);
    run<3, true>("mbarrier lane0 + work", d_out, d_sink);
    run<4, true>("mbarrier test_wait + work", d_out, d_sink);
    run<6, true>("all-to-all flags + work", d_out, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
