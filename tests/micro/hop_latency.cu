// Micro-benchmark (not a test): one-way latency of a {value, epoch} hand-off between
// two warps on different SMs through L2, for several store/load flavours.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_latency hop_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

struct W { unsigned int lo, f0, hi, f1; };

template <int MODE> __device__ __forceinline__ void put(W* p, unsigned v, unsigned e)
{
    if (MODE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v), "r"(e), "r"(v), "r"(e) : "memory");
    if (MODE == 1) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v), "r"(e), "r"(v), "r"(e) : "memory");
    if (MODE == 2) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v), "r"(e), "r"(v), "r"(e) : "memory");
    if (MODE == 3) { asm volatile("st.global.cg.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v), "r"(e) : "memory"); }
}
template <int MODE> __device__ __forceinline__ void get(const W* p, unsigned e)
{
    unsigned lo, f0, hi, f1;
    long long t0 = clock64();
    do {
        if (clock64() - t0 > 2000000000ll) return;   // give up, never hang
        if (MODE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
        if (MODE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
        if (MODE == 2) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(p) : "memory");
        if (MODE == 3) { asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(f0) : "l"(p) : "memory"); f1 = f0; }
    } while (f0 != e || f1 != e);
}

template <int MODE> __global__ void pingpong(W* x, W* y, int iters, long long* cycles)
{
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    if (blockIdx.x == 0) {
        for (int i = 1; i <= iters; i++) { put<MODE>(x, i, i); get<MODE>(y, i); }
        *cycles = clock64() - t0;
    } else if (blockIdx.x == gridDim.x - 1) {
        for (int i = 1; i <= iters; i++) { get<MODE>(x, i); put<MODE>(y, i, i); }
    }
}

// chain: block b waits for block b-1's word then publishes its own (one hop per block)
template <int MODE> __global__ void chain(W* w, int epoch, long long* cycles)
{
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    if (blockIdx.x > 0) get<MODE>(w + blockIdx.x - 1, epoch);
    put<MODE>(w + blockIdx.x, 1, epoch);
    if (blockIdx.x == gridDim.x - 1) *cycles = clock64() - t0;
}

int main()
{
    W *x, *y, *w; long long* c; long long h;
    cudaMalloc(&x, 256); cudaMalloc(&y, 256); cudaMalloc(&c, 8); cudaMalloc(&w, 148 * sizeof(W));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[] = {"volatile v4", "cg v4", "relaxed.gpu v4", "cg v2(8B)"};
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaMemset(x, 0, 256); cudaMemset(y, 0, 256);
            const int iters = 2000;
            if (mode == 0) pingpong<0><<<148, 32>>>(x, y, iters, c);
            if (mode == 1) pingpong<1><<<148, 32>>>(x, y, iters, c);
            if (mode == 2) pingpong<2><<<148, 32>>>(x, y, iters, c);
            if (mode == 3) pingpong<3><<<148, 32>>>(x, y, iters, c);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) { printf("%-16s one-way hop: %7.1f cycles = %6.3f us\n", names[mode], h / (2.0 * iters), h / (2.0 * iters) / (clk * 1e-3)); fflush(stdout); }
        }
    }
    for (int mode = 0; mode < 2; mode++) {
        cudaMemset(w, 0, 148 * sizeof(W));
        if (mode == 0) chain<0><<<148, 32>>>(w, 1, c); else chain<1><<<148, 32>>>(w, 1, c);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("chain of 148 blocks (%s): %lld cycles (incl. launch skew) = %.1f per hop\n", names[mode], h, h / 147.0); fflush(stdout);
    }
    printf("cudaError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
