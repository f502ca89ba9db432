// Micro-benchmark: exponentials per clock per SM for the flavours of ex2.approx on sm_100a (f32, f16, f16x2, bf16x2) and
// tanh.approx (f32, f16x2, bf16x2): does a packed MUFU op produce two results per issue slot of the unit?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ex2_bench tools/micro/ex2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512, 1) ex2_kernel(long long* out, int iters, uint32_t seed) {
    uint32_t acc = seed + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            uint32_t x = acc + j, y;
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
            if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
            if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
            if (MODE == 3) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
            if (MODE == 4) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
            if (MODE == 5) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
            if (MODE == 6) {
                unsigned short h, g;
                asm volatile("cvt.u16.u32 %0, %1;" : "=h"(h) : "r"(x));
                asm volatile("ex2.approx.f16 %0, %1;" : "=h"(g) : "h"(h));
                y = g;
            }
            acc ^= y;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[1000] = acc;
}

template <int MODE>
void run(const char* name, int per_op, long long* d) {
    const int iters = 400;
    for (int warps : {4, 8, 16}) {
        ex2_kernel<MODE><<<148, warps * 32>>>(d, iters, 12345u);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double clk = (double)h[0] / iters;
        printf("%-22s warps %2d: %8.1f clk / 64 ops / thread -> %6.2f ops/clk/SM = %6.2f results/clk/SM %s\n", name, warps, clk,
               warps * 32 * 64.0 / clk, per_op * warps * 32 * 64.0 / clk, cudaGetErrorString(e));
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16384);
    cudaMemset(d, 0, 16384);
    run<0>("ex2.approx.ftz.f32", 1, d);
    run<1>("ex2.approx.f16x2", 2, d);
    run<2>("ex2.approx.ftz.bf16x2", 2, d);
    run<6>("ex2.approx.f16", 1, d);
    run<3>("tanh.approx.f32", 1, d);
    run<4>("tanh.approx.f16x2", 2, d);
    run<5>("tanh.approx.bf16x2", 2, d);
    return 0;
}
