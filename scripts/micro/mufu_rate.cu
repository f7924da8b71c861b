// Micro-benchmark: MUFU.EX2 throughput, fp32 vs packed f16x2 vs bf16x2 (results per clock per SM).
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float tanhf_(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8]; uint32_t h[8];
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xB800B800u + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2f(a[i]) - 1.0f;
            if (MODE == 1) h[i] = ex2h2(h[i]) ^ 0x80008000u;
            if (MODE == 2) h[i] = ex2b2(h[i]) ^ 0x80008000u;
            if (MODE == 3) a[i] = tanhf_(a[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    for (int threads : {128, 256, 512}) {
        for (int mode = 0; mode < 4; ++mode) {
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            double ops = (double)iters * 8 * threads;  // MUFU lane-ops per SM
            const char* nm[4] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2", "tanh.f32"};
            printf("threads %3d %-10s cycles %lld  lane-ops/clk/SM %.2f  results/clk/SM %.2f\n", threads, nm[mode], c, ops / c, ops / c * (mode == 1 || mode == 2 ? 2 : 1));
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
