// Micro-benchmark: throughput of the non-MUFU instructions of the attention softmax (lane-ops per clock per SM):
// cvt.rn.bf16x2.f32 (F2FP pack), prmt truncation pack, fma.rn.f32x2 (FFMA2), 3-input max (FMNMX3), and F2FP mixed with MUFU.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_rn(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t pack_tr(float lo, float hi) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi))); return r; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm volatile("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
                 "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8]; uint32_t h[8];
    for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0x3F803F80u + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) h[i] = pack_rn(__uint_as_float(h[i]), a[i]);
            if (MODE == 1) h[i] = pack_tr(__uint_as_float(h[i]), a[i]);
            if (MODE == 2) { float2 r = ffma2(make_float2(a[i], __uint_as_float(h[i])), make_float2(0.999f, 0.999f), make_float2(0.001f, 0.001f)); a[i] = r.x; h[i] = __float_as_uint(r.y); }
            if (MODE == 3) a[i] = fmaxf(fmaxf(a[i], __uint_as_float(h[i])), a[(i + 1) & 7]);
            if (MODE == 4) { a[i] = ex2f(a[i]) - 1.0f; h[i] = pack_rn(__uint_as_float(h[i]), a[(i + 3) & 7]); }   // 1 MUFU + 1 F2FP
            if (MODE == 5) { a[i] = ex2f(a[i]) - 1.0f; h[i] = pack_tr(__uint_as_float(h[i]), a[(i + 3) & 7]); }   // 1 MUFU + 1 PRMT
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
    const char* nm[6] = {"cvt.rn.bf16x2.f32", "prmt pack", "fma.rn.f32x2", "max3", "ex2 + cvt.rn.bf16x2", "ex2 + prmt"};
    for (int threads : {256, 512}) {
        for (int mode = 0; mode < 6; ++mode) {
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
            if (mode == 4) k<4><<<148, threads>>>(out, cyc, iters);
            if (mode == 5) k<5><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            double ops = (double)iters * 8 * threads;
            printf("threads %3d %-22s cycles %lld  loop-bodies/clk/SM %.2f\n", threads, nm[mode], c, ops / c);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
