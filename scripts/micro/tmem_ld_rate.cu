// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM, 32x32b.x32 shape (4 KB per warp instruction).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NF>
__global__ void __launch_bounds__(512, 1) k(int warps_active, int iters, long long* cyc, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp < warps_active) {
        for (int it = 0; it < iters; ++it) {
            uint32_t v[NF][32];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const uint32_t addr = tmem + (uint32_t)(((it * 32 * NF) + f * 32 + (warp >> 2) * 64) & 511 & ~31);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[f][0]), "=r"(v[f][1]), "=r"(v[f][2]), "=r"(v[f][3]), "=r"(v[f][4]), "=r"(v[f][5]), "=r"(v[f][6]), "=r"(v[f][7]),
                      "=r"(v[f][8]), "=r"(v[f][9]), "=r"(v[f][10]), "=r"(v[f][11]), "=r"(v[f][12]), "=r"(v[f][13]), "=r"(v[f][14]), "=r"(v[f][15]),
                      "=r"(v[f][16]), "=r"(v[f][17]), "=r"(v[f][18]), "=r"(v[f][19]), "=r"(v[f][20]), "=r"(v[f][21]), "=r"(v[f][22]), "=r"(v[f][23]),
                      "=r"(v[f][24]), "=r"(v[f][25]), "=r"(v[f][26]), "=r"(v[f][27]), "=r"(v[f][28]), "=r"(v[f][29]), "=r"(v[f][30]), "=r"(v[f][31])
                    : "r"(addr) : "memory");
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int f = 0; f < NF; ++f)
#pragma unroll
                for (int i = 0; i < 32; ++i) acc ^= v[f][i];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}
int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 4096;
    for (int nf : {1, 2, 3})
        for (int w : {1, 4, 8, 16}) {
            if (nf == 1) k<1><<<148, 512>>>(w, iters, cyc, sink);
            else if (nf == 2) k<2><<<148, 512>>>(w, iters, cyc, sink);
            else k<3><<<148, 512>>>(w, iters, cyc, sink);
            cudaDeviceSynchronize();
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%d loads in flight, %2d warps issuing: %lld cycles, %.1f B/clk/SM, %.1f cycles per 4 KB warp-load (%s)\n", nf, w, c,
                   (double)iters * w * nf * 4096 / c, (double)c / iters / w / nf, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
