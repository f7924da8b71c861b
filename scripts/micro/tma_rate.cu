// Micro-benchmark: TMA tile-load throughput per SM (bytes/clk) for different box shapes / tensor-map ranks.
// One CTA per SM, one producer thread issues boxes into a ring of smem stages (mbarrier completion), a
// consumer thread waits and immediately frees the stage.  No MMA: measures what the TMA engine can deliver.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t d = 0;
    while (!d) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(d) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
constexpr int STAGES = 6;
// mode 0: 2-D map, box {64 bf16, rows}; mode 1: 4-D map {64, tw, th, 1} (rows = tw*th) from an NHWC image with C = 128
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap m, int mode, int rows, int tw, int th, int iters, int span, int nl, long long* cyc) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    const uint32_t bar = base + STAGES * 32768;
    if (threadIdx.x == 0) { for (int s = 0; s < 2 * STAGES; ++s) mbar_init(bar + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const uint32_t bytes = rows * 128;
    long long t0 = clock64();
    if (nl > 1 && threadIdx.x < nl) {
        // nl lanes of one warp issue in lockstep: lane l owns iterations l, l + nl, ... (one instruction = nl boxes)
        for (int i = threadIdx.x; i < iters; i += nl) {
            const int st = i % STAGES; const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(bar + 8 * (STAGES + st), ph ^ 1);
            mbar_expect(bar + 8 * st, bytes);
            const int t = (blockIdx.x * iters + i) % span;
            if (mode == 0) tma2(base + st * 32768, &m, bar + 8 * st, (t & 31) * 64, (t >> 5) * rows);
            else {
                const int nw = 256 / tw, nh = 256 / th;
                tma4(base + st * 32768, &m, bar + 8 * st, (t & 1) * 64, (t / 2 % nw) * tw, (t / 2 / nw % nh) * th, t / 2 / nw / nh);
            }
        }
    } else if (nl <= 1 && threadIdx.x == 0) {
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bar + 8 * (STAGES + st), ph ^ 1);
            mbar_expect(bar + 8 * st, bytes);
            const int t = (blockIdx.x * iters + i) % span;
            if (mode == 0) tma2(base + st * 32768, &m, bar + 8 * st, (t & 31) * 64, (t >> 5) * rows);  // 32 K-chunks x row tiles
            else {
                const int nw = 256 / tw, nh = 256 / th;
                tma4(base + st * 32768, &m, bar + 8 * st, (t & 1) * 64, (t / 2 % nw) * tw, (t / 2 / nw % nh) * th, t / 2 / nw / nh);
            }
            if (++st == STAGES) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) {
            mbar_wait(bar + 8 * st, ph);
            mbar_arrive(bar + 8 * (STAGES + st));
            if (++st == STAGES) { st = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}
int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const size_t elems = (size_t)256 * 1024 * 1024;  // 512 MB of bf16
    __nv_bfloat16* buf; cudaMalloc(&buf, elems * 2); cudaMemset(buf, 0, elems * 2);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * 32768 + 2048);
    const int iters = 2000;
    struct Cfg { int mode, rows, tw, th; const char* name; } cfgs[] = {
        {0, 64, 0, 0, "2-D  64 rows x 128 B"}, {0, 128, 0, 0, "2-D 128 rows x 128 B"}, {0, 256, 0, 0, "2-D 256 rows x 128 B"},
        {1, 128, 32, 4, "4-D 32x4 px x 128 B (pitch 256 B)"}, {1, 128, 16, 8, "4-D 16x8 px x 128 B (pitch 256 B)"},
        {1, 256, 32, 8, "4-D 32x8 px x 128 B (pitch 256 B)"}};
    for (auto& c : cfgs) {
        CUtensorMap m;
        if (c.mode == 0) {  // [K = 64 of 2304][rows]: weight-like, row pitch 4608 B
            cuuint64_t d[2] = {2304, elems / 2304}; cuuint64_t s[1] = {4608}; cuuint32_t b[2] = {64, (cuuint32_t)c.rows}; cuuint32_t e[2] = {1, 1};
            enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {  // NHWC image [N][256][256][128]
            cuuint64_t d[4] = {128, 256, 256, elems / (128 * 256 * 256)}; cuuint64_t s[3] = {256, 256 * 256, 256ull * 256 * 256};
            cuuint32_t b[4] = {64, (cuuint32_t)c.tw, (cuuint32_t)c.th, 1}; cuuint32_t e[4] = {1, 1, 1, 1};
            enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, buf, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        // distinct tiles in the 512 MB buffer: 2-D: 32 K-chunks x (rows available / rows); 4-D: 2 chunks x tiles per image x 32 images
        const int all_tiles = c.mode == 0 ? 32 * (int)((elems / 2304) / c.rows) : 2 * (256 / c.tw) * (256 / c.th) * 32;
        for (int nl : {1, 2, 4})
        for (int span : {64, all_tiles}) {  // 64: every CTA re-reads the same few tiles (L2 hits); all: streaming the buffer from HBM
            if (nl > 1 && span != 64) continue;
            k<<<148, 64, STAGES * 32768 + 2048>>>(m, c.mode, c.rows, c.tw, c.th, iters, span, nl, cyc);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
            const double bpc = (double)iters * c.rows * 128 / mx;
            printf("%-36s lanes %d span %6d: %7.1f cycles/box  %5.1f B/clk/SM  %.2f cycles/row  (%s)\n", c.name, nl, span, (double)mx / iters, bpc, (double)mx / iters / c.rows, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
