// Micro-benchmark: tcgen05.mma issue / execution rate per SM as a function of N (M = 128, K = 16 per instruction, bf16),
// operand sources (A from shared memory or TMEM, B K-major or MN-major) and tcgen05.commit frequency.
#include <cstdio>
#include <cstdint>
#include "common.cuh"
using namespace skb;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// Whole-warp variants: every lane executes the instruction stream, the MMA itself is predicated on lane 0 inside the asm
// (no divergent region around it, so the compiler needs no ELECT / BRA.U.ANY loop per instruction).
__device__ __forceinline__ void umma_ss_pred(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t is_issuer) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(is_issuer) : "memory");
}
__device__ __forceinline__ void umma_ts_pred(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t is_issuer) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(is_issuer) : "memory");
}
__device__ __forceinline__ void commit_pred(uint32_t bar, uint32_t is_issuer) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(is_issuer) : "memory");
}

// Four K-steps of one tile in ONE asm statement (descriptors advanced inside the asm): does the compiler wrap the block or each
// instruction in its ELECT / BRA.U.ANY sequence?
__device__ __forceinline__ void umma_ts_x4(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t a_step, uint64_t b_step) {
    asm volatile(
        "{\n\t.reg .b32 a1, a2, a3;\n\t.reg .b64 b1, b2, b3;\n\t"
        "add.u32 a1, %1, %4;\n\tadd.u32 a2, a1, %4;\n\tadd.u32 a3, a2, %4;\n\t"
        "add.u64 b1, %2, %5;\n\tadd.u64 b2, b1, %5;\n\tadd.u64 b3, b2, %5;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], b1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], b2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], b3, %3, 1;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(a_step), "l"(b_step) : "memory");
}

// mode 0: A smem (K-major), B K-major.  1: A TMEM, B K-major.  2: A TMEM, B MN-major.
// nd = number of independent accumulators the instruction stream rotates over (1 = every MMA accumulates into the same D).
__global__ void __launch_bounds__(128) k(int N, int mode, int commit_every, int reps, int tmem_cols, int nd, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, bar2 = bar + 8, slot = bar + 16;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
        __syncwarp();
        tmem_alloc(slot, (uint32_t)tmem_cols);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, (mode == 2 || mode == 5) ? 1 : 0);
        const uint32_t idesc_qk = umma_idesc_bf16(128, 64, 0, 0), idesc_pv = umma_idesc_bf16(128, 80, 0, 1);
        const uint32_t tA = tmem + (uint32_t)(tmem_cols - 32);
        const uint64_t ad = umma_desc(sA, 16, 1024, UMMA_SW128);
        const uint64_t bdk = umma_desc(sB, 16, 1024, UMMA_SW128);
        const uint64_t bdm = umma_desc(sB, 8192, 1024, UMMA_SW128);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t d = tmem + (uint32_t)((kk & (nd - 1)) * N);
                if (mode == 0) umma_bf16_ss(d, ad + (uint64_t)(kk * 2), bdk + (uint64_t)(kk * 2), idesc, 1);
                else if (mode == 1) umma_ts(d, tA + kk * 8, bdk + (uint64_t)(kk * 2), idesc, 1);
                else if (mode == 2) umma_ts(d, tA + kk * 8, bdm + (uint64_t)(kk * 128), idesc, 1);
                else if (mode == 4) { if (kk == 0) umma_ts_x4(tmem, tA, bdk, idesc, 8u, 2ull); }
                else if (mode == 5) { if (kk == 0) umma_ts_x4(tmem, tA, bdm, idesc, 8u, 128ull); }
                else {  // mode 3: the attention pattern -- a QK chain (N = 64, K-major B) interleaved with a PV chain (N = 80, MN-major B)
                    umma_ts(tmem, tA + kk * 8, bdk + (uint64_t)(kk * 2), idesc_qk, 1);
                    umma_ts(tmem + 128, tA + kk * 8, bdm + (uint64_t)(kk * 128), idesc_pv, 1);
                }
            }
            if (commit_every == 1) umma_commit(bar2);
            if (commit_every == 2) { umma_commit(bar2); umma_commit(bar2); }
        }
        umma_commit(bar);
        long long t1 = clock64();
        mbar_wait(bar, 0);
        long long t2 = clock64();
        out[blockIdx.x * 2] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

__global__ void __launch_bounds__(128) kp(int N, int mode, int reps, int tmem_cols, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
        __syncwarp();
        tmem_alloc(slot, (uint32_t)tmem_cols);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    if (warp == 0) {
        const uint32_t iss = lane == 0 ? 1u : 0u;
        const uint32_t idesc = umma_idesc_bf16(128, N, 0, mode == 2 ? 1 : 0);
        const uint32_t tA = tmem + (uint32_t)(tmem_cols - 32);
        const uint64_t ad = umma_desc(sA, 16, 1024, UMMA_SW128);
        const uint64_t bdk = umma_desc(sB, 16, 1024, UMMA_SW128);
        const uint64_t bdm = umma_desc(sB, 8192, 1024, UMMA_SW128);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (mode == 0) umma_ss_pred(tmem, ad + (uint64_t)(kk * 2), bdk + (uint64_t)(kk * 2), idesc, 1, iss);
                else if (mode == 1) umma_ts_pred(tmem, tA + kk * 8, bdk + (uint64_t)(kk * 2), idesc, 1, iss);
                else umma_ts_pred(tmem, tA + kk * 8, bdm + (uint64_t)(kk * 128), idesc, 1, iss);
            }
        }
        commit_pred(bar, iss);
        long long t1 = clock64();
        mbar_wait(bar, 0);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

int main() {
    long long* out;
    cudaMalloc(&out, 296 * 16);
    const int smem = 16384 + 32768 + 1024 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2048;
    const char* names[6] = {"A smem, B K-major ", "A TMEM, B K-major ", "A TMEM, B MN-major", "QK(64)+PV(80) pair", "TS K-major, 4 per asm", "TS MN-major, 4 per asm"};
    for (int ctas : {1, 2, 4})
        for (int mode = 0; mode < 6; ++mode)
            for (int N : {64, 128, 256})
                for (int nd : {1, 2, 4}) {
                    if (mode >= 4 && nd != 1) continue;
                    const int cols = ctas == 1 ? 512 : ctas == 2 ? 256 : 128;
                    if (nd * N > cols - 32) continue;
                    if (mode == 3 && ctas == 4) continue;
                    if (mode == 3 && (N != 64 || nd != 1)) continue;
                    k<<<148 * ctas, 128, smem>>>(N, mode, 0, reps, cols, nd, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long h[2];
                    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                    const double per = mode == 3 ? 8.0 : 4.0;
                    printf("%d CTA/SM %s N=%3d accumulators=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA per CTA (floor N/2 = %d)%s\n", ctas, names[mode], N,
                           nd, (double)h[0] / (per * reps), (double)h[1] / (per * reps), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
    cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {64, 128, 256}) {
            kp<<<148, 128, smem>>>(N, mode, reps, 512, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("whole-warp predicated issue, 1 CTA/SM %s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor N/2 = %d)%s\n", names[mode], N,
                   (double)h[0] / (4.0 * reps), (double)h[1] / (4.0 * reps), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
