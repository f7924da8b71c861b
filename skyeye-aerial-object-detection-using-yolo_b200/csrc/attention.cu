// Flash-style multi-head self-attention on tcgen05 (sm_100a), head_dim = 64, bf16 in / bf16 out.
// Replaces nn.MultiheadAttention's core inside TransformerLayer (skyeye/core/models/attention.py:298),
// which materialises the [B*h, N, N] weights (10.5 GB fp32 per image at N = 25600).
//
// One CTA = one (image, head, 128-query tile); it streams 64-key tiles of K and V:
//   S  = Q K^T        tcgen05.mma with A = Q held in TMEM (copied there once), B = K tile (K-major,
//                     SWIZZLE_128B, from TMA); S fp32 in TMEM, two buffers
//   P  = exp2(S*c-m)  128 softmax threads (one query row each): tcgen05.ld -> online softmax with a
//                     lazily updated reference max (O is rescaled only when a row max grows by > 2^8) ->
//                     bf16 P written back INTO the S buffer with tcgen05.st (no shared-memory round trip)
//   O += P V          tcgen05.mma with A = P from TMEM, B = V consumed MN-major straight from its TMA tile
// Shared memory therefore carries only the K/V stream (the smem port is shared by TMA writes and MMA
// operand reads and was the co-bottleneck with P and Q staged there).  The softmax is MUFU-bound
// (64 flop per exponential at head_dim 64); ATT_POLY of every 64 exponentials can be evaluated on the
// FMA pipes instead (Cody-Waite split + degree-3 polynomial, rel. error 7.5e-5 << bf16), in packed
// fp32x2 arithmetic.  The softmax denominator is produced by the PV MMA itself (16 all-ones B columns).  MMAs issued
// by one thread execute in issue order, which is what lets QK(j+2) reuse the S/P buffer of tile j right after PV(j)
// has been issued.
// The in_proj output [B, N, 3C] is read in place: one 3-D tensor map {channel, token, image} serves
// Q, K and V (different channel coordinates), ragged N is TMA zero fill + a -inf mask on the last tile.
// NQ = 1 (short sequences): 2 CTAs are resident per SM (89 KB smem, 256 TMEM columns each) so one CTA's softmax
// overlaps the other's MMAs; warp roles 0 = TMA producer, 1 = MMA issuer, 2..5 = softmax/epilogue.
// NQ = 2 (N >= 4096): ONE CTA per SM owns two query tiles that share the K/V stream; each tile has its own MMA-issuing
// warp (1, 2), softmax warpgroup (4..7, 8..11) and 256 TMEM columns, and the two run unsynchronised.  Timed alone this
// form is 0-5 % slower than two NQ = 1 CTAs (profiles/r1m_attention_pipeline.md, r1q_ncu_hot_kernels.md), but the K/V
// traffic L2 -> shared memory and the TMA writes are halved, and under the board's 1000 W power cap (which holds the SM
// clock near 1.7-1.8 GHz for the whole step) that buys clock: +2.1 % images/s on the full step, measured on the same box
// in one call.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace skb {

constexpr int ATT_BQ = 128, ATT_BKV = 64, ATT_D = 64;
constexpr int ATT_Q_BYTES = ATT_BQ * ATT_D * 2;     // 16 KB
constexpr int ATT_KV_BYTES = ATT_BKV * ATT_D * 2;   // 8 KB
constexpr int ATT_ONES_BYTES = ATT_BKV * 128;       // 8 KB of bf16 1.0: extra B columns that make the PV MMA emit the row sums
constexpr int ATT_ON = ATT_D + 16;                  // PV accumulator width: 64 output dims + 16 copies of sum_k P
constexpr int ATT_POLY_DEFAULT = 16;                // exponentials per 64 evaluated on the FMA pipes (0, 8, 16, 24, 32)

// NQ = query tiles (of 128 rows) per CTA.  NQ = 1: 192 threads, 2 CTAs per SM.  NQ = 2: ONE CTA per SM whose two
// query tiles share the K/V stream in shared memory (half the K/V bytes from L2 and half the TMA writes).  Each query
// tile has its own MMA-issuing warp, softmax warpgroup and 256 TMEM columns; the two run unsynchronised.
template <int NQ>
struct AttCfg {
    static constexpr int KVS = NQ == 1 ? 4 : 8;          // K/V ring stages (16 KB each)
    static constexpr int THREADS = NQ == 1 ? 192 : 384;  // warp 0 TMA, warps 1..NQ MMA, 4 softmax warps per query tile
    static constexpr int SW0 = NQ == 1 ? 2 : 4;          // first softmax warp
    static constexpr int CTAS_PER_SM = NQ == 1 ? 2 : 1;
    static constexpr int SMEM = NQ * ATT_Q_BYTES + 2 * KVS * ATT_KV_BYTES + ATT_ONES_BYTES + 1024 + 512;  // barriers, TMEM slot, PROF stamps
    static constexpr uint32_t TMEM_COLS = 256 * NQ;      // per query tile: S0/P0 [0,64) S1/P1 [64,128) O [128,208) Q [208,240)
};

// Diagnostics (SKB_ATT_PROF=1): cycle sums per role phase, accumulated with atomics by lane 0 of each role warp.
//  0 softmax: wait S   1 softmax: TMEM load   2 softmax: exponentials / max / pack   3 softmax: TMEM store + arrive
//  4 softmax iterations   5 MMA: wait P   6 MMA: issue PV   7 MMA: wait K/V   8 MMA: issue QK   9 MMA iterations
// 10 TMA: wait empty stage   11 TMA iterations   12 hop P-arrive -> MMA warp awake   13 hop S-commit -> softmax awake
__device__ unsigned long long g_att_prof[16];

struct AttnParams {
    int N, heads, C, T;
    float scale_log2;
    __nv_bfloat16* out;
    long out_pitch;
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ float2 fadd2_rm(float2 a, float2 b) {  // round toward -inf
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rm.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// 2^x for x <= ~100 on the FMA / ALU pipes: n = floor(x) from the mantissa of x + 1.5*2^23 (round down),
// f = x - n in [0,1), 2^f by a degree-3 minimax polynomial (max rel. error 7.5e-5), 2^n by adding n to
// the exponent field.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
    x.x = fmaxf(x.x, -126.0f);
    x.y = fmaxf(x.y, -126.0f);
    const float2 y = fadd2_rm(x, make_float2(12582912.0f, 12582912.0f));
    const float2 fl = fadd2(y, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = ffma2(fl, make_float2(-1.0f, -1.0f), x);
    float2 pz = ffma2(f, make_float2(0.0780244768f, 0.0780244768f), make_float2(0.2260672152f, 0.2260672152f));
    pz = ffma2(pz, f, make_float2(0.6958335042f, 0.6958335042f));
    pz = ffma2(pz, f, make_float2(0.9999251962f, 0.9999251962f));
    float2 r;
    r.x = __int_as_float(__float_as_int(pz.x) + (__float_as_int(y.x) << 23));
    r.y = __int_as_float(__float_as_int(pz.y) + (__float_as_int(y.y) << 23));
    return r;
}

template <int ATT_POLY, int NQ, bool PROF = false>
__global__ void __launch_bounds__(AttCfg<NQ>::THREADS, AttCfg<NQ>::CTAS_PER_SM)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
    using Cfg = AttCfg<NQ>;
    constexpr int KVS = Cfg::KVS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sK0 = sQ + NQ * ATT_Q_BYTES;
    const uint32_t sV0 = sK0 + KVS * ATT_KV_BYTES;
    const uint32_t sOnes = sV0 + KVS * ATT_KV_BYTES;
    const uint32_t bar0 = sOnes + ATT_ONES_BYTES;
    auto q_full = [&](int g) { return bar0 + 8u * g; };
    auto q_ready = [&](int g) { return bar0 + 8u * (NQ + g); };
    auto kv_full = [&](int s) { return bar0 + 8u * (2 * NQ + s); };
    auto kv_empty = [&](int s) { return bar0 + 8u * (2 * NQ + KVS + s); };
    auto s_full = [&](int g, int s) { return bar0 + 8u * (2 * NQ + 2 * KVS + 2 * g + s); };
    auto p_full = [&](int g, int s) { return bar0 + 8u * (4 * NQ + 2 * KVS + 2 * g + s); };
    auto o_done = [&](int g) { return bar0 + 8u * (6 * NQ + 2 * KVS + g); };
    // Completion of the LAST PV MMA only.  The epilogue must not wait on o_done: a softmax warp can reach the epilogue while
    // PV(T-2) is still waiting for a slower warp's P (one that took the rescale path late in the sequence); o_done is then two
    // phases behind and a parity wait for phase T-1 falls straight through (parity aliasing) -> O read before the last two
    // tiles were accumulated.  Found by the teacher-forced parity test on real activations (keys peaking in the last tile).
    auto o_final = [&](int g) { return bar0 + 8u * (7 * NQ + 2 * KVS + g); };
    const uint32_t slot = bar0 + 8u * (8 * NQ + 2 * KVS);
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // warp-uniform for the compiler
    const int qt0 = blockIdx.x * NQ, head = blockIdx.y, b = blockIdx.z;
    const int T = p.T;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int g = 0; g < NQ; ++g) {
                mbar_init(q_full(g), 1);
                mbar_init(q_ready(g), 4);
                mbar_init(o_done(g), 1);
                mbar_init(o_final(g), 1);
                for (int s = 0; s < 2; ++s) { mbar_init(s_full(g, s), 1); mbar_init(p_full(g, s), 4); }
            }
            for (int s = 0; s < KVS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), NQ); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot_ptr, 0);

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
        {
            const bool lead = elect_one();
            if (lead) {
                for (int g = 0; g < NQ; ++g) {
                    mbar_expect_tx(q_full(g), ATT_Q_BYTES);
                    tma_load_3d(sQ + g * ATT_Q_BYTES, &tmQ, q_full(g), head * ATT_D, (qt0 + g) * ATT_BQ, b);
                }
            }
            long long pf_tma = 0;
            for (int j = 0; j < T; ++j) {
                const int s = j % KVS;
                const uint32_t u = (uint32_t)(j / KVS);
                long long c0 = 0;
                if (PROF) c0 = clock64();
                mbar_wait(kv_empty(s), (u & 1u) ^ 1u);
                if (PROF) pf_tma += clock64() - c0;
                if (lead) {
                    mbar_expect_tx(kv_full(s), 2 * ATT_KV_BYTES);
                    tma_load_3d(sK0 + s * ATT_KV_BYTES, &tmKV, kv_full(s), p.C + head * ATT_D, j * ATT_BKV, b);
                    tma_load_3d(sV0 + s * ATT_KV_BYTES, &tmKV, kv_full(s), 2 * p.C + head * ATT_D, j * ATT_BKV, b);
                }
            }
            if (PROF && lead) {
                atomicAdd(&g_att_prof[10], (unsigned long long)pf_tma);
                atomicAdd(&g_att_prof[11], (unsigned long long)T);
            }
        }
    } else if (warp <= NQ) {
        // ===================== MMA issuer of query tile g =====================
        const int g = warp - 1;
        const uint32_t tmem = tmem_base + g * 256;
        const uint32_t tmem_O = tmem + 128;
        const uint32_t tmem_Q = tmem + 208;
        constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BQ, ATT_BKV, 0, 0);  // A = Q (TMEM), B = K (K-major)
        constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_ON, 0, 1);   // A = P (TMEM), B = [V | ones] (MN-major)
        long long pf_m[4] = {0, 0, 0, 0}, pf_hop = 0;
        volatile long long* stamp = reinterpret_cast<volatile long long*>(smem_raw + (slot + 16 - raw));  // [g][sb]: P arrive, [2NQ + ..]: S commit
        auto issue_qk = [&](int j) {
            const int s = j % KVS, sb = j & 1;
            long long c0 = 0, c1 = 0;
            if (PROF) c0 = clock64();
            mbar_wait(kv_full(s), (uint32_t)(j / KVS) & 1u);
            if (PROF) c1 = clock64();
            tc_fence_after();
            if (elect_one()) {  // a warp-uniform branch: the MMA operands stay in uniform registers (no per-instruction R2UR loop)
                const uint64_t bd = umma_desc(sK0 + s * ATT_KV_BYTES, 16, 1024, UMMA_SW128);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)  // 16 bf16 of A = 8 TMEM columns
                    umma_bf16_ts(tmem + sb * ATT_BKV, tmem_Q + k * 8, bd + (uint64_t)(k * 2), idesc_qk, k > 0);
                if (PROF) stamp[2 * NQ + 2 * g + sb] = clock64();
                umma_commit(s_full(g, sb));
            }
            __syncwarp();
            if (PROF) { pf_m[2] += c1 - c0; pf_m[3] += clock64() - c1; }
        };
        mbar_wait(q_ready(g), 0);
        tc_fence_after();
        issue_qk(0);
        if (T > 1) issue_qk(1);
        for (int j = 0; j < T; ++j) {
            const int s = j % KVS, sb = j & 1;
            long long c0 = 0, c1 = 0;
            if (PROF) c0 = clock64();
            mbar_wait(p_full(g, sb), (uint32_t)(j >> 1) & 1u);
            if (PROF) { c1 = clock64(); pf_hop += c1 - stamp[2 * g + sb]; }
            tc_fence_after();
            if (elect_one()) {
                // V tile [64 keys][64 d]: MN-major, 128-byte rows, 8-row groups 1024 B apart, 16 keys per MMA = 2048 B.
                // N = 80: the second 64-wide N atom (leading-dimension offset) is the all-ones tile, so columns
                // 64..79 of the accumulator receive sum_k P[q][k] -- the softmax denominator comes out of the
                // tensor core (and is the sum of exactly the bf16 values that multiply V).
                const uint32_t vaddr = sV0 + s * ATT_KV_BYTES;
                const uint64_t bd = umma_desc(vaddr, sOnes - vaddr, 1024, UMMA_SW128);
#pragma unroll
                for (int k = 0; k < ATT_BKV / 16; ++k)
                    umma_bf16_ts(tmem_O, tmem + sb * ATT_BKV + k * 8, bd + (uint64_t)(k * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(kv_empty(s));
                umma_commit(o_done(g));
                if (j == T - 1) umma_commit(o_final(g));
            }
            __syncwarp();
            if (PROF) { pf_m[0] += c1 - c0; pf_m[1] += clock64() - c1; }
            if (j + 2 < T) issue_qk(j + 2);  // reuses S/P buffer sb: ordered behind PV(j) by in-order MMA execution
        }
        if (PROF && lane == 0) {
            for (int i = 0; i < 4; ++i) atomicAdd(&g_att_prof[5 + i], (unsigned long long)pf_m[i]);
            atomicAdd(&g_att_prof[9], (unsigned long long)T);
            atomicAdd(&g_att_prof[12], (unsigned long long)pf_hop);
        }
    } else if (warp >= Cfg::SW0) {
        // ===================== softmax + epilogue of query tile g: thread <-> query row =====================
        const int g = (warp - Cfg::SW0) >> 2;
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t tmem = tmem_base + g * 256;
        const uint32_t tmem_O = tmem + 128;
        const uint32_t tmem_L = tmem_O + ATT_D;  // row sums (first of 16 identical columns)
        const uint32_t tmem_Q = tmem + 208;
        const int qt = qt0 + g;
        {   // Q tile (K-major SWIZZLE_128B rows in smem) -> TMEM columns [208,240): A operand of every QK MMA
            mbar_wait(q_full(g), 0);
            uint32_t qr[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 u = lds128(sQ + g * ATT_Q_BYTES + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4));
                qr[4 * c + 0] = u.x; qr[4 * c + 1] = u.y; qr[4 * c + 2] = u.z; qr[4 * c + 3] = u.w;
            }
            tmem_st32(tmem_Q + lane_addr, qr);
            {   // all-ones B tile (bf16 1.0 = 0x3F80): 128 threads x 64 bytes (every query tile's group writes the same values)
                const uint4 one4 = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
#pragma unroll
                for (int c = 0; c < 4; ++c) sts128(sOnes + (uint32_t)row * 64u + (uint32_t)(c << 4), one4);
                fence_proxy_async_smem();
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(q_ready(g));
        }
        float m_ref = -INFINITY;
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        long long pf_s[4] = {0, 0, 0, 0}, pf_hop = 0;
        volatile long long* stamp = reinterpret_cast<volatile long long*>(smem_raw + (slot + 16 - raw));
        for (int j = 0; j < T; ++j) {
            const int sb = j & 1;
            long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            if (PROF) c0 = clock64();
            mbar_wait(s_full(g, sb), (uint32_t)(j >> 1) & 1u);
            if (PROF) { c1 = clock64(); pf_hop += c1 - stamp[2 * NQ + 2 * g + sb]; }
            tc_fence_after();
            uint32_t sv[64];
            {
                uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
                uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
                tmem_ld32(tmem + lane_addr + sb * ATT_BKV, lo);
                tmem_ld32(tmem + lane_addr + sb * ATT_BKV + 32, hi);
            }
            tmem_ld_wait();
            if (PROF) c2 = clock64();
            // row max of the raw scores (keys beyond N masked on the ragged last tile)
            const int kbase = j * ATT_BKV;
            if (kbase + ATT_BKV > p.N) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (kbase + i >= p.N) sv[i] = 0xff800000u;  // -inf
            }
            // P = exp2(S*c - m_ref), bf16 pairs; every (64 / ATT_POLY)-th pair runs on the FMA pipes
            uint32_t pk[32];
            auto compute_p = [&](float mref) {
                const float2 nm2 = make_float2(-mref, -mref);
#pragma unroll
                for (int i = 0; i < 32; ++i) {  // pair i = keys 2i, 2i+1
                    const float2 x = ffma2(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), sc2, nm2);
                    float2 e;
                    if (ATT_POLY > 0 && ((i + 1) * ATT_POLY / 64) != (i * ATT_POLY / 64)) {
                        e = exp2_poly2(x);
                    } else {
                        e.x = ex2_approx(x.x);
                        e.y = ex2_approx(x.y);
                    }
                    pk[i] = pack_bf16x2(e.x, e.y);
                }
            };
            // Speculate that the reference max does not move (true for all but the first few tiles): the
            // exponentials start right after the TMEM load and the row max is computed alongside them instead
            // of in front of them (it was a ~120-cycle dependent chain ahead of the MUFU-bound phase).
            compute_p(m_ref);
            float mx;
            {   // 8 independent chains instead of one 64-deep dependent FMNMX chain
                float m8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) m8[i] = __uint_as_float(sv[i]);
#pragma unroll
                for (int i = 8; i < 64; ++i) m8[i & 7] = fmaxf(m8[i & 7], __uint_as_float(sv[i]));
                mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            }
            mx *= p.scale_log2;  // log2 domain (scale > 0)
            // lazy reference max: rescale only if some row's max grew by more than 8 (p stays <= 2^8)
            const bool need = mx > m_ref + 8.0f;
            if (__any_sync(0xffffffffu, need)) {  // rare: redo the tile against the new reference
                const float m_new = need ? mx : m_ref;
                if (j > 0) {
                    mbar_wait(o_done(g), (uint32_t)(j - 1) & 1u);  // PV(j-1) complete: O quiescent (PV(j) waits for our P)
                    tc_fence_after();
                    const float alpha = need ? exp2f(m_ref - m_new) : 1.0f;
                    {
                        const uint32_t lsum = tmem_ld1(tmem_L + lane_addr);
                        tmem_ld_wait();
                        tmem_st1(tmem_L + lane_addr, __float_as_uint(__uint_as_float(lsum) * alpha));
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t o[32];
                        tmem_ld32(tmem_O + lane_addr + h * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st32(tmem_O + lane_addr + h * 32, o);
                    }
                    tmem_st_wait();
                }
                m_ref = m_new;
                compute_p(m_ref);
            }
            // P (bf16, K-major: lane = query row, column c holds keys 2c, 2c+1) over the first half of S
            if (PROF) c3 = clock64();
            tmem_st32(tmem + lane_addr + sb * ATT_BKV, pk);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (PROF && lane == 0) stamp[2 * g + sb] = clock64();  // last writer = last arriving warp (approximately)
            if (lane == 0) mbar_arrive(p_full(g, sb));
            if (PROF) { pf_s[0] += c1 - c0; pf_s[1] += c2 - c1; pf_s[2] += c3 - c2; pf_s[3] += clock64() - c3; }
        }
        if (PROF && lane == 0) {
            for (int i = 0; i < 4; ++i) atomicAdd(&g_att_prof[i], (unsigned long long)pf_s[i]);
            atomicAdd(&g_att_prof[4], (unsigned long long)T);
            atomicAdd(&g_att_prof[13], (unsigned long long)pf_hop);
        }
        // ---- epilogue: O / l -> bf16 ----
        mbar_wait(o_final(g), 0);
        tc_fence_after();
        const uint32_t lsum = tmem_ld1(tmem_L + lane_addr);
        tmem_ld_wait();
        const float inv = 1.0f / __uint_as_float(lsum);
        const int token = qt * ATT_BQ + row;
        __nv_bfloat16* dst = p.out + ((long)b * p.N + token) * p.out_pitch + head * ATT_D;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t o[32];
            tmem_ld32(tmem_O + lane_addr + h * 32, o);
            tmem_ld_wait();
            if (token < p.N) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
                    u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
                    u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
                    u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + h * 32 + g * 8) = u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}


}  // namespace skb

using namespace skb;

extern "C" int skb_flash_attn_bf16(const skb_view* qkv, const skb_view* o, int32_t heads, float scale, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(qkv && o && qkv->ptr && o->ptr && qkv->dtype == SKB_BF16 && o->dtype == SKB_BF16, SKB_ERR_ARG, "flash_attn: bad views");
    const int C = o->c;
    SKB_REQUIRE(qkv->c == 3 * C && heads >= 1 && C == heads * ATT_D, SKB_ERR_UNSUPPORTED,
                "flash_attn: head_dim must be 64 (C=%d heads=%d qkv channels=%d)", C, heads, qkv->c);
    SKB_REQUIRE(qkv->n == o->n && qkv->h == o->h && qkv->w == o->w, SKB_ERR_ARG, "flash_attn: shape mismatch");
    SKB_REQUIRE(qkv->pitch % 8 == 0 && o->pitch % 8 == 0 && ((uintptr_t)qkv->ptr & 15) == 0 && ((uintptr_t)o->ptr & 15) == 0, SKB_ERR_ARG,
                "flash_attn: alignment");
    const int B = qkv->n, N = qkv->h * qkv->w;
    CUtensorMap tmQ, tmKV;
    uint64_t dims[3] = {(uint64_t)qkv->c, (uint64_t)N, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)qkv->pitch * 2, (uint64_t)qkv->pitch * 2 * N};
    uint32_t boxq[3] = {ATT_D, ATT_BQ, 1}, boxk[3] = {ATT_D, ATT_BKV, 1};
    rc = encode_tensor_map(&tmQ, qkv->ptr, 2, 3, dims, str, boxq, 128);
    if (rc != SKB_OK) return rc;
    rc = encode_tensor_map(&tmKV, qkv->ptr, 2, 3, dims, str, boxk, 128);
    if (rc != SKB_OK) return rc;
    AttnParams p;
    p.N = N; p.heads = heads; p.C = C; p.T = (N + ATT_BKV - 1) / ATT_BKV;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out = (__nv_bfloat16*)o->ptr; p.out_pitch = o->pitch;
    static int poly = -1, nq_force = 0;
    if (poly < 0) {  // tuning knobs (not part of the ABI): SKB_ATT_POLY in {0, 8, 16, 24, 32}, SKB_ATT_NQ in {1, 2}
        const char* e = getenv("SKB_ATT_POLY");
        int pv = e ? atoi(e) : ATT_POLY_DEFAULT;
        if (pv != 0 && pv != 8 && pv != 16 && pv != 24 && pv != 32) pv = ATT_POLY_DEFAULT;
        e = getenv("SKB_ATT_NQ");
        nq_force = e ? atoi(e) : 0;
        poly = pv;
    }
    static PerDeviceOnce attr_once;
    if (attr_once.first()) {  // the shared-memory opt-in is a per-device function attribute
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<24, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<24, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<2>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<2>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<2>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<2>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<2>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<8, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<1>::SMEM));
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttCfg<2>::SMEM));
    }
    static int prof = -1;
    if (prof < 0) {
        const char* e = getenv("SKB_ATT_PROF");
        prof = e ? atoi(e) : 0;
    }
    // two query tiles per CTA once the K/V stream is long enough to matter (and 256-row tiles waste little of N)
    const int nq = nq_force == 1 || nq_force == 2 ? nq_force : (N >= 4096 ? 2 : 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (prof) {
        if (nq == 2) flash_attn_kernel<8, 2, true><<<dim3((N + 2 * ATT_BQ - 1) / (2 * ATT_BQ), heads, B), AttCfg<2>::THREADS, AttCfg<2>::SMEM, st>>>(tmQ, tmKV, p);
        else flash_attn_kernel<8, 1, true><<<dim3((N + ATT_BQ - 1) / ATT_BQ, heads, B), AttCfg<1>::THREADS, AttCfg<1>::SMEM, st>>>(tmQ, tmKV, p);
    } else if (nq == 2) {
        dim3 grid((N + 2 * ATT_BQ - 1) / (2 * ATT_BQ), heads, B);
        switch (poly) {
            case 0: flash_attn_kernel<0, 2><<<grid, AttCfg<2>::THREADS, AttCfg<2>::SMEM, st>>>(tmQ, tmKV, p); break;
            case 16: flash_attn_kernel<16, 2><<<grid, AttCfg<2>::THREADS, AttCfg<2>::SMEM, st>>>(tmQ, tmKV, p); break;
            case 24: flash_attn_kernel<24, 2><<<grid, AttCfg<2>::THREADS, AttCfg<2>::SMEM, st>>>(tmQ, tmKV, p); break;
            case 32: flash_attn_kernel<32, 2><<<grid, AttCfg<2>::THREADS, AttCfg<2>::SMEM, st>>>(tmQ, tmKV, p); break;
            default: flash_attn_kernel<8, 2><<<grid, AttCfg<2>::THREADS, AttCfg<2>::SMEM, st>>>(tmQ, tmKV, p); break;
        }
    } else {
        dim3 grid((N + ATT_BQ - 1) / ATT_BQ, heads, B);
        switch (poly) {
            case 0: flash_attn_kernel<0, 1><<<grid, AttCfg<1>::THREADS, AttCfg<1>::SMEM, st>>>(tmQ, tmKV, p); break;
            case 16: flash_attn_kernel<16, 1><<<grid, AttCfg<1>::THREADS, AttCfg<1>::SMEM, st>>>(tmQ, tmKV, p); break;
            case 24: flash_attn_kernel<24, 1><<<grid, AttCfg<1>::THREADS, AttCfg<1>::SMEM, st>>>(tmQ, tmKV, p); break;
            case 32: flash_attn_kernel<32, 1><<<grid, AttCfg<1>::THREADS, AttCfg<1>::SMEM, st>>>(tmQ, tmKV, p); break;
            default: flash_attn_kernel<8, 1><<<grid, AttCfg<1>::THREADS, AttCfg<1>::SMEM, st>>>(tmQ, tmKV, p); break;
        }
    }
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

// Diagnostics: copies (and optionally clears) the 16 phase counters of the SKB_ATT_PROF=1 kernel variant.
extern "C" int skb_debug_attn_prof(unsigned long long* out16, int32_t reset) {
    if (out16) SKB_CUDA(cudaMemcpyFromSymbol(out16, g_att_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        SKB_CUDA(cudaMemcpyToSymbol(g_att_prof, z, sizeof(z)));
    }
    return SKB_OK;
}

// =============================================================================================
// WindowedSelfAttention core (attention.py:372-395): windows of <= 64 tokens, relative-position bias, optional
// additive window mask.  One CTA per (window, head), one thread per query row; K and V of the window live in
// shared memory as bf16 and every thread walks them with broadcast reads.  4*N^2*d flop per window-head is tiny
// (N <= 64): the kernel is bound by reading qkv and writing o once, so no tensor-core path is needed here.
// Two addressing modes: ws == 0, the reference class's own layout (x [B*nW, w*w, C]: window `win` is a contiguous run of
// N tokens); ws > 0, window partition and reverse folded into the addressing (NOT IN REFERENCE): qkv and
// o are feature maps [B, H, W, .], window `win` of image n covers pixels (wy*ws .. +ws, wx*ws .. +ws) and token t of it
// is pixel (wy*ws + t / ws, wx*ws + t % ws) -- the per-token qkv / proj GEMMs run on the unpartitioned map and no
// partitioned copy ever exists.
// =============================================================================================
namespace skb {

template <int HD>
__global__ void __launch_bounds__(64)
window_attn_kernel(const __nv_bfloat16* __restrict__ qkv, long qpitch, const float* __restrict__ bias, const float* __restrict__ mask,
                   int n_mask, __nv_bfloat16* __restrict__ out, long opitch, int N, int C, float scale, int ws, int Wimg, int wins_x,
                   int wins_per_img) {
    __shared__ __align__(16) __nv_bfloat16 sK[64 * HD];
    __shared__ __align__(16) __nv_bfloat16 sV[64 * HD];
    const int win = blockIdx.x, head = blockIdx.y, t = threadIdx.x;
    // token index of the window -> token (pixel) index of the tensor
    long tok0 = (long)win * N;
    if (ws > 0) {
        const int n = win / wins_per_img, w = win - n * wins_per_img;
        const int wy = w / wins_x, wx = w - wy * wins_x;
        tok0 = ((long)n * (wins_per_img / wins_x) * ws + (long)wy * ws) * Wimg + (long)wx * ws;  // pixel (n, wy*ws, wx*ws)
    }
    auto tok = [&](int r) { return ws > 0 ? tok0 + (long)(r / ws) * Wimg + (r % ws) : tok0 + r; };
    const __nv_bfloat16* base = qkv + head * HD;
    for (int i = t; i < N * (HD / 8); i += 64) {  // 16-byte pieces of the K and V rows of this head
        const int r = i / (HD / 8), c = i - r * (HD / 8);
        const __nv_bfloat16* rowp = base + tok(r) * qpitch;
        reinterpret_cast<uint4*>(sK)[i] = *reinterpret_cast<const uint4*>(rowp + C + c * 8);
        reinterpret_cast<uint4*>(sV)[i] = *reinterpret_cast<const uint4*>(rowp + 2 * C + c * 8);
    }
    __syncthreads();
    if (t >= N) return;
    base += tok(t) * qpitch;  // this thread's query row
    float q[HD];
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + c * 8);
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { q[c * 8 + 2 * k] = bf16_lo(w4[k]) * scale; q[c * 8 + 2 * k + 1] = bf16_hi(w4[k]) * scale; }
    }
    const float* brow = bias + ((long)head * N + t) * N;
    const float* mrow = mask ? mask + ((long)((ws > 0 ? win % wins_per_img : win) % n_mask) * N + t) * N : nullptr;
    float s[64];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        float a = -INFINITY;
        if (j < N) {
            a = 0.f;
            const uint4* kr = reinterpret_cast<const uint4*>(sK + j * HD);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                const uint4 u = kr[c];
                a += q[c * 8 + 0] * bf16_lo(u.x) + q[c * 8 + 1] * bf16_hi(u.x) + q[c * 8 + 2] * bf16_lo(u.y) + q[c * 8 + 3] * bf16_hi(u.y) +
                     q[c * 8 + 4] * bf16_lo(u.z) + q[c * 8 + 5] * bf16_hi(u.z) + q[c * 8 + 6] * bf16_lo(u.w) + q[c * 8 + 7] * bf16_hi(u.w);
            }
            a += brow[j];
            if (mrow) a += mrow[j];
        }
        s[j] = a;
        mx = fmaxf(mx, a);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        s[j] = j < N ? expf(s[j] - mx) : 0.f;
        l += s[j];
    }
    float o[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = 0.f;
#pragma unroll 4
    for (int j = 0; j < 64; ++j) {
        if (j < N) {
            const float pj = s[j];
            const uint4* vr = reinterpret_cast<const uint4*>(sV + j * HD);
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
                const uint4 u = vr[c];
                o[c * 8 + 0] += pj * bf16_lo(u.x); o[c * 8 + 1] += pj * bf16_hi(u.x); o[c * 8 + 2] += pj * bf16_lo(u.y); o[c * 8 + 3] += pj * bf16_hi(u.y);
                o[c * 8 + 4] += pj * bf16_lo(u.z); o[c * 8 + 5] += pj * bf16_hi(u.z); o[c * 8 + 6] += pj * bf16_lo(u.w); o[c * 8 + 7] += pj * bf16_hi(u.w);
            }
        }
    }
    const float inv = 1.0f / l;
    __nv_bfloat16* dst = out + tok(t) * opitch + head * HD;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
        uint4 u;
        u.x = pack_bf16x2(o[c * 8 + 0] * inv, o[c * 8 + 1] * inv); u.y = pack_bf16x2(o[c * 8 + 2] * inv, o[c * 8 + 3] * inv);
        u.z = pack_bf16x2(o[c * 8 + 4] * inv, o[c * 8 + 5] * inv); u.w = pack_bf16x2(o[c * 8 + 6] * inv, o[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c * 8) = u;
    }
}

}  // namespace skb

// =============================================================================================
// Tensor-core form of the same core for the shape the detector uses (64-token windows, head_dim 64, no mask): the CUDA-core
// kernel above runs ~14 TFLOP/s (1.95 ms for the P3 level of skyeye_lw at B16) where reading qkv and writing o once costs
// 0.13 ms.  One persistent CTA per SM owns ONE head (its relative-position bias, pre-multiplied by log2 e, stays in shared
// memory) and walks pairs of windows: the Q, K and V head slices of two windows are six 8 KB TMA boxes ({64 ch, 8 x, 8 y}
// out of the unpartitioned map, or 64 consecutive tokens of the class layout) stacked into three 128-row operand tiles.
//   S = Q K^T   one M128 x N128 x K64 tcgen05.mma (SS form); only the two diagonal 64 x 64 blocks are meaningful
//   P           thread <-> query row: tcgen05.ld of its window's 64 scores, s * scale + bias, exact softmax (the whole row is
//               here: no online rescaling), bf16 P written over S with tcgen05.st -- zeros in the other window's key columns
//   O = P V     one M128 x N64 x K128 tcgen05.mma (A = P from TMEM, B = both windows' V, MN-major)
// Half of the tensor work multiplies zeros: 576 MMA cycles per pair against ~2400 cycles of HBM time for its 64 KB, so the
// kernel stays HBM-bound.  Two softmax warpgroups alternate pairs (each owns one S/P and one O buffer in TMEM: 384 columns),
// a three-stage TMA ring (144 KB) keeps ~100 KB per SM in flight.
// =============================================================================================
namespace skb {

constexpr int WA_STAGES = 3;
constexpr int WA_TILE = 64 * 128;            // one window's Q, K or V slice of one head: 64 tokens x 128 B
constexpr int WA_STAGE = 6 * WA_TILE;        // Q[2 windows] K[2] V[2]
constexpr int WA_THREADS = 320;              // warp 0 TMA, warp 1 MMA, warps 2..5 / 6..9 softmax + epilogue groups
constexpr int WA_SMEM = WA_STAGES * WA_STAGE + 64 * 64 * 4 + 1024 + 256;

struct WinAttnParams {
    int n_windows, n_pairs, heads, C;
    int mode2d, ws, Wimg, Himg, wins_x, wins_per_img;
    float scale_log2;
    const float* bias;   // [heads][64][64]
    __nv_bfloat16* out;
    long opitch;
};

__global__ void __launch_bounds__(WA_THREADS, 1)
window_attn_tc_kernel(const __grid_constant__ CUtensorMap tm, const WinAttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sBias = base + WA_STAGES * WA_STAGE;   // fp32 [16 key quads][64 tokens][4]: conflict-free 16-byte reads
    const uint32_t bar0 = sBias + 64 * 64 * 4;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (WA_STAGES + s); };
    auto s_full = [&](int b) { return bar0 + 8u * (2 * WA_STAGES + b); };
    auto p_full = [&](int b) { return bar0 + 8u * (2 * WA_STAGES + 2 + b); };
    auto o_full = [&](int b) { return bar0 + 8u * (2 * WA_STAGES + 4 + b); };
    const uint32_t slot = bar0 + 8u * (2 * WA_STAGES + 6);
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int head = blockIdx.x % p.heads, slot_id = blockIdx.x / p.heads, n_slots = gridDim.x / p.heads;
    const int n_my = slot_id < p.n_pairs ? (p.n_pairs - slot_id + n_slots - 1) / n_slots : 0;  // pairs slot_id, slot_id + n_slots, ...

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm);
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < WA_STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
            for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(p_full(b), 4); mbar_init(o_full(b), 1); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot_ptr, 0);
    // TMEM columns: S/P buffer b at b * 128 (128 fp32 score columns; P = 64 columns of packed bf16 over them), O buffer b at 256 + b * 64

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool lead = elect_one();
        for (int i = 0; i < n_my; ++i) {
            const int s = i % WA_STAGES;
            mbar_wait(empty(s), ((uint32_t)(i / WA_STAGES) & 1u) ^ 1u);
            if (lead) {
                mbar_expect_tx(full(s), WA_STAGE);
                const int pair = slot_id + i * n_slots;
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int win = min(2 * pair + w, p.n_windows - 1);  // an odd tail reloads the last window (its rows are not stored)
                    const uint32_t dst = base + s * WA_STAGE + w * WA_TILE;
                    if (p.mode2d) {
                        const int n = win / p.wins_per_img, r = win - n * p.wins_per_img;
                        const int wy = r / p.wins_x, wx = r - wy * p.wins_x;
#pragma unroll
                        for (int m = 0; m < 3; ++m)
                            tma_load_4d(dst + m * 2 * WA_TILE, &tm, full(s), m * p.C + head * 64, wx * 8, wy * 8, n);
                    } else {
#pragma unroll
                        for (int m = 0; m < 3; ++m) tma_load_3d(dst + m * 2 * WA_TILE, &tm, full(s), m * p.C + head * 64, 0, win);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);  // A = Q (smem, K-major), B = K (smem, K-major)
        constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // A = P (TMEM), B = V (smem, MN-major)
        auto issue_pv = [&](int k) {
            const int b = k & 1, s = k % WA_STAGES;
            mbar_wait(p_full(b), (uint32_t)(k >> 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t vd = umma_desc(base + s * WA_STAGE + 4 * WA_TILE, 16, 1024, UMMA_SW128);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)  // 16 keys per MMA: 8 TMEM columns of A, 2048 B of V
                    umma_bf16_ts(tmem_base + 256 + b * 64, tmem_base + b * 128 + kk * 8, vd + (uint64_t)(kk * 128), idesc_pv, kk > 0);
                umma_commit(empty(s));
                umma_commit(o_full(b));
            }
            __syncwarp();
        };
        for (int i = 0; i < n_my; ++i) {
            const int s = i % WA_STAGES, b = i & 1;
            mbar_wait(full(s), (uint32_t)(i / WA_STAGES) & 1u);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t qd = umma_desc(base + s * WA_STAGE, 16, 1024, UMMA_SW128);
                const uint64_t kd = umma_desc(base + s * WA_STAGE + 2 * WA_TILE, 16, 1024, UMMA_SW128);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_base + b * 128, qd + (uint64_t)(kk * 2), kd + (uint64_t)(kk * 2), idesc_qk, kk > 0);
                umma_commit(s_full(b));
            }
            __syncwarp();
            if (i > 0) issue_pv(i - 1);  // in-order MMA execution: QK(i + 1) overwrites S/P buffer b^1 only after PV(i - 1) has read it
        }
        if (n_my > 0) issue_pv(n_my - 1);
    } else {
        // ===================== softmax + epilogue group grp: pairs i = grp, grp + 2, ... (S/P and O buffer grp) =====================
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                       // TMEM lane quadrant of this warp
        const int row = q * 32 + lane, w = row >> 6, t = row & 63;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t tS = tmem_base + grp * 128 + lane_addr, tO = tmem_base + 256 + grp * 64 + lane_addr;
        {   // bias[head][t][j] * log2(e) -> shared [j / 4][t][j % 4]
            const float* bsrc = p.bias + (long)head * 64 * 64;
            for (int i = threadIdx.x - 64; i < 64 * 64; i += 256) {
                const int tt = i >> 6, j = i & 63;
                sts32f(sBias + (uint32_t)(((j >> 2) * 64 + tt) * 4 + (j & 3)) * 4u, bsrc[i] * 1.4426950408889634f);
            }
            named_bar_sync(1, 256);
        }
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        for (int i = grp; i < n_my; i += 2) {
            const uint32_t ph = (uint32_t)(i >> 1) & 1u;
            mbar_wait(s_full(grp), ph);
            tc_fence_after();
            uint32_t sv[64];
            {
                uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
                uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
                tmem_ld32(tS + w * 64, lo);
                tmem_ld32(tS + w * 64 + 32, hi);
            }
            tmem_ld_wait();
            float x[64];
            float m8[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4) {
                const float4 bv = lds128f(sBias + (uint32_t)(j4 * 64 + t) * 16u);
                const float2 a = ffma2(make_float2(__uint_as_float(sv[4 * j4]), __uint_as_float(sv[4 * j4 + 1])), sc2, make_float2(bv.x, bv.y));
                const float2 c = ffma2(make_float2(__uint_as_float(sv[4 * j4 + 2]), __uint_as_float(sv[4 * j4 + 3])), sc2, make_float2(bv.z, bv.w));
                x[4 * j4] = a.x; x[4 * j4 + 1] = a.y; x[4 * j4 + 2] = c.x; x[4 * j4 + 3] = c.y;
                m8[0] = fmaxf(m8[0], a.x); m8[1] = fmaxf(m8[1], a.y); m8[2] = fmaxf(m8[2], c.x); m8[3] = fmaxf(m8[3], c.y);
            }
            const float mx = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
            uint32_t pk[32];
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float e0 = ex2_approx(x[2 * j] - mx), e1 = ex2_approx(x[2 * j + 1] - mx);
                l4[(2 * j) & 3] += e0;
                l4[(2 * j + 1) & 3] += e1;
                pk[j] = pack_bf16x2(e0, e1);
            }
            const float inv = 1.0f / ((l4[0] + l4[1]) + (l4[2] + l4[3]));
            {   // P: this window's 64 keys are columns [w * 32, w * 32 + 32) of the packed row, the other window's are zeros
                uint32_t z[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) z[j] = 0u;
                tmem_st32(tS + w * 32, pk);
                tmem_st32(tS + (w ^ 1) * 32, z);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(grp));
            // ---- epilogue: O / l -> bf16, 128 contiguous bytes per token ----
            const int pair = slot_id + i * n_slots, win = 2 * pair + w;
            long tok;
            if (p.mode2d) {
                const int n = win / p.wins_per_img, r = win - n * p.wins_per_img;
                const int wy = r / p.wins_x, wx = r - wy * p.wins_x;
                tok = ((long)n * p.Himg + wy * 8 + (t >> 3)) * p.Wimg + wx * 8 + (t & 7);
            } else {
                tok = (long)win * 64 + t;
            }
            __nv_bfloat16* dst = p.out + tok * p.opitch + head * 64;
            mbar_wait(o_full(grp), ph);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t o[32];
                tmem_ld32(tO + h * 32, o);
                tmem_ld_wait();
                if (win < p.n_windows) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(dst + h * 32 + g * 8) = u;
                    }
                }
            }
            tc_fence_before();  // the O loads are complete before the next p_full arrive lets PV(i + 2) overwrite the buffer
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace skb

// Launches the tensor-core window kernel when the shape is the detector's (64 tokens, head_dim 64, no mask); returns 1 if it did.
static int window_attn_tc_try(const skb_view* qkv, const float* bias, const float* mask, const skb_view* o, int heads, int window,
                              float scale, cudaStream_t st, int* launched) {
    *launched = 0;
    const int C = o->c;
    static int use_tc = -1;  // tuning knob (not part of the ABI): SKB_WATT_TC=0 keeps the CUDA-core kernel
    if (use_tc < 0) { const char* e = getenv("SKB_WATT_TC"); use_tc = e ? atoi(e) : 1; }
    const int n_tok = window > 0 ? window * window : qkv->w;
    if (!use_tc || mask || C != heads * 64 || n_tok != 64 || heads > num_sms()) return SKB_OK;
    WinAttnParams p;
    CUtensorMap tm;
    int rc;
    if (window > 0) {
        p.mode2d = 1; p.ws = window; p.Wimg = qkv->w; p.Himg = qkv->h; p.wins_x = qkv->w / window; p.wins_per_img = p.wins_x * (qkv->h / window);
        p.n_windows = qkv->n * p.wins_per_img;
        uint64_t dims[4] = {(uint64_t)qkv->c, (uint64_t)qkv->w, (uint64_t)qkv->h, (uint64_t)qkv->n};
        uint64_t str[3] = {(uint64_t)qkv->pitch * 2, (uint64_t)qkv->pitch * 2 * qkv->w, (uint64_t)qkv->pitch * 2 * qkv->w * qkv->h};
        uint32_t box[4] = {64, 8, 8, 1};
        rc = encode_tensor_map(&tm, qkv->ptr, 2, 4, dims, str, box, 128);
    } else {
        p.mode2d = 0; p.ws = 0; p.Wimg = 0; p.Himg = 0; p.wins_x = 1; p.wins_per_img = 1;
        p.n_windows = qkv->n;
        uint64_t dims[3] = {(uint64_t)qkv->c, 64, (uint64_t)qkv->n};
        uint64_t str[2] = {(uint64_t)qkv->pitch * 2, (uint64_t)qkv->pitch * 2 * 64};
        uint32_t box[3] = {64, 64, 1};
        rc = encode_tensor_map(&tm, qkv->ptr, 2, 3, dims, str, box, 128);
    }
    if (rc != SKB_OK) return rc;
    p.n_pairs = (p.n_windows + 1) / 2; p.heads = heads; p.C = C;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.bias = bias; p.out = (__nv_bfloat16*)o->ptr; p.opitch = o->pitch;
    static PerDeviceOnce once;
    if (once.first()) SKB_CUDA(cudaFuncSetAttribute(window_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WA_SMEM));
    int slots = num_sms() / heads;
    if (slots > p.n_pairs) slots = p.n_pairs;
    window_attn_tc_kernel<<<heads * slots, WA_THREADS, WA_SMEM, st>>>(tm, p);
    SKB_LAUNCH_CHECK();
    *launched = 1;
    return SKB_OK;
}

extern "C" int skb_window_attn_bf16(const skb_view* qkv, const float* bias, const float* mask, int32_t n_mask, const skb_view* o,
                                    int32_t heads, float scale, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(qkv && o && bias && qkv->ptr && o->ptr && qkv->dtype == SKB_BF16 && o->dtype == SKB_BF16, SKB_ERR_ARG, "window_attn: bad views");
    const int C = o->c;
    SKB_REQUIRE(heads >= 1 && qkv->c == 3 * C && C % heads == 0, SKB_ERR_ARG, "window_attn: C=%d heads=%d qkv channels=%d", C, heads, qkv->c);
    const int hd = C / heads;
    SKB_REQUIRE(hd == 16 || hd == 32 || hd == 64, SKB_ERR_UNSUPPORTED, "window_attn: head_dim %d (supported: 16, 32, 64)", hd);
    SKB_REQUIRE(qkv->h == 1 && o->h == 1 && qkv->w == o->w && qkv->n == o->n && qkv->w >= 1 && qkv->w <= 64, SKB_ERR_UNSUPPORTED,
                "window_attn: views must be [windows, 1, tokens <= 64, C] (got tokens %d)", qkv->w);
    SKB_REQUIRE(qkv->pitch % 8 == 0 && o->pitch % 8 == 0 && ((uintptr_t)qkv->ptr & 15) == 0 && ((uintptr_t)o->ptr & 15) == 0, SKB_ERR_ARG,
                "window_attn: alignment");
    SKB_REQUIRE(!mask || n_mask >= 1, SKB_ERR_ARG, "window_attn: mask given with n_mask=%d", n_mask);
    const int N = qkv->w;
    dim3 grid(qkv->n, heads);
    cudaStream_t st = (cudaStream_t)stream;
    {
        int launched = 0;
        rc = window_attn_tc_try(qkv, bias, mask, o, heads, 0, scale, st, &launched);
        if (rc != SKB_OK || launched) return rc;
    }
    const __nv_bfloat16* qp = (const __nv_bfloat16*)qkv->ptr;
    __nv_bfloat16* op = (__nv_bfloat16*)o->ptr;
    if (hd == 64) window_attn_kernel<64><<<grid, 64, 0, st>>>(qp, qkv->pitch, bias, mask, n_mask, op, o->pitch, N, C, scale, 0, 0, 1, 1);
    else if (hd == 32) window_attn_kernel<32><<<grid, 64, 0, st>>>(qp, qkv->pitch, bias, mask, n_mask, op, o->pitch, N, C, scale, 0, 0, 1, 1);
    else window_attn_kernel<16><<<grid, 64, 0, st>>>(qp, qkv->pitch, bias, mask, n_mask, op, o->pitch, N, C, scale, 0, 0, 1, 1);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}

extern "C" int skb_window_attn2d_bf16(const skb_view* qkv, const float* bias, const float* mask, int32_t n_mask, const skb_view* o,
                                      int32_t heads, int32_t window, float scale, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(qkv && o && bias && qkv->ptr && o->ptr && qkv->dtype == SKB_BF16 && o->dtype == SKB_BF16, SKB_ERR_ARG, "window_attn2d: bad views");
    const int C = o->c;
    SKB_REQUIRE(heads >= 1 && qkv->c == 3 * C && C % heads == 0, SKB_ERR_ARG, "window_attn2d: C=%d heads=%d qkv channels=%d", C, heads, qkv->c);
    const int hd = C / heads;
    SKB_REQUIRE(hd == 16 || hd == 32 || hd == 64, SKB_ERR_UNSUPPORTED, "window_attn2d: head_dim %d (supported: 16, 32, 64)", hd);
    SKB_REQUIRE(window >= 1 && window <= 8, SKB_ERR_UNSUPPORTED, "window_attn2d: window %d (supported: 1..8, i.e. <= 64 tokens)", window);
    SKB_REQUIRE(qkv->n == o->n && qkv->h == o->h && qkv->w == o->w && qkv->h % window == 0 && qkv->w % window == 0, SKB_ERR_ARG,
                "window_attn2d: maps must agree and be multiples of the window (%dx%d, window %d)", qkv->h, qkv->w, window);
    SKB_REQUIRE(qkv->pitch % 8 == 0 && o->pitch % 8 == 0 && ((uintptr_t)qkv->ptr & 15) == 0 && ((uintptr_t)o->ptr & 15) == 0, SKB_ERR_ARG,
                "window_attn2d: alignment");
    SKB_REQUIRE(!mask || n_mask >= 1, SKB_ERR_ARG, "window_attn2d: mask given with n_mask=%d", n_mask);
    const int N = window * window, wins_x = qkv->w / window, wpi = wins_x * (qkv->h / window);
    dim3 grid(qkv->n * wpi, heads);
    cudaStream_t st = (cudaStream_t)stream;
    {
        int launched = 0;
        rc = window_attn_tc_try(qkv, bias, mask, o, heads, window, scale, st, &launched);
        if (rc != SKB_OK || launched) return rc;
    }
    const __nv_bfloat16* qp = (const __nv_bfloat16*)qkv->ptr;
    __nv_bfloat16* op = (__nv_bfloat16*)o->ptr;
    if (hd == 64) window_attn_kernel<64><<<grid, 64, 0, st>>>(qp, qkv->pitch, bias, mask, n_mask, op, o->pitch, N, C, scale, window, qkv->w, wins_x, wpi);
    else if (hd == 32) window_attn_kernel<32><<<grid, 64, 0, st>>>(qp, qkv->pitch, bias, mask, n_mask, op, o->pitch, N, C, scale, window, qkv->w, wins_x, wpi);
    else window_attn_kernel<16><<<grid, 64, 0, st>>>(qp, qkv->pitch, bias, mask, n_mask, op, o->pitch, N, C, scale, window, qkv->w, wins_x, wpi);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
