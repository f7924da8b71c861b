// Conv-BN-SiLU as an implicit GEMM on tcgen05 (sm_100a).
//
//   D[M = pixels, N = Cout] = sum over (tap, cin-chunk) A_tap[M, 64] * W[N, tap*Cin + chunk]^T
//
// * A operand: activations stay NHWC bf16 in HBM.  One M tile is a box of tn x th x tw output pixels
//   (tn*th*tw <= 128); for every filter tap the TMA engine fetches the correspondingly shifted input
//   box (5-D tiled tensor map {c, w, parity, h, n}) straight into the canonical K-major SWIZZLE_128B
//   layout -- im2col never exists in memory and the conv zero padding is TMA out-of-bounds fill.
//   Stride-2 convs address the input as [N, H/2, 2, W/2, 2*pitch] so a tap selects a (row-parity,
//   column-parity) plane and the box is dense again.
// * B operand: weights [Cout_pad][taps*Cin] bf16 (BN folded), 2-D tensor map, same swizzle.
// * Accumulators: fp32 in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps
//   the MMAs of tile i+1.  Persistent CTAs, static tile schedule (Cout block fastest).
// * Warp roles: warp 0 = TMA producer (activations), warp 10 = TMA producer (weights), warp 1 = MMA
//   issuer (+TMEM alloc), warps 2..9 = two epilogue
//   groups of 4 warps.  Group g drains TMEM accumulator stage g, i.e. tiles alternate between the
//   groups and two epilogues are in flight while the MMAs of a third tile run.  Each warp owns its
//   TMEM lane quarter (32 pixels) and walks the columns in 32-column pieces:
//   tcgen05.ld -> +bias (LDS.128 broadcast) -> SiLU as h + h*tanh(h) (one MUFU) / ReLU -> +residual ->
//   bf16/fp32 (packed fp32x2 arithmetic), written into the group's ring of two 16 KB swizzled smem
//   sub-tiles (64 bf16 / 32 fp32 channels x 128 pixels) that one elected thread drains with TMA stores
//   (coalesced, asynchronous, edge clipping for free) into a channel slice of the destination buffer.
//   The residual sub-tile is TMA-prefetched into the same ring slot one sub-tile ahead and updated in
//   place.  The two 2x-upsampling lateral convs store the same sub-tile through four tensor maps, one per
//   (dy, dx) phase of the upsampled image.
// * Tile index -> (n-block, w, h, n) uses multiply-shift division: three runtime integer divisions per
//   tile were ~800 cycles of serial latency on every role.
// * K chunk = one swizzle span: 64 channels (SWIZZLE_128B), 32 (64B) or 16 (32B; the Focus conv).
//
// Replaces ConvolutionBlock.forward (skyeye/core/models/blocks.py:36-38) and friends, see
// include/skyeye_b200.h.
#include <stdlib.h>

#include "common.cuh"

namespace skb {

// n / d and n % d for a runtime divisor without the ~100-cycle integer-division sequence
// (the per-tile coordinate decode sits on every role's critical path): n / d = umulhi(n, mul) >> shr.
struct FastDiv {
    uint32_t div, mul, shr;
};
static FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.div = (uint32_t)d;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;  // ceil(log2 d)
    const int pw = 31 + lg;
    f.mul = (uint32_t)((((unsigned long long)1 << pw) + (unsigned long long)d - 1) / (unsigned long long)d);
    f.shr = (uint32_t)(pw - 32);
    return f;
}
__device__ __forceinline__ void fast_divmod(const FastDiv& f, int n, int& q, int& r) {  // 0 <= n < 2^31
    q = f.div == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
    r = n - q * (int)f.div;
}

struct ConvParams {
    int tiles_w, tiles_h, tiles_n;
    int tw, th, tn;
    int n_blocks, total_tiles;
    FastDiv fd_nb, fd_tw, fd_th;  // divisors n_blocks, tiles_w, tiles_h of the tile index decode
    int k_iters, cchunks;
    int a_box_bytes;
    int tap_dw[9], tap_dh[9], tap_ph[9], tap_coff[9];
    int out_f32, cout, up2;
    const float* bias;
    int act;
    int has_res, sub_cols, epi_box_bytes;
    long long* trace;  // debug timeline of CTA 0 (null = off): [role][event] = clock64
};

// debug timeline: role 0 = producer, 1 = MMA, 2 = epilogue group 0 thread 0; 4096 events per role
#define SKB_TR(role, ev)                                                                         \
    do {                                                                                         \
        if (p.trace && blockIdx.x == 0 && tr_n < 4096) { p.trace[(role) * 8192 + 2 * tr_n] = (ev); p.trace[(role) * 8192 + 2 * tr_n + 1] = clock64(); ++tr_n; } \
    } while (0)

template <int BN, int BK, int NCTA>
struct ConvCfg {
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_BYTES = (BN / NCTA) * BK * 2;  // a CTA pair splits the weight tile: N/2 rows each
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_GROUPS = 2;       // two 4-warp epilogue groups, one per TMEM accumulator stage
    static constexpr int EPI_NB = 2;           // staging slots per group (sub-tiles)
    static constexpr int EPI_BUF = 16384;      // 128 rows x 128 B
    static constexpr int EPI_BYTES = EPI_GROUPS * EPI_NB * EPI_BUF + EPI_GROUPS * BN * 4;
    static constexpr int MAIN_BUDGET = 227 * 1024 - 1024 - 512 - EPI_BYTES;
    static constexpr int STAGES = (MAIN_BUDGET / STAGE_BYTES) < 8 ? (MAIN_BUDGET / STAGE_BYTES) : 8;
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // 64..512, power of two
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 512;
};

// work item (one per CTA, or per CTA pair) -> n-block and this CTA's pixel box.  In a pair, CTA r takes
// M tile 2*mp + r; a phantom tile past the end decodes to n0 >= B: its loads are zero fill, its stores clip.
template <int NCTA>
__device__ __forceinline__ void decode_tile(const ConvParams& p, int tile, int cta_rank, int& nb, int& w0, int& h0, int& n0) {
    int mt, iw, ih, in;
    fast_divmod(p.fd_nb, tile, mt, nb);
    mt = mt * NCTA + cta_rank;
    fast_divmod(p.fd_tw, mt, mt, iw);
    fast_divmod(p.fd_th, mt, in, ih);
    w0 = iw * p.tw; h0 = ih * p.th; n0 = in * p.tn;
}

// bias + activation (+ residual) on 32 accumulator columns -> four 16-byte chunks of bf16 in the staging row.
// All shared-memory loads are issued first, then the 32 independent value chains, then the four stores: the
// per-chunk version (load, compute, store, next chunk) was a serial ~100-cycle latency chain per chunk because the
// volatile shared-memory accesses keep their program order (measured: ~2500 cycles per 128 x 64 sub-tile).
__device__ __forceinline__ void epi_piece_bf16(const uint32_t* v, uint32_t bias_addr, int act, bool has_res, uint32_t rowp,
                                               uint32_t chunk0, uint32_t sw) {
    float4 b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = lds128f(bias_addr + 16u * i);
    uint4 r[4];
    uint32_t a16[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a16[j] = rowp + (((chunk0 + (uint32_t)j) ^ sw) << 4);
    if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = lds128(a16[j]);
    }
    float2 x[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[2 * i] = fadd2(make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), make_float2(b[i].x, b[i].y));
        x[2 * i + 1] = fadd2(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), make_float2(b[i].z, b[i].w));
    }
    if (act == SKB_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = silu2_tanh(x[i]);
    } else if (act == SKB_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { x[i].x = fmaxf(x[i].x, 0.f); x[i].y = fmaxf(x[i].y, 0.f); }
    }
    if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[4 * j + 0] = fadd2(x[4 * j + 0], make_float2(bf16_lo(r[j].x), bf16_hi(r[j].x)));
            x[4 * j + 1] = fadd2(x[4 * j + 1], make_float2(bf16_lo(r[j].y), bf16_hi(r[j].y)));
            x[4 * j + 2] = fadd2(x[4 * j + 2], make_float2(bf16_lo(r[j].z), bf16_hi(r[j].z)));
            x[4 * j + 3] = fadd2(x[4 * j + 3], make_float2(bf16_lo(r[j].w), bf16_hi(r[j].w)));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 o4;
        o4.x = pack_bf16x2(x[4 * j + 0].x, x[4 * j + 0].y);
        o4.y = pack_bf16x2(x[4 * j + 1].x, x[4 * j + 1].y);
        o4.z = pack_bf16x2(x[4 * j + 2].x, x[4 * j + 2].y);
        o4.w = pack_bf16x2(x[4 * j + 3].x, x[4 * j + 3].y);
        sts128(a16[j], o4);
    }
}
// bias + activation on 4 accumulator columns -> one 16-byte chunk of fp32 (precise SiLU)
__device__ __forceinline__ void epi_chunk_f32(const uint32_t* v, uint32_t bias_addr, int act, uint32_t a16) {
    const float4 b0 = lds128f(bias_addr);
    float f[4];
    f[0] = __uint_as_float(v[0]) + b0.x; f[1] = __uint_as_float(v[1]) + b0.y;
    f[2] = __uint_as_float(v[2]) + b0.z; f[3] = __uint_as_float(v[3]) + b0.w;
    if (act == SKB_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = silu_f(f[i]);
    } else if (act == SKB_ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = fmaxf(f[i], 0.0f);
    }
    uint4 o4;
    o4.x = __float_as_uint(f[0]); o4.y = __float_as_uint(f[1]);
    o4.z = __float_as_uint(f[2]); o4.w = __float_as_uint(f[3]);
    sts128(a16, o4);
}

template <int BN, int BK, int NCTA>
__global__ void __launch_bounds__(352, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmU1, const __grid_constant__ CUtensorMap tmU2,
                 const __grid_constant__ CUtensorMap tmU3, const ConvParams p) {
    using Cfg = ConvCfg<BN, BK, NCTA>;
    // NCTA == 2: the two CTAs of a cluster form a tcgen05 CTA pair.  One MMA (issued by the leader) computes
    // a 256 x BN tile: each CTA stages ITS 128 pixels and HALF of the weight tile, so the shared-memory
    // traffic per MMA cycle drops from 128*(128+BN)/BN to 128*(128+BN/2)/BN B/clk (192 -> 128 at BN = 256,
    // 256 -> 192 at BN = 128) -- the smem port, not the tensor pipe, bounds the single-CTA kernel.
    const int cta_rank = NCTA == 2 ? (int)cluster_ctarank() : 0;
    const bool leader = cta_rank == 0;
    const int work0 = NCTA == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // first work item of this CTA (pair)
    const int wstride = NCTA == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t SW = BK == 64 ? UMMA_SW128 : (BK == 32 ? UMMA_SW64 : UMMA_SW32);
    constexpr uint32_t SBO = 8 * BK * 2;  // bytes between 8-row groups
    constexpr int NB = Cfg::EPI_NB;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA0 = base;
    const uint32_t sB0 = base + STAGES * Cfg::A_BYTES;
    const uint32_t ebuf0 = base + STAGES * Cfg::STAGE_BYTES;
    const uint32_t sbias = ebuf0 + Cfg::EPI_GROUPS * NB * Cfg::EPI_BUF;
    const uint32_t bar0 = sbias + Cfg::EPI_GROUPS * BN * 4;
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    auto res_full = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + s); };
    const uint32_t slot = bar0 + 8u * (2 * STAGES + 4 + Cfg::EPI_GROUPS * NB);
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));

    // Warp-uniform role index (a shuffle result is uniform to the compiler) and ONE elected lane per warp for every
    // asynchronous issue (TMA, tcgen05.mma, commits): inside an `if (lane == 0)` region the compiler cannot prove that a
    // single thread is active and wraps every UTMALDG / UTCHMMA in an ELECT + R2UR "waterfall" loop, which cost the issuing
    // thread 100-160 cycles per instruction (attention: 640 -> 260 cycles per 4 MMAs once removed, profiles/r2b_uniform_issue.md).
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const bool lead = elect_one();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmY);
        if (p.has_res) tma_prefetch_desc(&tmR);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(full(s), 1);
                mbar_init(empty(s), 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(tfull(s), 1);
                mbar_init(tempty(s), 4 * NCTA);  // every epilogue warp of the pair arrives on the leader's barrier
            }
            for (int s = 0; s < Cfg::EPI_GROUPS * NB; ++s) mbar_init(res_full(s), 1);
            fence_barrier_init();
        }
        __syncwarp();
        if (NCTA == 2) tmem_alloc_pair(slot, Cfg::TMEM_COLS);
        else tmem_alloc(slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();  // the peer's barriers are initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot_ptr, 0);
    // Everything above (barrier init, TMEM allocation, descriptor prefetch) touched no global data and may have
    // overlapped the previous kernel's tail; from here on its outputs are read.
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0 || warp == 10) {
        // ===================== TMA producers =====================
        // One cp.async.bulk.tensor costs its issuing thread ~300 cycles whatever the box size (measured:
        // a 6 KB and a 32 KB k-iteration both took ~630 cycles with one thread issuing both loads), so the
        // activation and the weight tile of a stage are issued by two threads in different warps.  Warp 0
        // arms the stage's barrier with the total byte count and loads A, warp 10 loads B; B's bytes may
        // complete first (a transiently negative tx-count), the phase cannot complete before the arm.
        const bool loads_a = warp == 0;
        {
            int stage = 0;
            uint32_t phase = 0;
            int tr_n = 0;
            for (int tile = work0; tile < p.total_tiles; tile += wstride) {
                int nb, w0, h0, n0;
                decode_tile<NCTA>(p, tile, cta_rank, nb, w0, h0, n0);
                // tap / chunk counters and the current tap's coordinates live in registers: the per-iteration
                // it / cchunks division and four dynamically indexed parameter loads cost ~200 of the ~630 cycles a
                // k-iteration takes this thread (scripts/trace_focus.py), and the producer paces every BN <= 256 layer.
                int tap = 0, cc = 0;
                int t_coff = p.tap_coff[0], t_dw = p.tap_dw[0], t_ph = p.tap_ph[0], t_dh = p.tap_dh[0];
                for (int it = 0; it < p.k_iters; ++it) {
                    mbar_wait(empty(stage), phase ^ 1);
                    // pair: both CTAs' bytes complete on the LEADER's full barrier (its single arrival carries the total)
                    const uint32_t fb = NCTA == 2 ? mapa_shared(full(stage), 0) : full(stage);
                    if (loads_a) {
                        if (lead) {
                            SKB_TR(0, it);
                            if (leader) mbar_expect_tx(full(stage), (uint32_t)NCTA * (uint32_t)(p.a_box_bytes + Cfg::B_BYTES));
                            if (NCTA == 2)
                                tma_load_5d_pair(sA0 + stage * Cfg::A_BYTES, &tmA, fb, t_coff + cc * BK, w0 + t_dw, t_ph, h0 + t_dh, n0);
                            else
                                tma_load_5d(sA0 + stage * Cfg::A_BYTES, &tmA, fb, t_coff + cc * BK, w0 + t_dw, t_ph, h0 + t_dh, n0);
                        }
                        if (++cc == p.cchunks) {  // next tap: its coordinates load under the next barrier wait
                            cc = 0;
                            if (++tap < 9) { t_coff = p.tap_coff[tap]; t_dw = p.tap_dw[tap]; t_ph = p.tap_ph[tap]; t_dh = p.tap_dh[tap]; }
                        }
                    } else if (lead) {
                        if (NCTA == 2) tma_load_2d_pair(sB0 + stage * Cfg::B_BYTES, &tmB, fb, it * BK, nb * BN + cta_rank * (BN / 2));
                        else tma_load_2d(sB0 + stage * Cfg::B_BYTES, &tmB, fb, it * BK, nb * BN);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(128 * NCTA, BN);
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        int tr_n = 0;
        for (int tile = work0; leader && tile < p.total_tiles; tile += wstride) {
            mbar_wait(tempty(as), aphase ^ 1);
            tc_fence_after();
            if (lead) SKB_TR(1, 1000);
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
            for (int it = 0; it < p.k_iters; ++it) {
                mbar_wait(full(stage), phase);
                tc_fence_after();
                if (lead) {
                    SKB_TR(1, it);
                    const uint64_t ad = umma_desc(sA0 + stage * Cfg::A_BYTES, 16, SBO, SW);
                    const uint64_t bd = umma_desc(sB0 + stage * Cfg::B_BYTES, 16, SBO, SW);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (NCTA == 2) umma_bf16_ss_pair(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
                        else umma_bf16_ss(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
                    }
                    if (NCTA == 2) {  // frees the stage / publishes the accumulator in BOTH CTAs
                        umma_commit_pair(empty(stage));
                        if (it == p.k_iters - 1) umma_commit_pair(tfull(as));
                    } else {
                        umma_commit(empty(stage));
                        if (it == p.k_iters - 1) umma_commit(tfull(as));
                    }
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    } else {
        // ===================== epilogue: two groups of 4 warps, group g drains TMEM stage g =====================
        // (tiles alternate between the groups, so the epilogue of one tile overlaps the epilogue of the
        // next one as well as the MMAs of the one after; each warp owns 32 accumulator rows = its TMEM
        // lane quarter and walks the tile's columns in 32-column pieces)
        constexpr int GT = 128;                    // threads per group
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;               // accumulator row = pixel of the tile = TMEM lane
        const int tg = (int)threadIdx.x - 64 - g * GT;
        const bool T0 = ((warp - 2) & 3) == 0 && lead;  // one elected lane of the group's first warp issues its TMA stores / residual prefetches
        const int barid = 1 + g;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t ebuf_g = ebuf0 + (uint32_t)(g * NB) * Cfg::EPI_BUF;
        const uint32_t sbias_g = sbias + (uint32_t)(g * BN) * 4u;
        const int sub_cols = p.sub_cols;
        const uint32_t row_bytes = (uint32_t)sub_cols * (p.out_f32 ? 4u : 2u);
        const uint32_t sw = row_bytes == 128 ? (uint32_t)(m & 7) : 0u;  // SWIZZLE_128B: 16 B chunk ^= row % 8
        const uint32_t row_off = (uint32_t)m * row_bytes;
        const uint32_t acc = tmem_base + lane_addr + (uint32_t)(g * BN);
        const int step = 2 * wstride;
        uint32_t qseq = 0;  // running sub-tile number of this group: staging slot = qseq % NB
        uint32_t aphase = 0;
        int tr_n = 0;
        int tile = work0 + g * wstride;
        // TMEM stage g is handed back on the LEADER's barrier (the leader issues the pair's MMAs)
        const uint32_t tempty_g = NCTA == 2 ? mapa_shared(tempty(g), 0) : tempty(g);
        auto release_tmem = [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (NCTA == 2) mbar_arrive_cluster(tempty_g);
                else mbar_arrive(tempty_g);
            }
        };

        auto n_sub = [&](int nb) {
            int nv = (p.cout - nb * BN + sub_cols - 1) / sub_cols;
            return nv < 0 ? 0 : (nv > BN / sub_cols ? BN / sub_cols : nv);
        };
        auto load_bias = [&](int nb) {
            for (int i = tg; i < BN; i += GT) sts32f(sbias_g + 4u * i, __ldg(p.bias + nb * BN + i));
        };
        if (p.n_blocks == 1) load_bias(0);
        if (T0 && p.has_res && tile < p.total_tiles) {  // residual of the very first sub-tile
            int nb, w0, h0, n0;
            decode_tile<NCTA>(p, tile, cta_rank, nb, w0, h0, n0);
            if (n_sub(nb) > 0) {
                mbar_expect_tx(res_full(g * NB), (uint32_t)p.epi_box_bytes);
                tma_load_4d(ebuf_g, &tmR, res_full(g * NB), nb * BN, w0, h0, n0);
            }
        }
        named_bar_sync(barid, GT);
        for (; tile < p.total_tiles; tile += step) {
            int nb, w0, h0, n0;
            decode_tile<NCTA>(p, tile, cta_rank, nb, w0, h0, n0);
            const int ncol0 = nb * BN;
            const int nvalid = n_sub(nb);
            if (T0 && g == 0) SKB_TR(2, 100);
            if (p.n_blocks > 1) {  // the previous tile's last barrier ordered every read of the old bias
                load_bias(nb);
                named_bar_sync(barid, GT);
            }
            mbar_wait(tfull(g), aphase);
            tc_fence_after();
            if (T0 && g == 0) SKB_TR(2, 103);
            for (int sub = 0; sub < nvalid; ++sub, ++qseq) {
                const uint32_t slot_i = qseq % NB;
                const uint32_t buf = ebuf_g + slot_i * Cfg::EPI_BUF;
                const int c0 = sub * sub_cols;
                const uint32_t rowp = buf + row_off;
                const bool last = sub == nvalid - 1;
                if (p.has_res) mbar_wait(res_full(g * NB + slot_i), (qseq / NB) & 1u);
                if (!p.out_f32) {
                    const int pieces = sub_cols >> 5;  // 32 accumulator columns = 4 chunks of 8 bf16
                    uint32_t v[64];
                    {   // both 32-column pieces of the sub-tile are loaded from TMEM before any arithmetic starts
                        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                        tmem_ld32(acc + (uint32_t)c0, lo);
                        if (pieces == 2) tmem_ld32(acc + (uint32_t)(c0 + 32), hi);
                        tmem_ld_wait();
                    }
                    if (last) release_tmem();  // accumulator fully read: hand the TMEM stage back
                    epi_piece_bf16(v, sbias_g + 4u * (uint32_t)c0, p.act, p.has_res != 0, rowp, 0u, sw);
                    if (pieces == 2) epi_piece_bf16(v + 32, sbias_g + 4u * (uint32_t)(c0 + 32), p.act, p.has_res != 0, rowp, 4u, sw);
                } else {  // fp32 store: 32 columns per sub-tile = 8 chunks of 4 floats
                    uint32_t v[32];
                    tmem_ld32(acc + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (last) release_tmem();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        epi_chunk_f32(v + j * 4, sbias_g + 4u * (uint32_t)(c0 + j * 4), p.act, rowp + ((((uint32_t)j) ^ sw) << 4));
                }
                if (T0) tma_store_wait_read<0>();  // the group's previous store has drained the OTHER slot
                fence_proxy_async_smem();          // generic-proxy smem writes -> visible to the TMA engine
                named_bar_sync(barid, GT);
                if (T0) {
                    tma_store_4d(&tmY, buf, ncol0 + c0, w0, h0, n0);
                    if (p.up2) {  // nearest 2x upsample (detector.py:214,218): the same sub-tile goes to all 4 phases
                        tma_store_4d(&tmU1, buf, ncol0 + c0, w0, h0, n0);
                        tma_store_4d(&tmU2, buf, ncol0 + c0, w0, h0, n0);
                        tma_store_4d(&tmU3, buf, ncol0 + c0, w0, h0, n0);
                    }
                    tma_store_commit();
                    if (g == 0) SKB_TR(2, 110);
                    if (p.has_res) {  // prefetch the residual of the group's next sub-tile into the other slot
                        const uint32_t s2 = (qseq + 1) % NB;
                        int nb2 = nb, w2 = w0, h2 = h0, n2 = n0, c2 = c0 + sub_cols;
                        bool have = !last;
                        if (last && tile + step < p.total_tiles) {
                            decode_tile<NCTA>(p, tile + step, cta_rank, nb2, w2, h2, n2);
                            c2 = 0;
                            have = n_sub(nb2) > 0;
                        }
                        if (have) {
                            mbar_expect_tx(res_full(g * NB + s2), (uint32_t)p.epi_box_bytes);
                            tma_load_4d(ebuf_g + s2 * Cfg::EPI_BUF, &tmR, res_full(g * NB + s2), nb2 * BN + c2, w2, h2, n2);
                        }
                    }
                }
                __syncwarp();
            }
            if (nvalid == 0) release_tmem();
            aphase ^= 1;
        }
        if (T0) tma_store_wait_all();
    }

    tc_fence_before();
    if (NCTA == 2) {
        cluster_sync_all();  // the peer may still be reading operands / signalling barriers in this CTA
        if (warp == 1) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// =============================================================================================
// 3x3 stride-1 conv with a shared-memory HALO tile (the bottleneck convs of every CSP block).
//
// The generic kernel above fetches the activation tile once per filter tap (9x) and its weight tile once
// per 128 pixels; on B200 it is bound by bytes entering the SM through TMA (~53 B/clk/SM measured), not
// by the tensor pipe (37-43 % busy).  Here one CTA owns a 16 x 16 pixel tile (two 128-row accumulators,
// left and right 8-pixel halves):
//   * per 64-channel chunk the TMA engine loads ONE 18 x 24 pixel halo box (54 KB, rows padded to 24 pixels
//     so a pixel row is 3 swizzle atoms) and all 9 taps read it in place: the A operand of tap (dy, dx) and
//     half a is the same smem tile addressed at +dy*3072 + (8a + dx)*128 bytes with SBO = 3072 -- a UMMA
//     descriptor whose start is not 1024-aligned (the hardware swizzle follows absolute address bits, so the
//     TMA-written pattern and the MMA read agree; the descriptor's base-offset field stays 0);
//   * every weight tile (BN x 64, one per tap and chunk) is used by both accumulators.
// Bytes into the SM per 256 pixels: 54 KB*chunks + 9*chunks*BN*128 B, vs 2*9*chunks*(16 KB + BN*128 B) before
// (128->128: 396 KB vs 1152 KB), which makes these layers tensor-bound.
// Warp roles: 0 = halo TMA producer, 10 = weight TMA producer, 1 / 11 = MMA issuers of the left / right accumulator,
// 2..5 / 6..9 = epilogue of the left / right accumulator.  TMEM: 2 tiles in flight x 2 halves x BN columns.
// =============================================================================================
struct HaloParams {
    int tiles_w, tiles_h;
    int n_blocks, total_items;
    FastDiv fd_nb, fd_tw, fd_th;
    int cchunks, cin;
    int cout;
    const float* bias;
    int act, has_res;
    long long* trace;
    // tap list: tap t reads the halo box at (dy, dx) = (t / tdiv, t % tdiv); the box's origin is (w0 + w_org, h0 + h_org).
    // 3x3: ntaps 9, tdiv 3, origin (-1, -1).  Focus (row taps over pre-shifted 4-pixel windows): ntaps 3, tdiv 1, origin (0, -1).
    int ntaps, tdiv, w_org, h_org;
};

// HW = pixels per halo-box row: 24 for the 3x3 convs (16 + 2 halo pixels, padded to 3 swizzle atoms), 16 for the Focus row-tap
// conv (its column taps live inside the 4-pixel windows).  The smaller Focus box leaves room for THREE halo stages: a Focus tile
// has only 3 taps x 4 MMAs of work per box (768 tensor cycles), so with two stages the ~1.5 us HBM latency of the single box in
// flight paced the kernel (3600 cycles per tile measured).
template <int BN, int HW = 24>
struct HaloCfg {
    static constexpr int HALO_W = HW, HALO_H = 18;
    static constexpr int HALO_BYTES = HALO_H * HALO_W * 128;  // 55296 / 36864
    static constexpr int ROW_BYTES = HALO_W * 128;            // one halo-box row = the stride between 8-pixel row groups
    static constexpr int HS = HW == 24 ? 2 : 3;               // halo stages (HW 16, BN 64: 3 x 36 KB + 6 weight stages)
    static constexpr int B_BYTES = BN * 128;
    static constexpr int EPI_NB = BN == 64 ? 2 : 1;           // staging slots per epilogue group
    static constexpr int EPI_BUF = 16384;
    static constexpr int EPI_BYTES = 2 * EPI_NB * EPI_BUF + 2 * BN * 4;
    static constexpr int BUDGET = 227 * 1024 - 1024 - 512 - EPI_BYTES - HS * HALO_BYTES;
    static constexpr int BS = (BUDGET / B_BYTES) < 8 ? (BUDGET / B_BYTES) : 8;  // weight stages
    static constexpr int TMEM_COLS = 4 * BN;                  // 256 or 512
    static constexpr int SMEM_BYTES = HS * HALO_BYTES + BS * B_BYTES + EPI_BYTES + 1024 + 512;
};

__device__ __forceinline__ void halo_decode(const HaloParams& p, int item, int& nb, int& w0, int& h0, int& n) {
    int t, iw, ih;
    fast_divmod(p.fd_nb, item, t, nb);
    fast_divmod(p.fd_tw, t, t, iw);
    fast_divmod(p.fd_th, t, n, ih);
    w0 = iw * 16; h0 = ih * 16;
}

template <int BN, int HW>
__global__ void __launch_bounds__(384, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR, const HaloParams p) {
    using Cfg = HaloCfg<BN, HW>;
    constexpr int HS = Cfg::HS, BS = Cfg::BS, NB = Cfg::EPI_NB;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sH0 = base;
    const uint32_t sB0 = base + HS * Cfg::HALO_BYTES;
    const uint32_t ebuf0 = sB0 + BS * Cfg::B_BYTES;
    const uint32_t sbias = ebuf0 + 2 * NB * Cfg::EPI_BUF;
    const uint32_t bar0 = sbias + 2 * BN * 4;
    auto h_full = [&](int s) { return bar0 + 8u * s; };
    auto h_empty = [&](int s) { return bar0 + 8u * (HS + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (2 * HS + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (2 * HS + BS + s); };
    auto tfull = [&](int buf, int half) { return bar0 + 8u * (2 * HS + 2 * BS + buf * 2 + half); };       // per accumulator half
    auto tempty = [&](int buf, int half) { return bar0 + 8u * (2 * HS + 2 * BS + 4 + buf * 2 + half); };
    auto res_full = [&](int s) { return bar0 + 8u * (2 * HS + 2 * BS + 8 + s); };
    const uint32_t slot = bar0 + 8u * (2 * HS + 2 * BS + 8 + 2 * NB);
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // see conv_gemm_kernel
    const bool lead = elect_one();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmY);
        if (p.has_res) tma_prefetch_desc(&tmR);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < HS; ++s) { mbar_init(h_full(s), 1); mbar_init(h_empty(s), 2); }  // released by both MMA issuers
            for (int s = 0; s < BS; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 2); }
            for (int s = 0; s < 4; ++s) { mbar_init(tfull(s >> 1, s & 1), 1); mbar_init(tempty(s >> 1, s & 1), 4); }
            for (int s = 0; s < 2 * NB; ++s) mbar_init(res_full(s), 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(slot, Cfg::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot_ptr, 0);
    pdl_launch_dependents();  // prologue done without touching global data: see conv_gemm_kernel
    pdl_wait();

    if (warp == 0) {
        // ===================== halo producer =====================
        {
            int hs = 0;
            uint32_t hph = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int nb, w0, h0, n;
                halo_decode(p, item, nb, w0, h0, n);
                for (int c = 0; c < p.cchunks; ++c) {
                    mbar_wait(h_empty(hs), hph ^ 1);
                    if (lead) {
                        mbar_expect_tx(h_full(hs), (uint32_t)Cfg::HALO_BYTES);
                        tma_load_4d(sH0 + hs * Cfg::HALO_BYTES, &tmA, h_full(hs), c * 64, w0 + p.w_org, h0 + p.h_org, n);  // borders: zero fill
                    }
                    if (++hs == HS) { hs = 0; hph ^= 1; }
                }
            }
        }
    } else if (warp == 10) {
        // ===================== weight producer =====================
        {
            int bs = 0;
            uint32_t bph = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int nb, w0, h0, n;
                halo_decode(p, item, nb, w0, h0, n);
                for (int c = 0; c < p.cchunks; ++c)
                    for (int t = 0; t < p.ntaps; ++t) {
                        mbar_wait(b_empty(bs), bph ^ 1);
                        if (lead) {
                            mbar_expect_tx(b_full(bs), (uint32_t)Cfg::B_BYTES);
                            tma_load_2d(sB0 + bs * Cfg::B_BYTES, &tmB, b_full(bs), t * p.cin + c * 64, nb * BN);
                        }
                        if (++bs == BS) { bs = 0; bph ^= 1; }
                    }
            }
        }
    } else if (warp == 1 || warp == 11) {
        // ===================== MMA issuers: warp 1 = left half accumulator, warp 11 = right half =====================
        // One thread issues a tcgen05.mma at most every ~80 cycles whatever its N (scripts/micro/umma_rate.cu), i.e. 8 MMAs
        // per tap cost one issuer ~680 cycles against 4*BN cycles of tensor time: with BN <= 128 a single issuer was the
        // bound.  Each half has its own issuer, accumulator and full/empty barriers; stages are released by both.
        const int half = warp == 11 ? 1 : 0;
        constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
        int hs = 0, bs = 0, buf = 0;
        uint32_t hph = 0, bph = 0, tph = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            mbar_wait(tempty(buf, half), tph ^ 1);
            tc_fence_after();
            const uint32_t d0 = tmem_base + (uint32_t)((buf * 2 + half) * BN);
            for (int c = 0; c < p.cchunks; ++c) {
                mbar_wait(h_full(hs), hph);
                const uint32_t hbase = sH0 + hs * Cfg::HALO_BYTES;
                for (int t = 0; t < p.ntaps; ++t) {
                    mbar_wait(b_full(bs), bph);
                    tc_fence_after();
                    if (lead) {
                        const int dy = t / p.tdiv, dx = t - dy * p.tdiv;
                        const uint32_t a0 = hbase + (uint32_t)(dy * Cfg::ROW_BYTES + dx * 128 + half * 1024);  // right half: 8 pixels further
                        // start not 1024-aligned when dx != 0: the 128B-swizzle XOR is taken from the absolute smem
                        // address bits [7,10) (measured: correct with the descriptor's base-offset field left 0)
                        const uint64_t ad0 = umma_desc(a0, 16, Cfg::ROW_BYTES, UMMA_SW128);
                        const uint64_t bd = umma_desc(sB0 + bs * Cfg::B_BYTES, 16, 1024, UMMA_SW128);
                        const uint32_t acc = (c > 0 || t > 0) ? 1u : 0u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ss(d0, ad0 + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (acc || k > 0) ? 1u : 0u);
                        umma_commit(b_empty(bs));
                        if (t == p.ntaps - 1) {
                            umma_commit(h_empty(hs));
                            if (c == p.cchunks - 1) umma_commit(tfull(buf, half));
                        }
                    }
                    __syncwarp();
                    if (++bs == BS) { bs = 0; bph ^= 1; }
                }
                if (++hs == HS) { hs = 0; hph ^= 1; }
            }
            buf ^= 1;
            if (buf == 0) tph ^= 1;
        }
    } else {
        // ===================== epilogue: group g = accumulator half g (pixels w0 + 8g .. w0 + 8g + 7) =====================
        constexpr int GT = 128;
        const int g = (warp - 2) >> 2;
        const int q = warp & 3;
        const int m = q * 32 + lane;               // accumulator row = (row m / 8, pixel m % 8) of the half tile
        const int tg = (int)threadIdx.x - 64 - g * GT;
        const bool T0 = ((warp - 2) & 3) == 0 && lead;
        const int barid = 1 + g;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t ebuf_g = ebuf0 + (uint32_t)(g * NB) * Cfg::EPI_BUF;
        const uint32_t sbias_g = sbias + (uint32_t)(g * BN) * 4u;
        const uint32_t sw = (uint32_t)(m & 7);
        const uint32_t row_off = (uint32_t)m * 128u;
        constexpr int NSUB = BN / 64;
        uint32_t qseq = 0;
        int buf = 0;
        uint32_t tph = 0;
        auto load_bias = [&](int nb) {
            for (int i = tg; i < BN; i += GT) sts32f(sbias_g + 4u * i, __ldg(p.bias + nb * BN + i));
        };
        int item = blockIdx.x;
        if (p.n_blocks == 1) load_bias(0);
        if (T0 && p.has_res && item < p.total_items) {
            int nb, w0, h0, n;
            halo_decode(p, item, nb, w0, h0, n);
            mbar_expect_tx(res_full(g * NB), (uint32_t)Cfg::EPI_BUF);
            tma_load_4d(ebuf_g, &tmR, res_full(g * NB), nb * BN, w0 + 8 * g, h0, n);
        }
        named_bar_sync(barid, GT);
        for (; item < p.total_items; item += gridDim.x) {
            int nb, w0, h0, n;
            halo_decode(p, item, nb, w0, h0, n);
            const int ncol0 = nb * BN;
            if (p.n_blocks > 1) {
                load_bias(nb);
                named_bar_sync(barid, GT);
            }
            const uint32_t acc = tmem_base + lane_addr + (uint32_t)((buf * 2 + g) * BN);
            mbar_wait(tfull(buf, g), tph);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < NSUB; ++sub, ++qseq) {
                const uint32_t slot_i = qseq % NB;
                const uint32_t bufa = ebuf_g + slot_i * Cfg::EPI_BUF;
                const int c0 = sub * 64;
                const uint32_t rowp = bufa + row_off;
                const bool last = sub == NSUB - 1;
                if (p.has_res) mbar_wait(res_full(g * NB + slot_i), (qseq / NB) & 1u);
                {
                    uint32_t v[64];
                    uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                    uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                    tmem_ld32(acc + (uint32_t)c0, lo);
                    tmem_ld32(acc + (uint32_t)(c0 + 32), hi);
                    tmem_ld_wait();
                    if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty(buf, g));
                    }
                    epi_piece_bf16(v, sbias_g + 4u * (uint32_t)c0, p.act, p.has_res != 0, rowp, 0u, sw);
                    epi_piece_bf16(v + 32, sbias_g + 4u * (uint32_t)(c0 + 32), p.act, p.has_res != 0, rowp, 4u, sw);
                }
                if (NB == 2 && T0) tma_store_wait_read<0>();  // store(q-1) has drained the other slot
                fence_proxy_async_smem();
                named_bar_sync(barid, GT);
                if (T0) {
                    tma_store_4d(&tmY, bufa, ncol0 + c0, w0 + 8 * g, h0, n);
                    tma_store_commit();
                    if (NB == 1) tma_store_wait_read<0>();    // single slot: it must drain before it is refilled
                    if (p.has_res) {  // residual of the group's next sub-tile
                        int nb2 = nb, w2 = w0, h2 = h0, n2 = n, c2 = c0 + 64;
                        bool have = !last;
                        if (last && item + (int)gridDim.x < p.total_items) {
                            halo_decode(p, item + (int)gridDim.x, nb2, w2, h2, n2);
                            c2 = 0;
                            have = true;
                        }
                        if (have) {
                            const uint32_t s2 = (qseq + 1) % NB;
                            mbar_expect_tx(res_full(g * NB + s2), (uint32_t)Cfg::EPI_BUF);
                            tma_load_4d(ebuf_g + s2 * Cfg::EPI_BUF, &tmR, res_full(g * NB + s2), nb2 * BN + c2, w2 + 8 * g, h2, n2);
                        }
                    }
                }
                if (NB == 1) named_bar_sync(barid, GT);       // nobody refills the slot before it has drained
                else __syncwarp();
            }
            buf ^= 1;
            if (buf == 0) tph ^= 1;
        }
        if (T0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int HW = 24>
static int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR, const HaloParams& p,
                       cudaStream_t stream) {
    using Cfg = HaloCfg<BN, HW>;
    static PerDeviceOnce attr_once;
    if (attr_once.first())
        SKB_CUDA((cudaFuncSetAttribute(conv3x3_halo_kernel<BN, HW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)));
    const int grid = p.total_items < num_sms() ? p.total_items : num_sms();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    SKB_CUDA((cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<BN, HW>, tmA, tmB, tmY, tmR, p)));
    return SKB_OK;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int BK, int NCTA>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmR, ConvParams p,
                       cudaStream_t stream, const CUtensorMap* tmU = nullptr) {
    using Cfg = ConvCfg<BN, BK, NCTA>;
    const CUtensorMap& u1 = tmU ? tmU[0] : tmY;
    const CUtensorMap& u2 = tmU ? tmU[1] : tmY;
    const CUtensorMap& u3 = tmU ? tmU[2] : tmY;
    static PerDeviceOnce attr_once;
    if (attr_once.first())
        SKB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, BK, NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    const int units = num_sms() / NCTA;  // CTAs (or CTA pairs) resident at once
    const int grid = (p.total_tiles < units ? p.total_tiles : units) * NCTA;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(352);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    SKB_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BN, BK, NCTA>, tmA, tmB, tmY, tmR, u1, u2, u3, p));
    return SKB_OK;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// pick the (tn, th, tw) pixel box with the fewest 128-row MMA tiles; ties -> wider tw (longer runs)
static void pick_tile(int B, int Ho, int Wo, int& tn, int& th, int& tw, int& tiles) {
    long best = -1;
    for (int w = 1; w <= Wo && w <= 128; ++w) {
        for (int h = 1; h <= Ho && h * w <= 128; ++h) {
            int n = 128 / (h * w);
            if (n > B) n = B;
            if (n < 1) continue;
            long t = (long)cdiv(Wo, w) * cdiv(Ho, h) * cdiv(B, n);
            long score = t * 1024 - w;  // fewer tiles first, then wider
            if (best < 0 || score < best) {
                best = score;
                tn = n; th = h; tw = w;
                tiles = (int)t;
            }
        }
    }
}

}  // namespace skb

using namespace skb;

static long long* g_conv_trace = nullptr;
// Debug aid (not part of the drop-in surface): device buffer of 3*8192 int64 that CTA 0 of every
// following conv launch fills with (event, clock64) pairs; pass NULL to switch tracing off.
extern "C" int skb_debug_conv_trace(void* device_buffer) {
    g_conv_trace = (long long*)device_buffer;
    return SKB_OK;
}

extern "C" int skb_conv2d_bf16(const skb_view* x, const void* w_packed, const float* bias, const skb_view* residual,
                               const skb_view* y, int32_t cout_pad, int32_t ksize, int32_t stride, int32_t act,
                               int32_t upsample2x, void* stream) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(x && y && w_packed && bias && x->ptr && y->ptr, SKB_ERR_ARG, "conv2d: null argument");
    SKB_REQUIRE(x->dtype == SKB_BF16, SKB_ERR_ARG, "conv2d: input must be bf16");
    SKB_REQUIRE(ksize == 1 || ksize == 3, SKB_ERR_UNSUPPORTED, "conv2d: kernel size %d (only 1 and 3 are on the path)", ksize);
    SKB_REQUIRE(stride == 1 || stride == 2, SKB_ERR_UNSUPPORTED, "conv2d: stride %d", stride);
    const int Cin = x->c;
    SKB_REQUIRE(Cin % 16 == 0 && Cin >= 16, SKB_ERR_ARG, "conv2d: Cin=%d must be a multiple of 16 (pad the producer)", Cin);
    SKB_REQUIRE(x->pitch % 8 == 0 && ((uintptr_t)x->ptr & 15) == 0, SKB_ERR_ARG, "conv2d: input view must be 16B aligned (pitch %d)", x->pitch);
    SKB_REQUIRE(cout_pad % 32 == 0 && (cout_pad == 32 || cout_pad % 64 == 0), SKB_ERR_ARG, "conv2d: cout_pad=%d", cout_pad);
    SKB_REQUIRE(y->c % 8 == 0 && y->c <= cout_pad, SKB_ERR_ARG, "conv2d: y->c=%d must be a multiple of 8 and <= cout_pad", y->c);
    SKB_REQUIRE(y->pitch % (y->dtype == SKB_F32 ? 4 : 8) == 0 && ((uintptr_t)y->ptr & 15) == 0, SKB_ERR_ARG, "conv2d: output view alignment");
    SKB_REQUIRE(((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)bias & 15) == 0, SKB_ERR_ARG, "conv2d: weight/bias alignment");
    if (stride == 2) SKB_REQUIRE(x->h % 2 == 0 && x->w % 2 == 0, SKB_ERR_UNSUPPORTED, "conv2d: stride 2 needs even H, W (%dx%d)", x->h, x->w);
    const int Ho = stride == 1 ? x->h : x->h / 2, Wo = stride == 1 ? x->w : x->w / 2;
    const int up = upsample2x ? 2 : 1;
    SKB_REQUIRE(y->n == x->n && y->h == Ho * up && y->w == Wo * up, SKB_ERR_ARG, "conv2d: output dims [%d,%d,%d] != expected [%d,%d,%d]",
                y->n, y->h, y->w, x->n, Ho * up, Wo * up);
    if (residual) {
        SKB_REQUIRE(!upsample2x, SKB_ERR_UNSUPPORTED, "conv2d: residual with upsample2x");
        SKB_REQUIRE(residual->dtype == SKB_BF16 && residual->n == y->n && residual->h == y->h && residual->w == y->w &&
                        residual->c >= y->c && residual->pitch % 8 == 0 && ((uintptr_t)residual->ptr & 15) == 0,
                    SKB_ERR_ARG, "conv2d: residual view mismatch");
    }
    const int BK = (Cin % 64 == 0) ? 64 : (Cin % 32 == 0 ? 32 : 16);  // K chunk = swizzle span (128 / 64 / 32 B rows)
    const int taps = ksize * ksize;

    // ---- 3x3 stride-1: halo-tile kernel when 16 x 16 pixel tiles cover the map with little waste ----
    {
        static int halo_mode = -1;  // tuning knob (not part of the ABI): SKB_CONV_HALO=0 forces the generic kernel
        if (halo_mode < 0) {
            const char* e = getenv("SKB_CONV_HALO");
            halo_mode = e ? atoi(e) : 1;
        }
        const int th16 = cdiv(Ho, 16), tw16 = cdiv(Wo, 16);
        const double cover = (double)th16 * 16 * tw16 * 16 / ((double)Ho * Wo);
        if (halo_mode && ksize == 3 && stride == 1 && !upsample2x && Cin % 64 == 0 && y->dtype == SKB_BF16 && cout_pad % 64 == 0 &&
            cover <= 1.2) {
            const int BN = cout_pad % 128 == 0 ? 128 : 64;
            HaloParams hp;
            memset(&hp, 0, sizeof(hp));
            hp.tiles_w = tw16; hp.tiles_h = th16;
            hp.n_blocks = cout_pad / BN;
            hp.total_items = tw16 * th16 * x->n * hp.n_blocks;
            hp.fd_nb = make_fastdiv(hp.n_blocks); hp.fd_tw = make_fastdiv(tw16); hp.fd_th = make_fastdiv(th16);
            hp.cchunks = Cin / 64; hp.cin = Cin; hp.cout = y->c;
            hp.bias = bias; hp.act = act; hp.has_res = residual ? 1 : 0; hp.trace = g_conv_trace;
            hp.ntaps = 9; hp.tdiv = 3; hp.w_org = -1; hp.h_org = -1;
            CUtensorMap tmA, tmB, tmY, tmR;
            const uint64_t pitchB = (uint64_t)x->pitch * 2;
            uint64_t ad[4] = {(uint64_t)Cin, (uint64_t)x->w, (uint64_t)x->h, (uint64_t)x->n};
            uint64_t as[3] = {pitchB, pitchB * x->w, pitchB * x->w * x->h};
            uint32_t ab[4] = {64u, 24u, 18u, 1u};
            rc = encode_tensor_map(&tmA, x->ptr, 2, 4, ad, as, ab, 128);
            if (rc != SKB_OK) return rc;
            uint64_t bd[2] = {(uint64_t)taps * Cin, (uint64_t)cout_pad};
            uint64_t bs[1] = {(uint64_t)taps * Cin * 2};
            uint32_t bb[2] = {64u, (uint32_t)BN};
            rc = encode_tensor_map(&tmB, w_packed, 2, 2, bd, bs, bb, 128);
            if (rc != SKB_OK) return rc;
            uint64_t yd[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
            uint64_t ys[3] = {(uint64_t)y->pitch * 2, (uint64_t)y->pitch * 2 * y->w, (uint64_t)y->pitch * 2 * y->w * y->h};
            uint32_t yb[4] = {64u, 8u, 16u, 1u};
            rc = encode_tensor_map(&tmY, y->ptr, 2, 4, yd, ys, yb, 128);
            if (rc != SKB_OK) return rc;
            if (residual) {
                uint64_t rd[4] = {(uint64_t)residual->c, (uint64_t)residual->w, (uint64_t)residual->h, (uint64_t)residual->n};
                uint64_t rs[3] = {(uint64_t)residual->pitch * 2, (uint64_t)residual->pitch * 2 * residual->w,
                                  (uint64_t)residual->pitch * 2 * residual->w * residual->h};
                rc = encode_tensor_map(&tmR, residual->ptr, 2, 4, rd, rs, yb, 128);
                if (rc != SKB_OK) return rc;
            } else {
                tmR = tmB;
            }
            return BN == 128 ? launch_halo<128>(tmA, tmB, tmY, tmR, hp, (cudaStream_t)stream)
                             : launch_halo<64>(tmA, tmB, tmY, tmR, hp, (cudaStream_t)stream);
        }
    }

    ConvParams p;
    memset(&p, 0, sizeof(p));
    pick_tile(x->n, Ho, Wo, p.tn, p.th, p.tw, p.total_tiles);
    p.tiles_w = cdiv(Wo, p.tw);
    p.tiles_h = cdiv(Ho, p.th);
    p.tiles_n = cdiv(x->n, p.tn);
    const int m_tiles = p.total_tiles;
    // Output-channel block BN and CTA pairing: minimise waves * per-tile cost.  Per k-iteration the TMA writes
    // its operand tiles into shared memory and the MMA reads them back, in 2*BN tensor cycles per SM:
    //   one CTA : (128+BN) box rows of 128 B per k-iteration    CTA pair: (128+BN/2) rows per CTA
    // and the TMA engine delivers ~one 128-byte row per 2.4 cycles per SM, so a k-iteration cannot be shorter
    // than 2.4*(rows) cycles while its MMAs need 2*BN: wide BN amortises the activation rows.
    int BN = 32, NCTA = 1;
    {
        // CTA pairs: timed alone (clocks near 1.9 GHz) they do not beat the single-CTA kernel, which is bound by bytes
        // entering the SM through TMA, not by the shared-memory port.  Inside the full step the board sits on its
        // 1000 W power cap (SM clock ~1.7 GHz) and the halved weight traffic per CTA buys clock: +0.8-1.8 % images/s
        // measured on the same box, so pairs are ON by default.  SKB_CONV_PAIR=0 switches them off (tuning knob, not
        // part of the ABI).
        static int pair_mode = -1;
        if (pair_mode < 0) {
            const char* e = getenv("SKB_CONV_PAIR");
            pair_mode = e ? atoi(e) : 1;
        }
        // K = 256 layers (qkv, ff0, the 256-channel 1x1s) run 5-12 % faster unpaired when timed alone (scripts/tune_conv.py,
        // profiles/r2d_tune_conv.txt): too few k-iterations per tile to amortise the pair's cross-CTA hand-offs
        static int pair_min_k = -1;
        if (pair_min_k < 0) {
            const char* e = getenv("SKB_CONV_PAIR_MIN_K");
            pair_min_k = e ? atoi(e) : 512;
        }
        const bool pair_ok = pair_mode != 0 && !upsample2x && taps * Cin >= pair_min_k && m_tiles >= 2;
        double best = 1e30;
        const int cands[4] = {256, 128, 64, 32};
        const double pen1[4] = {1.5, 2.0, 3.0, 5.0};
        const double pen2[4] = {1.0, 1.5, 2.5, 1e9};
        for (int nc = 1; nc <= (pair_ok ? 2 : 1); ++nc)
            for (int i = 0; i < 4; ++i) {
                if (cout_pad % cands[i]) continue;
                if (nc == 2 && cands[i] < 64) continue;
                const long items = (long)cdiv(m_tiles, nc) * (cout_pad / cands[i]);
                const double cost = (double)cdiv((int)items, num_sms() / nc) * cands[i] * (nc == 2 ? pen2[i] : pen1[i]);
                if (cost < best) {
                    best = cost;
                    BN = cands[i];
                    NCTA = nc;
                }
            }
    }
    {   // tuning override (not part of the ABI): SKB_CONV_FORCE="BN,NCTA" is re-read on every call (scripts/tune_conv.py)
        const char* e = getenv("SKB_CONV_FORCE");
        int fbn = 0, fnc = 0;
        if (e && sscanf(e, "%d,%d", &fbn, &fnc) == 2 && (fbn == 32 || fbn == 64 || fbn == 128 || fbn == 256) && (fnc == 1 || fnc == 2) &&
            cout_pad % fbn == 0 && !(fnc == 2 && (fbn < 64 || upsample2x || m_tiles < 2))) {
            BN = fbn;
            NCTA = fnc;
        }
    }
    p.n_blocks = cout_pad / BN;
    p.total_tiles = cdiv(m_tiles, NCTA) * p.n_blocks;  // work items: one per CTA, or per CTA pair
    p.fd_nb = make_fastdiv(p.n_blocks); p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);
    p.cchunks = Cin / BK;
    p.k_iters = taps * p.cchunks;
    p.a_box_bytes = p.tn * p.th * p.tw * BK * 2;
    const int pad = ksize / 2;
    for (int r = 0; r < ksize; ++r)
        for (int s = 0; s < ksize; ++s) {
            const int t = r * ksize + s;
            if (stride == 1) {
                p.tap_dh[t] = r - pad; p.tap_dw[t] = s - pad; p.tap_ph[t] = 0; p.tap_coff[t] = 0;
            } else if (ksize == 3) {  // input row 2*oh + r - 1 -> (half-res row, parity)
                p.tap_dh[t] = r == 0 ? -1 : 0; p.tap_ph[t] = r == 1 ? 0 : 1;
                p.tap_dw[t] = s == 0 ? -1 : 0; p.tap_coff[t] = (s == 1 ? 0 : 1) * x->pitch;
            } else {  // 1x1 stride 2: input (2*oh, 2*ow)
                p.tap_dh[t] = 0; p.tap_dw[t] = 0; p.tap_ph[t] = 0; p.tap_coff[t] = 0;
            }
        }
    p.out_f32 = y->dtype == SKB_F32; p.cout = y->c; p.up2 = upsample2x ? 1 : 0;
    p.bias = bias; p.act = act;
    p.trace = g_conv_trace;
    p.has_res = residual ? 1 : 0;
    if (residual) SKB_REQUIRE(y->dtype == SKB_BF16, SKB_ERR_UNSUPPORTED, "conv2d: residual needs a bf16 output");
    p.sub_cols = p.out_f32 ? 32 : (BN >= 64 ? 64 : 32);
    p.epi_box_bytes = p.tn * p.th * p.tw * 128;

    CUtensorMap tmA, tmB, tmY, tmR, tmU[3];
    {
        const uint64_t pitchB = (uint64_t)x->pitch * 2;
        uint64_t dims[5], str[4];
        uint32_t box[5] = {(uint32_t)BK, (uint32_t)p.tw, 1u, (uint32_t)p.th, (uint32_t)p.tn};
        if (stride == 1) {
            dims[0] = Cin; dims[1] = x->w; dims[2] = 1; dims[3] = x->h; dims[4] = x->n;
            str[0] = pitchB; str[1] = pitchB * x->w; str[2] = pitchB * x->w; str[3] = pitchB * x->w * x->h;
        } else {
            dims[0] = (uint64_t)x->pitch + Cin; dims[1] = x->w / 2; dims[2] = 2; dims[3] = x->h / 2; dims[4] = x->n;
            str[0] = 2 * pitchB; str[1] = pitchB * x->w; str[2] = 2 * pitchB * x->w; str[3] = pitchB * x->w * x->h;
        }
        rc = encode_tensor_map(&tmA, x->ptr, 2, 5, dims, str, box, BK * 2);
        if (rc != SKB_OK) return rc;
        uint64_t bd[2] = {(uint64_t)taps * Cin, (uint64_t)cout_pad};
        uint64_t bs[1] = {(uint64_t)taps * Cin * 2};
        uint32_t bb[2] = {(uint32_t)BK, (uint32_t)(BN / NCTA)};
        rc = encode_tensor_map(&tmB, w_packed, 2, 2, bd, bs, bb, BK * 2);
        if (rc != SKB_OK) return rc;
        // epilogue maps: 4-D {c, w, h, n} boxes of sub_cols channels x (tw, th, tn) pixels, 128-byte rows
        const int esz = p.out_f32 ? 4 : 2;
        const int row_bytes = p.sub_cols * esz;
        uint64_t yd[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
        uint64_t ys[3] = {(uint64_t)y->pitch * esz, (uint64_t)y->pitch * esz * y->w, (uint64_t)y->pitch * esz * y->w * y->h};
        uint32_t yb[4] = {(uint32_t)p.sub_cols, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
        if (!upsample2x) {
            rc = encode_tensor_map(&tmY, y->ptr, esz, 4, yd, ys, yb, row_bytes == 128 ? 128 : 0);
            if (rc != SKB_OK) return rc;
        } else {
            // one map per (dy, dx) phase of the upsampled image: pixel (n, 2h+dy, 2w+dx) = base + phase offset
            // + w * 2 pixels + h * 2 rows; extents are the conv's own (Wo, Ho, N), so edge clipping still works
            const uint64_t pix = (uint64_t)y->pitch * esz, rowb = pix * y->w;
            uint64_t ud[4] = {(uint64_t)y->c, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)y->n};
            uint64_t us[3] = {2 * pix, 2 * rowb, rowb * y->h};
            for (int ph = 0; ph < 4; ++ph) {
                const uint8_t* basep = (const uint8_t*)y->ptr + (ph >> 1) * rowb + (ph & 1) * pix;
                rc = encode_tensor_map(ph == 0 ? &tmY : &tmU[ph - 1], basep, esz, 4, ud, us, yb, row_bytes == 128 ? 128 : 0);
                if (rc != SKB_OK) return rc;
            }
        }
        if (residual) {
            uint64_t rd[4] = {(uint64_t)residual->c, (uint64_t)residual->w, (uint64_t)residual->h, (uint64_t)residual->n};
            uint64_t rs[3] = {(uint64_t)residual->pitch * 2, (uint64_t)residual->pitch * 2 * residual->w,
                              (uint64_t)residual->pitch * 2 * residual->w * residual->h};
            rc = encode_tensor_map(&tmR, residual->ptr, 2, 4, rd, rs, yb, row_bytes == 128 ? 128 : 0);
            if (rc != SKB_OK) return rc;
            p.epi_box_bytes = p.tn * p.th * p.tw * row_bytes;
        } else {
            tmR = tmB;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
#define SKB_CONV_CASE(bn, bk, nc) \
    if (BN == bn && BK == bk && NCTA == nc) return launch_conv<bn, bk, nc>(tmA, tmB, tmY, tmR, p, st, upsample2x ? tmU : nullptr);
    SKB_CONV_CASE(256, 64, 1) SKB_CONV_CASE(128, 64, 1) SKB_CONV_CASE(64, 64, 1) SKB_CONV_CASE(32, 64, 1)
    SKB_CONV_CASE(256, 32, 1) SKB_CONV_CASE(128, 32, 1) SKB_CONV_CASE(64, 32, 1) SKB_CONV_CASE(32, 32, 1)
    SKB_CONV_CASE(256, 16, 1) SKB_CONV_CASE(128, 16, 1) SKB_CONV_CASE(64, 16, 1) SKB_CONV_CASE(32, 16, 1)
    SKB_CONV_CASE(256, 64, 2) SKB_CONV_CASE(128, 64, 2) SKB_CONV_CASE(64, 64, 2)
    SKB_CONV_CASE(256, 32, 2) SKB_CONV_CASE(128, 32, 2) SKB_CONV_CASE(64, 32, 2)
#undef SKB_CONV_CASE
    set_error("conv2d: no kernel for BN=%d BK=%d NCTA=%d", BN, BK, NCTA);
    return SKB_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------------
// FocusBlock as (padded space-to-depth) + (row-tap implicit GEMM over sliding 128-byte windows)
// ---------------------------------------------------------------------------------------------
namespace skb {
int launch_focus_pad(const void* img, int img_dtype, int n, int h, int w, void* scratch, cudaStream_t st, const int* tiles, int frame_h,
                     int frame_w);
}

extern "C" size_t skb_focus_conv_workspace_bytes(int32_t n, int32_t h, int32_t w) {
    if (n <= 0 || h <= 0 || w <= 0) return 256;
    return (size_t)n * (h / 2) * (w / 2 + 4) * 16 * 2 + 256;
}

static int focus_conv_impl(const void* img, int32_t img_dtype, int32_t n, int32_t h, int32_t w, const void* w_rowtap,
                           const float* bias, const skb_view* y, int32_t cout_pad, int32_t act, void* workspace,
                           size_t workspace_bytes, void* stream, const int32_t* tiles, int32_t frame_h, int32_t frame_w);

extern "C" int skb_focus_conv_bf16(const void* img, int32_t img_dtype, int32_t n, int32_t h, int32_t w, const void* w_rowtap,
                                   const float* bias, const skb_view* y, int32_t cout_pad, int32_t act, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    return focus_conv_impl(img, img_dtype, n, h, w, w_rowtap, bias, y, cout_pad, act, workspace, workspace_bytes, stream, nullptr, 0, 0);
}

extern "C" int skb_focus_conv_tiles_bf16(const void* frames, int32_t img_dtype, int32_t frame_h, int32_t frame_w, const int32_t* tiles_dev,
                                         int32_t n, int32_t h, int32_t w, const void* w_rowtap, const float* bias, const skb_view* y,
                                         int32_t cout_pad, int32_t act, void* workspace, size_t workspace_bytes, void* stream) {
    SKB_REQUIRE(tiles_dev && frame_h >= h && frame_w >= w, SKB_ERR_ARG, "focus_conv_tiles: tile table / frame %dx%d smaller than the %dx%d tile",
                frame_h, frame_w, h, w);
    return focus_conv_impl(frames, img_dtype, n, h, w, w_rowtap, bias, y, cout_pad, act, workspace, workspace_bytes, stream, tiles_dev, frame_h,
                           frame_w);
}

static int focus_conv_impl(const void* img, int32_t img_dtype, int32_t n, int32_t h, int32_t w, const void* w_rowtap,
                           const float* bias, const skb_view* y, int32_t cout_pad, int32_t act, void* workspace,
                           size_t workspace_bytes, void* stream, const int32_t* tiles, int32_t frame_h, int32_t frame_w) {
    int rc = check_device();
    if (rc != SKB_OK) return rc;
    SKB_REQUIRE(img && w_rowtap && bias && y && y->ptr && workspace, SKB_ERR_ARG, "focus_conv: null argument");
    SKB_REQUIRE(img_dtype == SKB_F32 || img_dtype == SKB_U8, SKB_ERR_ARG, "focus_conv: image dtype must be SKB_F32 or SKB_U8");
    SKB_REQUIRE(n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, SKB_ERR_ARG, "focus_conv: H, W must be even (got %dx%d)", h, w);
    SKB_REQUIRE(tiles || ((uintptr_t)img & (img_dtype == SKB_F32 ? 7 : 1)) == 0, SKB_ERR_ARG, "focus_conv: image alignment");
    const int Ho = h / 2, Wo = w / 2, Wp = Wo + 4;
    SKB_REQUIRE(y->dtype == SKB_BF16 && y->n == n && y->h == Ho && y->w == Wo, SKB_ERR_ARG, "focus_conv: output view [%d,%d,%d] != [%d,%d,%d]",
                y->n, y->h, y->w, n, Ho, Wo);
    SKB_REQUIRE(cout_pad % 32 == 0 && (cout_pad == 32 || cout_pad % 64 == 0), SKB_ERR_ARG, "focus_conv: cout_pad=%d", cout_pad);
    SKB_REQUIRE(y->c % 8 == 0 && y->c <= cout_pad && y->pitch % 8 == 0 && ((uintptr_t)y->ptr & 15) == 0, SKB_ERR_ARG, "focus_conv: output view alignment");
    SKB_REQUIRE(((uintptr_t)w_rowtap & 15) == 0 && ((uintptr_t)bias & 15) == 0, SKB_ERR_ARG, "focus_conv: weight/bias alignment");
    SKB_REQUIRE(workspace_bytes >= skb_focus_conv_workspace_bytes(n, h, w), SKB_ERR_WORKSPACE, "focus_conv: workspace too small");
    void* scratch = (void*)(((uintptr_t)workspace + 127) & ~(uintptr_t)127);
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_focus_pad(img, img_dtype, n, h, w, scratch, st, tiles, frame_h, frame_w);
    if (rc != SKB_OK) return rc;

    // Halo form (default): a 16 x 16 output tile takes ONE TMA box of 18 rows x 16 windows from the padded scratch and its three
    // row taps read that box in place (conv3x3_halo_kernel with the tap list (0,0) (1,0) (2,0)): 36 KB of activations + 24 KB of
    // weights per 256 pixels instead of 144 KB with one box per row tap -- the per-tap form was bound by bytes entering the SM.
    {
        static int focus_halo = -1;  // tuning knob (not part of the ABI): SKB_FOCUS_HALO=0 keeps the per-tap generic kernel
        if (focus_halo < 0) {
            const char* e = getenv("SKB_FOCUS_HALO");
            focus_halo = e ? atoi(e) : 1;
        }
        const int th16 = cdiv(Ho, 16), tw16 = cdiv(Wo, 16);
        const double cover = (double)th16 * 16 * tw16 * 16 / ((double)Ho * Wo);
        if (focus_halo && cout_pad % 64 == 0 && cover <= 1.2) {
            const int BN = cout_pad % 128 == 0 ? 128 : 64;
            HaloParams hp;
            memset(&hp, 0, sizeof(hp));
            hp.tiles_w = tw16; hp.tiles_h = th16;
            hp.n_blocks = cout_pad / BN;
            hp.total_items = tw16 * th16 * n * hp.n_blocks;
            hp.fd_nb = make_fastdiv(hp.n_blocks); hp.fd_tw = make_fastdiv(tw16); hp.fd_th = make_fastdiv(th16);
            hp.cchunks = 1; hp.cin = 64; hp.cout = y->c;
            hp.bias = bias; hp.act = act; hp.has_res = 0; hp.trace = g_conv_trace;
            hp.ntaps = 3; hp.tdiv = 1; hp.w_org = 0; hp.h_org = -1;
            CUtensorMap tmA, tmB, tmY;
            // window (n, oy, ox) = 64 elements starting at scratch pixel ox of padded row oy (32 B per pixel: overlapping rows)
            uint64_t ad[4] = {64, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)n};
            uint64_t as[3] = {32, (uint64_t)Wp * 32, (uint64_t)Wp * 32 * Ho};
            uint32_t ab[4] = {64u, 16u, 18u, 1u};
            rc = encode_tensor_map(&tmA, scratch, 2, 4, ad, as, ab, 128);
            if (rc != SKB_OK) return rc;
            uint64_t bd[2] = {192, (uint64_t)cout_pad};
            uint64_t bs[1] = {192 * 2};
            uint32_t bb[2] = {64u, (uint32_t)BN};
            rc = encode_tensor_map(&tmB, w_rowtap, 2, 2, bd, bs, bb, 128);
            if (rc != SKB_OK) return rc;
            uint64_t yd[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
            uint64_t ys[3] = {(uint64_t)y->pitch * 2, (uint64_t)y->pitch * 2 * y->w, (uint64_t)y->pitch * 2 * y->w * y->h};
            uint32_t yb[4] = {64u, 8u, 16u, 1u};
            rc = encode_tensor_map(&tmY, y->ptr, 2, 4, yd, ys, yb, 128);
            if (rc != SKB_OK) return rc;
            return BN == 128 ? launch_halo<128, 16>(tmA, tmB, tmY, tmB, hp, st) : launch_halo<64, 16>(tmA, tmB, tmY, tmB, hp, st);
        }
    }

    ConvParams p;
    memset(&p, 0, sizeof(p));
    pick_tile(n, Ho, Wo, p.tn, p.th, p.tw, p.total_tiles);
    p.tiles_w = cdiv(Wo, p.tw); p.tiles_h = cdiv(Ho, p.th); p.tiles_n = cdiv(n, p.tn);
    const int m_tiles = p.total_tiles;
    const int BN = cout_pad % 128 == 0 ? 128 : (cout_pad % 64 == 0 ? 64 : 32);
    p.n_blocks = cout_pad / BN;
    p.total_tiles = m_tiles * p.n_blocks;
    p.fd_nb = make_fastdiv(p.n_blocks); p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);
    p.cchunks = 1; p.k_iters = 3;
    p.a_box_bytes = p.tn * p.th * p.tw * 64 * 2;
    for (int t = 0; t < 3; ++t) { p.tap_dh[t] = t - 1; p.tap_dw[t] = 0; p.tap_ph[t] = 0; p.tap_coff[t] = 0; }
    p.out_f32 = 0; p.cout = y->c; p.up2 = 0;
    p.bias = bias; p.act = act; p.trace = g_conv_trace;
    p.sub_cols = BN >= 64 ? 64 : 32;
    const int row_bytes = p.sub_cols * 2;
    p.epi_box_bytes = p.tn * p.th * p.tw * row_bytes;

    CUtensorMap tmA, tmB, tmY;
    {
        // window (n, oy, ox) = scratch pixels ox .. ox+3 of the padded row = image pixels ox-1 .. ox+2:
        // 64 elements whose start advances by ONE pixel (32 B) per w step -> overlapping 128-byte rows
        uint64_t dims[5] = {64, (uint64_t)Wo, 1, (uint64_t)Ho, (uint64_t)n};
        uint64_t str[4] = {32, (uint64_t)Wp * 32, (uint64_t)Wp * 32, (uint64_t)Wp * 32 * Ho};
        uint32_t box[5] = {64u, (uint32_t)p.tw, 1u, (uint32_t)p.th, (uint32_t)p.tn};
        rc = encode_tensor_map(&tmA, scratch, 2, 5, dims, str, box, 128);
        if (rc != SKB_OK) return rc;
        uint64_t bd[2] = {192, (uint64_t)cout_pad};
        uint64_t bs[1] = {192 * 2};
        uint32_t bb[2] = {64u, (uint32_t)BN};
        rc = encode_tensor_map(&tmB, w_rowtap, 2, 2, bd, bs, bb, 128);
        if (rc != SKB_OK) return rc;
        uint64_t yd[4] = {(uint64_t)y->c, (uint64_t)y->w, (uint64_t)y->h, (uint64_t)y->n};
        uint64_t ys[3] = {(uint64_t)y->pitch * 2, (uint64_t)y->pitch * 2 * y->w, (uint64_t)y->pitch * 2 * y->w * y->h};
        uint32_t yb[4] = {(uint32_t)p.sub_cols, (uint32_t)p.tw, (uint32_t)p.th, (uint32_t)p.tn};
        rc = encode_tensor_map(&tmY, y->ptr, 2, 4, yd, ys, yb, row_bytes == 128 ? 128 : 0);
        if (rc != SKB_OK) return rc;
    }
    if (BN == 128) return launch_conv<128, 64, 1>(tmA, tmB, tmY, tmB, p, st);
    if (BN == 64) return launch_conv<64, 64, 1>(tmA, tmB, tmY, tmB, p, st);
    return launch_conv<32, 64, 1>(tmA, tmB, tmY, tmB, p, st);
}
