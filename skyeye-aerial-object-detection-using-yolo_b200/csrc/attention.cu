// Flash-style multi-head self-attention on tcgen05 (sm_100a), head_dim = 64, bf16 in / bf16 out.
// Replaces nn.MultiheadAttention's core inside TransformerLayer (skyeye/core/models/attention.py:298),
// which materialises the [B*h, N, N] weights (10.5 GB fp32 per image at N = 25600).
//
// One CTA = one (image, head, 128-query tile); it streams 64-key tiles of K and V:
//   S  = Q K^T        tcgen05.mma SS, Q/K K-major SWIZZLE_128B tiles from TMA, S fp32 in TMEM (2 buffers)
//   P  = exp2(S*c-m)  128 softmax threads (one query row each): tcgen05.ld -> online softmax with a
//                     lazily updated reference max (rescale O only when the max grows by > 2^8) ->
//                     bf16 P written to smem in the canonical K-major SWIZZLE_128B layout
//   O += P V          tcgen05.mma SS, V consumed MN-major straight from its [key][d] TMA tile
// The in_proj output [B, N, 3C] is read in place: one 3-D tensor map {channel, token, image} serves
// Q, K and V (different channel coordinates), ragged N is TMA zero fill + a -inf mask on the last tile.
// 2 CTAs are resident per SM (96 KB smem, 256 TMEM columns each) so one CTA's MUFU-bound softmax
// overlaps the other's MMAs.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = softmax/epilogue.
#include <math.h>

#include "common.cuh"

namespace skb {

constexpr int ATT_BQ = 128, ATT_BKV = 64, ATT_D = 64, ATT_KVS = 4;
constexpr int ATT_Q_BYTES = ATT_BQ * ATT_D * 2;     // 16 KB
constexpr int ATT_KV_BYTES = ATT_BKV * ATT_D * 2;   // 8 KB
constexpr int ATT_P_BYTES = ATT_BQ * ATT_BKV * 2;   // 16 KB
constexpr int ATT_SMEM = ATT_Q_BYTES + 2 * ATT_KVS * ATT_KV_BYTES + ATT_P_BYTES + 1024 + 256;
constexpr uint32_t ATT_TMEM_COLS = 256;             // S0 [0,64) S1 [64,128) O [128,192)

struct AttnParams {
    int N, heads, C, T;
    float scale_log2;
    __nv_bfloat16* out;
    long out_pitch;
};

__global__ void __launch_bounds__(192, 2)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sQ = base;
    const uint32_t sK0 = sQ + ATT_Q_BYTES;
    const uint32_t sV0 = sK0 + ATT_KVS * ATT_KV_BYTES;
    const uint32_t sP = sV0 + ATT_KVS * ATT_KV_BYTES;
    const uint32_t bar0 = sP + ATT_P_BYTES;
    const uint32_t q_full = bar0;
    auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar0 + 8u * (1 + ATT_KVS + s); };
    auto s_full = [&](int s) { return bar0 + 8u * (1 + 2 * ATT_KVS + s); };
    auto s_empty = [&](int s) { return bar0 + 8u * (3 + 2 * ATT_KVS + s); };
    const uint32_t p_full = bar0 + 8u * (5 + 2 * ATT_KVS);
    const uint32_t pv_done = bar0 + 8u * (6 + 2 * ATT_KVS);
    const uint32_t slot = bar0 + 8u * (7 + 2 * ATT_KVS);
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
    uint8_t* sP_ptr = smem_raw + (sP - raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int T = p.T;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(q_full, 1);
            for (int s = 0; s < ATT_KVS; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(s_full(s), 1); mbar_init(s_empty(s), 4); }
            mbar_init(p_full, 4);
            mbar_init(pv_done, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(slot, ATT_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;
    const uint32_t tmem_O = tmem + 128;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_expect_tx(q_full, ATT_Q_BYTES);
            tma_load_3d(sQ, &tmQ, q_full, head * ATT_D, qt * ATT_BQ, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % ATT_KVS;
                const uint32_t u = (uint32_t)(j / ATT_KVS);
                mbar_wait(kv_empty(s), (u & 1u) ^ 1u);
                mbar_expect_tx(kv_full(s), 2 * ATT_KV_BYTES);
                tma_load_3d(sK0 + s * ATT_KV_BYTES, &tmKV, kv_full(s), p.C + head * ATT_D, j * ATT_BKV, b);
                tma_load_3d(sV0 + s * ATT_KV_BYTES, &tmKV, kv_full(s), 2 * p.C + head * ATT_D, j * ATT_BKV, b);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BQ, ATT_BKV, 0, 0);  // A = Q (K-major), B = K (K-major)
        constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BQ, ATT_D, 0, 1);    // A = P (K-major), B = V (MN-major)
        auto issue_qk = [&](int j) {
            const int s = j % ATT_KVS, sb = j & 1;
            mbar_wait(kv_full(s), (uint32_t)(j / ATT_KVS) & 1u);
            mbar_wait(s_empty(sb), ((uint32_t)(j >> 1) & 1u) ^ 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t ad = umma_desc(sQ, 16, 1024, UMMA_SW128);
                const uint64_t bd = umma_desc(sK0 + s * ATT_KV_BYTES, 16, 1024, UMMA_SW128);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_bf16_ss(tmem + sb * ATT_BKV, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_qk, k > 0);
                umma_commit(s_full(sb));
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_qk(0);
        for (int j = 0; j < T; ++j) {
            if (j + 1 < T) issue_qk(j + 1);
            const int s = j % ATT_KVS;
            mbar_wait(p_full, (uint32_t)j & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t ad = umma_desc(sP, 16, 1024, UMMA_SW128);
                // V tile [64 keys][64 d]: MN-major, 128-byte rows, 8-row groups 1024 B apart, 16 keys per MMA = 2048 B
                const uint64_t bd = umma_desc(sV0 + s * ATT_KV_BYTES, 1024, 1024, UMMA_SW128);
#pragma unroll
                for (int k = 0; k < ATT_BKV / 16; ++k)
                    umma_bf16_ss(tmem_O, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 128), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(kv_empty(s));
                umma_commit(pv_done);
            }
            __syncwarp();
        }
    } else {
        // ===================== softmax + epilogue: thread <-> query row =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float m_ref = -INFINITY, l = 0.f;
        for (int j = 0; j < T; ++j) {
            const int sb = j & 1;
            mbar_wait(s_full(sb), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            uint32_t sv[64];
            {
                uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[0]);
                uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sv[32]);
                tmem_ld32(tmem + lane_addr + sb * ATT_BKV, lo);
                tmem_ld32(tmem + lane_addr + sb * ATT_BKV + 32, hi);
            }
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(sb));
            // row max of the raw scores (keys beyond N masked on the ragged last tile)
            float mx = -INFINITY;
            const int kbase = j * ATT_BKV;
            if (kbase + ATT_BKV > p.N) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (kbase + i >= p.N) sv[i] = 0xff800000u;  // -inf
            }
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
            const bool any = __any_sync(0xffffffffu, need);
            const float m_new = need ? mx : m_ref;
            if (j > 0) mbar_wait(pv_done, (uint32_t)(j - 1) & 1u);  // P buffer free, O quiescent
            if (any && j > 0) {
                tc_fence_after();
                const float alpha = need ? exp2f(m_ref - m_new) : 1.0f;
                l *= alpha;
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
            float sum8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            uint8_t* prow = sP_ptr + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    e[i] = ex2_approx(fmaf(__uint_as_float(sv[c * 8 + i]), p.scale_log2, -m_ref));  // one FFMA + one MUFU
                    sum8[i] += e[i];
                }
                uint4 u;
                u.x = pack_bf16x2(e[0], e[1]); u.y = pack_bf16x2(e[2], e[3]);
                u.z = pack_bf16x2(e[4], e[5]); u.w = pack_bf16x2(e[6], e[7]);
                *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = u;  // SWIZZLE_128B: 16B chunk ^= row % 8
            }
            l += ((sum8[0] + sum8[1]) + (sum8[2] + sum8[3])) + ((sum8[4] + sum8[5]) + (sum8[6] + sum8[7]));
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 ----
        mbar_wait(pv_done, (uint32_t)(T - 1) & 1u);
        tc_fence_after();
        const float inv = 1.0f / l;
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
    if (warp == 1) tmem_dealloc(tmem, ATT_TMEM_COLS);
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
    static bool attr_set = false;
    if (!attr_set) {
        SKB_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid((N + ATT_BQ - 1) / ATT_BQ, heads, B);
    flash_attn_kernel<<<grid, 192, ATT_SMEM, (cudaStream_t)stream>>>(tmQ, tmKV, p);
    SKB_LAUNCH_CHECK();
    return SKB_OK;
}
