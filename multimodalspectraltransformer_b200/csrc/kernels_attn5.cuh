// Encoder self-attention of the 32-wide heads of encoder_cross (models_MMT_v15_4.py:531-533, 941-944) on the sm_100a tensor
// cores: tcgen05.mma with the scores and the output accumulators in TMEM.
//
// One CTA = one (head, sequence); 256 threads = two per TMEM lane = two per query row of a 128-row tile (each takes half of a
// chunk's keys in the softmax and half of the head's dims in the output); the sequence's queries are walked in tiles of 128 rows, its attendable keys in chunks of 64 (any key count: the 582-row memory at maximum peak
// counts, the 902 rows of the "MS" modes).  Per (query tile, key chunk):
//
//   stage    K chunk  -> smem tile [64 keys][Kh(32) | Kl(32)]           bf16, 128-byte swizzled rows (threads, generic stores)
//            V chunk  -> smem tile [Vh^T (32 dims) ; Vl^T (32 dims)][64 keys]
//   MMA 1    S (TMEM, 128 x 64 fp32) = [Qh | Ql] . [Kh | Kl]^T + [Qh | Ql] . [Kl | Kh]^T      8 x tcgen05.mma 128x64x16
//   softmax  thread = query row: tcgen05.ld the 64 scores, + key bias, running max / sum (log2 domain, ex2.approx),
//            P -> smem tiles Ph, Pl [128][64] (swizzled), running output (registers) rescaled
//   MMA 2    O (TMEM, 128 x 64) = Ph . [Vh^T ; Vl^T]^T (columns 0-31: Ph.Vh, 32-63: Ph.Vl), += Pl . Vh^T into columns 0-31
//   update   tcgen05.ld O, out += O[0:32] + O[32:64]
//
// Every operand is a two-term bf16 split (x = hi + lo): the two K tiles of MMA 1 carry the halves of K in both orders, so two
// 64-wide contractions against ONE Q tile yield all four products (Qh + Ql).(Kh + Kl) -- the scores are the fp32 kernel's to
// fp32 round-off, like the mma.sync kernel this replaces (which spends three MMAs per product and is issue-bound: HMMA is 11 %
// of its instructions).  The phases of a chunk run one after the other (all threads stage and do the softmax, thread 0
// issues the MMAs; the next chunk's K / V rows are loaded into registers meanwhile); three CTAs per SM (74 KB of shared memory,
// 128 TMEM columns each) overlap each other's phases.
#pragma once
#include "kernels_tc.cuh"

namespace mmt {

constexpr int A5_DH = 32;                 // head width
constexpr int A5_KC = 64;                 // keys per chunk
constexpr int A5_THREADS = 256;           // two threads per query row: thread t -> row t & 127, half t >> 7 (keys / dims of the chunk)
constexpr int A5_TILE = TC_SLAB_BYTES;    // [128 rows][128 B] = 16 KB
constexpr int A5_HALF = A5_TILE / 2;      // [64 rows][128 B] = 8 KB
// smem: A1 (Q tile) | Ph | Pl | K1 | K2 | V^T | key bias | row-sum exchange
constexpr int A5_SMEM_BYTES = 3 * A5_TILE + 3 * A5_HALF + A5_KC * 4 + 128 * 4 + 1024;
constexpr uint32_t A5_TMEM_COLS = 128;    // S chunk: columns 0-63, O chunk: 64-127

// byte offset of 16-byte chunk `c` (8 bf16) of row `r` in a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t a5_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void a5_split8(const float* x, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair(x[2 * i], x[2 * i + 1], h[i], l[i]);
    hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(A5_THREADS, 2) attn_encoder_tc5(const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t a5_raw[];
    __shared__ __align__(8) uint64_t bar_s, bar_o;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = a5_raw + ((1024u - (smem_u32(a5_raw) & 1023u)) & 1023u);
    uint8_t* sA1 = smem;                       // [128 rows][Qh | Ql]
    uint8_t* sPh = sA1 + A5_TILE;
    uint8_t* sPl = sPh + A5_TILE;
    uint8_t* sK1 = sPl + A5_TILE;              // [64 keys][Kh | Kl]
    uint8_t* sK2 = sK1 + A5_HALF;              // [64 keys][Kl | Kh]: A1 . K1^T + A1 . K2^T = (Qh + Ql) . (Kh + Kl)
    uint8_t* sV = sK2 + A5_HALF;               // rows 0-31: Vh^T, rows 32-63: Vl^T; 64 keys per row
    float* bs = reinterpret_cast<float*>(sV + A5_HALF);
    float* lsum = bs + A5_KC;                  // [128] row sums of the upper-half threads

    const AttnGroup& g = p.g[blockIdx.z];
    const int h = blockIdx.x, b = blockIdx.y;
    const int S = g.cnt ? g.cnt[b] : g.S;
    const int kstride = g.cnt ? g.kstride : g.S;
    const int64_t row0 = g.row_start ? (int64_t)g.row_start[b] : (int64_t)b * g.S;
    const int nk = g.nk[b];
    const float* base = g.qkv + row0 * (3 * D);
    constexpr float LOG2E = 1.4426950408889634f;
    const float qscale = p.scale * LOG2E;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, half = tid >> 7;

    if (tid == 0) { mbar_init(&bar_s, 1); mbar_init(&bar_o, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, A5_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_s = tmem_slot, tmem_o = tmem_slot + A5_KC;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;    // TMEM lane quarter of this warp (warp id % 4)
    constexpr uint32_t idesc64 = umma_idesc_bf16(TC_BM, A5_KC), idesc32 = umma_idesc_bf16(TC_BM, A5_DH);
    uint32_t it = 0;                                                   // chunk iterations so far (mbarrier phase)

    // K / V rows of a chunk: thread -> key (tid >> 2), dims 8 (tid & 3) .. + 7; loaded a chunk ahead (registers)
    const int sj = tid >> 2, sq = tid & 3;
    float kk[8], vv[8];
    int skey;
    auto load_chunk = [&](int k0) {
        skey = (k0 + sj < nk) ? g.kidx[(int64_t)b * kstride + k0 + sj] : -1;
        if (skey >= 0) {
            const float4* kr = reinterpret_cast<const float4*>(base + (int64_t)skey * (3 * D) + D + h * A5_DH + 8 * sq);
            const float4* vr = reinterpret_cast<const float4*>(base + (int64_t)skey * (3 * D) + 2 * D + h * A5_DH + 8 * sq);
            const float4 a0 = kr[0], a1 = kr[1], c0 = vr[0], c1 = vr[1];
            kk[0] = a0.x; kk[1] = a0.y; kk[2] = a0.z; kk[3] = a0.w; kk[4] = a1.x; kk[5] = a1.y; kk[6] = a1.z; kk[7] = a1.w;
            vv[0] = c0.x; vv[1] = c0.y; vv[2] = c0.z; vv[3] = c0.w; vv[4] = c1.x; vv[5] = c1.y; vv[6] = c1.z; vv[7] = c1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { kk[i] = 0.f; vv[i] = 0.f; }
        }
    };

    for (int q0 = 0; q0 < S; q0 += TC_BM) {
        load_chunk(0);
        // ---- Q tile: row `row`, dims 16 half .. + 15, scaled into the log2 domain
        {
            float q[16];
            const int r = q0 + row;
            if (r < S) {
                const float4* src = reinterpret_cast<const float4*>(base + (int64_t)r * (3 * D) + h * A5_DH + 16 * half);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float4 v = src[i]; q[4 * i] = v.x * qscale; q[4 * i + 1] = v.y * qscale; q[4 * i + 2] = v.z * qscale; q[4 * i + 3] = v.w * qscale; }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) q[i] = 0.f;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint4 hi, lo;
                a5_split8(q + 8 * c, hi, lo);
                *reinterpret_cast<uint4*>(sA1 + a5_off(row, 2 * half + c)) = hi;
                *reinterpret_cast<uint4*>(sA1 + a5_off(row, 4 + 2 * half + c)) = lo;
            }
        }
        float out[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) out[d] = 0.f;
        float m = MMT_NEG_INF, l = 0.f;

        for (int k0 = 0; k0 < nk; k0 += A5_KC, ++it) {
            // ---- stage the key chunk from the registers loaded a chunk ago
            {
                if (sq == 0) bs[sj] = skey >= 0 ? (g.kbias ? g.kbias[(int64_t)b * g.S + skey] * LOG2E : 0.f) : MMT_NEG_INF;
                uint4 hi, lo;
                a5_split8(kk, hi, lo);
                *reinterpret_cast<uint4*>(sK1 + a5_off(sj, sq)) = hi; *reinterpret_cast<uint4*>(sK1 + a5_off(sj, 4 + sq)) = lo;
                *reinterpret_cast<uint4*>(sK2 + a5_off(sj, sq)) = lo; *reinterpret_cast<uint4*>(sK2 + a5_off(sj, 4 + sq)) = hi;
#pragma unroll
                for (int i = 0; i < 8; ++i) {       // V^T: element (dim, key sj) of the hi / lo planes
                    const int d = 8 * sq + i;
                    const __nv_bfloat16 vh = __float2bfloat16_rn(vv[i]);
                    const __nv_bfloat16 vl = __float2bfloat16_rn(vv[i] - __bfloat162float(vh));
                    *reinterpret_cast<__nv_bfloat16*>(sV + a5_off(d, sj >> 3) + (sj & 7) * 2) = vh;
                    *reinterpret_cast<__nv_bfloat16*>(sV + a5_off(A5_DH + d, sj >> 3) + (sj & 7) * 2) = vl;
                }
            }
            fence_proxy_async_smem();
            __syncthreads();
            // ---- MMA 1: scores of the chunk
            if (tid == 0) {
                tc_fence_after();
                const uint64_t a1 = umma_desc_sw128(smem_u32(sA1)), k1 = umma_desc_sw128(smem_u32(sK1)), k2 = umma_desc_sw128(smem_u32(sK2));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, a1 + (uint64_t)(2 * k), k1 + (uint64_t)(2 * k), idesc64, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, a1 + (uint64_t)(2 * k), k2 + (uint64_t)(2 * k), idesc64, 1u);
                umma_commit(&bar_s);
            }
            if (k0 + A5_KC < nk) load_chunk(k0 + A5_KC);       // the next chunk's rows travel while the tensor core and the softmax work
            // only the issuing thread waits on the mbarrier; the others park at the hardware barrier (256 threads polling
            // try_wait take the issue slots the other CTAs of the SM need)
            if (tid == 0) mbar_wait(&bar_s, it & 1u);
            __syncthreads();
            tc_fence_after();
            // ---- softmax: both threads of a row take the row maximum over all 64 scores, each exponentiates its 32 keys
            {
                uint32_t own[32];
                float cm = MMT_NEG_INF;
                {
                    uint32_t oth[32];
                    tmem_ld_32x32(tmem_s + lane_base + 32 * half, own);
                    tmem_ld_32x32(tmem_s + lane_base + 32 * (1 - half), oth);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float a = __uint_as_float(own[j]) + bs[32 * half + j];
                        own[j] = __float_as_uint(a);
                        cm = fmaxf(cm, fmaxf(a, __uint_as_float(oth[j]) + bs[32 * (1 - half) + j]));
                    }
                }
                const float mn = fmaxf(m, cm);
                const float corr = ex2_approx(m - mn);          // first chunk: 2^-inf = 0
                m = mn;
                l *= corr;
#pragma unroll
                for (int d = 0; d < 16; ++d) out[d] *= corr;
#pragma unroll
                for (int c = 0; c < 4; ++c) {                   // 8 keys -> one 16-byte chunk of the P row, hi and lo
                    float e[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { e[u] = ex2_approx(__uint_as_float(own[8 * c + u]) - mn); l += e[u]; }
                    uint4 hi, lo;
                    a5_split8(e, hi, lo);
                    *reinterpret_cast<uint4*>(sPh + a5_off(row, 4 * half + c)) = hi;
                    *reinterpret_cast<uint4*>(sPl + a5_off(row, 4 * half + c)) = lo;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            // ---- MMA 2: the chunk's contribution to the output
            if (tid == 0) {
                tc_fence_after();
                const uint64_t ph = umma_desc_sw128(smem_u32(sPh)), pl = umma_desc_sw128(smem_u32(sPl)), vd = umma_desc_sw128(smem_u32(sV));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, ph + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc64, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, pl + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc32, 1u);
                umma_commit(&bar_o);
                mbar_wait(&bar_o, it & 1u);
            }
            __syncthreads();
            tc_fence_after();
            {   // dims 16 half .. + 15: Ph.Vh + Pl.Vh (columns 0-31) and Ph.Vl (columns 32-63)
                uint32_t o0[16], o1[16];
                tmem_ld_32x16(tmem_o + lane_base + 16 * half, o0);
                tmem_ld_32x16(tmem_o + lane_base + 32 + 16 * half, o1);
                tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < 16; ++d) out[d] += __uint_as_float(o0[d]) + __uint_as_float(o1[d]);
            }
            tc_fence_before();          // the next chunk's MMAs overwrite S / O: order this thread's TMEM reads before them
        }
        // ---- write the tile's rows: the row sum is the sum of the two threads' partial sums
        if (half == 1) lsum[row] = l;
        __syncthreads();
        if (half == 0) lsum[row] = l + lsum[row];
        __syncthreads();
        {
            const int r = q0 + row;
            if (r < S) {
                const float inv = 1.0f / lsum[row];
                if (g.out) {
                    float4* dst = reinterpret_cast<float4*>(g.out + (row0 + r) * D + h * A5_DH + 16 * half);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(out[4 * i] * inv, out[4 * i + 1] * inv, out[4 * i + 2] * inv, out[4 * i + 3] * inv);
                }
                if (g.out16) {
                    uint2* dst = reinterpret_cast<uint2*>(g.out16 + (row0 + r) * D + h * A5_DH + 16 * half);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = pack_bf16x4(make_float4(out[4 * i] * inv, out[4 * i + 1] * inv, out[4 * i + 2] * inv, out[4 * i + 3] * inv));
                }
            }
        }
        __syncthreads();        // the next tile rewrites the Q tile and lsum
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_s, A5_TMEM_COLS);
    }
}

}  // namespace mmt
