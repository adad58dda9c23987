// bf16 tensor-core kernels of the MMT path for sm_100a: tcgen05.mma with TMEM accumulators,
// operands staged in shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle), fused
// epilogues (bias / ReLU / bf16 cast, or bias + residual + LayerNorm over the 128-wide row).
//
//   C[M,N] = epi(A[M,K] . W[N,K]^T)      A, W bf16 with K contiguous, fp32 accumulate
//
// Weights may be given as a two-term bf16 split W ~= W_hi + W_lo (wsplit): the MMA pass runs
// twice per K slab on the same A slab, which removes the weight-rounding error (the dominant
// term of the bf16 logit error, DESIGN.md "bf16 numerics") for no extra HBM traffic.
//
// One CTA = one 128x128 output tile (UMMA 128x128x16, cta_group::1), 10 warps:
//   warp 0     TMA producer   (one lane): K slabs of 64 elements (=128 B swizzle rows)
//   warp 1     MMA issuer     (one lane) + TMEM allocation (128 columns)
//   warps 2-9  epilogue: TMEM -> registers -> padded smem tile -> coalesced global I/O
//              (two warps per TMEM lane quarter: with one warp per scheduler the epilogue is a pure
//              latency chain and paces the whole kernel)
// Several CTAs are resident per SM (68 KB smem at K=128), so one tile's epilogue overlaps
// the TMA/MMA of its neighbours without a persistent scheduler.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder entry point is fetched at run time)

#include "common.cuh"

namespace mmt {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
constexpr int TC_SLAB_BYTES = TC_BM * TC_BK * 2;          // 16 KB: one operand slab (128 rows x 128 B)
constexpr int TC_STAGE_BYTES = 2 * TC_SLAB_BYTES;          // A slab + W slab
constexpr int TC_STAGE_BYTES_WSPLIT = 3 * TC_SLAB_BYTES;   // A slab + W_hi slab + W_lo slab
constexpr int TC_LDS = TC_BN + 4;                          // padded fp32 staging row (conflict-free float4)
constexpr int TC_STAGING_BYTES = TC_BM * TC_LDS * 4;       // 67,584 B
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + TC_EPI_WARPS * 32;
constexpr int TC_MAX_STAGES = 4;

enum { TC_EPI_STORE = 0, TC_EPI_LN = 1 };

struct TcGemmParams {
    CUtensorMap tmA;          // A [M,K] bf16, box {64,128}, SWIZZLE_128B
    CUtensorMap tmW;          // W [N,K] bf16, box {64,128}, SWIZZLE_128B
    CUtensorMap tmW2;         // low-order term of the two-term weight split (wsplit): W ~= W_hi + W_lo, both bf16
    int wsplit;               // 1: accumulate A.W_hi^T + A.W_lo^T (weight rounding error removed; 2x MMAs, same HBM bytes)
    int M, N, K;
    int stages;               // smem ring depth (<= TC_MAX_STAGES)
    int splits;               // split-K over blockIdx.z: raw fp32 partials, no bias / act
    int64_t part_stride;      // floats between split partials
    const float* bias;        // [N] or nullptr
    int act;                  // 0 none, 1 ReLU
    float* out_f32; int64_t ld_f32;            // optional fp32 output
    __nv_bfloat16* out_b16; int64_t ld_b16;    // optional bf16 output
    // cross-attention K/V output, one contiguous block per memory (spectrum): row r = b * hm_rows + j, column c ->
    // C[(((b * 2 + c/128) * heads + (c%128)/dh) * hm_rows + j) * dh + c % dh]   (hm_rows = memory rows per spectrum)
    int head_major, hm_heads, hm_dh; int64_t hm_rows;
    // LayerNorm epilogue (N == 128): out = LN(acc + bias + res[r]) * gamma + beta
    const float* res; const float* gamma; const float* beta; float eps;
    // output row map (in rows): (r / S_in) * stride_b + (r % S_in) * stride_s + off
    int S_in; int64_t stride_b, stride_s, off;
    const int* out_rows;      // optional explicit output row per input row (ragged encoder), overrides the affine map
    // Decoder self-attention QKV projection (N = 384): columns 128.. (K, V of the new position) go straight into the paged bf16
    // KV cache of the layer -- page block_table[row * pps + t / 16], slot t % 16, t = *step -- instead of an fp32 row the
    // attention kernel would re-read and append; only the Q columns are stored to out_f32.
    int kv_append;   // 0 none, 1 head-major pages, 2 token-major pages
    __nv_bfloat16* kv_pool; const int* block_table; int pps; const int* step; int kv_heads;
    // Chained projection (LayerNorm epilogue only): chain_out[r] = bf16(LN output row r) . Wc^T + chain_bias, Wc [128,128] as a
    // two-term bf16 split.  The decoder's "out-proj + LN1" and "cross-attention query projection" as ONE launch: the
    // normalised rows go to shared memory as the A operand of a second MMA instead of round-tripping through HBM.
    int chain;                // 0 none, 1 Wc hi term only, 2 both terms
    CUtensorMap tmC, tmC2;    // Wc hi / lo, box {64,128}, SWIZZLE_128B
    const float* chain_bias; float* chain_out; int64_t ld_chain;
};
constexpr int TC_CHAIN_BYTES = 6 * TC_SLAB_BYTES;        // A2 (2 K slabs) | Wc hi (2) | Wc lo (2) = 96 KB

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("mmt: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::f16 (bf16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand slab, 128-byte swizzle: rows of 128 B, 8-row atoms 1024 B apart (SBO),
// descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).  LBO is unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

// ------------------------------------------------------------------ shared epilogue pieces
// The epilogue warps (4 consecutive warps, warp id % 4 = TMEM lane quarter q) first move their 32
// accumulator rows TMEM -> registers -> a padded fp32 smem tile (thread = row), then walk the rows
// with lane = 4 consecutive columns so that every global access is a coalesced 512 B row segment.
// Rows are processed 8 at a time so that the loads and the LayerNorm shuffle chains of
// independent rows overlap (a single row's chain is ~500 cycles of pure latency).
constexpr int EPI_ILP = 8;

// Epilogue warp e = warp - 2 owns TMEM lane quarter q = warp % 4 (hardware rule) and half hf = e / 4:
// it moves the hf-th half of the accumulator columns of its 32 rows into the staging tile, and after
// epi_bar_sync() processes rows [16*hf, 16*hf + 16) of the quarter.
// (NP = epilogue warps per lane quarter: 2 by default; the wide fused-FFN variant runs 4, each moving a quarter of the columns
// and then 8 of the quarter's rows)
template <int NCOLS, int NP = 2>
__device__ __forceinline__ void epi_tmem_to_stage(uint32_t tmem_acc, int q, int hf, int lane, float* stage_q) {
#pragma unroll
    for (int cc = 0; cc < NCOLS / 32 / NP; ++cc) {
        const int c = hf * (NCOLS / 32 / NP) + cc;
        uint32_t r[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        float* dst = stage_q + lane * TC_LDS + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + 4 * j) =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    }
}
template <int EW = TC_EPI_WARPS>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory"); }

__device__ __forceinline__ void warp_sum_ilp(float (&s)[EPI_ILP]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int u = 0; u < EPI_ILP; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
}

// Output row map r -> (r / S_in) * stride_b + (r % S_in) * stride_s + off, stepped row by row (one
// integer division per warp instead of one per row: the epilogue is instruction-bound otherwise).
struct RowStep {
    int rs, S_in; int64_t cur, stride_s, wrap;     // wrap = stride_b - S_in * stride_s: added when rs wraps
    template <class P>
    __device__ __forceinline__ RowStep(const P& p, int r) {
        S_in = p.S_in; stride_s = p.stride_s; wrap = p.stride_b - (int64_t)p.S_in * p.stride_s;
        const int rb = r / S_in;
        rs = r - rb * S_in;
        cur = (int64_t)rb * p.stride_b + (int64_t)rs * p.stride_s + p.off;
    }
    __device__ __forceinline__ int64_t next() {     // returns the mapped row, then advances by one input row
        const int64_t o = cur;
        cur += stride_s;
        if (++rs == S_in) { rs = 0; cur += wrap; }
        return o;
    }
};

// What a LayerNorm row pass reads from global memory that does not depend on the accumulator: bias / gamma / beta slices and
// the residual rows of the pass's first batch.  Issued BEFORE the wait for the accumulator (all epilogue warps otherwise sit
// through the same ~1-2 K cycles of load latency together, after the accumulator is ready and with nothing to overlap it).
struct LnPre { float4 bias, ga, be; float4 rs[EPI_ILP]; };
template <class P>
__device__ __forceinline__ void epi_ln_prefetch(const P& p, int row0, int nrows, int lane, LnPre& pre) {
    const int col = lane * 4;
    pre.bias = p.bias ? *reinterpret_cast<const float4*>(p.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    pre.ga = *reinterpret_cast<const float4*>(p.gamma + col);
    pre.be = *reinterpret_cast<const float4*>(p.beta + col);
    const int rows = min(nrows, p.M - row0);
    const float* res = p.res + (int64_t)row0 * D + col;
#pragma unroll
    for (int u = 0; u < EPI_ILP; ++u)
        pre.rs[u] = u < rows ? *reinterpret_cast<const float4*>(res + (int64_t)u * D) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// out[map(r)] = LN(stage[r] + bias + res[r]) * gamma + beta for the warp's 32 rows (N == 128)
// a2 != nullptr: the normalised rows are also written as bf16 into a 128B-swizzled K-major operand tile (two K slabs of
// [128 rows x 64]); a2_row0 = tile-local index of the warp's first row
template <class P>
__device__ __forceinline__ void epi_rows_ln(const P& p, const float* stage, int row0, int nrows, int lane, uint8_t* a2 = nullptr, int a2_row0 = 0,
                                            const LnPre* pre = nullptr) {
    const int col = lane * 4;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), ga, be;
    if (pre) { bias = pre->bias; ga = pre->ga; be = pre->be; }
    else {
        if (p.bias) bias = *reinterpret_cast<const float4*>(p.bias + col);
        ga = *reinterpret_cast<const float4*>(p.gamma + col);
        be = *reinterpret_cast<const float4*>(p.beta + col);
    }
    const int rows = min(nrows, p.M - row0);
    if (rows <= 0) return;
    const float* res = p.res + (int64_t)row0 * D + col;
    float* const out32 = p.out_f32 ? p.out_f32 + col : nullptr;
    __nv_bfloat16* const out16 = p.out_b16 ? p.out_b16 + col : nullptr;
    const int64_t ld32 = p.ld_f32, ld16 = p.ld_b16;
    const float eps = p.eps;
    RowStep map(p, row0);
    for (int i0 = 0; i0 < rows; i0 += EPI_ILP) {
        float4 v[EPI_ILP];
        float s[EPI_ILP];
#pragma unroll
        for (int u = 0; u < EPI_ILP; ++u) {
            const bool ok = i0 + u < rows;
            const float4 rs = (pre && i0 == 0) ? pre->rs[u]
                            : ok ? *reinterpret_cast<const float4*>(res + (int64_t)(i0 + u) * D) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a = *reinterpret_cast<const float4*>(stage + (i0 + u) * TC_LDS + col);
            v[u] = make_float4(a.x + bias.x + rs.x, a.y + bias.y + rs.y, a.z + bias.z + rs.z, a.w + bias.w + rs.w);
            s[u] = v[u].x + v[u].y + v[u].z + v[u].w;
        }
        warp_sum_ilp(s);
#pragma unroll
        for (int u = 0; u < EPI_ILP; ++u) {
            const float mean = s[u] * (1.0f / D);
            v[u].x -= mean; v[u].y -= mean; v[u].z -= mean; v[u].w -= mean;
            s[u] = v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
        }
        warp_sum_ilp(s);
#pragma unroll
        for (int u = 0; u < EPI_ILP; ++u) {
            if (i0 + u < rows) {
                const float rstd = rsqrtf(s[u] * (1.0f / D) + eps);
                const float4 o = make_float4(v[u].x * rstd * ga.x + be.x, v[u].y * rstd * ga.y + be.y, v[u].z * rstd * ga.z + be.z, v[u].w * rstd * ga.w + be.w);
                const int64_t orow = p.out_rows ? (int64_t)p.out_rows[row0 + i0 + u] : map.next();
                if (out32) *reinterpret_cast<float4*>(out32 + orow * ld32) = o;
                if (out16) *reinterpret_cast<uint2*>(out16 + orow * ld16) = pack_bf16x4(o);
                if (a2) {     // column 4 lane .. 4 lane + 3: slab lane / 16, 16-byte chunk (lane % 16) / 2 (XOR row & 7), half lane & 1
                    const int rt = a2_row0 + i0 + u;
                    *reinterpret_cast<uint2*>(a2 + (lane >> 4) * TC_SLAB_BYTES + rt * 128 + (((((lane & 15) >> 1) ^ (rt & 7))) << 4) + ((lane & 1) << 3)) = pack_bf16x4(o);
                }
            }
        }
    }
}

// the KV-append fields exist in TcGemmParams only (the fused FFN shares this epilogue and has none)
__device__ __forceinline__ bool epi_kv_append(const TcGemmParams& p) { return p.kv_append != 0; }
__device__ __forceinline__ bool epi_kv_tok_major(const TcGemmParams& p) { return p.kv_append == 2; }
__device__ __forceinline__ const int* epi_kv_step(const TcGemmParams& p) { return p.step; }
__device__ __forceinline__ int epi_kv_heads(const TcGemmParams& p) { return p.kv_heads; }
__device__ __forceinline__ __nv_bfloat16* epi_kv_pool(const TcGemmParams& p) { return p.kv_pool; }
__device__ __forceinline__ const int* epi_kv_bt(const TcGemmParams& p) { return p.block_table; }
__device__ __forceinline__ int epi_kv_pps(const TcGemmParams& p) { return p.pps; }
template <class P> __device__ __forceinline__ bool epi_kv_append(const P&) { return false; }
template <class P> __device__ __forceinline__ bool epi_kv_tok_major(const P&) { return false; }
template <class P> __device__ __forceinline__ const int* epi_kv_step(const P&) { return nullptr; }
template <class P> __device__ __forceinline__ int epi_kv_heads(const P&) { return 1; }
template <class P> __device__ __forceinline__ __nv_bfloat16* epi_kv_pool(const P&) { return nullptr; }
template <class P> __device__ __forceinline__ const int* epi_kv_bt(const P&) { return nullptr; }
template <class P> __device__ __forceinline__ int epi_kv_pps(const P&) { return 0; }

// out[map(r)][n0 + ...] = act(stage[r] + bias) for the warp's 32 rows; split > 0 partials are raw sums
template <class P>
__device__ __forceinline__ void epi_rows_store(const P& p, const float* stage, int row0, int nrows, int n0, int split, int lane) {
    const int col = n0 + lane * 4;
    if (col >= p.N) return;
    const int rows = min(nrows, p.M - row0);
    if (rows <= 0) return;
    const bool raw = p.splits > 1;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias && !raw) bias = *reinterpret_cast<const float4*>(p.bias + col);
    const bool relu = p.act == 1 && !raw;
    const float* st = stage + lane * 4;
    if (p.head_major) {
        const int kv = col / D, cc = col % D;
        const int h = cc / p.hm_dh, hd = cc % p.hm_dh, dh = p.hm_dh;
        const int S = (int)p.hm_rows;
        int b = row0 / S, j = row0 - b * S;
        const int64_t plane = (int64_t)S * dh;                         // one (spectrum, K|V, head) plane
        int64_t o = (((int64_t)b * 2 + kv) * p.hm_heads + h) * plane + (int64_t)j * dh + hd;
        const int64_t wrap = 2 * (int64_t)p.hm_heads * plane - plane;   // from the end of one spectrum's plane to the next spectrum's
#pragma unroll 4
        for (int i = 0; i < rows; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(st + i * TC_LDS);
            const float4 v = make_float4(a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w);
            if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + o) = v;
            if (p.out_b16) *reinterpret_cast<uint2*>(p.out_b16 + o) = pack_bf16x4(v);
            o += dh;
            if (++j == S) { j = 0; o += wrap; }
        }
        return;
    }
    if (epi_kv_append(p) && col >= D) {      // K | V of position t -> cache page (4 consecutive head dims per lane: 8 bytes)
        const int t = *epi_kv_step(p);
        const int kv = (col - D) / D, cc = col % D, H = epi_kv_heads(p), dh = D / H;
        const int h = cc / dh, d0 = cc % dh;
        __nv_bfloat16* const pool = epi_kv_pool(p);
        const int* bt = epi_kv_bt(p);
        const int pps = epi_kv_pps(p);
        // head-major page [K|V][H][16][dh]; token-major page [16][K|V][H][dh] (a warp then writes one 256-byte run per row)
        const int off = epi_kv_tok_major(p) ? ((t % PAGE_TOKENS) * 2 + kv) * D + cc : ((kv * H + h) * PAGE_TOKENS + (t % PAGE_TOKENS)) * dh + d0;
        // the rows' page numbers in one load (lane i holds row i's): a block-table load per row put an L2 round trip in front of
        // every batch of stores
        const int my_page = lane < rows ? bt[(int64_t)(row0 + lane) * pps + t / PAGE_TOKENS] : 0;
        const unsigned live = __activemask();      // (N = 384 here: all lanes; lanes past N have returned above)
#pragma unroll 4
        for (int i = 0; i < rows; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(st + i * TC_LDS);
            const float4 v = make_float4(a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w);
            const int64_t page = __shfl_sync(live, my_page, i);
            *reinterpret_cast<uint2*>(pool + page * (2 * PAGE_TOKENS * D) + off) = pack_bf16x4(v);
        }
        return;
    }
    float* const out32 = p.out_f32 ? p.out_f32 + (int64_t)split * p.part_stride + col : nullptr;
    __nv_bfloat16* const out16 = p.out_b16 ? p.out_b16 + col : nullptr;
    const int64_t ld32 = p.ld_f32, ld16 = p.ld_b16;
    RowStep map(p, row0);
#pragma unroll 4
    for (int i = 0; i < rows; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(st + i * TC_LDS);
        float4 o = make_float4(a.x + bias.x, a.y + bias.y, a.z + bias.z, a.w + bias.w);
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        const int64_t orow = map.next();
        if (out32) *reinterpret_cast<float4*>(out32 + orow * ld32) = o;
        if (out16) *reinterpret_cast<uint2*>(out16 + orow * ld16) = pack_bf16x4(o);
    }
}

// ------------------------------------------------------------------ the GEMM
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS) gemm_bf16_tc(const __grid_constant__ TcGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ __align__(8) uint64_t chain_w_bar, chain_a_bar, chain_acc_bar;
    __shared__ uint32_t tmem_slot;

    // 1024-byte alignment by pointer arithmetic on the shared array (keeps the shared address space: LDS/STS, not generic)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const bool chain = EPI == TC_EPI_LN && p.chain;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    const int split = blockIdx.z;
    const int kb_total = p.K / TC_BK;
    const int kb_per = (kb_total + p.splits - 1) / p.splits;
    const int kb_begin = split * kb_per;
    const int kb_end = min(kb_total, kb_begin + kb_per);
    const int num_kb = kb_end - kb_begin;      // host guarantees >= 1

    const int stage_bytes = p.wsplit ? TC_STAGE_BYTES_WSPLIT : TC_STAGE_BYTES;
    // chained projection: its operands live behind the stage ring / staging tile
    uint8_t* sC = smem + ((max(p.stages * stage_bytes, TC_STAGING_BYTES) + 1023) & ~1023);
    const uint32_t tmem_cols = chain ? 2 * TC_BN : TC_BN;
    pdl_launch_dependents();     // programmatic dependent launch (un-fused decode step): barrier / TMEM set-up overlaps the predecessor's tail
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmW);
        if (p.wsplit) tma_prefetch_desc(&p.tmW2);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        if (chain) { tma_prefetch_desc(&p.tmC); if (p.chain == 2) tma_prefetch_desc(&p.tmC2); mbar_init(&chain_w_bar, 1); mbar_init(&chain_a_bar, TC_EPI_WARPS); mbar_init(&chain_acc_bar, 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            if (chain) {             // the chained weights are decode-loop constants: in flight before the PDL wait
                mbar_arrive_expect_tx(&chain_w_bar, (p.chain == 2 ? 4 : 2) * TC_SLAB_BYTES);
                for (int ks = 0; ks < 2; ++ks) {
                    tma_load_2d(sC + (2 + ks) * TC_SLAB_BYTES, &p.tmC, &chain_w_bar, ks * TC_BK, 0);
                    if (p.chain == 2) tma_load_2d(sC + (4 + ks) * TC_SLAB_BYTES, &p.tmC2, &chain_w_bar, ks * TC_BK, 0);
                }
            }
            pdl_wait();              // A is the predecessor's output
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % p.stages;
                const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                uint8_t* a = smem + (size_t)s * stage_bytes;
                tma_load_2d(a, &p.tmA, &full_bar[s], (kb_begin + i) * TC_BK, m0);
                tma_load_2d(a + TC_SLAB_BYTES, &p.tmW, &full_bar[s], (kb_begin + i) * TC_BK, n0);
                if (p.wsplit) tma_load_2d(a + 2 * TC_SLAB_BYTES, &p.tmW2, &full_bar[s], (kb_begin + i) * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(TC_BM, TC_BN);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % p.stages;
                const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t adesc = umma_desc_sw128(a_addr);
                const uint64_t bdesc = umma_desc_sw128(a_addr + TC_SLAB_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)   // +32 B per K step inside the 128 B swizzle row (encoded >> 4)
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (i > 0 || k > 0) ? 1u : 0u);
                if (p.wsplit) {
                    const uint64_t bdesc2 = umma_desc_sw128(a_addr + 2 * TC_SLAB_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc2 + (uint64_t)(2 * k), idesc, 1u);
                }
                umma_commit(&empty_bar[s]);            // smem slot reusable once these MMAs retire
            }
            umma_commit(&tmem_full_bar);               // accumulator complete
            if (chain) {     // second product on the normalised rows the epilogue warps put into shared memory (same K-slab / term order as a stand-alone launch)
                mbar_wait(&chain_w_bar, 0);
                mbar_wait(&chain_a_bar, 0);
                tc_fence_after();
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sC + ks * TC_SLAB_BYTES));
                    for (int t2 = 0; t2 < p.chain; ++t2) {
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(sC + (2 + 2 * t2 + ks) * TC_SLAB_BYTES));
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            umma_bf16(tmem_base + TC_BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ks > 0 || t2 > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&chain_acc_bar);
            }
        }
    } else {
        // ---------------- epilogue: 8 warps, warp (id % 4) owns TMEM lanes [32*(id%4), +32), two warps per quarter
        const int q = warp & 3, hf = (warp - 2) >> 2;
        float* stage_q = reinterpret_cast<float*>(smem) + (q * 32) * TC_LDS;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        pdl_wait();                  // residual reads / output writes below: the predecessor has completed (satisfied: A was loaded after it)
        epi_tmem_to_stage<TC_BN>(tmem_base, q, hf, lane, stage_q);
        epi_bar_sync();
        const float* st = stage_q + (hf * 16) * TC_LDS;
        // (no LnPre here: measured, the prefetch registers cost this kernel its second resident CTA -- 28.9 -> 42 us per launch)
        if (EPI == TC_EPI_LN) epi_rows_ln(p, st, m0 + q * 32 + hf * 16, 16, lane, chain ? sC : nullptr, q * 32 + hf * 16);
        else epi_rows_store(p, st, m0 + q * 32 + hf * 16, 16, n0, split, lane);
        if (chain) {
            fence_proxy_async_smem();          // the A2 rows (generic-proxy stores) -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(&chain_a_bar);
            mbar_wait(&chain_acc_bar, 0);      // (implies every epilogue warp is done with the staging tile)
            tc_fence_after();
            epi_tmem_to_stage<TC_BN>(tmem_base + TC_BN, q, hf, lane, stage_q);
            epi_bar_sync();
            const int r0 = m0 + q * 32 + hf * 16, rows = min(16, p.M - r0);
            const float4 cb = *reinterpret_cast<const float4*>(p.chain_bias + lane * 4);
            for (int i = 0; i < rows; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(st + i * TC_LDS + lane * 4);
                *reinterpret_cast<float4*>(p.chain_out + (int64_t)(r0 + i) * p.ld_chain + lane * 4) = make_float4(a.x + cb.x, a.y + cb.y, a.z + cb.z, a.w + cb.w);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// fp32 -> bf16 row pack with an optional gather (row r of the source starts at src + row_off[r])
__global__ void __launch_bounds__(256) pack_rows_bf16(const float* src, const int64_t* row_off, int64_t lds, int64_t rows, __nv_bfloat16* dst) {
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* s = row_off ? src + row_off[r] : src + r * lds;
    const float4 v = *reinterpret_cast<const float4*>(s + lane * 4);
    *reinterpret_cast<uint2*>(dst + r * D + lane * 4) = pack_bf16x4(v);
}

// two-term bf16 split of an fp32 array: hi = bf16(x), lo = bf16(x - hi)
__global__ void f32_to_bf16_split(const float* in, int64_t n, __nv_bfloat16* hi, __nv_bfloat16* lo) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float x = in[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        hi[i] = h;
        lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}

}  // namespace mmt
