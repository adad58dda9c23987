// Fused transformer FFN on the sm_100a tensor cores:
//
//   out = epi( relu(X . W1^T + b1) . W2^T )        X [M,128] bf16, W1 [F,128], W2 [128,F] bf16 (K contiguous)
//
// The F-wide hidden activation never leaves the SM.  A CTA owns 128 rows of X (resident in shared
// memory) and walks its share of F in chunks of 64 columns:
//
//   GEMM1   acc1[c&1] (TMEM)  = X . [W1_hi[c] ; W1_lo[c]]^T         tcgen05.mma 128x128x16, K = 128
//   convert H[c&1] (smem, bf16, 128B-swizzled K-major) = relu(acc1_hi + acc1_lo + b1[c])    8 epilogue warps
//   GEMM2   acc2 (TMEM)      += H[c&1] . [W2_hi[:, c] ; W2_lo[:, c]]^T   tcgen05.mma 128x256x16, K = 64
//
// Weights are the two-term bf16 split W_hi + W_lo (kernels_tc.cuh).  The two terms are stacked along the MMA's
// N dimension (hi rows, then lo rows, of one shared-memory tile) instead of being issued as two MMA passes: the
// chunk loop is bound by shared-memory operand reads (measured 2.1 K cycles per chunk vs 1.0 K of tensor time;
// halving the L2 weight stream with two row tiles per CTA changed nothing), and stacking reads every A slice
// (X, H) once instead of twice.  The hi and lo halves of the accumulators are summed by the epilogue warps.
//
// Software pipeline: GEMM1 of chunk c+1 runs on the tensor core while the epilogue warps convert chunk c
// (acc1 and H are double-buffered).  W1 and W2 chunks travel through two independent two-stage TMA rings fed by
// two producer lanes: a W1 slot is free as soon as GEMM1 of its chunk retires, so the next chunks' weights are
// always in flight while the tensor core works.
//
// grid = (splits, ceil(M/128)).  splits == 1: the epilogue is bias + residual + LayerNorm over the
// 128-wide row (encoder layers, large decode batches).  splits > 1 (small decode batches: more CTAs
// than M/128): each CTA handles F/64/splits chunks and writes a raw fp32 partial; the consumer
// (decode_attn / sample_tokens prologue, or bias_res_layernorm) reduces them in a fixed order.
//
// warp 0: TMA producer (one lane) | warp 1: GEMM1 issuer (one lane) + TMEM alloc (512 cols) | warps 2-9: epilogue
// (two epilogue warps per TMEM lane quarter: with one warp per scheduler the conversion is a latency chain) | warp 10:
// GEMM2 issuer (one lane).  Two issuing threads because the chunk loop is bound by the ISSUER's own serial control
// instructions, not by the tensor pipe: per 64-column chunk one thread spent ~1.1 K cycles in three mbarrier waits, two
// tcgen05 fences and the commits (measured with every MMA, TMA load and conversion removed, profiles/r02_ffn_pipeline.md)
// around ~0.5 K cycles of MMA time; GEMM1 and GEMM2 write different accumulators, so their issue streams are independent.
//
// Two variants (template WS).  WS = 1: two-term weights as above, acc1 / H double-buffered, two-stage weight rings
// (TMEM: 2 x 128 + 256 columns; 192 KB of shared memory).  WS = 0: the hi term alone (decoder FFN).  The
// freed TMEM columns and shared memory buy a deeper pipeline: acc1 / H triple-buffered and GEMM1 issued TWO chunks
// ahead of GEMM2, so the ~900-cycle conversion of a chunk (TMEM -> bias + ReLU -> bf16 -> swizzled smem -> proxy fence)
// no longer sits between the two GEMMs of the same chunk on the tensor pipe, and four-stage weight rings of 16 KB
// stages keep two more chunks of weights in flight (measured chunk cadence of the two-buffer pipeline: ~1.6 K cycles
// against ~0.7 K of operand-read / tensor time, profiles/r02_ffn_pipeline.md).
#pragma once
#include "kernels_tc.cuh"

namespace mmt {

constexpr int FF_CH = 64;                                  // hidden columns per chunk
constexpr int FF_THREADS = TC_THREADS + 32;                // producer + GEMM1 issuer + 8 epilogue warps + GEMM2 issuer
// (the epilogue warps start at warp 2: TMEM lane quarter = warp id % 4; the GEMM2 issuer is the warp after them)
template <int WS, int WIDE> constexpr int ff_threads();    // 64 + 32 * epilogue warps + 32
constexpr int FF_X_BYTES = 2 * TC_SLAB_BYTES;              // X: two K slabs of [128 rows x 64]
constexpr int FF_H_BYTES = TC_SLAB_BYTES;                  // one H buffer: [128 rows x 64] bf16
constexpr int FF_W1_HALF = FF_CH * 128;                    // 8 KB: 64 rows x 128 B (one term of one K slab)
constexpr int FF_W1_STAGE = 2 * TC_SLAB_BYTES;             // W1 chunk: 2 K slabs of [64 hi rows ; 64 lo rows] x 64 k = 32 KB
constexpr int FF_W2_STAGE = 2 * TC_SLAB_BYTES;             // W2 chunk: [128 hi rows ; 128 lo rows] x 64 k = 32 KB
constexpr int FF_MAX_F = 2048;
// per-variant pipeline geometry.  WS: two-term weights (1) or the hi term alone (0).  WIDE: chunks of 128 hidden columns instead of
// 64 (hi term + LayerNorm epilogue only): the chunk loop is paced by the issuing threads' control instructions per chunk (two
// waits, a fence, a commit: ~0.7 K cycles), so twice the columns per chunk halves that cost per column -- at 128 columns a chunk
// carries ~1.0 K cycles of tensor work and the loop becomes tensor-paced.  Needs the TMEM columns and shared memory the lo term
// vacates, and whole 128-column chunks per CTA (no split-F).
template <int WS, int WIDE> struct FfCfg;
template <> struct FfCfg<1, 0> {       // two-term weights
    static constexpr int CH = FF_CH, NB = 2, NS = 2;                       // chunk width, acc1 / H buffers, weight ring stages
    static constexpr int EW = TC_EPI_WARPS;                                // epilogue warps
    static constexpr int W1_STAGE = FF_W1_STAGE, W2_STAGE = FF_W2_STAGE;   // 32 KB each (hi | lo)
    static constexpr int W1_SLAB = TC_SLAB_BYTES;                          // K slab stride inside a W1 stage
    static constexpr int W2_KSLAB = 2 * TC_SLAB_BYTES;                     // 64-k slab stride inside a W2 stage (one slab: hi rows | lo rows)
    static constexpr int H_BYTES = FF_H_BYTES;
    static constexpr int ACC1_STRIDE = 2 * FF_CH, ACC2_COL = 256;          // TMEM columns
};
template <> struct FfCfg<0, 0> {       // hi term only, 64-column chunks (split-F partial path): deeper pipeline in the freed TMEM / shared memory
    static constexpr int CH = FF_CH, NB = 3, NS = 4;
    static constexpr int EW = TC_EPI_WARPS;
    static constexpr int W1_STAGE = FF_W1_STAGE / 2, W2_STAGE = FF_W2_STAGE / 2;   // 16 KB each
    static constexpr int W1_SLAB = FF_W1_HALF;                             // 8 KB: [64 rows x 64 k]
    static constexpr int W2_KSLAB = TC_SLAB_BYTES;
    static constexpr int H_BYTES = FF_H_BYTES;
    static constexpr int ACC1_STRIDE = FF_CH, ACC2_COL = 256;
};
template <> struct FfCfg<0, 1> {       // hi term only, 128-column chunks (LayerNorm epilogue, whole F per CTA)
    static constexpr int CH = 2 * FF_CH, NB = 2, NS = 2;
    // 16 epilogue warps (four per TMEM lane quarter): per 128-row tile the epilogue warps carry ~32 K cycles of latency-bound
    // work (two LayerNorm passes of 8 K cycles at 16 rows per warp, 16 conversions) against ~17 K cycles of tensor time --
    // with twice the warps every one of those chains is half as long
    static constexpr int EW = 2 * TC_EPI_WARPS;
    static constexpr int W1_STAGE = 2 * TC_SLAB_BYTES, W2_STAGE = 2 * TC_SLAB_BYTES;   // [128 rows x 128 k] each: two 64-k slabs of 16 KB
    static constexpr int W1_SLAB = TC_SLAB_BYTES;
    static constexpr int W2_KSLAB = TC_SLAB_BYTES;
    static constexpr int H_BYTES = 2 * FF_H_BYTES;                         // two 64-k slabs
    static constexpr int ACC1_STRIDE = 2 * FF_CH, ACC2_COL = 256;
};
template <int WS, int WIDE> constexpr int ff_threads() { return 64 + 32 * FfCfg<WS, WIDE>::EW + 32; }
// dynamic smem: X (32 KB) | H (NB buffers) | W1 ring | W2 ring + alignment slack
template <int WS, int WIDE = 0> constexpr int ff_smem_bytes() {
    typedef FfCfg<WS, WIDE> C;
    return FF_X_BYTES + C::NB * C::H_BYTES + C::NS * (C::W1_STAGE + C::W2_STAGE) + 1024;
}
constexpr int FF_SMEM_BYTES = ff_smem_bytes<1>();
static_assert(FfCfg<1, 0>::NS * (FfCfg<1, 0>::W1_STAGE + FfCfg<1, 0>::W2_STAGE) >= TC_STAGING_BYTES, "final staging tile aliases the weight rings");
static_assert(FfCfg<0, 0>::NS * (FfCfg<0, 0>::W1_STAGE + FfCfg<0, 0>::W2_STAGE) >= TC_STAGING_BYTES, "final staging tile aliases the weight rings");
static_assert(FfCfg<0, 1>::NS * (FfCfg<0, 1>::W1_STAGE + FfCfg<0, 1>::W2_STAGE) >= TC_STAGING_BYTES, "final staging tile aliases the weight rings");
static_assert(ff_smem_bytes<0>() <= 227 * 1024 && ff_smem_bytes<1>() <= 227 * 1024 && ff_smem_bytes<0, 1>() <= 227 * 1024, "shared memory budget");
constexpr uint32_t FF_TMEM_COLS = 512;                     // WS=1: acc1 2 x (64 hi + 64 lo) | acc2 128 hi + 128 lo;  WS=0: acc1 3 x 64 | acc2 128 at column 256

struct FfnParams {
    CUtensorMap tmX;                 // X  [M,128] bf16, box {64,128}
    CUtensorMap tmW1, tmW1lo;        // W1 [F,128] bf16, box {64,64}
    CUtensorMap tmW2, tmW2lo;        // W2 [128,F] bf16, box {64,128}
    int wsplit;                      // 0: single-term weights (the lo halves are skipped)
    int M, N, F;                     // N == 128
    int splits;
    const float* b1;                 // [F]
    int64_t part_stride;             // floats between split partials
    const float* bias; int act;      // b2 (LayerNorm epilogue only); act unused (0)
    float* out_f32; int64_t ld_f32;
    __nv_bfloat16* out_b16; int64_t ld_b16;
    int head_major, hm_heads, hm_dh; int64_t hm_rows;   // unused (0); keeps the shared epilogue templates happy
    const float* res; const float* gamma; const float* beta; float eps;
    int S_in; int64_t stride_b, stride_s, off;
    const int* out_rows;             // optional explicit output rows (LayerNorm epilogue, ragged encoder)
    long long* dbg;                  // optional [CTA][16] phase timestamps (MMT_DA_DEBUG)
    // Prologue (LayerNorm-epilogue variant only): X is not the FFN input but the attention output of the block before it;
    // the kernel first computes  x = LN(res + X . Wp^T + pro_bias) * pro_gamma + pro_beta  (the decoder's cross-attention
    // out-projection + norm2), writes it to pro_out (fp32, the residual of the FFN's own LayerNorm) and to the X tile in shared
    // memory (bf16), then runs the FFN on it: one launch and one HBM round trip of the activations less per layer.
    int pro;                         // 0 none, 1 Wp hi term only, 2 both terms
    CUtensorMap tmP, tmPlo;          // Wp [128,128] bf16 hi / lo, box {64,128}
    const float *pro_bias, *pro_gamma, *pro_beta; float* pro_out;
    int knock;                       // timing experiments only (MMT_FFN_KNOCK, results garbage): 1 no weight TMA after the first ring fill,
                                     // 2 no conversion work, 4 no MMAs
};


// TMEM -> staging tile for an accumulator stored as two 128-column halves (hi | lo): the halves are summed on the way
__device__ __forceinline__ void epi_tmem2_to_stage(uint32_t tmem_acc, int q, int hf, int lane, float* stage_q) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int c = hf * 2 + cc;
        uint32_t a[32], b[32];
        tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), a);
        tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(TC_BN + c * 32), b);
        tmem_ld_wait();
        float* dst = stage_q + lane * TC_LDS + c * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(dst + 4 * j) =
                make_float4(__uint_as_float(a[4 * j]) + __uint_as_float(b[4 * j]), __uint_as_float(a[4 * j + 1]) + __uint_as_float(b[4 * j + 1]),
                            __uint_as_float(a[4 * j + 2]) + __uint_as_float(b[4 * j + 2]), __uint_as_float(a[4 * j + 3]) + __uint_as_float(b[4 * j + 3]));
    }
}

// the fields epi_rows_ln reads, for the prologue's LayerNorm
struct LnView {
    const float *bias, *gamma, *beta, *res; float* out_f32; __nv_bfloat16* out_b16; int64_t ld_f32, ld_b16; int M; float eps;
    const int* out_rows; int S_in; int64_t stride_b, stride_s, off;
};

template <int EPI, int WS, int WIDE = 0>
__global__ void __launch_bounds__((ff_threads<WS, WIDE>()), 1) ffn_fused_tc(const __grid_constant__ FfnParams p) {
    static_assert(!WIDE || (WS == 0 && EPI == TC_EPI_LN), "128-column chunks: hi term + LayerNorm epilogue only");
    typedef FfCfg<WS, WIDE> C;
    constexpr int NB = C::NB, NS = C::NS, CH = C::CH;
    constexpr int EW = C::EW, NP = EW / 4, RPW = 32 / NP;      // epilogue warps, warps per TMEM lane quarter, rows per warp in the row passes
    constexpr int FF_G2_WARP = 2 + EW;
    extern __shared__ uint8_t smem_raw[];
    // "GEMM1 of chunk i retired" frees a W1 stage AND publishes acc1; "GEMM2 of chunk j retired" frees a W2 stage AND an H
    // buffer: one tcgen05.commit each (a commit costs the issuing thread a few hundred cycles), on barrier rings of
    // R = lcm(NS, NB) slots so that every waiter (whatever its buffer count) finds chunk i's barrier at slot i % R
    constexpr int R = 12;
    static_assert(R % NS == 0 && R % NB == 0, "barrier ring must be a multiple of both buffer counts");
    __shared__ __align__(8) uint64_t x_full, w1_full[NS], w2_full[NS], h_full[NB], acc1_free[NB], g1_done[R], g2_done[R], acc2_full;
    __shared__ __align__(8) uint64_t pro_w_full, pro_acc_full, x2_ready;
    const bool pro = EPI == TC_EPI_LN && p.pro;
    __shared__ uint32_t tmem_slot;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays in the shared address space
    uint8_t* sX = smem;
    uint8_t* sH = sX + FF_X_BYTES;
    uint8_t* sW = sH + NB * C::H_BYTES;   // W1 ring; the two rings together also hold the final staging tile
    uint8_t* sW2 = sW + NS * C::W1_STAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, m0 = blockIdx.y * TC_BM;
    const int n = (p.F / CH) / p.splits;         // chunks of this CTA (host guarantees divisibility, n >= 1)
    const int c0 = split * n;
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
#define FF_STAMP(i) do { if (p.dbg) p.dbg[cta * 16 + (i)] = clock64(); } while (0)
    if (threadIdx.x == 64) FF_STAMP(0);
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmX); tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmW2);
        if (WS) { tma_prefetch_desc(&p.tmW1lo); tma_prefetch_desc(&p.tmW2lo); }
        mbar_init(&x_full, 1); mbar_init(&acc2_full, 1);
        for (int s = 0; s < NS; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w2_full[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&acc1_free[s], EW); mbar_init(&h_full[s], EW); }     // one arrival per epilogue warp (256 arrivals on one barrier serialise)
        for (int s = 0; s < R; ++s) { mbar_init(&g1_done[s], 1); mbar_init(&g2_done[s], 1); }
        if (pro) { tma_prefetch_desc(&p.tmP); if (p.pro == 2) tma_prefetch_desc(&p.tmPlo); mbar_init(&pro_w_full, 1); mbar_init(&pro_acc_full, 1); mbar_init(&x2_ready, EW); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, FF_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tmem_acc2 = tmem_base + C::ACC2_COL;   // acc1[b] at columns [ACC1_STRIDE b, +ACC1_STRIDE): hi 64 (| lo 64)
    // N of the MMAs: both terms stacked, or the hi half alone
    constexpr int n1 = WS ? 2 * CH : CH, n2 = WS ? 2 * TC_BN : TC_BN;
    if (threadIdx.x == 64) FF_STAMP(1);

    if (warp == 0) {
        auto load_x = [&]() {
            mbar_arrive_expect_tx(&x_full, FF_X_BYTES);
            tma_load_2d(sX, &p.tmX, &x_full, 0, m0);
            tma_load_2d(sX + TC_SLAB_BYTES, &p.tmX, &x_full, TC_BK, m0);
        };
        // ONE producer lane feeds both weight rings.  (Two lanes of this warp each running its own loop of blocking mbarrier
        // waits stall each other: a lane parked in try_wait holds the warp, measured ~1.1 K cycles per chunk with nothing
        // else in the loop.)  Order per chunk: W1(i) as soon as GEMM1(i - NS) retired, then W2(i) once GEMM2(i - NS) retired.
        if (lane == 0) {
            if (pro) {
                // prologue weights (decode-loop constants) into the still idle W2 ring, then the attention output tile; the
                // weight rings start only when the prologue is over: its staging tile lies over them
                mbar_arrive_expect_tx(&pro_w_full, (p.pro == 2 ? 4 : 2) * TC_SLAB_BYTES);
                for (int ks = 0; ks < 2; ++ks) {
                    tma_load_2d(sW2 + ks * TC_SLAB_BYTES, &p.tmP, &pro_w_full, ks * TC_BK, 0);
                    if (p.pro == 2) tma_load_2d(sW2 + (2 + ks) * TC_SLAB_BYTES, &p.tmPlo, &pro_w_full, ks * TC_BK, 0);
                }
                pdl_wait(); load_x();
                mbar_wait(&x2_ready, 0);
            }
            // W1(i) is needed LA = NB - 1 chunks before W2(i) (GEMM1 runs that far ahead of GEMM2), and its stage frees earlier
            // (GEMM1(i - NS) retires before GEMM2(i - NS - LA)): the loop issues W1 LA chunks ahead so that no wait of one ring
            // delays a load of the other.
            constexpr int LA_P = NB - 1;
            for (int it = 0; it < n + LA_P; ++it) {
                if (it < n) {
                    const int i = it, s = i % NS, c = c0 + i;
                    if (!pro && i == min(n, NS)) { pdl_wait(); load_x(); }      // the first ring fill is decode-loop constants only: ahead of the PDL wait
                    if (i >= NS) mbar_wait(&g1_done[(i - NS) % R], ((uint32_t)((i - NS) / R)) & 1u);     // GEMM1 of the stage's previous chunk retired
                    if (!((p.knock & 1) && i >= NS)) {
                        mbar_arrive_expect_tx(&w1_full[s], C::W1_STAGE);
                        uint8_t* w = sW + (size_t)s * C::W1_STAGE;          // K slab ks at ks * W1_SLAB: rows 0-63 hi (, rows 64-127 lo)
#pragma unroll
                        for (int rh = 0; rh < CH / FF_CH; ++rh) {           // the W1 map's box is 64 rows: a 128-column chunk takes two per K slab
                            tma_load_2d(w + rh * FF_W1_HALF, &p.tmW1, &w1_full[s], 0, c * CH + rh * FF_CH);
                            tma_load_2d(w + C::W1_SLAB + rh * FF_W1_HALF, &p.tmW1, &w1_full[s], TC_BK, c * CH + rh * FF_CH);
                        }
                        if (WS) {
                            tma_load_2d(w + FF_W1_HALF, &p.tmW1lo, &w1_full[s], 0, c * FF_CH);
                            tma_load_2d(w + C::W1_SLAB + FF_W1_HALF, &p.tmW1lo, &w1_full[s], TC_BK, c * FF_CH);
                        }
                    } else mbar_arrive(&w1_full[s]);
                }
                if (it >= LA_P) {
                    const int i = it - LA_P, s = i % NS, c = c0 + i;
                    if (i >= NS) mbar_wait(&g2_done[(i - NS) % R], ((uint32_t)((i - NS) / R)) & 1u);     // GEMM2 of the stage's previous chunk retired
                    if (!((p.knock & 1) && i >= NS)) {
                        mbar_arrive_expect_tx(&w2_full[s], C::W2_STAGE);
                        uint8_t* w = sW2 + (size_t)s * C::W2_STAGE;         // rows 0-127 hi (, rows 128-255 lo)
#pragma unroll
                        for (int ks = 0; ks < CH / FF_CH; ++ks) tma_load_2d(w + ks * C::W2_KSLAB, &p.tmW2, &w2_full[s], c * CH + ks * FF_CH, 0);
                        if (WS) tma_load_2d(w + TC_SLAB_BYTES, &p.tmW2lo, &w2_full[s], c * FF_CH, 0);
                    } else mbar_arrive(&w2_full[s]);
                }
            }
            if (!pro && n <= NS) { pdl_wait(); load_x(); }     // short loops never reached the in-loop wait
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_bf16(TC_BM, n1);
            const uint32_t x_addr = smem_u32(sX);
            // GEMM1 issue stream: acc1[i % NB] = X . [W1_hi (; W1_lo)][chunk i]^T.  The accumulator buffer is free once the
            // conversion of chunk i - NB has read it (h_full of that chunk).
            mbar_wait(&x_full, 0);
            if (pro) {       // acc2 = X . [Wp_hi ; Wp_lo]^T as two accumulating passes per K slab (the order of a stand-alone launch)
                mbar_wait(&pro_w_full, 0);
                tc_fence_after();
                const uint32_t idescP = umma_idesc_bf16(TC_BM, TC_BN);
                for (int ks = 0; ks < 2; ++ks) {
                    const uint64_t adesc = umma_desc_sw128(x_addr + ks * TC_SLAB_BYTES);
                    for (int t2 = 0; t2 < p.pro; ++t2) {
                        const uint64_t bdesc = umma_desc_sw128(smem_u32(sW2 + (2 * t2 + ks) * TC_SLAB_BYTES));
#pragma unroll
                        for (int k = 0; k < TC_BK / 16; ++k)
                            umma_bf16(tmem_acc2, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idescP, (ks > 0 || t2 > 0 || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(&pro_acc_full);
                mbar_wait(&x2_ready, 0);       // the epilogue warps have replaced the X tile by the normalised rows
                tc_fence_after();
            }
            for (int i = 0; i < n; ++i) {
                const int s = i % NS;
                mbar_wait(&w1_full[s], ((uint32_t)(i / NS)) & 1u);
                // acc1[i % NB] is free once every epilogue warp has READ chunk i - NB out of it (acc1_free), well before that
                // chunk's H tile is written and fenced (h_full): the round trip "conversion done -> GEMM1 two chunks on -> its
                // commit -> next conversion" was the chunk loop's critical path (1.1 K cycles per chunk with every MMA, TMA
                // load and conversion removed)
                if (i >= NB) mbar_wait(&acc1_free[i % NB], ((uint32_t)((i - NB) / NB)) & 1u);
                tc_fence_after();
                const uint32_t w1 = smem_u32(sW + (size_t)s * C::W1_STAGE);
                const uint32_t acc1 = tmem_base + (uint32_t)((i % NB) * C::ACC1_STRIDE);
#pragma unroll
                for (int k = 0; k < D / 16; ++k) {
                    const uint64_t adesc = umma_desc_sw128(x_addr + (k >> 2) * TC_SLAB_BYTES) + (uint64_t)(2 * (k & 3));
                    const uint64_t bdesc = umma_desc_sw128(w1 + (k >> 2) * C::W1_SLAB) + (uint64_t)(2 * (k & 3));
                    if (!(p.knock & 4) || i < NB) umma_bf16(acc1, adesc, bdesc, idesc1, k > 0 ? 1u : 0u);
                }
                umma_commit(&g1_done[i % R]);  // acc1 of chunk i complete; its W1 stage reusable
            }
        }
    } else if (warp == FF_G2_WARP) {
        if (lane == 0) {
            const uint32_t idesc2 = umma_idesc_bf16(TC_BM, n2);
            // GEMM2 issue stream: acc2 += H[j % NB] . [W2_hi (; W2_lo)][:, chunk j]^T
            for (int j = 0; j < n; ++j) {
                const int b = j % NB, sw = j % NS;
                mbar_wait(&w2_full[sw], ((uint32_t)(j / NS)) & 1u);
                mbar_wait(&h_full[b], ((uint32_t)(j / NB)) & 1u);
                tc_fence_after();
                const uint32_t h_addr = smem_u32(sH + (size_t)b * C::H_BYTES), w2_addr = smem_u32(sW2 + (size_t)sw * C::W2_STAGE);
#pragma unroll
                for (int k = 0; k < CH / 16; ++k) {
                    const uint64_t adesc = umma_desc_sw128(h_addr + (k >> 2) * TC_SLAB_BYTES) + (uint64_t)(2 * (k & 3));
                    const uint64_t bdesc = umma_desc_sw128(w2_addr + (k >> 2) * C::W2_KSLAB) + (uint64_t)(2 * (k & 3));
                    if (!(p.knock & 4) || j == 0) umma_bf16(tmem_acc2, adesc, bdesc, idesc2, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&g2_done[j % R]);  // W2 stage and H buffer of chunk j reusable
            }
            umma_commit(&acc2_full);
        }
    } else {
        // ---------------- epilogue warps: warp (id % 4) owns TMEM lanes [32*(id%4), +32); thread = row;
        // the NP warps of a quarter split the chunk's columns (hf = 0 .. NP - 1)
        static_assert(!WS || NP == 2, "epi_tmem2_to_stage splits the columns in halves");
        const int q = warp & 3, hf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        if (pro) {
            LnView v{p.pro_bias, p.pro_gamma, p.pro_beta, p.res, p.pro_out, nullptr, D, D, p.M, p.eps, nullptr, p.M > 0 ? p.M : 1, 0, 1, 0};
            pdl_wait();                        // the residual rows are the predecessor's output
            LnPre pre;
            epi_ln_prefetch(v, m0 + q * 32 + hf * RPW, RPW, lane, pre);     // in flight under the X load and the prologue MMAs
            mbar_wait(&pro_acc_full, 0);
            tc_fence_after();
            float* stage_p = reinterpret_cast<float*>(sW) + (q * 32) * TC_LDS;
            epi_tmem_to_stage<TC_BN, NP>(tmem_acc2, q, hf, lane, stage_p);
            epi_bar_sync<EW>();
            epi_rows_ln(v, stage_p + (hf * RPW) * TC_LDS, m0 + q * 32 + hf * RPW, RPW, lane, sX, q * 32 + hf * RPW, &pre);
            fence_proxy_async_smem();          // the rewritten X tile -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x2_ready);
        }
        for (int i = 0; i < n; ++i) {
            const int b = i % NB;
            // this warp converts columns [hf * CH / 2, + CH / 2) of its 32 rows, 32 columns at a time
            constexpr int CPW = CH / NP;
            // bias slice of this chunk (a decode-loop constant): in registers before the accumulator is ready
            float4 bias[CPW / 32][8];
#pragma unroll
            for (int cc = 0; cc < CPW / 32; ++cc) {
                const float4* bsrc = reinterpret_cast<const float4*>(p.b1 + (size_t)(c0 + i) * CH + hf * CPW + cc * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) bias[cc][j] = __ldg(bsrc + j);
            }
            mbar_wait(&g1_done[i % R], ((uint32_t)(i / R)) & 1u);
            tc_fence_after();
            if (threadIdx.x == 64 && i == 0) FF_STAMP(2);
            uint32_t pk[CPW / 2];
            if ((p.knock & 2) && i > 0) {
#pragma unroll
                for (int j = 0; j < CPW / 2; ++j) pk[j] = 0x3c003c00u;
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_free[b]);
            } else {
                // every TMEM load of the warp's columns is issued before the one wait (a load -> wait -> convert sequence per 32
                // columns put the TMEM latency on the chunk's critical path once per 32 columns)
                uint32_t r[CPW / 32][32], rl[WS ? CPW / 32 : 1][32];
#pragma unroll
                for (int cc = 0; cc < CPW / 32; ++cc) {
                    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * C::ACC1_STRIDE + hf * CPW + cc * 32);
                    tmem_ld_32x32(t0, r[cc]);
                    if (WS) tmem_ld_32x32(t0 + FF_CH, rl[cc]);
                }
                tmem_ld_wait();
                tc_fence_before();                 // the accumulator is in registers: GEMM1 of chunk i + NB may overwrite it
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_free[b]);
#pragma unroll
                for (int cc = 0; cc < CPW / 32; ++cc) {
                    const float4* bb = bias[cc];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v0 = __uint_as_float(r[cc][4 * j]), v1 = __uint_as_float(r[cc][4 * j + 1]), v2 = __uint_as_float(r[cc][4 * j + 2]), v3 = __uint_as_float(r[cc][4 * j + 3]);
                        if (WS) { v0 += __uint_as_float(rl[cc][4 * j]); v1 += __uint_as_float(rl[cc][4 * j + 1]); v2 += __uint_as_float(rl[cc][4 * j + 2]); v3 += __uint_as_float(rl[cc][4 * j + 3]); }
                        v0 = fmaxf(v0 + bb[j].x, 0.f); v1 = fmaxf(v1 + bb[j].y, 0.f); v2 = fmaxf(v2 + bb[j].z, 0.f); v3 = fmaxf(v3 + bb[j].w, 0.f);
                        __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
                        pk[cc * 16 + 2 * j] = *reinterpret_cast<uint32_t*>(&lo);
                        pk[cc * 16 + 2 * j + 1] = *reinterpret_cast<uint32_t*>(&hi);
                    }
                }
            }
            if (i >= NB) mbar_wait(&g2_done[(i - NB) % R], ((uint32_t)((i - NB) / R)) & 1u);      // GEMM2 of the buffer's previous chunk retired
#pragma unroll
            for (int cc = 0; cc < CPW / 32; ++cc) {
                const int col0 = hf * CPW + cc * 32;              // column of the chunk -> 64-k slab col0 / 64, 16-byte chunk (col0 % 64) / 8 + c4
                uint8_t* hrow = sH + (size_t)b * C::H_BYTES + (size_t)(col0 >> 6) * TC_SLAB_BYTES + (size_t)row * 128;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {    // 16-byte chunk cj of the row lands at chunk (cj ^ (row & 7)): SWIZZLE_128B
                    const int cj = ((col0 & 63) >> 3) + c4;
                    *reinterpret_cast<uint4*>(hrow + ((cj ^ (row & 7)) << 4)) = make_uint4(pk[cc * 16 + 4 * c4], pk[cc * 16 + 4 * c4 + 1], pk[cc * 16 + 4 * c4 + 2], pk[cc * 16 + 4 * c4 + 3]);
                }
            }
            if (!(p.knock & 8)) fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
            if (!(p.knock & 16)) tc_fence_before();
            __syncwarp();                      // every lane's stores are fenced before the warp's single arrival
            if (lane == 0) mbar_arrive(&h_full[b]);
            if (threadIdx.x == 64 && i == 0) FF_STAMP(3);
            if (threadIdx.x == 64 && i >= 8 && i < 16) FF_STAMP(i);     // steady-state chunk cadence
        }
        float* stage_q = reinterpret_cast<float*>(sW) + (q * 32) * TC_LDS;
        pdl_wait();                            // (already satisfied: acc2 depends on X) orders the stores below explicitly
        LnPre pre;
        // residual rows of the final LayerNorm (with the prologue: rows this very thread wrote) in flight while the last GEMM2 retires
        if (EPI == TC_EPI_LN) epi_ln_prefetch(p, m0 + q * 32 + hf * RPW, RPW, lane, pre);
        mbar_wait(&acc2_full, 0);
        tc_fence_after();
        if (threadIdx.x == 64) FF_STAMP(4);
        if (WS) epi_tmem2_to_stage(tmem_acc2, q, hf, lane, stage_q);
        else epi_tmem_to_stage<TC_BN, NP>(tmem_acc2, q, hf, lane, stage_q);
        epi_bar_sync<EW>();
        if (threadIdx.x == 64) FF_STAMP(5);
        const float* st = stage_q + (hf * RPW) * TC_LDS;
        if (EPI == TC_EPI_LN) epi_rows_ln(p, st, m0 + q * 32 + hf * RPW, RPW, lane, nullptr, 0, &pre);
        else epi_rows_store(p, st, m0 + q * 32 + hf * RPW, RPW, 0, split, lane);
        if (threadIdx.x == 64) FF_STAMP(6);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 64) FF_STAMP(7);
#undef FF_STAMP
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FF_TMEM_COLS);
    }
}

}  // namespace mmt
