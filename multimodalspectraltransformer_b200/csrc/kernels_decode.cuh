// Fused, row-local kernels of the KV-cached decoder step for SMALL batches (a few hundred to a
// few thousand sequences): the step is then a chain of tiny dependent launches whose fixed
// latency (launch + ramp + drain, several us each) dominates.  Every operation of a decoder layer
// except the FFN is local to one sequence, so ONE kernel per layer does everything between the FFNs:
//
//   decode_attn   x  = LN3_prev(x + b2 + sum FFN2 partials) | E_tok[token] + E_pos[t]   (prologue)
//                 qkv = W_in x + b ; append K,V to the paged cache ; causal self-attention
//                 x1 = LN1(x + W_o att + b_o) ; qc = W_q^cross x1 + b
//                 att = cross-attention of qc over the projected encoder memory
//                 x2 = LN2(x1 + W_o^cross att + b)      (fp32 + bf16 operand copy for the FFN)
//
// The per-step working set (cross K/V of every sequence, 40+ MB per layer) streams through L2 and
// evicts the weights between steps, so every dependent global access pays a DRAM round trip and the
// kernel is a chain of such round trips.  It is organised to keep that chain short:
//   * weights never pass through registers: the 192 KB QKV matrix is pulled into shared memory by
//     bulk async copies (cp.async.bulk + mbarrier) issued before anything else, and the three
//     128x128 matrices of the later phases replace it while the self-attention runs;
//   * the cross-attention K/V rows a warp will read in the second half are prefetched into L2 at
//     kernel start (no registers held), so that HBM stream overlaps the self-attention phases;
//   * one warp per (row, head): every per-warp index / page-table load is issued up front.
//
// Arithmetic is fp32 FMA with fp32 activations in both precision modes.  In the tensor-core mode the four
// projection matrices are read as bf16 (the hi term of the two-term split: measured logit error unchanged,
// DESIGN.md "bf16 numerics"): half the bytes per CTA (worth 2.5 % of the step) and all four fit in shared memory
// at once, so every weight copy is issued before the PDL wait.  Later same-box knock-outs
// (profiles/r01_decode_phase_cycles.md) showed the phases to be fixed latency chains -- neither the arithmetic, nor
// the issue slots, nor the weight stream, nor (with the L2 prefetch below) the K/V stream bounds them.
//
// Shape of the work: the kernels are pure latency chains (L2 / HBM round trips), so a CTA is a full
// 1024-thread SM's worth of warps for DA_R rows: one warp per (row, head) in the attention phases,
// and the 128-wide matrix-vector products split K over the lanes of a warp (conflict-free 512 B
// weight rows in shared memory) with 16 outputs per warp pass, finished by a shuffle reduce-scatter.
#pragma once
#include "common.cuh"
#include "kernels_simt.cuh"   // ldmatrix_x4, mma_bf16_16816
#include "kernels_tc.cuh"     // TMA / mbarrier wrappers (the cluster variant stages FFN weight tiles with cp.async.bulk.tensor)

namespace mmt {

constexpr int DA_R = 2;            // sequences per CTA
constexpr int DA_WARPS = 32;
constexpr int DA_THREADS = DA_WARPS * 32;
constexpr int DA_H = 16;           // decoder heads (config_V8 num_heads): one warp per (row, head)
static_assert(DA_R * DA_H == DA_WARPS, "one warp per (row, head)");
// dynamic shared memory: weight buffer + small parameter vectors
constexpr int DA_W_FLOATS = 3 * D * D;                      // in_proj (384x128); later out_w | cq_w | co_w
enum { DA_P_INB = 0, DA_P_OUTB = 3 * D, DA_P_N1W = 4 * D, DA_P_N1B = 5 * D, DA_P_CQB = 6 * D, DA_P_COB = 7 * D,
       DA_P_N2W = 8 * D, DA_P_N2B = 9 * D, DA_P_PB = 10 * D, DA_P_PG = 11 * D, DA_P_PBETA = 12 * D, DA_P_FLOATS = 13 * D };
constexpr int DA_SMEM_BYTES = (DA_W_FLOATS + DA_P_FLOATS) * 4 + 128;

// v[0..15] per lane -> every lane returns sum over the 32 lanes of v[(lane >> 1) & 15]
// (reduce-scatter butterfly: 8 + 4 + 2 + 1 + 1 shuffles)
__device__ __forceinline__ float reduce_scatter16(float (&v)[16], int lane) {
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1) {
        const int half = off >> 1;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// v[0..7] per lane -> every lane returns the sum over the 32 lanes of v[(lane >> 2) & 7]   (4 + 2 + 1 + 1 + 1 shuffles
// instead of 8 x 5 for eight separate butterflies)
__device__ __forceinline__ float reduce_scatter8(float (&v)[8], int lane) {
#pragma unroll
    for (int off = 16; off >= 4; off >>= 1) {
        const int half = off >> 2;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? v[i] : v[i + half];
            const float keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// out[r][n] = bias[n] + W[n,:] . xs[r,:]   for n in [0, n_out), all DA_R rows; W row-major [n_out][128]
// and bias in SHARED memory.  Warps take groups of 16 outputs round-robin.
__device__ __forceinline__ void gemv_rows(const float* W, const float* bias, int n_out,
                                          const float (*xs)[D], float* out, int ldo, int warp, int lane) {
    float4 xr[DA_R];
#pragma unroll
    for (int r = 0; r < DA_R; ++r) xr[r] = *reinterpret_cast<const float4*>(&xs[r][4 * lane]);
    for (int n0 = warp * 16; n0 < n_out; n0 += DA_WARPS * 16) {
        float acc[DA_R][16];
        const float4* wp = reinterpret_cast<const float4*>(W + n0 * D) + lane;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float4 w = wp[j * (D / 4)];
#pragma unroll
            for (int r = 0; r < DA_R; ++r)
                acc[r][j] = fmaf(w.w, xr[r].w, fmaf(w.z, xr[r].z, fmaf(w.y, xr[r].y, w.x * xr[r].x)));
        }
        const int n = n0 + ((lane >> 1) & 15);
        const float b = bias[n];
#pragma unroll
        for (int r = 0; r < DA_R; ++r) {
            const float v = reduce_scatter16(acc[r], lane) + b;
            if (!(lane & 1)) out[r * ldo + n] = v;
        }
    }
}

// the same with bf16 weights in shared memory (row-major [n_out][128]): lane owns k = 4 lane .. 4 lane + 3
__device__ __forceinline__ void gemv_rows_w16(const __nv_bfloat16* W, const float* bias, int n_out,
                                              const float (*xs)[D], float* out, int ldo, int warp, int lane) {
    float4 xr[DA_R];
#pragma unroll
    for (int r = 0; r < DA_R; ++r) xr[r] = *reinterpret_cast<const float4*>(&xs[r][4 * lane]);
    for (int n0 = warp * 16; n0 < n_out; n0 += DA_WARPS * 16) {
        float acc[DA_R][16];
        const uint2* wp = reinterpret_cast<const uint2*>(W + n0 * D) + lane;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint2 u = wp[j * (D / 4)];
            const float w0 = __uint_as_float(u.x << 16), w1 = __uint_as_float(u.x & 0xffff0000u);
            const float w2 = __uint_as_float(u.y << 16), w3 = __uint_as_float(u.y & 0xffff0000u);
#pragma unroll
            for (int r = 0; r < DA_R; ++r)
                acc[r][j] = fmaf(w3, xr[r].w, fmaf(w2, xr[r].z, fmaf(w1, xr[r].y, w0 * xr[r].x)));
        }
        const int n = n0 + ((lane >> 1) & 15);
        const float b = bias[n];
#pragma unroll
        for (int r = 0; r < DA_R; ++r) {
            const float v = reduce_scatter16(acc[r], lane) + b;
            if (!(lane & 1)) out[r * ldo + n] = v;
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- bulk async copy global -> shared, completion on an mbarrier (same primitives as kernels_tc.cuh)
__device__ __forceinline__ uint32_t da_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void da_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(da_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void da_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(da_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void da_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(da_smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 4000000000LL) { printf("mmt: decode_attn mbarrier wait timed out\n"); __trap(); }
    }
}
__device__ __forceinline__ void da_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(da_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(da_smem_u32(bar)) : "memory");
}
// copy `floats` fp32 in pieces of at most 16 KB
__device__ __forceinline__ void da_bulk_matrix(float* dst, const float* src, int floats, uint64_t* bar) {
    for (int o = 0; o < floats; o += 4096) da_bulk_g2s(dst + o, src + o, (uint32_t)min(4096, floats - o) * 4u, bar);
}

struct DecAttnParams {
    // ---- source of the layer input x (exactly one of the two)
    const int64_t* tokens; int tok_shift; int sos; int64_t ldn;      // layer 0: x = E_tok[token(t)] + E_pos[t]
    const float* E_tok; const float* E_pos; int vocab;
    const float* x_in;                                               // layers > 0: [M][D] (x2 of the previous layer)
    const float* part; int splits; int64_t part_stride;              // + FFN2 partial sums of the previous layer
    const float* pbias; const float* pgamma; const float* pbeta;     //   x = LN3(x_in + pbias + sum_s part[s])
    // ---- self-attention block
    const float *in_w, *in_b, *out_w, *out_b, *n1_w, *n1_b;
    void* kv_pool; const int* block_table; int pps;                  // paged self-attention cache of this layer (KVT elements)
    const int* step;
    // ---- cross-attention block
    const float *cq_w, *cq_b, *co_w, *co_b, *n2_w, *n2_b;
    const void* ckv; int64_t rows_total;       // projected memory of this layer: [spectrum][K|V][H][rows_total][DH] (KVT elements),
                                               // rows_total = rows reserved per spectrum; one contiguous block per spectrum keeps
                                               // a CTA's K/V inside one or two 2 MB pages (32 head planes would thrash the TLB)
    const int* nk; const int* row_start; const float* kbias_c; int n_cand;
    float* x2; __nv_bfloat16* x2_16;           // out [M][D] fp32 (+ bf16 operand copy for the FFN, optional)
    int64_t M; float scale; float eps;
    // tensor-core mode: the four projection matrices as bf16 (hi term of the two-term split), same shapes as the fp32 ones
    const __nv_bfloat16 *in_w16, *out_w16, *cq_w16, *co_w16;
    long long* dbg;                            // optional [gridDim.x][16] phase timestamps (MMT_DA_DEBUG)
    // FFN = true (cluster variant, tensor-core mode): the layer's FFN + norm3 run inside this kernel
    CUtensorMap tmW1, tmW2;                    // linear1 [F][128], linear2 [128][F] (hi terms, F = DA_FF), box {64 k, 128 rows}, SWIZZLE_128B
    const float *b1, *b2, *n3_w, *n3_b;
    float* x_out;                              // [M][D] layer output (the next layer's x_in / the sampler's x)
};

// ---- thread-block-cluster primitives (the FFN variant runs as clusters of DA_CL CTAs)
constexpr int DA_CL = 4;                       // CTAs per cluster: 8 rows share one pass over the FFN weights
constexpr int DA_FF = 2048;                    // d_ff
constexpr int DA_FS = DA_FF / DA_CL;           // hidden columns per CTA
constexpr int DA_XG_LD = D + 8;                // padded bf16 row of the gathered FFN input (conflict-free B fragments)
constexpr int DA_HS_LD = DA_FS + 8;            // padded bf16 row of the hidden activation
constexpr int DA_RED_LD = D + 4;               // padded fp32 row of the K-split partial outputs
constexpr int DA_BOX = 16384;                  // one weight box: [128 rows][64 k] bf16, 128-byte rows, 16-byte chunks XOR (row & 7)
constexpr int DA_SMEM_BYTES_CL = DA_SMEM_BYTES + 1024;     // + slack to align the weight buffer to the swizzle period
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_map(const void* local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(da_smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_st_u32x2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void cluster_st_f32x2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
// FFN = true: the kernel is launched as clusters of DA_CL CTAs (8 rows) and ends with the layer's FFN and norm3 instead
// of handing x2 to a separate FFN kernel:  the LN2 rows of the cluster are gathered (bf16) into every CTA through
// distributed shared memory; CTA c computes hidden columns [c * 512, +512) for all 8 rows and its K-slice of the second
// product on mma.sync m16n8k16 with the weight fragments read straight from L2 (256 KB per CTA and layer); the partial
// outputs are scattered to the row owners, which add them in rank order, apply bias + residual + LayerNorm and write the
// layer output.  This removes the FFN kernel, its two hand-offs and the split-F partial exchange through HBM from the
// latency chain of a small wave (13 -> 7 launches per position).
template <int DH, typename KVT, bool FFN = false>
__global__ void __launch_bounds__(DA_THREADS, 1) decode_attn(const __grid_constant__ DecAttnParams p) {
    static_assert(DH == 8 && DH * DA_H == D, "8-element key rows, 16 heads");
    static_assert(!FFN || sizeof(KVT) == 2, "the in-kernel FFN exists in the tensor-core mode only");
    typedef KvRow<KVT> KV;
    extern __shared__ __align__(128) uint8_t da_smem[];
    constexpr bool W16 = sizeof(KVT) == 2;            // tensor-core mode: bf16 projection weights, all four resident at once
    // (cluster variant: the weight buffer later receives 128B-swizzled TMA boxes -> 1024-byte aligned)
    uint8_t* const da_base = FFN ? da_smem + ((1024u - (da_smem_u32(da_smem) & 1023u)) & 1023u) : da_smem;
    float* Wbuf = reinterpret_cast<float*>(da_base);
    __nv_bfloat16* Wb16 = reinterpret_cast<__nv_bfloat16*>(da_base);    // W16 layout: in_w [384][128] | out_w | cq_w | co_w (192 KB)
    float* Ps = Wbuf + DA_W_FLOATS;
    __shared__ __align__(8) uint64_t bars[4];      // 0: in_w + vectors, 1: out_w, 2: cq_w, 3: co_w
    // cluster variant: FFN weight boxes land in the regions the attention weights vacate --
    // fb[0]: W1 boxes 0-5 -> in_w region | fb[1]: W1 boxes 6, 7 -> out_w | fb[2]: W2 boxes 0, 1 -> cq_w | fb[3]: W2 boxes 2, 3 -> co_w |
    // fb[4], fb[5]: W2 boxes 4, 5 / 6, 7 -> in_w region once the first product has consumed W1
    __shared__ __align__(8) uint64_t fb[6];
    const uint32_t ffn_rank = FFN ? cluster_ctarank() : 0u;
    const int f_base = (int)ffn_rank * DA_FS;
    // W1 box j = (128-row block j / 2, K slab j % 2) of this CTA's 512 hidden columns; W2 box b = hidden columns [64 b, +64), all 128 rows
    auto ffn_load_w1 = [&](int j, uint64_t* bar) { tma_load_2d(da_base + (size_t)j * DA_BOX, &p.tmW1, bar, (j & 1) * 64, f_base + (j >> 1) * 128); };
    auto ffn_load_w2 = [&](int b, uint8_t* dst, uint64_t* bar) { tma_load_2d(dst, &p.tmW2, bar, f_base + b * 64, 0); };
    __shared__ __align__(16) float xs[DA_R][D];
    __shared__ __align__(16) float qkv[DA_R][3 * D];
    __shared__ __align__(16) float att[DA_R][D];
    __shared__ __align__(16) float x1s[DA_R][D];
    __shared__ __align__(16) float psum[DA_WARPS][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * DA_R;
    const int r = warp / DA_H, h = warp % DA_H;      // this warp's (row, head) in the attention phases
    const int64_t n = row0 + r;
    const bool live = n < p.M;
#define DA_STAMP(i) do { if (p.dbg && threadIdx.x == 0) p.dbg[blockIdx.x * 16 + (i)] = clock64(); } while (0)
#define DA_STAMP2(i) do { if (p.dbg && threadIdx.x == 5 * 32) p.dbg[(1024 + blockIdx.x) * 16 + (i)] = clock64(); } while (0)
    DA_STAMP(0);
    pdl_launch_dependents();
    if (FFN) cluster_arrive();     // "this CTA has started": waited for before the first remote shared-memory store

    // ---- weights of the first phase + every small vector: bulk async copies, issued before anything else
    // (one copy per lane of warp 0: a single thread issuing ~25 copies costs microseconds)
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < 4; ++i) da_mbar_init(&bars[i], 1);
            if (FFN) { for (int i = 0; i < 6; ++i) da_mbar_init(&fb[i], 1); tma_prefetch_desc(&p.tmW1); tma_prefetch_desc(&p.tmW2); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            const uint32_t vec_bytes = (uint32_t)(3 * D + 7 * D + (p.part ? 3 * D : 0)) * 4u;
            da_mbar_expect_tx(&bars[0], (W16 ? 3u * D * D * 2u : (uint32_t)DA_W_FLOATS * 4u) + vec_bytes);
            if (W16) for (int i = 1; i < 4; ++i) da_mbar_expect_tx(&bars[i], D * D * 2);
        }
        __syncwarp();
        if (W16) {
            if (lane < 12) {     // 12 pieces of 16 KB: in_w (6), out_w (2), cq_w (2), co_w (2) -- all before the PDL wait
                const int m = lane < 6 ? 0 : (lane - 4) >> 1;                        // matrix 0..3
                const int piece = lane < 6 ? lane : (lane & 1);
                const __nv_bfloat16* src = (m == 0 ? p.in_w16 : (m == 1 ? p.out_w16 : (m == 2 ? p.cq_w16 : p.co_w16))) + piece * 8192;
                da_bulk_g2s(Wb16 + (size_t)lane * 8192, src, 8192 * 2, &bars[m]);
            }
        } else if (lane == 0) {
            // every CTA reads the same matrix in the same order: requests for one line from many SMs that
            // arrive close together are merged by L2 (measured: a per-CTA rotated order is 2x slower)
            for (int piece = 0; piece < 12; ++piece)
                da_bulk_g2s(Wbuf + piece * 4096, p.in_w + piece * 4096, 4096 * 4, &bars[0]);
        }
        if (lane >= 12) {
            const float* src = nullptr; int dst = 0, nf = D;
            switch (lane) {
                case 12: src = p.in_b; dst = DA_P_INB; nf = 3 * D; break;
                case 13: src = p.out_b; dst = DA_P_OUTB; break;
                case 14: src = p.n1_w; dst = DA_P_N1W; break;
                case 15: src = p.n1_b; dst = DA_P_N1B; break;
                case 16: src = p.cq_b; dst = DA_P_CQB; break;
                case 17: src = p.co_b; dst = DA_P_COB; break;
                case 18: src = p.n2_w; dst = DA_P_N2W; break;
                case 19: src = p.n2_b; dst = DA_P_N2B; break;
                case 20: if (p.part) { src = p.pbias; dst = DA_P_PB; } break;
                case 21: if (p.part) { src = p.pgamma; dst = DA_P_PG; } break;
                case 22: if (p.part) { src = p.pbeta; dst = DA_P_PBETA; } break;
                default: break;
            }
            if (src) da_bulk_g2s(Ps + dst, src, (uint32_t)nf * 4u, &bars[0]);
        }
    }
    // ---- per-warp indices, issued up front: cross-attention key range, self-attention page table
    int cnt = 0;
    int64_t r0 = 0, bmem = 0;
    int my_page = 0;
    if (live) {
        bmem = n / p.n_cand;
        cnt = p.nk[bmem];
        r0 = p.row_start[bmem];
        my_page = (lane < p.pps) ? p.block_table[n * p.pps + lane] : 0;
    }
    const KVT* Kc = reinterpret_cast<const KVT*>(p.ckv) + ((bmem * 2 + 0) * DA_H + h) * p.rows_total * DH;
    const KVT* Vc = reinterpret_cast<const KVT*>(p.ckv) + ((bmem * 2 + 1) * DA_H + h) * p.rows_total * DH;
    constexpr int PF = 64 / (int)sizeof(KVT);     // elements per 64-byte prefetch granule
    // L2 prefetch of this warp's cross-attention K/V rows (64-byte granules; consumed in the second half)
    for (int g = lane; g * PF < cnt * DH; g += 32) { prefetch_l2(Kc + g * PF); prefetch_l2(Vc + g * PF); }
#ifdef DA_TLB_WARM
    if (live) {   // experiment: touch the pages this warp will read later (address translation warmed up off the critical path)
        const int pg0 = __shfl_sync(0xffffffffu, my_page, 0);
        const KVT* kp = reinterpret_cast<const KVT*>(p.kv_pool) + (int64_t)pg0 * (2 * PAGE_TOKENS * D) + (h * PAGE_TOKENS) * DH;
        unsigned d0, d1;
        asm volatile("ld.global.u32 %0, [%1];" : "=r"(d0) : "l"(kp + lane * 8) : "memory");
        asm volatile("ld.global.u32 %0, [%1];" : "=r"(d1) : "l"(Kc + lane * 8) : "memory");
        if ((d0 ^ d1) == 0x12345679u) printf("x");     // keep the loads
    }
#endif
    // ---- everything above reads only decode-loop constants; from here on the previous kernel's results are needed
    pdl_wait();
    const int t = *p.step;
    DA_STAMP(1);

    // ---- prologue: layer input
    if (p.tokens) {
        if (warp < DA_R) {
            const int64_t nn = row0 + warp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nn < p.M) {
                int64_t tok;
                if (p.tok_shift) tok = (t == 0) ? p.sos : p.tokens[(int64_t)(t - 1) * p.ldn + nn];
                else tok = p.tokens[(int64_t)t * p.ldn + nn];
                if (tok < 0 || tok >= p.vocab) tok = 0;
                const float4 a = *reinterpret_cast<const float4*>(p.E_tok + tok * D + lane * 4);
                const float4 b = *reinterpret_cast<const float4*>(p.E_pos + (int64_t)t * D + lane * 4);
                v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            }
            *reinterpret_cast<float4*>(&xs[warp][lane * 4]) = v;
        }
        __syncthreads();
        da_mbar_wait(&bars[0], 0);
    } else if (FFN) {
        // the previous layer's kernel has applied its norm3: x_in is the layer input
        if (warp < DA_R) {
            const int64_t nn = row0 + warp;
            *reinterpret_cast<float4*>(&xs[warp][lane * 4]) =
                nn < p.M ? *reinterpret_cast<const float4*>(p.x_in + nn * D + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        da_mbar_wait(&bars[0], 0);
    } else {
        // x = LN3(x_in + pbias + sum_s part[s]): the partial sums are spread over all warps (one L2
        // round trip), reduced through shared memory in a fixed order (deterministic)
        const int pr = warp % DA_R, k0 = warp / DA_R;
        const int64_t nn = row0 + pr;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 xin = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nn < p.M) {
            if (warp < DA_R) xin = *reinterpret_cast<const float4*>(p.x_in + nn * D + lane * 4);
            if (p.part) {
                float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
                if (k0 < p.splits) q0 = *reinterpret_cast<const float4*>(p.part + (int64_t)k0 * p.part_stride + nn * D + lane * 4);
                if (k0 + DA_WARPS / DA_R < p.splits)
                    q1 = *reinterpret_cast<const float4*>(p.part + (int64_t)(k0 + DA_WARPS / DA_R) * p.part_stride + nn * D + lane * 4);
                s = make_float4(q0.x + q1.x, q0.y + q1.y, q0.z + q1.z, q0.w + q1.w);
                for (int k = k0 + 2 * (DA_WARPS / DA_R); k < p.splits; k += DA_WARPS / DA_R) {
                    const float4 q = *reinterpret_cast<const float4*>(p.part + (int64_t)k * p.part_stride + nn * D + lane * 4);
                    s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
                }
            }
        }
        *reinterpret_cast<float4*>(&psum[warp][lane * 4]) = s;
        __syncthreads();
        da_mbar_wait(&bars[0], 0);
        if (warp < DA_R) {
            float4 v = xin;     // warp < DA_R => pr == warp
            if (p.part && row0 + warp < p.M) {
                float4 a = *reinterpret_cast<const float4*>(Ps + DA_P_PB + lane * 4);
#pragma unroll
                for (int k = 0; k < DA_WARPS / DA_R; ++k) {
                    const float4 q = *reinterpret_cast<const float4*>(&psum[k * DA_R + warp][lane * 4]);
                    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
                }
                v = ln_row(make_float4(a.x + v.x, a.y + v.y, a.z + v.z, a.w + v.w), Ps + DA_P_PG, Ps + DA_P_PBETA, p.eps, lane);
            }
            *reinterpret_cast<float4*>(&xs[warp][lane * 4]) = v;
        }
        __syncthreads();
    }

    DA_STAMP(2);
    // ---- QKV projection (384 outputs) from shared memory
    if (W16) gemv_rows_w16(Wb16, Ps + DA_P_INB, 3 * D, xs, &qkv[0][0], 3 * D, warp, lane);
    else gemv_rows(Wbuf, Ps + DA_P_INB, 3 * D, xs, &qkv[0][0], 3 * D, warp, lane);
    __syncthreads();
    DA_STAMP(3);
    // the QKV matrix is dead after the barrier above: the three 128x128 matrices of the later phases are pulled
    // into its place (issued by warp 0 from inside the attention phase, see DA_ISSUE_ROUND2)
#define DA_ISSUE_ROUND2() do {                                                                              \
        if (!W16 && warp == 0) {                                                                                  \
            if (lane < 3) da_mbar_expect_tx(&bars[1 + lane], D * D * 4);                                    \
            __syncwarp();                                                                                   \
            if (lane < 12) {                                                                                \
                const int m_ = lane >> 2, piece_ = lane & 3;                                                \
                const float* src_ = (m_ == 0 ? p.out_w : (m_ == 1 ? p.cq_w : p.co_w)) + piece_ * 4096;      \
                da_bulk_g2s(Wbuf + m_ * D * D + piece_ * 4096, src_, 4096 * 4, &bars[1 + m_]);              \
            }                                                                                               \
        }                                                                                                   \
    } while (0)
    // cluster variant: the first six boxes of this CTA's W1 slice replace the (dead) QKV matrix while the self-attention runs
#define DA_ISSUE_FFN_W1A() do {                                                                             \
        if (FFN && warp == 0 && lane == 0) {                                                                \
            mbar_arrive_expect_tx(&fb[0], 6 * DA_BOX);                                                      \
            for (int j_ = 0; j_ < 6; ++j_) ffn_load_w1(j_, &fb[0]);                                         \
        }                                                                                                   \
    } while (0)
    // ---- KV append + causal self-attention: one warp per (row, head)
    constexpr int PAGE_ELEMS = 2 * PAGE_TOKENS * D;
    KVT* const pool = reinterpret_cast<KVT*>(p.kv_pool);
    if (live) {
        const float* row = &qkv[r][h * DH];
        {
            const int pg = __shfl_sync(0xffffffffu, my_page, t / PAGE_TOKENS);
            if (lane < 2 * DH) {
                const int kv = lane / DH, d = lane % DH;
                KVT* page = pool + (int64_t)pg * PAGE_ELEMS;
                KV::st(page + ((kv * DA_H + h) * PAGE_TOKENS + (t % PAGE_TOKENS)) * DH + d, row[(1 + kv) * D + d]);
            }
        }
        __syncwarp();
        DA_STAMP2(0);
        float q[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) q[d] = row[d] * p.scale;
        constexpr int MAXK = 4;   // max_len 128 / 32
        float s[MAXK];
        float m = MMT_NEG_INF;
        typename KV::Raw kraw[MAXK];
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {     // all K rows of this lane in flight together; V rows prefetched to L2
            const int j = lane + i * 32;
            const int pg = __shfl_sync(0xffffffffu, my_page, (j / PAGE_TOKENS) & 31);
            if (j < t) {
                const KVT* page = pool + (int64_t)pg * PAGE_ELEMS;
                kraw[i] = KV::ld(page + ((0 * DA_H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH);
                prefetch_l2(page + ((1 * DA_H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH);   // V row: an L2 hit by the time the softmax needs it
            } else if (j == t) {
                kraw[i] = KV::pack(row + D);      // this position's own K: straight from the projection (no global round trip)
            }
        }
        DA_ISSUE_ROUND2();   // after this warp's K loads: the 192 KB of weights must not queue ahead of them on the SM's ingress
        DA_ISSUE_FFN_W1A();
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {
            const int j = lane + i * 32;
            s[i] = MMT_NEG_INF;
            if (j <= t) {
                float k[DH];
                KV::unpack(kraw[i], k);
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) a = fmaf(q[d], k[d], a);
                s[i] = a;
                m = fmaxf(m, a);
            }
        }
        m = warp_max(m);
        DA_STAMP2(1);
        float l = 0.f, acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {
            const int j = lane + i * 32;
            const int pg = __shfl_sync(0xffffffffu, my_page, (j / PAGE_TOKENS) & 31);
            if (j < t) {
                const KVT* page = pool + (int64_t)pg * PAGE_ELEMS;
                kraw[i] = KV::ld(page + ((1 * DA_H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH);
            } else if (j == t) {
                kraw[i] = KV::pack(row + 2 * D);
            }
        }
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {
            const int j = lane + i * 32;
            if (j <= t) {
                const float e = expf(s[i] - m);
                l += e;
                float v[DH];
                KV::unpack(kraw[i], v);
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] = fmaf(e, v[d], acc[d]);
            }
        }
        DA_STAMP2(2);
        l = warp_sum(l);
        {
            const float v = reduce_scatter8(acc, lane);       // lane holds dimension (lane >> 2) & 7
            if (!(lane & 3)) att[r][h * DH + (lane >> 2)] = v / l;
        }
        DA_STAMP2(3);
    } else {
        DA_ISSUE_ROUND2();
        DA_ISSUE_FFN_W1A();
        if (lane < DH) att[r][h * DH + lane] = 0.f;
    }
#undef DA_ISSUE_ROUND2
#undef DA_ISSUE_FFN_W1A
    __syncthreads();
    DA_STAMP2(4);
    DA_STAMP(4);

    // ---- out-projection -> qkv[r][0..127] (reused as scratch), then LN1
    da_mbar_wait(&bars[1], 0);
    DA_STAMP(5);
    if (W16) gemv_rows_w16(Wb16 + 3 * D * D, Ps + DA_P_OUTB, D, att, &qkv[0][0], 3 * D, warp, lane);
    else gemv_rows(Wbuf, Ps + DA_P_OUTB, D, att, &qkv[0][0], 3 * D, warp, lane);
    __syncthreads();
    if (FFN && warp == 1 && lane == 0) {       // out_w is dead: the last two W1 boxes
        mbar_arrive_expect_tx(&fb[1], 2 * DA_BOX);
        ffn_load_w1(6, &fb[1]); ffn_load_w1(7, &fb[1]);
    }
    if (warp < DA_R) {
        const float4 a = *reinterpret_cast<const float4*>(&xs[warp][lane * 4]);
        const float4 y = *reinterpret_cast<const float4*>(&qkv[warp][lane * 4]);
        *reinterpret_cast<float4*>(&x1s[warp][lane * 4]) =
            ln_row(make_float4(a.x + y.x, a.y + y.y, a.z + y.z, a.w + y.w), Ps + DA_P_N1W, Ps + DA_P_N1B, p.eps, lane);
    }
    __syncthreads();

    DA_STAMP(6);
    // ---- cross-attention query projection -> xs (the layer input is no longer needed)
    da_mbar_wait(&bars[2], 0);
    if (W16) gemv_rows_w16(Wb16 + 4 * D * D, Ps + DA_P_CQB, D, x1s, &xs[0][0], D, warp, lane);
    else gemv_rows(Wbuf + D * D, Ps + DA_P_CQB, D, x1s, &xs[0][0], D, warp, lane);
    __syncthreads();
    if (FFN && warp == 0 && lane == 0) {       // cq_w is dead: W2 boxes 0, 1 (hidden columns 0..127 of the slice)
        mbar_arrive_expect_tx(&fb[2], 2 * DA_BOX);
        ffn_load_w2(0, da_base + 8 * DA_BOX, &fb[2]); ffn_load_w2(1, da_base + 9 * DA_BOX, &fb[2]);
    }
    DA_STAMP(7);

    // ---- cross-attention over the projected memory: one warp per (row, head)
    if (live) {
        const float* bias = p.kbias_c + r0;
        float q[DH], acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) { q[d] = xs[r][h * DH + d] * p.scale; acc[d] = 0.f; }
        float m = MMT_NEG_INF, l = 0.f;
        DA_STAMP2(5);
        // two keys per lane per pass, K and V of both in flight together
        for (int j0 = lane; j0 < cnt; j0 += 64) {
            const int j1 = j0 + 32;
            const bool has1 = j1 < cnt;
            const int j1c = has1 ? j1 : j0;
            const typename KV::Raw rk0 = KV::ld(Kc + (int64_t)j0 * DH), rk1 = KV::ld(Kc + (int64_t)j1c * DH);
            const typename KV::Raw rv0 = KV::ld(Vc + (int64_t)j0 * DH), rv1 = KV::ld(Vc + (int64_t)j1c * DH);
            float s0 = bias[j0], s1 = has1 ? bias[j1] : MMT_NEG_INF;
            float k0[DH], k1[DH];
            KV::unpack(rk0, k0); KV::unpack(rk1, k1);
#pragma unroll
            for (int d = 0; d < DH; ++d) { s0 = fmaf(q[d], k0[d], s0); s1 = fmaf(q[d], k1[d], s1); }
            const float mn = fmaxf(m, fmaxf(s0, s1));
            if (mn > m) {
                const float corr = expf(m - mn);   // m = -inf on the first pass -> 0
                l *= corr;
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] *= corr;
                m = mn;
            }
            const float e0 = expf(s0 - m), e1 = has1 ? expf(s1 - m) : 0.f;
            l += e0 + e1;
            float v0[DH], v1[DH];
            KV::unpack(rv0, v0); KV::unpack(rv1, v1);
#pragma unroll
            for (int d = 0; d < DH; ++d) acc[d] = fmaf(e1, v1[d], fmaf(e0, v0[d], acc[d]));
        }
        DA_STAMP2(6);
        const float Mx = warp_max(m);
        const float corr = (m == MMT_NEG_INF) ? 0.f : expf(m - Mx);
        l = warp_sum(l * corr);
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] *= corr;
        {
            const float v = reduce_scatter8(acc, lane);
            if (!(lane & 3)) att[r][h * DH + (lane >> 2)] = v / l;
        }
        DA_STAMP2(7);
    }
    __syncthreads();
    DA_STAMP2(8);
    DA_STAMP(8);

    // ---- cross out-projection -> qkv scratch, then LN2 -> x2
    da_mbar_wait(&bars[3], 0);
    if (W16) gemv_rows_w16(Wb16 + 5 * D * D, Ps + DA_P_COB, D, att, &qkv[0][0], 3 * D, warp, lane);
    else gemv_rows(Wbuf + 2 * D * D, Ps + DA_P_COB, D, att, &qkv[0][0], 3 * D, warp, lane);
    __syncthreads();
    if (FFN && warp == 2 && lane == 0) {       // co_w is dead: W2 boxes 2, 3
        mbar_arrive_expect_tx(&fb[3], 2 * DA_BOX);
        ffn_load_w2(2, da_base + 10 * DA_BOX, &fb[3]); ffn_load_w2(3, da_base + 11 * DA_BOX, &fb[3]);
    }
    DA_STAMP(9);
    if constexpr (FFN) {
        // ================= in-kernel FFN over the cluster's 8 rows =================
        // Both products are computed transposed -- out^T = W . in^T -- so that the 8 rows of the cluster are exactly the N = 8
        // of mma.sync m16n8k16 and the weights are the A operand (ldmatrix from the swizzled boxes).
        __shared__ __align__(16) float yp[DA_CL][DA_R][D];        // partial outputs of this CTA's rows, one slot per source CTA (remote stores)
        // psum (16 KB) is unused in this variant: hidden activation | gathered FFN input (remote stores)
        uint8_t* scratch = reinterpret_cast<uint8_t*>(&psum[0][0]);
        __nv_bfloat16* hs = reinterpret_cast<__nv_bfloat16*>(scratch);                        // [8][DA_HS_LD]
        __nv_bfloat16* xg = reinterpret_cast<__nv_bfloat16*>(scratch + 8 * DA_HS_LD * 2);      // [8][DA_XG_LD]
        static_assert(8 * DA_HS_LD * 2 + 8 * DA_XG_LD * 2 <= DA_WARPS * D * 4, "FFN scratch must fit the psum buffer");
        float* red = reinterpret_cast<float*>(da_base + 6 * DA_BOX);                          // [4 K quarters][8][DA_RED_LD]: the out_w region, once W1 is consumed
        static_assert(4 * 8 * DA_RED_LD * 4 <= 2 * DA_BOX, "K-split scratch must fit the out_w region");
        const uint32_t rank = ffn_rank;
        const int g = lane >> 2, tq = lane & 3;
        // ldmatrix.x4 row address of this lane inside a 16-row weight tile: matrices (rows 0-7 | 8-15) x (chunk c | c + 1)
        const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8, lm_chunk = lane >> 4;
        cluster_wait();                       // every CTA of the cluster is running: remote stores may begin
        if (warp < DA_R) {
            const int64_t nn = row0 + warp;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nn < p.M) {
                const float4 a = *reinterpret_cast<const float4*>(&x1s[warp][lane * 4]);
                const float4 y = *reinterpret_cast<const float4*>(&qkv[warp][lane * 4]);
                o = ln_row(make_float4(a.x + y.x, a.y + y.y, a.z + y.z, a.w + y.w), Ps + DA_P_N2W, Ps + DA_P_N2B, p.eps, lane);
            }
            *reinterpret_cast<float4*>(&x1s[warp][lane * 4]) = o;       // fp32 residual of norm3
            const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
            const __nv_bfloat16* dst = xg + (rank * DA_R + warp) * DA_XG_LD + lane * 4;
#pragma unroll
            for (uint32_t d = 0; d < DA_CL; ++d)
                cluster_st_u32x2(cluster_map(dst, d), *reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
        // bias of this warp's 16 hidden columns (rows g and g + 8 of its tile)
        const float b1_lo = p.b1[f_base + warp * 16 + g], b1_hi = p.b1[f_base + warp * 16 + g + 8];
        cluster_arrive();
        cluster_wait();                       // the 8 input rows are in every CTA
        DA_STAMP(10);
        {   // ---- h^T[f_base + 16 warp .., :] = relu(W1 . x^T + b1): one 16-row tile of W1 per warp, K = 128
            mbar_wait(&fb[warp < 24 ? 0 : 1], 0);
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            const uint32_t tile = smem_u32(da_base) + (uint32_t)((warp >> 3) * 2 * DA_BOX + ((warp & 7) * 16 + lm_row) * 128);
            const __nv_bfloat16* xrow = xg + g * DA_XG_LD + 2 * tq;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t a[4];
                ldmatrix_x4(a, tile + (uint32_t)((ks >> 2) * DA_BOX + ((((ks & 3) * 2 + lm_chunk) ^ (lane & 7)) << 4)));
                mma_bf16_16816(c, a, *reinterpret_cast<const uint32_t*>(xrow + ks * 16), *reinterpret_cast<const uint32_t*>(xrow + ks * 16 + 8));
            }
            // c[0], c[1]: hidden column 16 warp + g of rows 2 tq, 2 tq + 1; c[2], c[3]: column + 8
            __nv_bfloat16* hcol = hs + warp * 16 + g;
            hcol[(2 * tq) * DA_HS_LD] = __float2bfloat16_rn(fmaxf(c[0] + b1_lo, 0.f));
            hcol[(2 * tq + 1) * DA_HS_LD] = __float2bfloat16_rn(fmaxf(c[1] + b1_lo, 0.f));
            hcol[(2 * tq) * DA_HS_LD + 8] = __float2bfloat16_rn(fmaxf(c[2] + b1_hi, 0.f));
            hcol[(2 * tq + 1) * DA_HS_LD + 8] = __float2bfloat16_rn(fmaxf(c[3] + b1_hi, 0.f));
        }
        DA_STAMP(11);
        __syncthreads();                      // hs complete; W1 consumed
        if (warp == 0 && lane == 0) {         // the second half of the W2 slice into the in_w region
            mbar_arrive_expect_tx(&fb[4], 2 * DA_BOX);
            ffn_load_w2(4, da_base + 0 * DA_BOX, &fb[4]); ffn_load_w2(5, da_base + 1 * DA_BOX, &fb[4]);
            mbar_arrive_expect_tx(&fb[5], 2 * DA_BOX);
            ffn_load_w2(6, da_base + 2 * DA_BOX, &fb[5]); ffn_load_w2(7, da_base + 3 * DA_BOX, &fb[5]);
        }
        {   // ---- y^T partial = W2[:, K quarter] . h^T: warp = (16-row output tile mt, K quarter kq of 128 hidden columns)
            const int mt = warp & 7, kq = warp >> 3;
            mbar_wait(&fb[2 + kq], 0);
            const uint32_t boxes = smem_u32(da_base) + (uint32_t)((kq < 2 ? 8 + 2 * kq : 2 * (kq - 2)) * DA_BOX + (mt * 16 + lm_row) * 128);
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            const __nv_bfloat16* hrow = hs + g * DA_HS_LD + kq * 128 + 2 * tq;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t a[4];
                ldmatrix_x4(a, boxes + (uint32_t)((ks >> 2) * DA_BOX + ((((ks & 3) * 2 + lm_chunk) ^ (lane & 7)) << 4)));
                mma_bf16_16816(c, a, *reinterpret_cast<const uint32_t*>(hrow + ks * 16), *reinterpret_cast<const uint32_t*>(hrow + ks * 16 + 8));
            }
            // c[0], c[1]: output column 16 mt + g of rows 2 tq, 2 tq + 1; c[2], c[3]: column + 8
            float* rq = red + (size_t)kq * 8 * DA_RED_LD + mt * 16 + g;
            rq[(2 * tq) * DA_RED_LD] = c[0]; rq[(2 * tq + 1) * DA_RED_LD] = c[1];
            rq[(2 * tq) * DA_RED_LD + 8] = c[2]; rq[(2 * tq + 1) * DA_RED_LD + 8] = c[3];
        }
        DA_STAMP(12);
        __syncthreads();
        {   // K quarters added in a fixed order; row i / 128 of the cluster belongs to CTA (row / DA_R), local row row % DA_R
            const int row = threadIdx.x >> 7, col = threadIdx.x & (D - 1);
            const float* rq = red + row * DA_RED_LD + col;
            const float v = ((rq[0] + rq[8 * DA_RED_LD]) + rq[2 * 8 * DA_RED_LD]) + rq[3 * 8 * DA_RED_LD];
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_map(&yp[rank][row % DA_R][col], (uint32_t)(row / DA_R))), "f"(v) : "memory");
        }
        cluster_arrive();
        cluster_wait();                       // the four partial outputs of this CTA's rows have landed
        DA_STAMP(13);
        if (warp < DA_R) {
            const int64_t nn = row0 + warp;
            if (nn < p.M) {
                float4 a = *reinterpret_cast<const float4*>(p.b2 + lane * 4);
#pragma unroll
                for (int sr = 0; sr < DA_CL; ++sr) {          // fixed order: deterministic
                    const float4 q = *reinterpret_cast<const float4*>(&yp[sr][warp][lane * 4]);
                    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
                }
                const float4 res = *reinterpret_cast<const float4*>(&x1s[warp][lane * 4]);
                const float4 o = ln_row(make_float4(a.x + res.x, a.y + res.y, a.z + res.z, a.w + res.w), p.n3_w, p.n3_b, p.eps, lane);
                *reinterpret_cast<float4*>(p.x_out + nn * D + lane * 4) = o;
            }
        }
        DA_STAMP(14);
        return;
    }
    if (warp < DA_R) {
        const int64_t nn = row0 + warp;
        if (nn < p.M) {
            const float4 a = *reinterpret_cast<const float4*>(&x1s[warp][lane * 4]);
            const float4 y = *reinterpret_cast<const float4*>(&qkv[warp][lane * 4]);
            const float4 o = ln_row(make_float4(a.x + y.x, a.y + y.y, a.z + y.z, a.w + y.w), Ps + DA_P_N2W, Ps + DA_P_N2B, p.eps, lane);
            *reinterpret_cast<float4*>(p.x2 + nn * D + lane * 4) = o;
            if (p.x2_16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                *reinterpret_cast<uint2*>(p.x2_16 + nn * D + lane * 4) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
    DA_STAMP(10);
#undef DA_STAMP
#undef DA_STAMP2
}


}  // namespace mmt
