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
// evicts the weights between steps, so every dependent phase would pay a DRAM round trip.  The kernel
// therefore starts by prefetching into L2 (no registers held): its slice of this layer's weights
// (all CTAs together cover them once) and the cross-attention K/V rows its own warps will read in
// the second half -- the HBM stream of the cross attention overlaps the self-attention phases.
//
// Arithmetic is fp32 in both precision modes (these projections are 20 % of the decoder's weights
// and the larger share of the bf16 logit error, DESIGN.md "bf16 numerics").
//
// Shape of the work: the kernels are pure latency chains (L2 / HBM round trips), so a CTA is a full
// 1024-thread SM's worth of warps for DA_R rows: one warp per (row, head) in the attention phases,
// and the 128-wide matrix-vector products split K over the lanes of a warp (coalesced 512 B weight
// rows straight from L2) with 16 outputs per warp pass, finished by a shuffle reduce-scatter.
#pragma once
#include "common.cuh"

namespace mmt {

constexpr int DA_R = 2;            // sequences per CTA
constexpr int DA_WARPS = 32;
constexpr int DA_THREADS = DA_WARPS * 32;

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

// out[r][n] = bias[n] + W[n,:] . xs[r,:]   for n in [0, n_out), all DA_R rows; W row-major [n_out][128]
// in global memory.  Warps take groups of 16 outputs round-robin.
__device__ __forceinline__ void gemv_rows(const float* __restrict__ W, const float* __restrict__ bias, int n_out,
                                          const float (*xs)[D], float* out, int ldo, int warp, int lane) {
    float4 xr[DA_R];
#pragma unroll
    for (int r = 0; r < DA_R; ++r) xr[r] = *reinterpret_cast<const float4*>(&xs[r][4 * lane]);
    for (int n0 = warp * 16; n0 < n_out; n0 += DA_WARPS * 16) {
        float acc[DA_R][16];
        const float4* wp = reinterpret_cast<const float4*>(W + (int64_t)n0 * D) + lane;
#pragma unroll
        for (int jb = 0; jb < 16; jb += 4) {
            float4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = __ldg(wp + (int64_t)(jb + j) * (D / 4));
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int r = 0; r < DA_R; ++r)
                    acc[r][jb + j] = fmaf(w[j].w, xr[r].w, fmaf(w[j].z, xr[r].z, fmaf(w[j].y, xr[r].y, w[j].x * xr[r].x)));
        }
        const int n = n0 + ((lane >> 1) & 15);
        const float b = bias ? bias[n] : 0.f;
#pragma unroll
        for (int r = 0; r < DA_R; ++r) {
            const float v = reduce_scatter16(acc[r], lane) + b;
            if (!(lane & 1)) out[r * ldo + n] = v;
        }
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// every CTA prefetches its 1/gridDim share of a weight matrix (64-byte granules), so that the grid
// as a whole pulls the matrix into L2 exactly once
__device__ __forceinline__ void prefetch_slice(const float* w, int n_floats) {
    const int granules = n_floats / 16;
    const int per = (granules + gridDim.x - 1) / gridDim.x;
    const int g0 = blockIdx.x * per;
    for (int g = g0 + threadIdx.x; g < min(g0 + per, granules); g += blockDim.x) prefetch_l2(w + (int64_t)g * 16);
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
    float* kv_pool; const int* block_table; int pps;                 // paged self-attention cache of this layer
    const int* step;
    // ---- cross-attention block
    const float *cq_w, *cq_b, *co_w, *co_b, *n2_w, *n2_b;
    const float* ckv; int64_t rows_total;      // projected memory of this layer, head-major [2][H][rows_total][DH]
    const int* nk; const int* row_start; const float* kbias_c; int n_cand;
    float* x2; __nv_bfloat16* x2_16;           // out [M][D] fp32 (+ bf16 operand copy for the FFN, optional)
    int64_t M; int H; float scale; float eps;
};

template <int DH>
__global__ void __launch_bounds__(DA_THREADS, 1) decode_attn(const __grid_constant__ DecAttnParams p) {
    static_assert(DH == 8, "two float4 per key row");
    __shared__ __align__(16) float xs[DA_R][D];
    __shared__ __align__(16) float qkv[DA_R][3 * D];
    __shared__ __align__(16) float att[DA_R][D];
    __shared__ __align__(16) float x1s[DA_R][D];
    __shared__ __align__(16) float psum[DA_WARPS][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * DA_R;

    // ---- L2 prefetch: weights of this layer (grid-wide, once) and this CTA's cross-attention K/V rows
    prefetch_slice(p.in_w, 3 * D * D);
    prefetch_slice(p.out_w, D * D);
    prefetch_slice(p.cq_w, D * D);
    prefetch_slice(p.co_w, D * D);
    for (int pair = warp; pair < DA_R * p.H; pair += DA_WARPS) {
        const int r = pair / p.H, h = pair % p.H;
        const int64_t n = row0 + r;
        if (n >= p.M) continue;
        const int64_t b = n / p.n_cand;
        const int cnt = p.nk[b];
        const int64_t r0 = p.row_start[b];
        const float* Kh = p.ckv + ((int64_t)(0 * p.H + h) * p.rows_total + r0) * DH;
        const float* Vh = p.ckv + ((int64_t)(1 * p.H + h) * p.rows_total + r0) * DH;
        for (int g = lane; g * 16 < cnt * DH; g += 32) { prefetch_l2(Kh + g * 16); prefetch_l2(Vh + g * 16); }
    }
    const int t = *p.step;

    // ---- prologue: layer input
    if (p.tokens) {
        if (warp < DA_R) {
            const int64_t n = row0 + warp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < p.M) {
                int64_t tok;
                if (p.tok_shift) tok = (t == 0) ? p.sos : p.tokens[(int64_t)(t - 1) * p.ldn + n];
                else tok = p.tokens[(int64_t)t * p.ldn + n];
                if (tok < 0 || tok >= p.vocab) tok = 0;
                const float4 a = *reinterpret_cast<const float4*>(p.E_tok + tok * D + lane * 4);
                const float4 b = *reinterpret_cast<const float4*>(p.E_pos + (int64_t)t * D + lane * 4);
                v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            }
            *reinterpret_cast<float4*>(&xs[warp][lane * 4]) = v;
        }
    } else {
        // x = LN3(x_in + pbias + sum_s part[s]): the partial sums are spread over all warps (one L2
        // round trip), reduced through shared memory in a fixed order (deterministic)
        const int r = warp % DA_R, k0 = warp / DA_R;
        const int64_t n = row0 + r;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.M && p.part) {
            for (int k = k0; k < p.splits; k += DA_WARPS / DA_R) {
                const float4 q = *reinterpret_cast<const float4*>(p.part + (int64_t)k * p.part_stride + n * D + lane * 4);
                s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
            }
        }
        *reinterpret_cast<float4*>(&psum[warp][lane * 4]) = s;
        __syncthreads();
        if (warp < DA_R) {
            const int64_t n2 = row0 + warp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n2 < p.M) {
                v = *reinterpret_cast<const float4*>(p.x_in + n2 * D + lane * 4);
                if (p.part) {
                    float4 a = *reinterpret_cast<const float4*>(p.pbias + lane * 4);
#pragma unroll
                    for (int k = 0; k < DA_WARPS / DA_R; ++k) {
                        const float4 q = *reinterpret_cast<const float4*>(&psum[k * DA_R + warp][lane * 4]);
                        a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
                    }
                    v = ln_row(make_float4(a.x + v.x, a.y + v.y, a.z + v.z, a.w + v.w), p.pgamma, p.pbeta, p.eps, lane);
                }
            }
            *reinterpret_cast<float4*>(&xs[warp][lane * 4]) = v;
        }
    }
    __syncthreads();

    // ---- QKV projection (384 outputs)
    gemv_rows(p.in_w, p.in_b, 3 * D, xs, &qkv[0][0], 3 * D, warp, lane);
    __syncthreads();

    // ---- KV append + causal self-attention: one warp per (row, head)
    constexpr int PAGE_FLOATS = 2 * PAGE_TOKENS * D;
    for (int pair = warp; pair < DA_R * p.H; pair += DA_WARPS) {
        const int r = pair / p.H, h = pair % p.H;
        const int64_t n = row0 + r;
        if (n >= p.M) continue;
        // page ids of this sequence: lane i holds page i (max_len 128 / PAGE_TOKENS 16 = 8 pages)
        const int my_page = (lane < p.pps) ? p.block_table[n * p.pps + lane] : 0;
        const float* row = &qkv[r][h * DH];
        {
            const int pg = __shfl_sync(0xffffffffu, my_page, t / PAGE_TOKENS);
            if (lane < 2 * DH) {
                const int kv = lane / DH, d = lane % DH;
                float* page = p.kv_pool + (int64_t)pg * PAGE_FLOATS;
                page[((kv * p.H + h) * PAGE_TOKENS + (t % PAGE_TOKENS)) * DH + d] = row[(1 + kv) * D + d];
            }
        }
        __syncwarp();
        float q[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) q[d] = row[d] * p.scale;
        constexpr int MAXK = 4;   // max_len 128 / 32
        float s[MAXK];
        float m = MMT_NEG_INF;
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {
            const int j = lane + i * 32;
            const int pg = __shfl_sync(0xffffffffu, my_page, (j / PAGE_TOKENS) & 31);
            s[i] = MMT_NEG_INF;
            if (j <= t) {
                const float* page = p.kv_pool + (int64_t)pg * PAGE_FLOATS;
                const float4* k = reinterpret_cast<const float4*>(page + ((0 * p.H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH);
                const float4* v = reinterpret_cast<const float4*>(page + ((1 * p.H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH);
                prefetch_l2(v);
                float a = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < DH / 4; ++d4) {
                    const float4 kk = k[d4];
                    a = fmaf(q[d4 * 4], kk.x, a); a = fmaf(q[d4 * 4 + 1], kk.y, a);
                    a = fmaf(q[d4 * 4 + 2], kk.z, a); a = fmaf(q[d4 * 4 + 3], kk.w, a);
                }
                s[i] = a;
                m = fmaxf(m, a);
            }
        }
        m = warp_max(m);
        float l = 0.f, acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll
        for (int i = 0; i < MAXK; ++i) {
            const int j = lane + i * 32;
            const int pg = __shfl_sync(0xffffffffu, my_page, (j / PAGE_TOKENS) & 31);
            if (j <= t) {
                const float e = expf(s[i] - m);
                l += e;
                const float* page = p.kv_pool + (int64_t)pg * PAGE_FLOATS;
                const float4* v = reinterpret_cast<const float4*>(page + ((1 * p.H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH);
#pragma unroll
                for (int d4 = 0; d4 < DH / 4; ++d4) {
                    const float4 vv = v[d4];
                    acc[d4 * 4] = fmaf(e, vv.x, acc[d4 * 4]); acc[d4 * 4 + 1] = fmaf(e, vv.y, acc[d4 * 4 + 1]);
                    acc[d4 * 4 + 2] = fmaf(e, vv.z, acc[d4 * 4 + 2]); acc[d4 * 4 + 3] = fmaf(e, vv.w, acc[d4 * 4 + 3]);
                }
            }
        }
        l = warp_sum(l);
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = warp_sum(acc[d]);
        if (lane < DH) {
            float v = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) if (lane == d) v = acc[d];
            att[r][h * DH + lane] = v / l;
        }
    }
    __syncthreads();

    // ---- out-projection -> qkv[r][0..127] (reused as scratch), then LN1
    gemv_rows(p.out_w, p.out_b, D, att, &qkv[0][0], 3 * D, warp, lane);
    __syncthreads();
    if (warp < DA_R) {
        const float4 a = *reinterpret_cast<const float4*>(&xs[warp][lane * 4]);
        const float4 y = *reinterpret_cast<const float4*>(&qkv[warp][lane * 4]);
        *reinterpret_cast<float4*>(&x1s[warp][lane * 4]) =
            ln_row(make_float4(a.x + y.x, a.y + y.y, a.z + y.z, a.w + y.w), p.n1_w, p.n1_b, p.eps, lane);
    }
    __syncthreads();

    // ---- cross-attention query projection -> xs (the layer input is no longer needed)
    gemv_rows(p.cq_w, p.cq_b, D, x1s, &xs[0][0], D, warp, lane);
    __syncthreads();

    // ---- cross-attention over the projected memory: one warp per (row, head)
    for (int pair = warp; pair < DA_R * p.H; pair += DA_WARPS) {
        const int r = pair / p.H, h = pair % p.H;
        const int64_t n = row0 + r;
        if (n >= p.M) continue;
        const int64_t b = n / p.n_cand;
        const int cnt = p.nk[b];
        const int64_t r0 = p.row_start[b];
        const float4* Kh = reinterpret_cast<const float4*>(p.ckv + ((int64_t)(0 * p.H + h) * p.rows_total + r0) * DH);
        const float4* Vh = reinterpret_cast<const float4*>(p.ckv + ((int64_t)(1 * p.H + h) * p.rows_total + r0) * DH);
        const float* bias = p.kbias_c + r0;
        float q[DH], acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) { q[d] = xs[r][h * DH + d] * p.scale; acc[d] = 0.f; }
        float m = MMT_NEG_INF, l = 0.f;
        // two keys per lane per pass, K and V of both in flight together
        for (int j0 = lane; j0 < cnt; j0 += 64) {
            const int j1 = j0 + 32;
            const bool has1 = j1 < cnt;
            const int j1c = has1 ? j1 : j0;
            const float4 ka0 = Kh[(int64_t)j0 * 2], kb0 = Kh[(int64_t)j0 * 2 + 1];
            const float4 ka1 = Kh[(int64_t)j1c * 2], kb1 = Kh[(int64_t)j1c * 2 + 1];
            const float4 va0 = Vh[(int64_t)j0 * 2], vb0 = Vh[(int64_t)j0 * 2 + 1];
            const float4 va1 = Vh[(int64_t)j1c * 2], vb1 = Vh[(int64_t)j1c * 2 + 1];
            float s0 = bias[j0], s1 = has1 ? bias[j1] : MMT_NEG_INF;
            s0 = fmaf(q[0], ka0.x, s0); s0 = fmaf(q[1], ka0.y, s0); s0 = fmaf(q[2], ka0.z, s0); s0 = fmaf(q[3], ka0.w, s0);
            s0 = fmaf(q[4], kb0.x, s0); s0 = fmaf(q[5], kb0.y, s0); s0 = fmaf(q[6], kb0.z, s0); s0 = fmaf(q[7], kb0.w, s0);
            s1 = fmaf(q[0], ka1.x, s1); s1 = fmaf(q[1], ka1.y, s1); s1 = fmaf(q[2], ka1.z, s1); s1 = fmaf(q[3], ka1.w, s1);
            s1 = fmaf(q[4], kb1.x, s1); s1 = fmaf(q[5], kb1.y, s1); s1 = fmaf(q[6], kb1.z, s1); s1 = fmaf(q[7], kb1.w, s1);
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
            acc[0] = fmaf(e0, va0.x, acc[0]); acc[1] = fmaf(e0, va0.y, acc[1]); acc[2] = fmaf(e0, va0.z, acc[2]); acc[3] = fmaf(e0, va0.w, acc[3]);
            acc[4] = fmaf(e0, vb0.x, acc[4]); acc[5] = fmaf(e0, vb0.y, acc[5]); acc[6] = fmaf(e0, vb0.z, acc[6]); acc[7] = fmaf(e0, vb0.w, acc[7]);
            acc[0] = fmaf(e1, va1.x, acc[0]); acc[1] = fmaf(e1, va1.y, acc[1]); acc[2] = fmaf(e1, va1.z, acc[2]); acc[3] = fmaf(e1, va1.w, acc[3]);
            acc[4] = fmaf(e1, vb1.x, acc[4]); acc[5] = fmaf(e1, vb1.y, acc[5]); acc[6] = fmaf(e1, vb1.z, acc[6]); acc[7] = fmaf(e1, vb1.w, acc[7]);
        }
        const float Mx = warp_max(m);
        const float corr = (m == MMT_NEG_INF) ? 0.f : expf(m - Mx);
        l = warp_sum(l * corr);
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = warp_sum(acc[d] * corr);
        if (lane < DH) {
            float v = 0.f;
#pragma unroll
            for (int d = 0; d < DH; ++d) if (lane == d) v = acc[d];
            att[r][h * DH + lane] = v / l;
        }
    }
    __syncthreads();

    // ---- cross out-projection -> qkv scratch, then LN2 -> x2
    gemv_rows(p.co_w, p.co_b, D, att, &qkv[0][0], 3 * D, warp, lane);
    __syncthreads();
    if (warp < DA_R) {
        const int64_t n = row0 + warp;
        if (n < p.M) {
            const float4 a = *reinterpret_cast<const float4*>(&x1s[warp][lane * 4]);
            const float4 y = *reinterpret_cast<const float4*>(&qkv[warp][lane * 4]);
            const float4 o = ln_row(make_float4(a.x + y.x, a.y + y.y, a.z + y.z, a.w + y.w), p.n2_w, p.n2_b, p.eps, lane);
            *reinterpret_cast<float4*>(p.x2 + n * D + lane * 4) = o;
            if (p.x2_16) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                *reinterpret_cast<uint2*>(p.x2_16 + n * D + lane * 4) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
        }
    }
}

}  // namespace mmt
