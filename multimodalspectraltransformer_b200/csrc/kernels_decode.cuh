// Fused, row-local kernels of the KV-cached decoder step for SMALL batches (a few hundred
// sequences): the step is then a chain of ~50 tiny dependent launches whose fixed latency
// (launch + ramp + drain, 5-20 us each) dominates.  Every operation of a decoder layer except the
// FFN is local to one sequence, so two kernels per layer do everything between the FFNs:
//
//   decode_attn_self   x  = LN3_prev(x + b2 + sum FFN2 partials) | E_tok[token] + E_pos[t]   (prologue)
//                      qkv = W_in x + b ; append K,V to the paged cache ; causal self-attention
//                      x1 = LN1(x + W_o att + b_o) ; qc = W_q^cross x1 + b
//   decode_attn_cross  att = cross-attention of qc over the projected encoder memory
//                      x2 = LN2(x1 + W_o^cross att + b)      (fp32 + bf16 operand copy for the FFN)
//
// Arithmetic is fp32 in both precision modes (these projections are 20 % of the decoder's weights
// and the larger share of the bf16 logit error, DESIGN.md "bf16 numerics").  A CTA owns DA_R rows;
// the 128-wide matrix-vector products split K over the lanes of a warp (coalesced 512 B weight rows
// straight from L2, no staging) and finish with a 31-shuffle reduce-scatter per 32 outputs.
#pragma once
#include "common.cuh"

namespace mmt {

constexpr int DA_R = 2;            // sequences per CTA
constexpr int DA_WARPS = 8;
constexpr int DA_THREADS = DA_WARPS * 32;

// v[0..31] per lane -> returns sum over lanes of v[lane]  (reduce-scatter butterfly, 31 shuffles)
__device__ __forceinline__ float reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = up ? v[i] : v[i + off];
            const float keep = up ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// out[r][n] = bias[n] + W[n,:] . xs[r,:]   for n in [0, n_out), all DA_R rows; W row-major [n_out][128]
// in global memory.  Warps take groups of 32 outputs round-robin.
__device__ __forceinline__ void gemv_rows(const float* __restrict__ W, const float* __restrict__ bias, int n_out,
                                          const float (*xs)[D], float* out, int ldo, int warp, int lane) {
    float4 xr[DA_R];
#pragma unroll
    for (int r = 0; r < DA_R; ++r) xr[r] = *reinterpret_cast<const float4*>(&xs[r][4 * lane]);
    for (int n0 = warp * 32; n0 < n_out; n0 += DA_WARPS * 32) {
        float acc[DA_R][32];
        const float4* wp = reinterpret_cast<const float4*>(W + (int64_t)n0 * D) + lane;
#pragma unroll
        for (int jb = 0; jb < 32; jb += 8) {
            float4 w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = __ldg(wp + (int64_t)(jb + j) * (D / 4));
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int r = 0; r < DA_R; ++r)
                    acc[r][jb + j] = fmaf(w[j].w, xr[r].w, fmaf(w[j].z, xr[r].z, fmaf(w[j].y, xr[r].y, w[j].x * xr[r].x)));
        }
        const float b = bias ? bias[n0 + lane] : 0.f;
#pragma unroll
        for (int r = 0; r < DA_R; ++r) out[r * ldo + n0 + lane] = reduce_scatter32(acc[r], lane) + b;
    }
}

// one warp: xs_out[row] = LN(a[row] + b[row]) * gamma + beta, lane owns 4 columns; returns the value
__device__ __forceinline__ float4 ln_row(float4 v, const float* gamma, const float* beta, float eps, int lane) {
    const float mean = warp_sum(v.x + v.y + v.z + v.w) * (1.0f / D);
    const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
    const float var = warp_sum(dx * dx + dy * dy + dz * dz + dw * dw) * (1.0f / D);
    const float rstd = rsqrtf(var + eps);
    const float4 ga = *reinterpret_cast<const float4*>(gamma + lane * 4);
    const float4 be = *reinterpret_cast<const float4*>(beta + lane * 4);
    return make_float4(dx * rstd * ga.x + be.x, dy * rstd * ga.y + be.y, dz * rstd * ga.z + be.z, dw * rstd * ga.w + be.w);
}

struct DecSelfParams {
    // ---- source of the layer input x (exactly one of the two)
    const int64_t* tokens; int tok_shift; int sos; int64_t ldn;      // layer 0: x = E_tok[token(t)] + E_pos[t]
    const float* E_tok; const float* E_pos; int vocab;
    const float* x_in;                                               // layers > 0: [M][D] (x2 of the previous layer)
    const float* part; int splits; int64_t part_stride;              // + FFN2 partial sums of the previous layer
    const float* pbias; const float* pgamma; const float* pbeta;     //   x = LN3(x_in + pbias + sum_s part[s])
    // ---- this layer
    const float *in_w, *in_b, *out_w, *out_b, *n1_w, *n1_b, *cq_w, *cq_b;
    float* kv_pool; const int* block_table; int pps;                 // paged self-attention cache of this layer
    const int* step;
    float* x1;        // out [M][D]: LN1 output (residual input of the cross block)
    float* qc;        // out [M][D]: cross-attention query
    int64_t M; int H; float scale; float eps;
};

template <int DH>
__global__ void __launch_bounds__(DA_THREADS) decode_attn_self(const __grid_constant__ DecSelfParams p) {
    __shared__ __align__(16) float xs[DA_R][D];
    __shared__ __align__(16) float qkv[DA_R][3 * D];
    __shared__ __align__(16) float att[DA_R][D];
    __shared__ __align__(16) float x1s[DA_R][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = *p.step;
    const int64_t row0 = (int64_t)blockIdx.x * DA_R;

    // ---- prologue: layer input
    if (warp < DA_R) {
        const int64_t n = row0 + warp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.M) {
            if (p.tokens) {
                int64_t tok;
                if (p.tok_shift) tok = (t == 0) ? p.sos : p.tokens[(int64_t)(t - 1) * p.ldn + n];
                else tok = p.tokens[(int64_t)t * p.ldn + n];
                if (tok < 0 || tok >= p.vocab) tok = 0;
                const float4 a = *reinterpret_cast<const float4*>(p.E_tok + tok * D + lane * 4);
                const float4 b = *reinterpret_cast<const float4*>(p.E_pos + (int64_t)t * D + lane * 4);
                v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            } else {
                v = *reinterpret_cast<const float4*>(p.x_in + n * D + lane * 4);
                if (p.part) {
                    float4 s = *reinterpret_cast<const float4*>(p.part + n * D + lane * 4);
                    for (int k = 1; k < p.splits; ++k) {
                        const float4 q = *reinterpret_cast<const float4*>(p.part + (int64_t)k * p.part_stride + n * D + lane * 4);
                        s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w;
                    }
                    const float4 b = *reinterpret_cast<const float4*>(p.pbias + lane * 4);
                    s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
                    v = ln_row(make_float4(s.x + v.x, s.y + v.y, s.z + v.z, s.w + v.w), p.pgamma, p.pbeta, p.eps, lane);
                }
            }
        }
        *reinterpret_cast<float4*>(&xs[warp][lane * 4]) = v;
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
        const int* bt = p.block_table + n * p.pps;
        const float* row = &qkv[r][h * DH];
        if (lane < 2 * DH) {
            const int kv = lane / DH, d = lane % DH;
            float* page = p.kv_pool + (int64_t)bt[t / PAGE_TOKENS] * PAGE_FLOATS;
            page[((kv * p.H + h) * PAGE_TOKENS + (t % PAGE_TOKENS)) * DH + d] = row[(1 + kv) * D + d];
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
            s[i] = MMT_NEG_INF;
            if (j <= t) {
                const float* page = p.kv_pool + (int64_t)bt[j / PAGE_TOKENS] * PAGE_FLOATS;
                const float* k = page + ((0 * p.H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH;
                float a = 0.f;
#pragma unroll
                for (int d = 0; d < DH; ++d) a = fmaf(q[d], k[d], a);
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
            if (j <= t) {
                const float e = expf(s[i] - m);
                l += e;
                const float* page = p.kv_pool + (int64_t)bt[j / PAGE_TOKENS] * PAGE_FLOATS;
                const float* v = page + ((1 * p.H + h) * PAGE_TOKENS + (j % PAGE_TOKENS)) * DH;
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] = fmaf(e, v[d], acc[d]);
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
        const int64_t n = row0 + warp;
        const float4 a = *reinterpret_cast<const float4*>(&xs[warp][lane * 4]);
        const float4 y = *reinterpret_cast<const float4*>(&qkv[warp][lane * 4]);
        const float4 o = ln_row(make_float4(a.x + y.x, a.y + y.y, a.z + y.z, a.w + y.w), p.n1_w, p.n1_b, p.eps, lane);
        *reinterpret_cast<float4*>(&x1s[warp][lane * 4]) = o;
        if (n < p.M) *reinterpret_cast<float4*>(p.x1 + n * D + lane * 4) = o;
    }
    __syncthreads();

    // ---- cross-attention query projection
    gemv_rows(p.cq_w, p.cq_b, D, x1s, &att[0][0], D, warp, lane);
    __syncthreads();
    if (warp < DA_R) {
        const int64_t n = row0 + warp;
        if (n < p.M) *reinterpret_cast<float4*>(p.qc + n * D + lane * 4) = *reinterpret_cast<const float4*>(&att[warp][lane * 4]);
    }
}

struct DecCrossParams {
    const float* qc; const float* x1;          // [M][D]
    const float* ckv; int64_t rows_total;      // projected memory of this layer, head-major [2][H][rows_total][DH]
    const int* nk; const int* row_start; const float* kbias_c; int n_cand;
    const float *co_w, *co_b, *n2_w, *n2_b;
    float* x2; __nv_bfloat16* x2_16;           // out [M][D] fp32 (+ bf16 operand copy, optional)
    int64_t M; int H; float scale; float eps;
};

template <int DH>
__global__ void __launch_bounds__(DA_THREADS) decode_attn_cross(const __grid_constant__ DecCrossParams p) {
    __shared__ __align__(16) float qs[DA_R][D];
    __shared__ __align__(16) float att[DA_R][D];
    __shared__ __align__(16) float ys[DA_R][D];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * DA_R;
    if (warp < DA_R) {
        const int64_t n = row0 + warp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < p.M) v = *reinterpret_cast<const float4*>(p.qc + n * D + lane * 4);
        *reinterpret_cast<float4*>(&qs[warp][lane * 4]) = v;
    }
    __syncthreads();
    for (int pair = warp; pair < DA_R * p.H; pair += DA_WARPS) {
        const int r = pair / p.H, h = pair % p.H;
        const int64_t n = row0 + r;
        if (n >= p.M) continue;
        const int64_t b = n / p.n_cand;
        const int cnt = p.nk[b];
        const int64_t r0 = p.row_start[b];
        const float* Kh = p.ckv + ((int64_t)(0 * p.H + h) * p.rows_total + r0) * DH;
        const float* Vh = p.ckv + ((int64_t)(1 * p.H + h) * p.rows_total + r0) * DH;
        const float* bias = p.kbias_c + r0;
        float q[DH], acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) { q[d] = qs[r][h * DH + d] * p.scale; acc[d] = 0.f; }
        float m = MMT_NEG_INF, l = 0.f;
        for (int j = lane; j < cnt; j += 32) {
            float s = bias[j];
            const float4* kp = reinterpret_cast<const float4*>(Kh + (int64_t)j * DH);
#pragma unroll
            for (int d4 = 0; d4 < DH / 4; ++d4) {
                const float4 k = kp[d4];
                s = fmaf(q[d4 * 4], k.x, s); s = fmaf(q[d4 * 4 + 1], k.y, s);
                s = fmaf(q[d4 * 4 + 2], k.z, s); s = fmaf(q[d4 * 4 + 3], k.w, s);
            }
            if (s > m) {
                const float corr = expf(m - s);
                l *= corr;
#pragma unroll
                for (int d = 0; d < DH; ++d) acc[d] *= corr;
                m = s;
            }
            const float e = expf(s - m);
            l += e;
            const float4* vp = reinterpret_cast<const float4*>(Vh + (int64_t)j * DH);
#pragma unroll
            for (int d4 = 0; d4 < DH / 4; ++d4) {
                const float4 v = vp[d4];
                acc[d4 * 4] = fmaf(e, v.x, acc[d4 * 4]); acc[d4 * 4 + 1] = fmaf(e, v.y, acc[d4 * 4 + 1]);
                acc[d4 * 4 + 2] = fmaf(e, v.z, acc[d4 * 4 + 2]); acc[d4 * 4 + 3] = fmaf(e, v.w, acc[d4 * 4 + 3]);
            }
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
    gemv_rows(p.co_w, p.co_b, D, att, &ys[0][0], D, warp, lane);
    __syncthreads();
    if (warp < DA_R) {
        const int64_t n = row0 + warp;
        if (n < p.M) {
            const float4 a = *reinterpret_cast<const float4*>(p.x1 + n * D + lane * 4);
            const float4 y = *reinterpret_cast<const float4*>(&ys[warp][lane * 4]);
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
