// Batched beam search bookkeeping (reference validate_generate_MMT_v15_4.py:995-1086).
//
// The reference keeps, per memory column, a Python list of <= K beams (score, sequence, prob_sequence)
// and per step re-runs the decoder on every beam's full prefix, one beam and one item at a time.  Here
// the K beams of every item are K slots of one KV-cached decode wave (slot n = item * K + beam, cross
// K/V shared by index): one decoder step serves all items x beams, and these kernels do the
// reference's list manipulation on the device:
//   beam_select      one CTA per item: softmax of the newest logits (no temperature, :1038), top-K per
//                    live beam, carried-over finished beams, stable sort by the double-precision score
//                    product, new histories / lengths / scores, and the KV page plan of the new slots;
//   beam_copy_pages  copy-on-write of the one open self-attention page per slot and layer.
// KV inheritance: a slot's closed pages (16 tokens) are immutable, so children inherit them by block
// table entry; only the open page is copied, into the slot's own page of the other step parity
// (every live beam is at the same position, so sources [parity t&1] and destinations [(t+1)&1] never
// overlap).  Physical page of (slot n, page p, parity a) = (a * Nw + n) * pps + p.
#pragma once
#include "common.cuh"

namespace mmt {

constexpr int BEAM_MAX = 32;

struct BeamParams {
    const float* logits;     // [Nw][V] logits of the newest position
    int V, K, T, eos, pps;
    int64_t Nw;              // slots of this wave (items * K)
    const int* step;         // device step counter, already advanced by the sampler: t = *step - 1
    double* score;           // [Nw]
    int* len;                // [Nw] tokens in the slot's sequence, <SOS> included
    int64_t* hist;           // [Nw][T+1]
    float* probs;            // [Nw][T]
    int64_t* cur_tok;        // [Nw] input token of the next step
    int* block_table;        // [Nw][pps]
    int* copy_src; int* copy_dst;   // [Nw] page to copy for the next step (-1: none)
    int* unfinished;         // [T] selected beams still open after step t (early exit)
};

__global__ void beam_init(const BeamParams p, int sos) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= p.Nw) return;
    p.score[n] = 1.0;
    p.len[n] = 1;
    p.cur_tok[n] = sos;
    p.copy_src[n] = -1; p.copy_dst[n] = -1;
    for (int i = 0; i <= p.T; ++i) p.hist[n * (p.T + 1) + i] = (i == 0) ? sos : 0;
    for (int i = 0; i < p.T; ++i) p.probs[n * p.T + i] = 0.f;
    for (int q = 0; q < p.pps; ++q) p.block_table[n * p.pps + q] = (int)(n * p.pps + q);
}

__global__ void __launch_bounds__(256) beam_select(const __grid_constant__ BeamParams p) {
    __shared__ double c_score[BEAM_MAX * BEAM_MAX];
    __shared__ float c_prob[BEAM_MAX * BEAM_MAX];
    __shared__ int8_t c_tok[BEAM_MAX * BEAM_MAX];      // -1: finished beam carried over unchanged
    __shared__ uint8_t c_par[BEAM_MAX * BEAM_MAX];
    __shared__ float p_old[BEAM_MAX][128];
    __shared__ uint8_t h_old[BEAM_MAX][132];
    __shared__ int bt_old[BEAM_MAX][8];
    __shared__ int len_old[BEAM_MAX], base[BEAM_MAX + 1], sel[BEAM_MAX];
    __shared__ uint8_t fin[BEAM_MAX];

    const int K = p.K, T = p.T, t = *p.step - 1;
    const int64_t n0 = (int64_t)blockIdx.x * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nalive = (t == 0) ? 1 : K;              // the search starts from the single [<SOS>] beam

    // ---- old state of the item's slots
    if (threadIdx.x < K) {
        const int L = p.len[n0 + threadIdx.x];
        len_old[threadIdx.x] = L;
        fin[threadIdx.x] = p.hist[(n0 + threadIdx.x) * (T + 1) + L - 1] == p.eos;
    }
    for (int i = threadIdx.x; i < K * (T + 1); i += blockDim.x) h_old[i / (T + 1)][i % (T + 1)] = (uint8_t)p.hist[n0 * (T + 1) + i];
    for (int i = threadIdx.x; i < K * T; i += blockDim.x) p_old[i / T][i % T] = p.probs[n0 * T + i];
    for (int i = threadIdx.x; i < K * p.pps; i += blockDim.x) bt_old[i / p.pps][i % p.pps] = p.block_table[n0 * p.pps + i];
    __syncthreads();
    if (threadIdx.x == 0) {
        base[0] = 0;
        for (int k = 0; k < nalive; ++k) base[k + 1] = base[k] + (fin[k] ? 1 : K);
    }
    __syncthreads();
    const int ncand = base[nalive];

    // ---- candidates in the reference's order: beam by beam, children by descending probability
    for (int k = warp; k < nalive; k += 8) {
        const double sc = p.score[n0 + k];
        if (fin[k]) {
            if (lane == 0) { c_score[base[k]] = sc; c_prob[base[k]] = 0.f; c_tok[base[k]] = -1; c_par[base[k]] = (uint8_t)k; }
            continue;
        }
        const float* lg = p.logits + (n0 + k) * p.V;
        const int v0 = lane, v1 = lane + 32;
        const bool ok0 = v0 < p.V, ok1 = v1 < p.V;
        const float z0 = ok0 ? lg[v0] : MMT_NEG_INF, z1 = ok1 ? lg[v1] : MMT_NEG_INF;
        const float mx = warp_max(fmaxf(z0, z1));
        const float e0 = ok0 ? expf(z0 - mx) : 0.f, e1 = ok1 ? expf(z1 - mx) : 0.f;
        const float sum = warp_sum(e0 + e1);
        float p0 = ok0 ? e0 / sum : -1.f, p1 = ok1 ? e1 / sum : -1.f;
        for (int r = 0; r < K; ++r) {             // top-K: K rounds of a warp arg-max, lowest index on ties
            float best = p0; int bi = v0;
            if (p1 > best) { best = p1; bi = v1; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (bi == v0) p0 = -1.f;
            if (bi == v1) p1 = -1.f;
            if (lane == 0) {
                const int c = base[k] + r;
                c_score[c] = sc * (double)best;   // Python float product (:1050)
                c_prob[c] = best; c_tok[c] = (int8_t)bi; c_par[c] = (uint8_t)k;
            }
        }
    }
    __syncthreads();

    // ---- new_beam.sort(key=score, reverse=True)[:K] -- stable: ties keep candidate order
    for (int i = threadIdx.x; i < ncand; i += blockDim.x) {
        const double s = c_score[i];
        int rank = 0;
        for (int j = 0; j < ncand; ++j) rank += (c_score[j] > s) || (c_score[j] == s && j < i);
        if (rank < K) sel[rank] = i;
    }
    __syncthreads();

    // ---- new state of the K slots
    const int pg = t / PAGE_TOKENS, r16 = t % PAGE_TOKENS;
    int open = 0;
    if (threadIdx.x < K) {
        const int j = threadIdx.x, c = sel[j], q = c_par[c], tok = c_tok[c];
        const int64_t n = n0 + j;
        p.len[n] = len_old[q] + (tok >= 0 ? 1 : 0);
        p.score[n] = c_score[c];
        p.cur_tok[n] = tok >= 0 ? tok : p.eos;
        open = tok >= 0 && tok != p.eos;
        for (int pp = 0; pp < pg; ++pp) p.block_table[n * p.pps + pp] = bt_old[q][pp];
        if (pg < p.pps) {
            if (r16 == PAGE_TOKENS - 1) {             // the page just closed: inherit it, the next page starts empty
                p.block_table[n * p.pps + pg] = bt_old[q][pg];
                p.copy_src[n] = -1; p.copy_dst[n] = -1;
            } else {
                const int dst = (int)((((int64_t)((t + 1) & 1)) * p.Nw + n) * p.pps + pg);
                p.block_table[n * p.pps + pg] = dst;
                p.copy_src[n] = bt_old[q][pg]; p.copy_dst[n] = dst;
            }
        }
    }
    for (int i = threadIdx.x; i < K * (T + 1); i += blockDim.x) {
        const int j = i / (T + 1), pos = i % (T + 1), c = sel[j], q = c_par[c], tok = c_tok[c], L = len_old[q];
        p.hist[n0 * (T + 1) + i] = pos < L ? (int64_t)h_old[q][pos] : ((pos == L && tok >= 0) ? (int64_t)tok : 0);
    }
    for (int i = threadIdx.x; i < K * T; i += blockDim.x) {
        const int j = i / T, pos = i % T, c = sel[j], q = c_par[c], tok = c_tok[c], L = len_old[q];
        p.probs[n0 * T + i] = pos < L - 1 ? p_old[q][pos] : ((pos == L - 1 && tok >= 0) ? c_prob[c] : 0.f);
    }
    const int n_open = __syncthreads_count(open);
    if (threadIdx.x == 0 && n_open) atomicAdd(p.unfinished + t, n_open);
}

// grid (Nw, layers): whole-page copy (K and V of 16 tokens, all heads) inside one layer's pool
__global__ void __launch_bounds__(256) beam_copy_pages(char* kv_pool, size_t layer_bytes, int page_bytes, const int* copy_src, const int* copy_dst) {
    const int src = copy_src[blockIdx.x];
    if (src < 0) return;
    const int dst = copy_dst[blockIdx.x];
    char* pool = kv_pool + (size_t)blockIdx.y * layer_bytes;
    const uint4* s = reinterpret_cast<const uint4*>(pool + (size_t)src * page_bytes);
    uint4* d = reinterpret_cast<uint4*>(pool + (size_t)dst * page_bytes);
    for (int i = threadIdx.x; i < page_bytes / 16; i += blockDim.x) d[i] = s[i];
}

}  // namespace mmt
