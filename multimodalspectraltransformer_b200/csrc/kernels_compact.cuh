// Ragged encoder: compute every DISTINCT token row once.
//
// The reference pads each peak list / formula to 64 tokens and runs all 582 rows of a spectrum
// through the six encoder stacks (models_MMT_v15_4.py:549-731, 846-944).  Under a bool key-padding
// mask the padded rows are never attended to, and -- being identical inputs (zero peak rows ->
// relu(b); MF/MS id 0 -> relu(E[0])) with identical keys -- all padded rows of one segment yield
// identical outputs in every layer, including encoder_cross.  So a spectrum is
// encoded on its COMPACT rows: per modality sequence [valid X | one padded X | valid MF | one padded
// MF | (MS likewise) | MW], and the padded outputs are replicated when the (582, B, 128) memory is
// written.  The result is the same tensor (every padded row carries the value the dense computation
// gives it), at 160-230 instead of 582 rows per spectrum for realistic peak counts.
//
// The kernels below build the index maps; `flag` is raised if some padded row of a segment does
// not carry the same raw input as the segment's first padded row (a caller that hides data under
// the mask), in which case the host takes the dense path for that batch.
#pragma once
#include "common.cuh"

namespace mmt {

struct CompactParams {
    const float* mask[4]; const float* src[4]; int cols[4];      // peak-list modalities (B,P[,cols]); nullptr if absent
    const uint8_t* mask_MF; const int64_t* src_MF;
    const uint8_t* mask_MS; const int64_t* src_MS;
    int present[5], n_x[5];
    int has_MF, has_MS, has_MW, P, B;
    int* cnt;     // [5][B]           compact rows of (modality, spectrum)
    int* nkeys;   // [5][B]           of which attendable (valid) rows
    int* d2c;     // [5][B][CP_SMAX]  dense position -> compact-local row
    int* kidx;    // [5][B][CP_SMAX]  compact-local indices of the valid rows
    int* flag;    // [1]
};

// one warp per (spectrum, modality)
__global__ void __launch_bounds__(32) compact_index(const __grid_constant__ CompactParams p) {
    const int b = blockIdx.x, m = blockIdx.y, lane = threadIdx.x;
    if (!p.present[m]) { if (lane == 0) { p.cnt[m * p.B + b] = 0; p.nkeys[m * p.B + b] = 0; } return; }
    int* d2c = p.d2c + ((int64_t)m * p.B + b) * CP_SMAX;
    int* kidx = p.kidx + ((int64_t)m * p.B + b) * CP_SMAX;
    const unsigned lt = (1u << lane) - 1u;
    int dpos = 0, crow = 0, nk = 0;
    // segment kinds: 0 peak list of modality m, 1 MF, 2 MS, 3 always-valid single token (IR row, MW)
    auto is_pad = [&](int kind, int i) -> bool {
        if (kind == 0) return p.mask[m][(int64_t)b * p.P + i] != 0.f;
        if (kind == 1) return p.mask_MF[(int64_t)b * p.P + i] != 0;
        if (kind == 2) return p.mask_MS[(int64_t)b * p.P + i] != 0;
        return false;
    };
    auto same = [&](int kind, int i, int j) -> bool {      // raw inputs of positions i and j identical?
        if (kind == 0) {
            const float* x = p.src[m] + (int64_t)b * p.P * p.cols[m];
            bool eq = true;
            for (int c = 0; c < p.cols[m]; ++c) eq = eq && (__float_as_uint(x[i * p.cols[m] + c]) == __float_as_uint(x[j * p.cols[m] + c]));
            return eq;
        }
        if (kind == 1) return p.src_MF[(int64_t)b * p.P + i] == p.src_MF[(int64_t)b * p.P + j];
        if (kind == 2) return p.src_MS[(int64_t)b * p.P + i] == p.src_MS[(int64_t)b * p.P + j];
        return true;
    };
    auto segment = [&](int kind, int len) {
        int nvalid = 0, first_pad = -1;
        for (int base = 0; base < len; base += 32) {
            const int i = base + lane;
            const bool in = i < len;
            const bool pad = in && is_pad(kind, i);
            const bool valid = in && !pad;
            const unsigned mv = __ballot_sync(0xffffffffu, valid), mp = __ballot_sync(0xffffffffu, pad);
            if (valid) {
                const int c = crow + nvalid + __popc(mv & lt);
                d2c[dpos + i] = c;
                kidx[nk + nvalid + __popc(mv & lt)] = c;
            }
            if (first_pad < 0 && mp) first_pad = base + __ffs(mp) - 1;
            nvalid += __popc(mv);
        }
        if (first_pad >= 0) {
            const int rep = crow + nvalid;
            bool bad = false;
            for (int base = 0; base < len; base += 32) {
                const int i = base + lane;
                if (i < len && is_pad(kind, i)) {
                    d2c[dpos + i] = rep;
                    if (!same(kind, i, first_pad)) bad = true;
                }
            }
            if (bad) atomicOr(p.flag, 1);
        }
        dpos += len;
        crow += nvalid + (first_pad >= 0 ? 1 : 0);
        nk += nvalid;
    };
    if (m < 4) segment(0, p.n_x[m]); else segment(3, 1);
    if (p.has_MF) segment(1, p.P);
    if (p.has_MS) segment(2, p.P);
    if (p.has_MW) segment(3, 1);
    if (lane == 0) { p.cnt[m * p.B + b] = crow; p.nkeys[m * p.B + b] = nk; }
}

struct CompactScanParams {
    const int* cnt; const int* nkeys; int B;
    int* row_start;   // [5][B+1]  prefix over spectra of cnt[m][.]
    int* cstart;      // [B+1]     prefix over spectra of the per-spectrum total (cross-encoder rows)
    int* moff;        // [5][B]    offset of modality m inside a spectrum's cross rows
    int* nk_c;        // [B]       attendable cross rows of a spectrum
    int* ccnt;        // [B]       cross rows of a spectrum
    int* totals;      // [16]: rows of modality 0..4, cross rows, flag copy, max cross keys / rows per spectrum,
                      //       [9+m] max rows per spectrum of modality m (host reads these)
    const int* flag;
};
// single CTA: six serial prefix sums over <= 256 spectra
__global__ void __launch_bounds__(256) compact_scan(const CompactScanParams p) {
    const int t = threadIdx.x;
    for (int b = t; b < p.B; b += blockDim.x) {
        int off = 0, nk = 0;
        for (int m = 0; m < 5; ++m) { p.moff[m * p.B + b] = off; off += p.cnt[m * p.B + b]; nk += p.nkeys[m * p.B + b]; }
        p.nk_c[b] = nk;
        p.ccnt[b] = off;
    }
    __syncthreads();
    if (t < 5) {
        int run = 0;
        for (int b = 0; b < p.B; ++b) { p.row_start[t * (p.B + 1) + b] = run; run += p.cnt[t * p.B + b]; }
        p.row_start[t * (p.B + 1) + p.B] = run;
        p.totals[t] = run;
    } else if (t == 5) {
        int run = 0;
        for (int b = 0; b < p.B; ++b) {
            p.cstart[b] = run;
            for (int m = 0; m < 5; ++m) run += p.cnt[m * p.B + b];
        }
        p.cstart[p.B] = run;
        p.totals[5] = run;
        p.totals[6] = *p.flag;
    } else if (t == 6) {
        int mk = 0, mr = 0;
        for (int b = 0; b < p.B; ++b) { mk = max(mk, p.nk_c[b]); mr = max(mr, p.ccnt[b]); }
        p.totals[7] = mk; p.totals[8] = mr;
    } else if (t >= 7 && t < 12) {
        const int m = t - 7;
        int mr = 0;
        for (int b = 0; b < p.B; ++b) mr = max(mr, p.cnt[m * p.B + b]);
        p.totals[9 + m] = mr;
    }
}

struct CompactCrossParams {
    const int* cnt; const int* nkeys; const int* kidx; const int* row_start; const int* cstart; const int* moff; int B;
    int* out_rows;    // [5][R_m] modality-compact row -> cross-compact row (global)
    int64_t out_rows_stride;
    int* kidx_c;      // [B][kc_stride] cross-compact-local indices of attendable rows
    int kc_stride;
};
// one warp per spectrum
__global__ void __launch_bounds__(32) compact_cross_index(const CompactCrossParams p) {
    const int b = blockIdx.x, lane = threadIdx.x;
    int nk = 0;
    for (int m = 0; m < 5; ++m) {
        const int c = p.cnt[m * p.B + b], k = p.nkeys[m * p.B + b], mo = p.moff[m * p.B + b];
        const int rs = p.row_start[m * (p.B + 1) + b], cs = p.cstart[b];
        for (int j = lane; j < c; j += 32) p.out_rows[m * p.out_rows_stride + rs + j] = cs + mo + j;
        const int* ki = p.kidx + ((int64_t)m * p.B + b) * CP_SMAX;
        for (int j = lane; j < k; j += 32) p.kidx_c[(int64_t)b * p.kc_stride + nk + j] = mo + ki[j];
        nk += k;
    }
}

// memory[s][b0 + b][:] = Y[cstart[b] + moff[m][b] + d2c[m][b][s - off_m]]   (one warp per (b, s) row)
struct ExpandParams {
    const float* Y; const int* cstart; const int* moff; const int* d2c;
    int B, S_total; int off[5], S_m[5];
    float* memory; int64_t B_total; int b0;
};
__global__ void __launch_bounds__(256) expand_memory(const __grid_constant__ ExpandParams p) {
    const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (w >= (int64_t)p.B * p.S_total) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(w / p.S_total), s = (int)(w % p.S_total);
    int m = 0;
#pragma unroll
    for (int k = 1; k < 5; ++k) if (s >= p.off[k]) m = k;
    const int sl = s - p.off[m];
    const int src = p.cstart[b] + p.moff[m * p.B + b] + p.d2c[((int64_t)m * p.B + b) * CP_SMAX + sl];
    const float4 v = *reinterpret_cast<const float4*>(p.Y + (int64_t)src * D + lane * 4);
    *reinterpret_cast<float4*>(p.memory + ((int64_t)s * p.B_total + p.b0 + b) * D + lane * 4) = v;
}

}  // namespace mmt
