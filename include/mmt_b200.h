/*
 * mmt_b200.h -- C ABI of the B200-native MMT candidate-generation engine.
 *
 * Drop-in boundary for ONE path of mpriessner/MultiModalSpectralTransformer:
 * spectra -> encoder memory -> greedy / multinomial SMILES token ids.
 * The reference has no FFI of its own (pure Python); each entry point below
 * names the reference function whose arithmetic it replaces (paths relative to
 * the reference's utils_MMT/).  The Python host in
 * multimodalspectraltransformer_b200/ binds these with ctypes and re-exposes the
 * reference's Python signatures.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer named d_* is DEVICE memory on
 *     the engine's device, h_* is HOST memory.
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*)
 *     and never synchronises unless documented.
 *   - threads and streams: an engine owns ONE workspace, so calls on one handle
 *     are serialised -- the library takes a per-engine mutex for the duration of
 *     each call (callers may use a handle from several threads), and a call
 *     issued on a different stream than the previous call on that handle first
 *     waits, on the device, for the previous call's work (an event recorded at
 *     the end of every call).  Two handles never share state.
 *   - return value: 0 on success, non-zero on failure; mmt_last_error() then
 *     returns a thread-local, NUL-terminated description.  There is no CPU
 *     fallback: without a usable sm_100 device mmt_create fails.
 */
#ifndef MMT_B200_H
#define MMT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMT_ABI_VERSION 2

typedef struct mmt_engine mmt_engine;

/* Hyper-parameters read from the reference's config (config_V8.json). */
typedef struct mmt_model_desc {
    int32_t d_model;        /* hidden_size        = 128 (only value supported) */
    int32_t n_heads;        /* num_heads          = 16  (modality encoders, decoder) */
    int32_t n_heads_cross;  /* int(num_heads / 4) = 4   (encoder_cross, models_MMT_v15_4.py:532) */
    int32_t d_ff;           /* 2048, torch default dim_feedforward (never overridden, :510-541) */
    int32_t n_enc_layers;   /* num_encoder_layers = 6 */
    int32_t n_dec_layers;   /* num_decoder_layers = 6 */
    int32_t vocab;          /* in_size = out_size = 43 */
    int32_t max_len;        /* rows of pe_trg     = 128 */
    int32_t mf_vocab;       /* MF_vocab_size      = 212 */
    int32_t ms_vocab;       /* MS_vocab_size      = 43 */
    int32_t ir_bins;        /* input_dim_IR       = 1000 */
    int32_t fp_size;        /* fingerprint_size   = 512 */
    int32_t pad_points;     /* padding_points_number = 64 */
} mmt_model_desc;

/* training_mode -> bit set ("1H" in config.training_mode, ...). */
enum {
    MMT_MODE_1H = 1 << 0, MMT_MODE_13C = 1 << 1, MMT_MODE_HSQC = 1 << 2, MMT_MODE_COSY = 1 << 3,
    MMT_MODE_IR = 1 << 4, MMT_MODE_MF = 1 << 5, MMT_MODE_MS = 1 << 6, MMT_MODE_MW = 1 << 7
};

/* Arithmetic of the GEMM-shaped work.  FP32 = SIMT FMA "check mode" (greedy ids
 * reproduce the reference's fp32 path); BF16 = tcgen05 tensor-core mode (bf16
 * operands, fp32 accumulate / residual stream / LayerNorm / softmax). */
enum { MMT_PREC_FP32 = 0, MMT_PREC_BF16 = 1 };

enum { MMT_SAMPLE_GREEDY = 0, MMT_SAMPLE_MULTINOMIAL = 1 };

/* Collated spectra, the tensor contract of dataloaders_pl_v15_4.py:665-712
 * (SURVEY.md A.1).  A pointer may be NULL iff its modality bit is not set. */
typedef struct mmt_spectra {
    const float*   d_src_1H;   const float* d_mask_1H;    /* (B,64,2) , (B,64) 0 valid / non-zero pad */
    const float*   d_src_13C;  const float* d_mask_13C;   /* (B,64)   , (B,64) */
    const float*   d_src_HSQC; const float* d_mask_HSQC;  /* (B,64,2) , (B,64) */
    const float*   d_src_COSY; const float* d_mask_COSY;  /* (B,64,2) , (B,64) */
    const float*   d_src_IR;                              /* (B,ir_bins); mask_IR is ignored by the model (:766) */
    const int64_t* d_src_MF;   const uint8_t* d_mask_MF;  /* (B,64) ids, (B,64) 1 = pad */
    const int64_t* d_src_MS;   const uint8_t* d_mask_MS;  /* (B,64) ids, (B,64) 1 = pad */
    const float*   d_trg_MW;                              /* (B) */
} mmt_spectra;

/* ---- lifetime ------------------------------------------------------------- */

int32_t mmt_abi_version(void);
const char* mmt_last_error(void);

/* Weight slots of the fp32 blob.  Names are the reference's state_dict keys
 * (models_MMT_v15_4.py:494-546) so the host packs a checkpoint by
 * name: blob[mmt_weight_offset(i) : +mmt_weight_numel(i)] =
 * state_dict[mmt_weight_name(i)].float().flatten(); gaps (alignment) are ignored. */
int32_t     mmt_weight_count(const mmt_model_desc* desc);
const char* mmt_weight_name(const mmt_model_desc* desc, int32_t i);
int64_t     mmt_weight_numel(const mmt_model_desc* desc, int32_t i);
int64_t     mmt_weight_offset(const mmt_model_desc* desc, int32_t i);
int64_t     mmt_weight_total(const mmt_model_desc* desc);

/* Replaces: MultimodalTransformer.__init__ + load_MMT_model (.to("cuda"))
 * (models_MMT_v15_4.py:488-546, mmt_result_test_functions_15_4.py:467-475).
 * Copies the fp32 blob (h_weights, n_floats == mmt_weight_total) to `device`
 * and derives the bf16 operand copies.  Synchronous. */
int32_t mmt_create(const mmt_model_desc* desc, const float* h_weights, int64_t n_floats,
                   int32_t device, mmt_engine** out);
void    mmt_destroy(mmt_engine* e);

/* ---- encoder ---------------------------------------------------------------- */

/* Length S of the encoder memory for a mode (582 for all five spectra + MF + MW;
 * blank-modality and MS rules of models_MMT_v15_4.py:834-939). */
int32_t mmt_memory_len(const mmt_model_desc* desc, uint32_t mode_bits);
/* 1 when the reference's concatenated key-padding mask is promoted to float
 * (a 1H/13C/HSQC/COSY modality absent from training_mode). */
int32_t mmt_mask_is_float(uint32_t mode_bits);

/* Replaces: vgmmt.run_model / MultimodalTransformer.forward(trg=None)
 * (validate_generate_MMT_v15_4.py:95-267, models_MMT_v15_4.py:803-953):
 * embedders -> 5 modality encoders -> encoder_cross -> mean -> fp1.
 *   d_memory        out (S,B,128) f32, sequence-first like the reference
 *   d_embedding_src out (S,B,128) f32 or NULL (forward()'s 2nd return)
 *   d_key_bias      out (B,S) f32: additive attention bias of each memory row,
 *                   0 valid, -inf padded, or the float mask value (0 / 1) in
 *                   float-mask modes
 *   d_pad_mask      out (B,S) u8: 1 where the reference's mask is True / 1.0
 *   d_fingerprint   out (B,fp_size) f32
 *   d_avg_memory    out (B,128) f32 or NULL (mean over S, input of fp1)        */
int32_t mmt_encode(mmt_engine* e, const mmt_spectra* in, int32_t B, uint32_t mode_bits,
                   int32_t precision, float* d_memory, float* d_embedding_src, float* d_key_bias,
                   uint8_t* d_pad_mask, float* d_fingerprint, float* d_avg_memory, void* stream);

/* Bitwise equality of the collated inputs of two encode calls (all arrays of the modalities in mode_bits, B spectra).
 * Serves the repeated-encode pattern of the callers -- CLIP's forward(trg=None) on a batch run_model has just encoded
 * (models_CLIP_v15_4.py:278-285, 337-347): the host keeps the last encode's inputs and outputs and
 * skips the second encode when this returns 1.  One small kernel, a 4-byte device-to-host read, synchronises `stream`.
 * mmt_encode itself accepts d_memory == NULL with d_embedding_src != NULL for an embedding-only pass (forward()'s
 * second output when the rest comes from the cache). */
int32_t mmt_spectra_equal(mmt_engine* e, const mmt_spectra* a, const mmt_spectra* b, int32_t B, uint32_t mode_bits,
                          int32_t* h_equal, void* stream);

/* ---- decoder ---------------------------------------------------------------- */

typedef struct mmt_decode_args {
    /* encoder memory as the reference hands it to the decode loops: element
     * (s,b,d) at d_memory[s*stride_s + b*stride_b + d]; any caller slicing /
     * duplication is expressible (validate_generate_MMT_v15_4.py:1083-1084). */
    const float* d_memory; int64_t stride_s; int64_t stride_b;
    const float* d_key_bias;       /* (Bm,S) additive bias, row stride S */
    int32_t S;                     /* memory rows */
    int32_t Bm;                    /* distinct memories */
    int32_t n_cand;                /* sequences per memory; sequence n reads memory n / n_cand.
                                      The reference gets 128 by tensor duplication
                                      (run_batch_gen_val_MMT_v15_4.py:93-107); here K/V are shared. */
    int32_t max_len;               /* config.max_len at call time (<= desc.max_len) */
    float   temperature;           /* config.temperature at call time */
    int32_t sampling;              /* MMT_SAMPLE_* */
    int32_t stop_on_all_pad;       /* greedy_sequence's early exit (validate_generate_MMT_v15_4.py:763) */
    int32_t precision;             /* MMT_PREC_* */
    /* Philox state of torch's CUDA generator for torch.multinomial parity
     * (SURVEY.md appendix D).  One "torch.multinomial((N_total,43),1)" call per
     * step: step t uses offset philox_offset + t*mmt_philox_increment(N_total*vocab).
     * seq_index_base / N_total place this call's sequences inside a larger
     * logical batch so that sharded runs draw the numbers of the unsharded run. */
    uint64_t philox_seed; uint64_t philox_offset;
    int64_t  seq_index_base; int64_t N_total;
    int32_t  rng_sm_count; int32_t rng_max_threads_per_sm;   /* 0 = this device's */
} mmt_decode_args;

/* Replaces the decode loops greedy_sequence / multinomial_sequence /
 * multinomial_sequence_multi (validate_generate_MMT_v15_4.py:723-775, 841-880;
 * run_batch_gen_val_MMT_v15_4.py:121-158): KV-cached, one step per position.
 *   d_tokens out (max_len, N) i64  -- generated ids, WITHOUT the <SOS> row
 *   d_probs  out (max_len, N) f32  -- softmax(logits/T) of the chosen id, every step
 *   h_steps  out host int: steps executed (max_len, or the step at which every
 *            sequence emitted <PAD> when stop_on_all_pad).  Synchronises `stream`
 *            once at the end when h_steps != NULL and stop_on_all_pad is set.   */
int32_t mmt_decode(mmt_engine* e, const mmt_decode_args* a, int64_t* d_tokens, float* d_probs,
                   int32_t* h_steps, void* stream);

/* Replaces the decoder tail of MultimodalTransformer.forward(..., trg_SMI_input)
 * (models_MMT_v15_4.py:955-976) and the teacher-forced scorers' inner call:
 * d_trg (T,N) i64 -> d_logits (T,N,vocab) f32 (fc_out, before softmax). */
int32_t mmt_teacher_forced(mmt_engine* e, const mmt_decode_args* a, const int64_t* d_trg, int32_t T,
                           float* d_logits, void* stream);

/* Replaces the teacher-forced scorers predict_prop_correct_max_sequence(_2) (validate_generate_MMT_v15_4.py:309-509)
 * and mrtf.predict_prop_correct_max_sequence_3 (mmt_result_test_functions_15_4.py:340-400): T positions with the given
 * input tokens d_trg_in (T,N) (row 0 = <SOS>), one KV-cached step per position instead of a full-prefix decoder run
 * per position.  At every position: d_pick / d_pick_prob (T,N) = the greedy (a->sampling == GREEDY: argmax, first max)
 * or multinomial (Philox contract of mmt_decode) pick under softmax(logits / a->temperature) and its probability --
 * never fed back; d_target_prob (T,N) = probability of d_target (T,N) (both NULL to skip).  a->stop_on_all_pad is
 * ignored. */
int32_t mmt_teacher_forced_scores(mmt_engine* e, const mmt_decode_args* a, const int64_t* d_trg_in, const int64_t* d_target,
                                  int32_t T, int64_t* d_pick, float* d_pick_prob, float* d_target_prob, void* stream);

/* Replaces vgmmt.beam_search / beam_search_step (validate_generate_MMT_v15_4.py:995-1086), batched: the
 * beam_size beams of each of the a->Bm memory columns are slots of one KV-cached decode wave (a->n_cand, max_len,
 * temperature and sampling are ignored: the reference ranks with softmax(logits) without temperature, :1038).
 * gen_len steps (config.gen_len) from the single [<SOS>] beam; beams ending in `eos` are carried unchanged;
 * score = product of the chosen probabilities in double (Python floats); per step a stable sort by score.
 * Slot n = column * beam_size + rank (rank 0 = best):
 *   d_seq   out (N, gen_len+1) i64  sequence incl. <SOS>, zero-filled past d_len
 *   d_len   out (N) i32             tokens in the sequence (<SOS> included)
 *   d_score out (N) f64
 *   d_probs out (N, gen_len) f32    probability of every chosen token, zero-filled past d_len-1
 *   h_steps out host int or NULL: steps executed (fewer than gen_len once every beam has finished; the skipped
 *           steps are fixed points of the reference's loop).  Synchronises `stream` every 16 steps. */
int32_t mmt_beam_search(mmt_engine* e, const mmt_decode_args* a, int32_t beam_size, int32_t gen_len, int32_t eos,
                        int64_t* d_seq, int32_t* d_len, double* d_score, float* d_probs, int32_t* h_steps, void* stream);

/* Offset increment torch applies to its Philox generator for one
 * exponential_() over `numel` elements (DistributionTemplates.h calc_execution_policy). */
uint64_t mmt_philox_increment(int64_t numel, int32_t sm_count, int32_t max_threads_per_sm);

/* Device-to-device gather of token ids into bytes: (T,N) i64 -> (T,N) u8
 * (vocab 43 < 256); the payload the data-parallel scheduler all-gathers. */
int32_t mmt_pack_tokens_u8(const int64_t* d_tokens, int64_t n, uint8_t* d_out, void* stream);
int32_t mmt_unpack_tokens_u8(const uint8_t* d_in, int64_t n, int64_t* d_tokens, void* stream);

/* The same payload in SEQUENCE-MAJOR order: (T,N) i64 ids -> (N,T) u8, and back.  Ranks own contiguous blocks of
 * sequences, so an all-gather of (N_local,T) blocks lands directly in the final (N_total,T) order -- no re-layout copy
 * after the collective (the time-major form interleaves the ranks' columns).  The inverse reads `N` rows of the
 * gathered (>= N, T) buffer and writes the caller-facing (T,N) i64 tensor (tiled transposes through shared memory). */
int32_t mmt_pack_tokens_u8_seqmajor(const int64_t* d_tokens, int32_t T, int64_t N, uint8_t* d_out, void* stream);
int32_t mmt_unpack_tokens_u8_seqmajor(const uint8_t* d_in, int32_t T, int64_t N, int64_t* d_tokens, void* stream);

/* ---- ragged ingest (the data format in front of the path) -------------------------- */

/* Replaces MultimodalData._zero_pad + the per-modality normalisation of __getitem__
 * (dataloaders_pl_v15_4.py:267-299, 352-365, 456-460, 481-485, 505-509, 546-550): CSR peak lists
 * (d_values: nnz x cols doubles, d_offsets: B+1 int64) -> d_src (B,P[,cols]) f32 = value / div_c rounded like
 * torch.tensor(python floats), zero padded, truncated at P; d_mask (B,P) f32, 0 valid / 1 pad.  cols == 1
 * reproduces the reference's all-ones mask for lists with >= P entries. */
int32_t mmt_ingest_peaks(const double* d_values, const int64_t* d_offsets, int32_t B, int32_t cols,
                         double div0, double div1, int32_t pad_points, float* d_src, float* d_mask, void* stream);
/* Replaces MultimodalData._load_IR_data's binning (dataloaders_pl_v15_4.py:324-346): every spectrum of
 * arbitrary length -> `bins` mean-binned values divided by the spectrum's maximum, (B,bins) f32. */
int32_t mmt_ingest_ir(const double* d_values, const int64_t* d_offsets, int32_t B, int32_t bins,
                      float* d_src_IR, void* stream);

/* Replaces the per-element .item() scans of tensor_to_smiles / tensor_to_smiles_and_prob(_2)
 * (helper_functions_pl_v15_4.py:255-301, 390-419): d_len[n] = position of the first `eos` id in column n
 * of d_tokens (T,N) i64, or T if the sequence never emits it.  The host then slices tokens / probabilities
 * and joins the strings after ONE device-to-host copy. */
int32_t mmt_first_eos(const int64_t* d_tokens, int32_t T, int64_t N, int32_t eos, int32_t* d_len, void* stream);

/* ---- building blocks exported for unit parity tests --------------------------- */

/* fused fc_out + softmax(logits/T) + greedy|multinomial pick on hidden states
 * d_x (N,128): the sampling kernel of the decode step, stand-alone.
 * d_logits_out (N,vocab) optional. */
int32_t mmt_sample(mmt_engine* e, const float* d_x, int64_t N, float temperature, int32_t sampling,
                   uint64_t philox_seed, uint64_t philox_offset, int64_t seq_index_base, int64_t N_total,
                   int32_t rng_sm_count, int32_t rng_max_threads_per_sm,
                   int64_t* d_token, float* d_prob, float* d_logits_out, void* stream);

/* The Exp(1) variates torch's CUDA `exponential_` writes (ATen DistributionTemplates.h:427-441 through
 * torch.multinomial's fast path): d_q[i] = the variate at linear element elem_base + i of a
 * tensor of numel_total elements under generator state (seed, offset) on a device with sm_count SMs of
 * max_threads_per_sm threads.  Lets a test assert BIT equality with torch.empty(...).exponential_(). */
int32_t mmt_exponential(uint64_t philox_seed, uint64_t philox_offset, int64_t elem_base, int64_t n, int64_t numel_total,
                        int32_t sm_count, int32_t max_threads_per_sm, float* d_q, void* stream);

/* torch.multinomial(p, 1) on caller-supplied probabilities d_p (N,V) f32: token = first arg-max of p / q with q as
 * above (rows seq_index_base .. seq_index_base+N of a logical (N_total,V) call).  With torch's own p as input the
 * ids must equal torch.multinomial's exactly -- separates the RNG stream from probability round-off. */
int32_t mmt_sample_probs(const float* d_p, int64_t N, int32_t V, uint64_t philox_seed, uint64_t philox_offset,
                         int64_t seq_index_base, int64_t N_total, int32_t sm_count, int32_t max_threads_per_sm,
                         int64_t* d_token, void* stream);

/* C[M,N] = act(A[M,K] . W[N,K]^T + bias) in the requested precision
 * (fp32 SIMT or bf16 tcgen05); act: 0 none, 1 ReLU. */
int32_t mmt_linear(mmt_engine* e, const float* d_A, const float* d_W, const float* d_bias, float* d_C,
                   int64_t M, int32_t N, int32_t K, int32_t act, int32_t precision, void* stream);

/* The fused feed-forward block of one transformer layer (linear1 -> ReLU -> linear2 -> +residual
 * -> LayerNorm; torch TransformerEncoderLayer._ff_block + norm2, models_MMT_v15_4.py:510-541) in
 * the bf16 tensor-core mode:  d_out = LN(d_x + W2 relu(W1 bf16(d_x) + b1) + b2) * gamma + beta,
 * d_x / d_out (M,128) f32, W1 (F,128), W2 (128,F).  splits > 1 exercises the split-F partial path
 * the small-batch decoder uses.  weight_terms: 2 = W as the two-term bf16 split W_hi + W_lo (encoder layers),
 * 1 = W_hi alone (decoder layers; the kernel variant with the deeper GEMM1 / convert / GEMM2 pipeline). */
int32_t mmt_ffn(mmt_engine* e, const float* d_x, const float* d_w1, const float* d_b1, const float* d_w2,
                const float* d_b2, const float* d_gamma, const float* d_beta, float* d_out,
                int64_t M, int32_t F, int32_t splits, int32_t weight_terms, void* stream);

/* launches issued by this engine since creation (bench.py's gpu_launches). */
int64_t mmt_launch_count(const mmt_engine* e);

/* Per-kernel-class device time for the roofline report: while enabled, every
 * launch is bracketed by CUDA events recorded on the launching stream.
 * mmt_profile_report synchronises the device and writes a JSON object
 * {"kernel": {"launches": n, "ms": total, "flops": algorithmic GEMM flops}, ...}. */
int32_t mmt_profile_enable(mmt_engine* e, int32_t on);
int32_t mmt_profile_report(mmt_engine* e, char* buf, int64_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* MMT_B200_H */
