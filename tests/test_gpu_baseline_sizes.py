"""Parity at the sizes BASELINE.json names (run on the B200 box with -m gpu, all through the C ABI).

    config 2   greedy, 256 spectra x 128 tokens: fp32 check mode ids == the CPU oracle for all 256 columns;
               bf16 tensor-core mode under north_star's margin rule
    config 3   multinomial, 1024 spectra x 128 candidates (131,072 sequences in waves of 65,536 (bf16) / 32,768 (fp32),
               Philox offset increment 20 per step), 16 positions: offset bookkeeping, shard invariance, and spectra
               of the first, a middle and the last wave against the oracle loop run on the device with the bit-exact
               Exp(1) variates
    config 5   max peak counts (582 attended memory rows per spectrum) + 128 tokens, 64 spectra, fp32 and bf16
    ADVICE r1  bf16 runs whose last wave is short (16,384 + 1,024 sequences; 600 x 16 beam slots)

Near-tie rule (fp32): the engine and the oracle are two fp32 evaluations that differ by ~1e-6 relative in the logits,
so a column may leave the oracle's ids only at a position where the oracle's own top-2 logit margin is below 1e-4 of
the row's max |logit|; what follows in that column is a different (equally valid) continuation and is not compared.
Margin rule (bf16, BASELINE.json north_star): logits within 1e-2 of the fp32 ones relative to the row scale, so ids
must agree wherever the fp32 top-2 margin exceeds 2e-2 of the row scale.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}
_S = {}


def setup():
    if "model" not in _S:
        import multimodalspectraltransformer_b200 as M
        from oracle import mmt_oracle as O
        cfg = M.default_config(device="cuda")
        torch.manual_seed(0)
        model = M.MultimodalTransformer(cfg).eval()
        _S.update(M=M, O=O, model=model, P=O.random_init_state_dict(O.default_config(), seed=0))
    return _S


def cfg_for(**over):
    return setup()["M"].default_config(device="cuda", **over)


def model_with(monkeypatch, **env):
    """A second model with the same seeded weights whose engine is created under the given MMT_* knobs."""
    M = setup()["M"]
    from multimodalspectraltransformer_b200.engine import engine_for
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    torch.manual_seed(0)
    m = M.MultimodalTransformer(cfg_for()).eval()
    engine_for(m, cfg_for())                  # the engine reads its knobs at creation
    for k in env:
        monkeypatch.delenv(k)
    return m


def oracle_greedy(tag, B, seed, peaks):
    """CPU oracle (full-prefix loop, fp32) for B synthetic spectra x 128 steps, with the relative top-2 margins; cached
    across the fp32 and bf16 tests of one config."""
    s = setup()
    if tag not in _S:
        from multimodalspectraltransformer_b200 import synthetic
        O = s["O"]
        data = synthetic.make_spectra(B, seed=seed, peaks=peaks)
        ocfg = O.default_config()
        with torch.no_grad():
            omem, omask, _, _ = O.encode(s["P"], data, ocfg)
            otok, opr, omargin = O.greedy_sequence_with_margins(s["P"], omem, omask, ocfg)
        _S[tag] = (data, omem, otok, opr, omargin)
    return _S[tag]


def first_mismatch(tok, otok):
    """per column: first position where the ids differ, or T"""
    T = tok.shape[0]
    ne = (tok != otok)
    return torch.where(ne.any(dim=0), ne.to(torch.uint8).argmax(dim=0), torch.full((tok.shape[1],), T))


def check_columns(tok, otok, omargin, tol):
    """Every column equals the oracle's up to the first position whose oracle margin is below tol; returns the number of
    columns identical over the whole length."""
    T, N = otok.shape
    assert tuple(tok.shape) == (T, N)
    fm = first_mismatch(tok.cpu(), otok)
    bad = torch.nonzero(fm < T)[:, 0]
    for n in bad.tolist():
        t = int(fm[n])
        assert float(omargin[t, n]) < tol, f"column {n} leaves the oracle at step {t} where the top-2 margin is {float(omargin[t, n]):.3e} >= {tol}"
    return N - bad.numel()


# ------------------------------------------------------------------------------------------------ config 2
def test_config2_fp32_256_spectra_128_tokens_ids_equal_oracle():
    s = setup()
    data, omem, otok, opr, omargin = oracle_greedy("c2", 256, 2002, "realistic")
    cfg = cfg_for(precision="fp32")
    memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
    torch.testing.assert_close(memory.cpu(), omem, atol=5e-5, rtol=0)
    tok, pr = s["M"].greedy_sequence(s["model"], STOI, None, memory, mask, cfg)
    assert tuple(tok.shape) == (128, 256) and tuple(pr.shape) == (127, 256)
    same = check_columns(tok, otok, omargin, 1e-4)
    assert same >= 254, same                       # a genuine fp32 near-tie is rare: all, or all but one or two, columns
    ok = (tok.cpu() == otok).all(dim=0)
    torch.testing.assert_close(pr.cpu()[:, ok], opr[:, ok], atol=2e-5, rtol=0)


def test_config2_bf16_256_spectra_128_tokens_margin_rule():
    s = setup()
    data, omem, otok, opr, omargin = oracle_greedy("c2", 256, 2002, "realistic")
    cfg = cfg_for(precision="bf16")
    memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
    rel = ((memory.cpu() - omem).pow(2).mean().sqrt() / omem.pow(2).mean().sqrt()).item()
    assert rel < 1.5e-2, rel
    tok, pr = s["M"].greedy_sequence(s["model"], STOI, None, memory, mask, cfg)
    same = check_columns(tok, otok, omargin, 2e-2)
    # free-running: a column that met a near-tie continues differently; most never meet one
    assert same >= 128, same
    # and where the ids agree the chosen-token probabilities are the fp32 ones to bf16 accuracy
    ok = (tok.cpu() == otok).all(dim=0)
    assert float((pr.cpu()[:, ok] - opr[:, ok]).abs().max()) < 2e-2


# ------------------------------------------------------------------------------------------------ config 5
def test_config5_max_peaks_128_tokens_fp32_and_bf16():
    s = setup()
    data, omem, otok, opr, omargin = oracle_greedy("c5", 64, 5005, "max")
    for prec, tol, floor in (("fp32", 1e-4, 63), ("bf16", 2e-2, 32)):
        cfg = cfg_for(precision=prec)
        memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
        assert int((~mask).sum()) == 64 * 582                       # every memory row attended: nothing for the ragged paths to skip
        if prec == "fp32":
            torch.testing.assert_close(memory.cpu(), omem, atol=5e-5, rtol=0)
        tok, pr = s["M"].greedy_sequence(s["model"], STOI, None, memory, mask, cfg)
        same = check_columns(tok, otok, omargin, tol)
        assert same >= floor, (prec, same)


# ------------------------------------------------------------------------------------------------ config 3
@pytest.mark.parametrize("precision,floor", [("fp32", 0.95), ("bf16", 0.6)])
def test_config3_eight_waves_multinomial_131072_sequences(precision, floor):
    """1024 spectra x 128 candidates, multinomial, 16 positions: two (bf16) / four (fp32) waves, numel = 131,072 x 43
    > 4 x 1,212,416 -> Philox loop iterations 0..4, offset += 20 per step.  fp32 check mode and the bf16 bench mode."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    from multimodalspectraltransformer_b200.engine import engine_for
    B, K, T = 1024, 128, 16
    data = synthetic.make_spectra(B, seed=3003)
    cfg = cfg_for(precision=precision, max_len=T)
    eng = engine_for(s["model"], cfg)
    memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(4242)
    torch.empty(5, device="cuda").uniform_()
    seed, off0 = gen.initial_seed(), gen.get_offset()
    tok, pr = s["M"].multinomial_sequence_multi(s["model"], memory, mask, STOI, cfg, n_candidates=K)
    assert tuple(tok.shape) == (T, B * K) and tuple(pr.shape) == (T, B * K)
    assert eng.philox_increment(B * K) == 20 and gen.get_offset() == off0 + T * 20
    assert int(tok.min()) >= 0 and int(tok.max()) < 43 and bool(torch.isfinite(pr).all()) and float(pr.min()) > 0
    # shard invariance at full size: spectra [512, 640) decoded alone (one wave) inside the logical 131,072-row call
    gen.set_offset(off0)
    lo, hi = 512, 640
    tok_s, pr_s = s["M"].multinomial_sequence_multi(s["model"], memory[:, lo:hi], mask[lo:hi], STOI, cfg, n_candidates=K,
                                                    seq_index_base=lo * K, n_total=B * K)
    assert torch.equal(tok_s, tok[:, lo * K:hi * K]) and torch.equal(pr_s, pr[:, lo * K:hi * K])
    # the oracle's loop (full prefix, torch fp32 ops on this GPU) for one spectrum of waves 0, 3 and 7 with the draws of
    # torch.multinomial on the full (131072, 43) tensor (Exp(1) variates bit-exact, test_exponential_variates_bit_equal_torch)
    O = s["O"]
    Pd = {k: v.cuda() for k, v in s["P"].items()}
    ocfg = O.default_config(max_len=T)
    mem32, mask32 = memory, mask
    if precision != "fp32":
        mem32, mask32, *_ = s["M"].run_model(s["model"], data, cfg_for(precision="fp32"))
    for b in (5, 3 * 128 + 77, 7 * 128 + 127):
        cols = slice(b * K, (b + 1) * K)
        qs = [eng.exponential(K * 43, seed=seed, offset=off0 + 20 * t, elem_base=b * K * 43, numel_total=B * K * 43).view(K, 43) for t in range(T)]
        step = iter(range(T))
        pick = lambda p: torch.argmax(p / qs[next(step)], dim=1)
        with torch.no_grad():
            otok, opr = O._decode_loop(Pd, mem32[:, b:b + 1].expand(-1, K, -1).contiguous(), mask32[b:b + 1].expand(K, -1).contiguous(), ocfg, pick, False, max_len=T)
        same = (tok[:, cols] == otok).all(dim=0)
        assert bool((tok[0, cols] == otok[0]).all())
        assert same.float().mean().item() >= floor, (b, same.float().mean().item())     # near-ties of p/q flip (more of them under bf16 probabilities)
        assert float((pr[:, cols][:, same] - opr[:, same]).abs().max()) < (2e-5 if precision == "fp32" else 2e-2)


def test_config3_waves_equal_single_wave(monkeypatch):
    """Wave splitting is invisible: 300 spectra x 16 candidates (4,800 sequences) in waves of 1,024 sequences (4 full + 1
    short wave) == the same run in one wave.  Bit for bit where both runs execute the same kernels with the same tiling
    (fp32 mode, un-fused kernels, the four full waves: >= 1,024 rows select the same GEMM tiles as 4,800 rows); elsewhere
    the wave size selects other tile shapes / split-K factors / the fused small-wave kernels, whose fp32 round-off differs
    in the last bits: ids equal except at near-ties, probabilities close."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    data = synthetic.make_spectra(300, seed=3113)
    T, K = 12, 16

    m_one = model_with(monkeypatch, MMT_FUSED_DECODE_ROWS="0")
    m_waves = model_with(monkeypatch, MMT_FUSED_DECODE_ROWS="0", MMT_MAX_WAVE_SEQS="1024")
    m_fused_waves = model_with(monkeypatch, MMT_MAX_WAVE_SEQS="1024")
    for prec in ("fp32", "bf16"):
        cfg = cfg_for(precision=prec, max_len=T)
        memory, mask, *_ = M.run_model(s["model"], data, cfg)
        res = []
        for m in (m_one, m_waves, m_fused_waves):
            torch.manual_seed(77)
            mt, mp_ = M.multinomial_sequence_multi(m, memory, mask, STOI, cfg, n_candidates=K)
            gt, gp = M.greedy_sequence(m, STOI, None, memory, mask, cfg, n_candidates=K)
            res.append((mt, mp_, gt, gp))
        a, b, c = res
        if prec == "fp32":
            assert all(torch.equal(x[:, :4096], y[:, :4096]) for x, y in zip(a, b)), prec
        tol = 2e-5 if prec == "fp32" else 2e-2
        for x_tok, x_pr, y_tok, y_pr in ((a[0], a[1], c[0], c[1]), (a[2], a[3], c[2], c[3]), (a[0], a[1], b[0], b[1]), (a[2], a[3], b[2], b[3])):
            same = (x_tok == y_tok).all(dim=0)
            assert same.float().mean().item() >= (0.98 if prec == "fp32" else 0.7), (prec, same.float().mean().item())
            assert bool((x_tok[0] == y_tok[0]).all()) or prec == "bf16"
            assert float((x_pr[:, same] - y_pr[:, same]).abs().max()) < tol


# ------------------------------------------------------------------------------------------------ ADVICE r1 (high)
def test_bf16_short_last_wave_16384_plus_1024(monkeypatch):
    """136 spectra x 128 candidates in waves of 16,384 = 16,384 + 1,024 sequences in bf16: the short last wave takes the
    fused split-F path whose partial buffer used to be sized for the planned (large) wave only.  Its columns must equal
    the same 8 spectra decoded alone."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    B, K, T = 136, 128, 8
    data = synthetic.make_spectra(B, seed=1361)
    cfg = cfg_for(precision="bf16", max_len=T)
    memory, mask, *_ = M.run_model(s["model"], data, cfg)
    m16k = model_with(monkeypatch, MMT_MAX_WAVE_SEQS="16384")
    torch.manual_seed(5)
    tok, pr = M.multinomial_sequence_multi(m16k, memory, mask, STOI, cfg, n_candidates=K)
    torch.manual_seed(5)
    tok_s, pr_s = M.multinomial_sequence_multi(s["model"], memory[:, 128:], mask[128:], STOI, cfg, n_candidates=K,
                                               seq_index_base=128 * K, n_total=B * K)
    assert torch.equal(tok[:, 128 * K:], tok_s) and torch.equal(pr[:, 128 * K:], pr_s)
    torch.manual_seed(5)
    tok_f, pr_f = M.multinomial_sequence_multi(s["model"], memory[:, :128], mask[:128], STOI, cfg, n_candidates=K,
                                               seq_index_base=0, n_total=B * K)
    assert torch.equal(tok[:, :128 * K], tok_f) and torch.equal(pr[:, :128 * K], pr_f)


def test_bf16_beam_600_items_16_beams_short_last_wave():
    """600 items x 16 beams = 8,192 + 1,408 slots in bf16 (the repo's own beam case): the items of the short
    second wave must come out as when searched alone."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    data = synthetic.make_spectra(600, seed=6001)
    cfg = cfg_for(precision="bf16")
    cfg.gen_len = 10
    memory, mask, *_ = M.run_model(s["model"], data, cfg)
    beams = M.beam_search(s["model"], STOI, memory, mask, cfg, 16)
    alone = M.beam_search(s["model"], STOI, memory[:, 512:], mask[512:], cfg, 16)
    assert len(beams) == 600 and len(alone) == 88
    for i in range(88):
        a, b = beams[512 + i], alone[i]
        assert [x[1] for x in a] == [x[1] for x in b], i
        np.testing.assert_allclose([x[0] for x in a], [x[0] for x in b], rtol=1e-6)


def test_chained_projection_equals_two_launches(monkeypatch):
    """The un-fused decode step issues "out-proj + LN1" and the cross-attention query projection as ONE tcgen05 launch (the
    normalised rows feed a second MMA from shared memory), and for >= 2048 rows the cross-attention out-projection + norm2
    as a prologue of the fused FFN kernel: bit for bit the separate-launch result, for a ragged row count (partial last
    tile, split-F FFN path), a multi-tile wave and a wave whose last tile is partial (19 x 128 = 2432 rows)."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    # (head-major pages on both sides: the token-major self-attention kernel partitions the keys of a head over four lanes
    # instead of eight, another summation order -- covered by test_token_major_pages_equal_head_major_pages)
    m_chain = model_with(monkeypatch, MMT_FUSED_DECODE_ROWS="0", MMT_KV_HEAD_MAJOR="1")
    m_plain = model_with(monkeypatch, MMT_FUSED_DECODE_ROWS="0", MMT_NO_GEMM_CHAIN="1", MMT_NO_FFN_PROLOGUE="1", MMT_NO_KV_EPILOGUE="1")
    for B, K, T in ((41, 7, 10), (300, 16, 6), (19, 128, 5)):
        data = synthetic.make_spectra(B, seed=700 + B)
        cfg = cfg_for(precision="bf16", max_len=T)
        memory, mask, *_ = M.run_model(s["model"], data, cfg)
        out = []
        for m in (m_chain, m_plain):
            torch.manual_seed(11)
            out.append(M.multinomial_sequence_multi(m, memory, mask, STOI, cfg, n_candidates=K))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]), (B, K)


def test_token_major_pages_equal_head_major_pages(monkeypatch):
    """Large bf16 waves keep the self-attention cache in TOKEN-major pages ([16][K|V][H][8]: the QKV projection's epilogue
    appends a token as contiguous rows, `decode_self_attention_tm` reads only the valid prefix of the open page); the
    head-major pages + `decode_self_attention_g8` remain behind MMT_KV_HEAD_MAJOR=1.  Same values in the cache, another
    partition of a head's keys over lanes (fp32 round-off of the attention row before its bf16 rounding): positions 0..3
    have at most one key per lane in both kernels -> the first sampled tokens are identical; afterwards ids equal except
    at near-ties, probabilities to bf16 accuracy.  T = 40 crosses two page boundaries (closed pages + an open page); the
    last case is a wave with a partial last tile."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    m_tm = model_with(monkeypatch, MMT_FUSED_DECODE_ROWS="0")
    m_hm = model_with(monkeypatch, MMT_FUSED_DECODE_ROWS="0", MMT_KV_HEAD_MAJOR="1")
    for B, K, T in ((24, 16, 40), (19, 128, 20), (5, 3, 33)):
        data = synthetic.make_spectra(B, seed=900 + B)
        cfg = cfg_for(precision="bf16", max_len=T)
        memory, mask, *_ = M.run_model(s["model"], data, cfg)
        out = []
        for m in (m_tm, m_hm):
            torch.manual_seed(13)
            mt, mp_ = M.multinomial_sequence_multi(m, memory, mask, STOI, cfg, n_candidates=K)
            gt, gp = M.greedy_sequence(m, STOI, None, memory, mask, cfg, n_candidates=K)
            out.append((mt, mp_, gt, gp))
        a, b = out
        for x_tok, x_pr, y_tok, y_pr in ((a[0], a[1], b[0], b[1]), (a[2], a[3], b[2], b[3])):
            assert torch.equal(x_tok[:2], y_tok[:2]), (B, K)
            same = (x_tok == y_tok).all(dim=0)
            assert same.float().mean().item() >= 0.7, (B, K, same.float().mean().item())
            assert float((x_pr[:, same] - y_pr[:, same]).abs().max()) < 2e-2



def test_cluster_ffn_decode_equals_separate_ffn_kernel(monkeypatch):
    """Opt-in (`MMT_CLUSTER_FFN=1`): small bf16 waves run the decoder FFN + norm3 INSIDE `decode_attn`, launched as clusters of
    four CTAs (the 8 rows of a cluster are gathered through distributed shared memory, every CTA owns 512 hidden columns on
    mma.sync, partial outputs scattered to the row owners) instead of the separate tcgen05 FFN kernel + split-F partials.  Same
    operands (bf16 x and h, hi weight term), another accumulation order: teacher-forced logits agree to 2e-3 of the row
    scale (the two differ from the fp32 reference by ~5e-3), free-running ids agree except at near-ties.  Row counts: one
    partly filled cluster (5 rows), several clusters with a dead tail CTA (37 x 1), candidates sharing a memory (19 x 7 =
    133 rows), and 256 rows in two concurrent lanes (the bench geometry)."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    m_cl = model_with(monkeypatch, MMT_CLUSTER_FFN="1")
    m_sep = model_with(monkeypatch)
    for B, K, T in ((5, 1, 12), (37, 1, 20), (19, 7, 12), (256, 1, 24)):
        data = synthetic.make_spectra(B, seed=1500 + B)
        cfg = cfg_for(precision="bf16", max_len=T)
        memory, mask, *_ = M.run_model(s["model"], data, cfg)
        g = torch.Generator().manual_seed(B)
        trg = torch.randint(4, 43, (T, B * K), generator=g)
        trg[0] = 3
        la = M.teacher_forced_logits(m_cl, memory, mask, trg.cuda(), cfg, n_candidates=K)
        lb = M.teacher_forced_logits(m_sep, memory, mask, trg.cuda(), cfg, n_candidates=K)
        scale = lb.abs().amax(dim=-1, keepdim=True)
        assert float(((la - lb).abs() / scale).max()) < 2e-3, (B, K, float(((la - lb).abs() / scale).max()))
        out = []
        for m in (m_cl, m_sep):
            torch.manual_seed(23)
            mt, mp_ = M.multinomial_sequence_multi(m, memory, mask, STOI, cfg, n_candidates=K)
            gt, gp = M.greedy_sequence(m, STOI, None, memory, mask, cfg, n_candidates=K)
            out.append((mt, mp_, gt, gp))
        a, b = out
        for x_tok, x_pr, y_tok, y_pr in ((a[0], a[1], b[0], b[1]), (a[2], a[3], b[2], b[3])):
            same = (x_tok == y_tok).all(dim=0)
            assert same.float().mean().item() >= 0.7, (B, K, same.float().mean().item())
            assert float((x_pr[:, same] - y_pr[:, same]).abs().max()) < 2e-2


def test_large_wave_split_is_bit_invisible_in_bf16(monkeypatch):
    """96 spectra x 128 candidates (12,288 sequences) as one wave against three waves of 4,096: the un-fused bf16 step tiles
    every GEMM by 128 rows whatever the wave size, the token-major self-attention and the candidate cross attention are per
    sequence / per spectrum, and the sampler's four-rows-per-warp form (waves >= 8,192 rows) computes every logit in the same
    order as the one-row form (smaller waves) -- tokens and probabilities are bit-identical."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    M = s["M"]
    m_one = model_with(monkeypatch)
    m_waves = model_with(monkeypatch, MMT_MAX_WAVE_SEQS="4096")
    data = synthetic.make_spectra(96, seed=4711)
    cfg = cfg_for(precision="bf16", max_len=20)
    memory, mask, *_ = M.run_model(s["model"], data, cfg)
    out = []
    for m in (m_one, m_waves):
        torch.manual_seed(29)
        mt, mp_ = M.multinomial_sequence_multi(m, memory, mask, STOI, cfg, n_candidates=128)
        gt, gp = M.greedy_sequence(m, STOI, None, memory, mask, cfg, n_candidates=128)
        out.append((mt, mp_, gt, gp))
    for x, y in zip(*out):
        assert torch.equal(x, y)
