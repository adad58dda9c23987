"""GPU parity tests of the bf16 tensor-core mode (tcgen05 GEMMs, fp32 accumulate / residual /
LayerNorm / softmax), through the C ABI.

Tolerance (BASELINE.json north_star): logits within 1e-2 relative error of the fp32 reference,
relative to the row's logit scale max|logit| (element-wise relative error is unbounded at
near-zero logits, SURVEY.md 7 "bf16 tolerance"), and identical greedy ids wherever the
reference's top-2 logit margin exceeds twice that tolerance (each of the two logits may move
by one tolerance)."""
import numpy as np
import pytest
import torch

from golden_util import CASES, load_case
from test_gpu_parity import STOI, cfg_for, setup

pytestmark = pytest.mark.gpu
REL_TOL = 1e-2


@pytest.mark.parametrize("M,N,K,act", [(5, 128, 128, 0), (128, 128, 64, 0), (256, 384, 128, 0), (300, 2048, 128, 1),
                                        (33, 128, 2048, 0), (1000, 256, 128, 0), (4097, 384, 128, 0), (2500, 128, 2048, 1),
                                        (131, 512, 192, 0)])
def test_linear_bf16_tcgen05(M, N, K, act):
    """tcgen05 GEMM == fp64 product of the bf16-rounded activations with the two-term bf16 weights
    W_hi + W_lo (only the accumulation order differs)."""
    s = setup()
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    out = s["eng"].linear(A, W, b, act=act, precision="bf16")
    W_hi = W.bfloat16()
    W_lo = (W - W_hi.float()).bfloat16()
    ref = torch.nn.functional.linear(A.bfloat16().double(), W_hi.double() + W_lo.double(), b.double())
    if act:
        ref = torch.relu(ref)
    torch.testing.assert_close(out.double(), ref, atol=2e-4, rtol=1e-4)


@pytest.mark.parametrize("name", CASES)
def test_encoder_bf16_vs_reference_golden(name):
    s = setup()
    case, data, z = load_case(name)
    cfg = cfg_for(case, precision="bf16")
    memory, mask, trg, fp, hs, co = s["M"].run_model(s["model"], data, cfg)
    assert torch.equal(mask.cpu(), torch.from_numpy(z["mask"]))
    stride = int(z["memory_stride"])
    got, ref = memory[::stride].cpu().numpy(), z["memory_sample"]
    assert np.isfinite(got).all()
    rms = float(np.sqrt(np.mean(ref ** 2)))
    err = np.abs(got - ref)
    assert float(np.sqrt(np.mean(err ** 2))) < 1.5e-2 * rms, (float(np.sqrt(np.mean(err ** 2))), rms)
    assert float(err.max()) < 0.12 * rms * 4, float(err.max())
    fp_ref = z["fingerprint"]
    assert np.abs(fp.cpu().numpy() - fp_ref).max() < 2e-2 * max(1.0, float(np.abs(fp_ref).max()))


@pytest.mark.parametrize("name", CASES)
def test_teacher_forced_bf16_logits_and_greedy_ids(name):
    s = setup()
    case, data, z = load_case(name)
    cfg = cfg_for(case, precision="bf16")
    memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
    logits = s["M"].teacher_forced_logits(s["model"], memory, mask, torch.from_numpy(z["tf_tokens"]), cfg).cpu().numpy()
    ref = z["tf_logits"]
    scale = np.abs(ref).max(axis=-1, keepdims=True)
    rel = np.abs(logits - ref) / scale
    assert float(rel.max()) < REL_TOL, float(rel.max())
    top2 = np.sort(ref, axis=-1)[..., -2:]
    margin = top2[..., 1] - top2[..., 0]
    decided = margin > 2 * REL_TOL * scale[..., 0]
    assert decided.mean() > 0.5
    assert np.array_equal(logits.argmax(-1)[decided], ref.argmax(-1)[decided])


def test_greedy_bf16_decode_runs_and_tracks_fp32():
    """Free-running greedy, 16 spectra x 128 steps: shapes / dtypes of the reference API and ids equal to
    the fp32 engine's up to each sequence's first low-margin position (most sequences entirely)."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    data = synthetic.make_spectra(16, seed=202)
    c32, c16 = cfg_for(), cfg_for(precision="bf16")
    m32, k32, *_ = s["M"].run_model(s["model"], data, c32)
    t32, p32 = s["M"].greedy_sequence(s["model"], STOI, None, m32, k32, c32)
    m16, k16, *_ = s["M"].run_model(s["model"], data, c16)
    t16, p16 = s["M"].greedy_sequence(s["model"], STOI, None, m16, k16, c16)
    assert tuple(t16.shape) == (128, 16) and t16.dtype == torch.int64 and tuple(p16.shape) == (127, 16)
    assert torch.isfinite(p16).all() and int(t16.min()) >= 0 and int(t16.max()) < 43
    same = (t16 == t32)
    first_diff = torch.where(same.all(0), torch.full((16,), 128, device=same.device), (~same).float().argmax(0))
    assert float((first_diff >= 16).float().mean()) >= 0.75, first_diff.tolist()
    assert (t16[0] == t32[0]).all()


def test_multinomial_bf16_candidates():
    """bf16 multinomial decode with shared cross-attention K/V: runs, reference layouts, generator offset advanced."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    data = synthetic.make_spectra(4, seed=9)
    cfg = cfg_for(precision="bf16", max_len=24)
    memory, mask, *_ = s["M"].run_model(s["model"], data, cfg)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(11)
    off0 = gen.get_offset()
    tok, pr = s["M"].multinomial_sequence_multi(s["model"], memory, mask, STOI, cfg, n_candidates=32)
    assert tuple(tok.shape) == (24, 128) and tuple(pr.shape) == (24, 128)
    assert gen.get_offset() - off0 == 24 * s["eng"].philox_increment(128)
    assert torch.isfinite(pr).all() and float(pr.min()) > 0 and float(pr.max()) <= 1
    torch.manual_seed(11)
    tok2, _ = s["M"].multinomial_sequence_multi(s["model"], memory, mask, STOI, cfg, n_candidates=32)
    assert torch.equal(tok, tok2)        # deterministic for a fixed seed


@pytest.mark.parametrize("terms", [2, 1])
@pytest.mark.parametrize("M,F,splits", [(256, 2048, 1), (256, 2048, 16), (256, 2048, 32), (77, 2048, 8), (1000, 2048, 1),
                                         (4097, 2048, 1), (130, 512, 2), (128, 64, 1), (300, 128, 1), (300, 192, 1), (300, 320, 1),
                                         (40001, 2048, 1)])     # >= 2 tiles per SM: two row tiles per CTA, odd tile count
def test_ffn_fused_tcgen05(M, F, splits, terms):
    """Fused FFN kernel == fp64 evaluation of the same contract: bf16(x) . (W1_hi + W1_lo), bias, ReLU, hidden
    rounded to bf16, . (W2_hi + W2_lo), + b2 + x (fp32 residual), LayerNorm.  terms = 1: the hi weight term only (the
    decoder's kernel variant: triple-buffered acc1 / H, GEMM1 two chunks ahead, four-stage weight rings; F = 64 ... 320
    covers chunk counts below, at and above the pipeline depth)."""
    s = setup()
    g = torch.Generator().manual_seed(M * 3 + F + splits)
    x = torch.randn(M, 128, generator=g).cuda()
    w1 = (torch.randn(F, 128, generator=g) / 128 ** 0.5).cuda()
    b1 = (0.1 * torch.randn(F, generator=g)).cuda()
    w2 = (torch.randn(128, F, generator=g) / F ** 0.5).cuda()
    b2 = (0.1 * torch.randn(128, generator=g)).cuda()
    gamma = (1 + 0.1 * torch.randn(128, generator=g)).cuda()
    beta = (0.1 * torch.randn(128, generator=g)).cuda()
    out = s["eng"].ffn(x, w1, b1, w2, b2, gamma, beta, splits=splits, weight_terms=terms)

    def two_term(w):
        hi = w.bfloat16()
        return hi.double() + ((w - hi.float()).bfloat16().double() if terms == 2 else 0.0)
    h = torch.relu(x.bfloat16().double() @ two_term(w1).T + b1.double())
    h = h.float().bfloat16().double()
    y = x.double() + h @ two_term(w2).T + b2.double()
    ref = torch.nn.functional.layer_norm(y, (128,), gamma.double(), beta.double(), 1e-5)
    # the hidden activation is rounded to bf16 from an fp32 accumulator whose last bits depend on the
    # accumulation order: a handful of elements may round the other way (1 bf16 ulp of h ~ 4e-3 * |h|)
    torch.testing.assert_close(out.double(), ref, atol=3e-3, rtol=1e-3)
    assert float((out.double() - ref).abs().mean()) < 2e-4


@pytest.mark.parametrize("tc5", [False, True])
@pytest.mark.parametrize("name,dense", [("full_b5", False), ("full_b5", True), ("maxpeaks_b2", False), ("mode_hsqc_b2", True),
                                        ("blank_hsqc_only_b3", True), ("mode_ms_max_b2", True)])
def test_cross_encoder_tensor_core_attention_equals_simt(name, dense, tc5, monkeypatch):
    """attn_encoder_tc / attn_encoder_tc8 (mma.sync, two-term operand splits, exp2; 32-wide heads of encoder_cross and
    8-wide heads of the modality encoders) reproduce the fp32 SIMT attention kernel to fp32 round-off.  Checked in the fp32 mode (test hook MMT_TC_ATTENTION_FP32), where no bf16 rounding downstream
    amplifies last-bit differences: ragged and dense key lists, bool and float key masks.  In the bf16 mode the two
    variants must agree within that mode's own rounding noise."""
    from multimodalspectraltransformer_b200.engine import Engine
    s = setup()
    case, data, z = load_case(name)
    cfg = cfg_for(case)
    dev = torch.device("cuda", torch.cuda.current_device())
    if dense:
        monkeypatch.setenv("MMT_DENSE_ENCODER", "1")
    monkeypatch.setenv("MMT_TC_ATTENTION_FP32", "1")
    if tc5:       # the tcgen05 / TMEM kernel for the 32-wide heads (kernels_attn5.cuh) instead of the mma.sync one
        monkeypatch.setenv("MMT_TC5_ATTENTION", "1")
    eng_tc = Engine(s["model"].state_dict(), cfg, dev)
    monkeypatch.delenv("MMT_TC_ATTENTION_FP32")
    if tc5:
        monkeypatch.delenv("MMT_TC5_ATTENTION")
    monkeypatch.setenv("MMT_NO_TC_ATTENTION", "1")
    eng_simt = Engine(s["model"].state_dict(), cfg, dev)
    a = eng_tc.encode(data, case["mode"], "fp32", False)
    b = eng_simt.encode(data, case["mode"], "fp32", False)
    assert torch.equal(a[1], b[1])
    scale = b[0].abs().max().item()
    # 12 layers of two-term-split products (lo.lo dropped at 2^-18) and ex2.approx: a few 1e-5 of the activation scale
    assert (a[0] - b[0]).abs().max().item() <= 6e-5 * scale
    assert (a[0] - b[0]).abs().mean().item() <= 3e-6 * scale
    torch.testing.assert_close(a[3], b[3], atol=5e-5, rtol=0)
    a16 = eng_tc.encode(data, case["mode"], "bf16", False)
    b16 = eng_simt.encode(data, case["mode"], "bf16", False)
    assert (a16[0] - b16[0]).abs().max().item() <= 2e-2 * scale
    assert (a16[0] - b16[0]).abs().mean().item() <= 1e-3 * scale


@pytest.mark.parametrize("n_cand,name", [(16, "full_b5"), (40, "mode_hsqc_b2"), (9, "maxpeaks_b2")])
def test_candidate_wave_cross_attention_on_tensor_cores(n_cand, name, monkeypatch):
    """decode_cross_attention_tc (candidates of a spectrum share K/V: mma.sync tiles of 16 candidates) against the SIMT
    kernel it replaces on the un-fused decode path, bf16 mode: teacher-forced logits of random targets agree within the
    mode's rounding noise and both stay within the tolerance of the fp32 mode; ragged candidate counts, bool and float masks."""
    from multimodalspectraltransformer_b200.engine import Engine
    from multimodalspectraltransformer_b200.generate import _mask_to_bias
    s = setup()
    case, data, z = load_case(name)
    cfg = cfg_for(case, precision="bf16")
    dev = torch.device("cuda", torch.cuda.current_device())
    monkeypatch.setenv("MMT_FUSED_DECODE_ROWS", "0")
    eng_tc = Engine(s["model"].state_dict(), cfg, dev)
    monkeypatch.setenv("MMT_NO_TC_ATTENTION", "1")
    eng_simt = Engine(s["model"].state_dict(), cfg, dev)
    memory, pad, key_bias, *_ = eng_simt.encode(data, case["mode"], "bf16", False)
    B = memory.shape[1]
    g = torch.Generator().manual_seed(n_cand)
    trg = torch.randint(0, 43, (6, B * n_cand), generator=g)
    trg[0] = 3
    a = eng_tc.teacher_forced(memory, key_bias, trg, n_cand=n_cand, precision="bf16")
    b = eng_simt.teacher_forced(memory, key_bias, trg, n_cand=n_cand, precision="bf16")
    ref = eng_simt.teacher_forced(memory, key_bias, trg, n_cand=n_cand, precision="fp32")
    scale = ref.abs().amax(dim=-1, keepdim=True)
    assert ((a - b).abs() / scale).max().item() <= 4e-3
    assert ((a - b).abs() / scale).mean().item() <= 4e-4
    assert ((a - ref).abs() / scale).max().item() <= REL_TOL and ((b - ref).abs() / scale).max().item() <= REL_TOL
