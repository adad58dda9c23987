"""GPU parity of the batched beam search (mmt_beam_search) against the reference's own output
(tests/golden/beam_b2.json, oracle/make_golden_beam.py) and the oracle restatement."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN

pytestmark = pytest.mark.gpu

STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}
MODE = "1H_13C_HSQC_COSY_IR_MF_MW"
_S = {}


def setup():
    if "model" not in _S:
        import multimodalspectraltransformer_b200 as M
        from oracle import mmt_oracle as O
        cfg = M.default_config(device="cuda")
        cfg.training_mode = MODE
        torch.manual_seed(0)
        model = M.MultimodalTransformer(cfg)
        model.eval()
        _S.update(M=M, O=O, cfg=cfg, model=model, P=O.random_init_state_dict(O.default_config(), seed=0))
    return _S


def run(model, data, beam, gen_len, precision="fp32", **cfg_over):
    s = setup()
    cfg = s["M"].default_config(device="cuda", **cfg_over)
    cfg.training_mode = MODE
    cfg.gen_len = gen_len
    cfg.precision = precision
    memory, mask, *_ = s["M"].run_model(model, data, cfg)
    return s["M"].beam_search(model, STOI, memory, mask, cfg, beam), memory, mask


def compare(got, want, score_rtol=5e-4, prob_atol=5e-6):
    assert len(got) == len(want)
    for g_item, w_item in zip(got, want):
        assert [g[1] for g in g_item] == [w[1] for w in w_item]
        np.testing.assert_allclose([g[0] for g in g_item], [w[0] for w in w_item], rtol=score_rtol)
        for g, w in zip(g_item, w_item):
            assert len(g[2]) == len(g[1]) - 1
            np.testing.assert_allclose(g[2], w[2], atol=prob_atol, rtol=0)


def test_beam_search_matches_reference_golden():
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    gold = json.load(open(os.path.join(GOLDEN, "beam_b2.json")))
    for case in gold["cases"]:
        data = synthetic.make_spectra(case["B"], seed=case["seed"])
        got, _, _ = run(s["model"], data, case["beam"], case["gen_len"])
        want = [[(w["score"], w["seq"], w["probs"]) for w in item] for item in case["beams"]]
        compare(got, want)


def eos_heavy():
    """Weights whose <EOS> logit is raised so that beams finish early: exercises the carried-over finished beams,
    mixed sequence lengths and the all-finished early exit."""
    s = setup()
    if "model_eos" not in _S:
        m = copy.deepcopy(s["model"])
        with torch.no_grad():
            m.fc_out.bias[2] += 1.5
        P = {k: v.clone() for k, v in s["P"].items()}
        P["fc_out.bias"][2] += 1.5
        _S.update(model_eos=m, P_eos=P)
    return _S["model_eos"], _S["P_eos"]


@pytest.mark.parametrize("B,beam,gen_len,seed", [(3, 4, 20, 5), (2, 1, 9, 6), (1, 7, 33, 7)])
def test_beam_search_matches_oracle_with_finished_beams(B, beam, gen_len, seed):
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    model, P = eos_heavy()
    data = synthetic.make_spectra(B, seed=seed)
    got, memory, mask = run(model, data, beam, gen_len)
    ocfg = s["O"].default_config(training_mode=MODE)
    with torch.no_grad():
        want = s["O"].beam_search(P, memory.cpu(), mask.cpu(), ocfg, beam, gen_len)
    compare(got, want)
    lens = {len(b[1]) for item in got for b in item}
    if beam > 1 and gen_len >= 20:
        assert len(lens) > 1, "expected beams of different lengths (some finished)"
    for item in got:
        for sc, seq, pr in item:
            assert seq[0] == 3 and (2 not in seq[:-1])
            assert sc == pytest.approx(float(np.prod(np.array(pr, dtype=np.float64))), rel=1e-12)


def test_beam_search_invariant_to_graphs_batching_and_kernel_family(monkeypatch):
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    from multimodalspectraltransformer_b200.engine import Engine
    model, _ = eos_heavy()
    data = synthetic.make_spectra(4, seed=11)
    base, memory, mask = run(model, data, 5, 48)
    cfg = s["M"].default_config(device="cuda")
    cfg.training_mode = MODE
    cfg.gen_len = 48
    # every item alone (the reference's own batching) gives the item's beams of the batched run
    for i in range(4):
        one = s["M"].beam_search(model, STOI, memory[:, i:i + 1], mask[i:i + 1], cfg, 5)
        assert one[0] == base[i]
    for env in ({"MMT_NO_GRAPH": "1"}, {"MMT_NO_PDL": "1"}, {"MMT_GRAPH_STEPS": "3"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = Engine(model.state_dict(), cfg, torch.device("cuda", torch.cuda.current_device()))
        from multimodalspectraltransformer_b200.generate import _mask_to_bias
        seq, ln, score, probs, _ = eng.beam_search(memory, _mask_to_bias(mask), beam_size=5, gen_len=48, eos=2)
        got = [[(score[i, k].item(), seq[i, k, :ln[i, k]].tolist(), probs[i, k, :ln[i, k] - 1].tolist()) for k in range(5)] for i in range(4)]
        assert got == base, env
        for k in env:
            monkeypatch.delenv(k)
    # the un-fused kernel family (large waves) ranks the same sequences
    monkeypatch.setenv("MMT_FUSED_DECODE_ROWS", "0")
    eng = Engine(model.state_dict(), cfg, torch.device("cuda", torch.cuda.current_device()))
    seq, ln, score, probs, _ = eng.beam_search(memory, _mask_to_bias(mask), beam_size=5, gen_len=48, eos=2)
    got = [[(score[i, k].item(), seq[i, k, :ln[i, k]].tolist(), probs[i, k, :ln[i, k] - 1].tolist()) for k in range(5)] for i in range(4)]
    compare(got, base)


def test_beam_search_early_exit_and_bf16():
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    from multimodalspectraltransformer_b200.engine import engine_for
    from multimodalspectraltransformer_b200.generate import _mask_to_bias
    model, _ = eos_heavy()
    with torch.no_grad():
        m2 = copy.deepcopy(model)
        m2.fc_out.bias[2] += 6.0            # <EOS> wins everywhere: every beam finishes within a few steps
    data = synthetic.make_spectra(2, seed=3)
    cfg = s["M"].default_config(device="cuda")
    cfg.training_mode = MODE
    memory, mask, *_ = s["M"].run_model(m2, data, cfg)
    eng = engine_for(m2, cfg)
    seq, ln, score, probs, steps = eng.beam_search(memory, _mask_to_bias(mask), beam_size=3, gen_len=64, eos=2)
    assert steps < 64 and int(ln.max()) <= steps + 1
    assert all(seq[i, k, ln[i, k] - 1].item() == 2 for i in range(2) for k in range(3))
    assert torch.all(score[:, :-1] >= score[:, 1:])
    # bf16 mode: same machinery on the tensor-core decoder; scores within the bf16 logit tolerance of fp32
    got32, _, _ = run(model, data, 4, 12)
    got16, _, _ = run(model, data, 4, 12, precision="bf16")
    for a_item, b_item in zip(got32, got16):
        assert a_item[0][0] == pytest.approx(b_item[0][0], rel=0.1)
        assert all(len(b[1]) == len(b[2]) + 1 for b in b_item)


def test_beam_search_argument_errors():
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    data = synthetic.make_spectra(1, seed=1)
    with pytest.raises(RuntimeError, match="beam_size"):
        run(s["model"], data, 44, 4)
    with pytest.raises(RuntimeError, match="gen_len"):
        run(s["model"], data, 2, 129)
    cfg = s["M"].default_config(device="cuda")
    cfg.gen_len = 0
    assert s["M"].beam_search(s["model"], STOI, torch.zeros(582, 2, 128, device="cuda"), None, cfg, 3) == [[(1, [3], [])]] * 2


def test_beam_search_waves_and_large_slot_counts():
    """600 items x 16 beams = 9600 slots: two waves (8192 slots on the un-fused kernel family, 1408 on the fused one);
    every probed item equals its own single-item search (the reference's batching)."""
    s = setup()
    from multimodalspectraltransformer_b200 import synthetic
    model, _ = eos_heavy()
    data = synthetic.make_spectra(600, seed=17)
    got, memory, mask = run(model, data, 16, 5)
    cfg = s["M"].default_config(device="cuda")
    cfg.training_mode = MODE
    cfg.gen_len = 5
    assert len(got) == 600 and all(len(item) == 16 for item in got)
    for i in (0, 255, 511, 512, 599):
        one = s["M"].beam_search(model, STOI, memory[:, i:i + 1], mask[i:i + 1], cfg, 16)
        compare([got[i]], one)
