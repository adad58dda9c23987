"""Small end-to-end pass over every kernel family (both precisions, fused and un-fused decode, ragged and dense encoder);
also usable under a memory checker where one is available:
compute-sanitizer --tool memcheck python profiles/all_paths_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic, ingest
STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}
torch.manual_seed(0)
for prec in ("bf16", "fp32"):
    cfg = M.default_config(device="cuda", precision=prec)
    model = M.MultimodalTransformer(cfg).eval()
    cfg.max_len = 12
    data = synthetic.make_spectra(3, seed=7)
    memory, mask, trg, fp, *_ = M.run_model(model, data, cfg)
    tok, pr = M.greedy_sequence(model, STOI, None, memory, mask, cfg)
    tok2, pr2 = M.multinomial_sequence_multi(model, memory, mask, STOI, cfg, n_candidates=9)
    cfg.gen_len = 6
    beams = M.beam_search(model, STOI, memory, mask, cfg, 3)
    out = M.predict_prop_correct_max_sequence_2(model, STOI, memory, mask, trg, cfg)
    keys = ("src_1H", "mask_1H", "src_13C", "mask_13C", "src_HSQC", "mask_HSQC", "src_COSY", "mask_COSY", "src_IR", "mask_IR",
            "src_MF", "mask_MF", "src_MS", "mask_MS", "trg_MW")
    model.config = cfg
    mem2 = model(*[data[k] for k in keys])[0]
    smi = M.tensor_to_smiles(tok, {**{str(i): "C" for i in range(43)}, "0": "<PAD>", "1": "<UNK>", "2": "<EOS>", "3": "<SOS>", "4": "<MASK>"})
    torch.cuda.synchronize()
    print(prec, "fused ok", tuple(tok.shape), tuple(tok2.shape), len(beams), tuple(out[1].shape), torch.equal(mem2, memory))
os.environ["MMT_FUSED_DECODE_ROWS"] = "0"
os.environ["MMT_DENSE_ENCODER"] = "1"
for prec in ("bf16", "fp32"):
    cfg = M.default_config(device="cuda", precision=prec)
    cfg.training_mode = "HSQC_MF_MW"
    model = M.MultimodalTransformer(cfg).eval()
    cfg.max_len = 20
    data = synthetic.make_spectra(2, seed=8)
    memory, mask, *_ = M.run_model(model, data, cfg)
    tok, pr = M.greedy_sequence(model, STOI, None, memory, mask, cfg)
    tok2, pr2 = M.multinomial_sequence_multi(model, memory, mask, STOI, cfg, n_candidates=17)
    cfg.gen_len = 18
    beams = M.beam_search(model, STOI, memory, mask, cfg, 9)
    torch.cuda.synchronize()
    print(prec, "unfused/dense ok", tuple(tok.shape), tuple(tok2.shape), len(beams))
