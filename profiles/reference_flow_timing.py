"""The reference's production loop for one spectrum (run_batch_gen_val_MMT_v15_4.py:93-158 + helper_functions_pl_v15_4.py:272):
duplicate_dict(1 spectrum, 128) -> run_model -> multinomial_sequence_multi -> tensor_to_smiles_and_prob, with the unmodified
call sequence, and the same through the n_candidates extension.  python profiles/reference_flow_timing.py [precision]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}
ITOS = {**{str(i): "C" for i in range(43)}, "0": "<PAD>", "1": "<UNK>", "2": "<EOS>", "3": "<SOS>", "4": "<MASK>"}
cfg = M.default_config(device="cuda", precision=prec)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
spectra = [synthetic.make_spectra(1, seed=100 + i) for i in range(12)]

def flow_reference_style(d):
    dd = M.duplicate_dict(d, 128)
    memory, mask, *_ = M.run_model(model, dd, cfg)
    tok, pr = M.multinomial_sequence_multi(model, memory, mask, STOI, cfg)
    return M.tensor_to_smiles_and_prob(tok.squeeze(0), pr, ITOS)

def flow_candidates(d):
    memory, mask, *_ = M.run_model(model, d, cfg)
    tok, pr = M.multinomial_sequence_multi(model, memory, mask, STOI, cfg, n_candidates=128)
    return M.tensor_to_smiles_and_prob(tok.squeeze(0), pr, ITOS)

for name, fn in (("duplicate_dict(…,128) flow", flow_reference_style), ("n_candidates=128 flow", flow_candidates)):
    for d in spectra[:3]:
        fn(d)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for d in spectra[3:]:
        out = fn(d)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / len(spectra[3:])
    print(f"{prec} {name}: {dt*1e3:.1f} ms per spectrum (128 candidates x 128 tokens, ids -> SMILES included), {128*128/dt/1e6:.2f} M tok/s")
