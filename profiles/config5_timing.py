"""BASELINE.json config 5, one GPU's share: 512 spectra with every peak slot valid (no masked keys: 582 attended memory rows per
spectrum, nothing for the ragged encoder to skip), greedy, 128 tokens.  python profiles/config5_timing.py [B] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
cfg = M.default_config(device="cuda", precision=prec, max_len=128)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000, peaks="max").items()}
ev = lambda: torch.cuda.Event(enable_timing=True)
enc, dec = [], []
for it in range(5):
    a, b, c = ev(), ev(), ev()
    a.record()
    memory, mask, *_ = M.run_model(model, data, cfg)
    b.record()
    tok, pr = M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
    c.record(); torch.cuda.synchronize()
    if it >= 2: enc.append(a.elapsed_time(b)); dec.append(b.elapsed_time(c))
e, d = sum(enc) / len(enc), sum(dec) / len(dec)
print(f"config5 share {prec}: {B} max-peak spectra x {tok.shape[0]} tokens: encode {e:.2f} ms, decode {d:.1f} ms ({1e3*d/tok.shape[0]:.0f} us/position) "
      f"-> {B * tok.shape[0] / ((e + d) * 1e-3):.3e} tokens/s; attended keys per spectrum {int((~mask).sum().item()) // B}")
