"""BASELINE.json config 3 on one GPU: multinomial sampling, n_cand candidates per spectrum sharing one encode
(cross-attention K/V shared), in waves of 16,384 sequences.  Usage: config3_timing.py [spectra] [n_cand] [max_len] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
T = int(sys.argv[3]) if len(sys.argv) > 3 else 128
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
cfg = M.default_config(device="cuda", precision=prec, max_len=T)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000).items()}
ev = lambda: torch.cuda.Event(enable_timing=True)
enc, dec = [], []
for it in range(4):
    torch.manual_seed(5)
    a, b, c = ev(), ev(), ev()
    a.record()
    memory, mask, *_ = M.run_model(model, data, cfg)
    b.record()
    tok, pr = M.multinomial_sequence_multi(model, memory, mask, {"<SOS>": 3}, cfg, n_candidates=K)
    c.record()
    torch.cuda.synchronize()
    if it >= 1:
        enc.append(a.elapsed_time(b)); dec.append(b.elapsed_time(c))
e, d = sum(enc) / len(enc), sum(dec) / len(dec)
print(f"config3 {prec}: {B} spectra x {K} candidates x {T} tokens: encode {e:.2f} ms, decode {d:.1f} ms "
      f"({1e3 * d / T:.0f} us/step) -> {B * K * T / ((e + d) * 1e-3):.3e} tokens/s", flush=True)
