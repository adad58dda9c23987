import os, sys
sys.path.insert(0, "/root/repo")
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
cfg = M.default_config(device="cuda", precision="bf16", max_len=128)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(256, seed=1000).items()}
memory, mask, *_ = M.run_model(model, data, cfg)
ev = lambda: torch.cuda.Event(enable_timing=True)
ts = []
for it in range(6):
    a, b = ev(), ev()
    a.record()
    tok, pr = M.multinomial_sequence_multi(model, memory, mask, {"<SOS>": 3}, cfg)
    b.record(); torch.cuda.synchronize()
    if it >= 3: ts.append(a.elapsed_time(b))
print("multinomial decode 256 x 128: %.2f ms (%.1f us/position)" % (sum(ts)/len(ts), 1e3*sum(ts)/len(ts)/128))
