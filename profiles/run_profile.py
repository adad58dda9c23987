"""Short single-GPU run for ncu: one warm-up pass and one measured pass of
encode(B spectra) + greedy decode(max_len steps).  Prints the number of engine launches of
one pass so the ncu -s/-c window can be chosen."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import engine_for

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 4
prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
cfg = M.default_config(device="cuda", precision=prec, max_len=max_len)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000).items()}
eng = engine_for(model, cfg)
for it in range(2):
    l0 = eng.launch_count()
    memory, mask, *_ = M.run_model(model, data, cfg)
    tok, pr = M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
    torch.cuda.synchronize()
    print("pass", it, "launches", eng.launch_count() - l0, flush=True)
