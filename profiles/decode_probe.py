"""Short decode for ncu captures: encode B spectra once, then T greedy positions (no graph replay under
MMT_NO_GRAPH=1 so that every kernel is a separate launch ncu can see)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
peaks = sys.argv[4] if len(sys.argv) > 4 else "realistic"
cfg = M.default_config(device="cuda", precision=prec, max_len=T)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1000, peaks=peaks).items()}
for _ in range(2):
    memory, mask, *_ = M.run_model(model, data, cfg)
    tok, pr = M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
torch.cuda.synchronize()
print("ok", tuple(tok.shape))
