import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
cfg = M.default_config(device="cuda", precision="bf16", max_len=128)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(256, seed=1000).items()}
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    memory, mask, *_ = M.run_model(model, data, cfg)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    tok, pr = M.greedy_sequence(model, {"<SOS>": 3}, None, memory, mask, cfg)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"encode: host enqueue {1e3*(t1-t0):.2f} ms, +sync {1e3*(t2-t0):.2f} ms | decode: host {1e3*(t3-t2):.2f} ms, +sync {1e3*(t4-t2):.2f} ms")
