"""Ad-hoc timing helper used during bring-up (not a test, not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import engine_for

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
cfg = M.default_config(device="cuda", precision=prec)
torch.manual_seed(0)
model = M.MultimodalTransformer(cfg).eval()
data = {k: v.cuda() for k, v in synthetic.make_spectra(B, seed=1).items()}
stoi = {"<SOS>": 3}
eng = engine_for(model, cfg)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time(); l0 = eng.launch_count()
    memory, mask, *_ = M.run_model(model, data, cfg)
    torch.cuda.synchronize(); t1 = time.time(); l1 = eng.launch_count()
    tok, pr = M.greedy_sequence(model, stoi, None, memory, mask, cfg)
    torch.cuda.synchronize(); t2 = time.time(); l2 = eng.launch_count()
    print(f"B={B} {prec} encode {1e3*(t1-t0):.2f} ms ({l1-l0} launches)  decode {1e3*(t2-t1):.2f} ms ({l2-l1} launches)  "
          f"tok/s {tok.numel()/(t2-t0):.0f}", flush=True)
