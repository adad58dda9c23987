"""A/B of the cross-encoder attention kernels inside the whole encode (alternating order, same box):
python profiles/tc_attention_ab.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multimodalspectraltransformer_b200 as M
from multimodalspectraltransformer_b200 import synthetic
from multimodalspectraltransformer_b200.engine import Engine
cfg = M.default_config(device="cuda"); cfg.training_mode = "1H_13C_HSQC_COSY_IR_MF_MW"
torch.manual_seed(0); model = M.MultimodalTransformer(cfg).eval()
data = synthetic.make_spectra(256, seed=1)
dev = torch.device("cuda", 0)
def timeit(eng, n=20):
    for _ in range(3): eng.encode(data, cfg.training_mode, "bf16", False)
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True); a.record()
    for _ in range(n): eng.encode(data, cfg.training_mode, "bf16", False)
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
e_tc = Engine(model.state_dict(), cfg, dev)
os.environ["MMT_NO_TC_ATTENTION"] = "1"
e_simt = Engine(model.state_dict(), cfg, dev)
for rep in range(3):
    print("encode ms (256 realistic spectra, bf16): simt %.3f tc %.3f" % (timeit(e_simt), timeit(e_tc)))
