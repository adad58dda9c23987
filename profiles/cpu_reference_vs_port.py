"""CPU timing of the UNMODIFIED reference (through oracle/ref_shim.py, needs /root/reference: build container only) beside
the oracle port on the same workloads -- the evidence that bench.py's `cpu_baseline` / `--impl reference` arm (kind "port",
the reference cannot travel to the GPU box) times the same work as the reference's own code would:

    python profiles/cpu_reference_vs_port.py      ->  profiles/r02_cpu_reference_vs_port.md (paste the table)
"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import ref_shim, mmt_oracle as O
from multimodalspectraltransformer_b200 import synthetic

warnings.simplefilter("ignore")
ref = ref_shim.load_reference()
cfg = ref_shim.load_reference_config("cpu")
stoi = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "stoi.json")))
itos = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "itos.json")))
torch.manual_seed(0)
model = ref.models.MultimodalTransformer(cfg).eval()
cfg.training_mode = "1H_13C_HSQC_COSY_IR_MF_MW"; cfg.temperature = 1; cfg.max_len = 128
ocfg = O.default_config()
P = {k: v.detach() for k, v in model.state_dict().items()}
print("threads", torch.get_num_threads())

def timed(f, n=1):
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); out = f(); best = min(best, time.perf_counter() - t0)
    return best, out

# (a) BASELINE.json config 1: greedy, 8 spectra x 128 tokens
d8 = synthetic.make_spectra(8, seed=1000)
def ref_greedy():
    with torch.no_grad():
        memory, mask, *_ = ref.vgmmt.run_model(model, d8, cfg)
        return ref.vgmmt.greedy_sequence(model, stoi, itos, memory, mask, cfg)[0]
def port_greedy():
    with torch.no_grad():
        mem, mask, _, _ = O.encode(P, d8, ocfg)
        return O.greedy_sequence(P, mem, mask, ocfg)[0]
tr, a = timed(ref_greedy, 2); tp, b = timed(port_greedy, 2)
print(f"| config 1: greedy, 8 spectra x 128 tokens | {a.numel() / tr:.0f} | {b.numel() / tp:.0f} | ids equal: {bool(torch.equal(a, b))} |")

# (b) the bench sample of config 3: one spectrum, 16 candidates (duplicate -> encode the copies -> multinomial loop)
d1 = synthetic.make_spectra(1, seed=1000)
dup = {k: v.repeat(*([16] + [1] * (v.dim() - 1))) for k, v in d1.items()}
def ref_mn():
    torch.manual_seed(7)
    with torch.no_grad():
        memory, mask, *_ = ref.vgmmt.run_model(model, dup, cfg)
        return ref.rbgvm.multinomial_sequence_multi(model, memory, mask, stoi, cfg)[0]
def port_mn():
    torch.manual_seed(7)
    with torch.no_grad():
        mem, mask, _, _ = O.encode(P, dup, ocfg)
        return O.multinomial_sequence_multi(P, mem, mask, ocfg)[0]
tr, a = timed(ref_mn, 1); tp, b = timed(port_mn, 1)
print(f"| config 3 sample: multinomial, 1 spectrum x 16 candidates x 128 tokens | {a.numel() / tr:.0f} | {b.numel() / tp:.0f} | ids equal: {bool(torch.equal(a, b))} |")
