"""Golden vectors for the beam search (reference validate_generate_MMT_v15_4.py:995-1086), produced by the UNMODIFIED
reference on the seeded random-init weights:  python -m oracle.make_golden_beam  ->  tests/golden/beam_b2.json"""
import json
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from multimodalspectraltransformer_b200 import synthetic  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = ref_shim.load_reference()
    cfg = ref_shim.load_reference_config("cpu")
    stoi = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "stoi.json")))
    torch.manual_seed(0)
    model = ref.models.MultimodalTransformer(cfg)
    model.eval()
    cfg.training_mode = "1H_13C_HSQC_COSY_IR_MF_MW"
    cfg.temperature = 1
    out = {"cases": []}
    for B, seed, beam, gen_len in ((2, 21, 3, 10), (1, 22, 5, 6)):
        data = synthetic.make_spectra(B, seed=seed)
        cfg.gen_len = gen_len
        with torch.no_grad():
            memory, mask, *_ = ref.vgmmt.run_model(model, data, cfg)
            beams = ref.vgmmt.beam_search(model, stoi, memory, mask, cfg, beam)
        out["cases"].append(dict(B=B, seed=seed, beam=beam, gen_len=gen_len,
                                 beams=[[dict(score=float(s), seq=[int(t) for t in q], probs=[float(x) for x in pr]) for s, q, pr in item]
                                        for item in beams]))
        print("case", B, seed, beam, gen_len, [[round(float(s), 6) for s, _, _ in item] for item in beams])
    with open(os.path.join(ROOT, "tests", "golden", "beam_b2.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
