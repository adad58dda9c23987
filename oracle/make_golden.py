"""Generate tests/golden/*.npz by running the UNMODIFIED reference (shimmed
imports only, see ref_shim.py) on seeded random-init weights and synthetic
spectra.  Run in the build container, where /root/reference is mounted:

    python -m oracle.make_golden

The fixtures pin ``oracle/mmt_oracle.py`` (tests/test_oracle_golden.py) and,
through it and directly, the CUDA engine (tests/test_*gpu*.py).  Every array
in a fixture is an OUTPUT OF THE REFERENCE'S OWN CODE (models_MMT_v15_4.py,
validate_generate_MMT_v15_4.py, run_batch_gen_val_MMT_v15_4.py); inputs are
regenerated from the recorded seeds by multimodalspectraltransformer_b200.synthetic.
"""
from __future__ import annotations

import json
import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from multimodalspectraltransformer_b200 import synthetic  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 0

CASES = [
    # name, B, data seed, peaks, blank, training_mode, greedy max_len, multinomial max_len
    dict(name="full_b2", B=2, seed=11, peaks="realistic", blank=(), mode="1H_13C_HSQC_COSY_IR_MF_MW", glen=128, mlen=32, full_memory=True),
    dict(name="full_b5", B=5, seed=12, peaks="realistic", blank=(), mode="1H_13C_HSQC_COSY_IR_MF_MW", glen=48, mlen=16),
    dict(name="maxpeaks_b2", B=2, seed=13, peaks="max", blank=(), mode="1H_13C_HSQC_COSY_IR_MF_MW", glen=24, mlen=0),
    dict(name="blank_hsqc_only_b3", B=3, seed=14, peaks="realistic", blank=("1H", "13C", "COSY"), mode="1H_13C_HSQC_COSY_IR_MF_MW", glen=24, mlen=0),
    dict(name="mode_hsqc_b2", B=2, seed=15, peaks="realistic", blank=(), mode="HSQC_MF_MW", glen=24, mlen=8),
    dict(name="mode_1h13c_b2", B=2, seed=16, peaks="realistic", blank=(), mode="1H_13C_MF_MW", glen=24, mlen=0),
    dict(name="mode_noir_b2", B=2, seed=17, peaks="realistic", blank=(), mode="1H_13C_HSQC_COSY_MF_MW", glen=16, mlen=0),
    # "MS" in training_mode: every modality sequence gains a 64-token MS block (193 / 130 rows, 902-row memory)
    dict(name="mode_ms_b2", B=2, seed=18, peaks="realistic", blank=(), mode="1H_13C_HSQC_COSY_IR_MF_MS_MW", glen=16, mlen=8),
    dict(name="mode_ms_max_b2", B=2, seed=19, peaks="max", blank=(), mode="1H_13C_HSQC_COSY_IR_MF_MS_MW", glen=12, mlen=0),
    # MS + ablation: float masks, all ~900 keys of the memory attended (more than the attention kernels stage at once)
    dict(name="mode_ms_hsqc_b2", B=2, seed=20, peaks="realistic", blank=(), mode="HSQC_MF_MS_MW", glen=12, mlen=0),
    # BASELINE.json config 4 by data blanking (dataloaders_pl_v15_4.py:369-392, 468-470): IR-only = all four NMR peak lists
    # blanked (zeros + all-ones bool mask; MF + MW keep every row attended), and 1H+13C = both 2-D spectra blanked
    dict(name="blank_ir_only_b3", B=3, seed=21, peaks="realistic", blank=("1H", "13C", "HSQC", "COSY"), mode="1H_13C_HSQC_COSY_IR_MF_MW", glen=24, mlen=8),
    dict(name="blank_1h13c_b2", B=2, seed=22, peaks="realistic", blank=("HSQC", "COSY"), mode="1H_13C_HSQC_COSY_IR_MF_MW", glen=16, mlen=0),
]


def main():
    warnings.simplefilter("ignore")
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_shim.load_reference()
    cfg = ref_shim.load_reference_config("cpu")
    stoi = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "stoi.json")))
    itos = json.load(open(os.path.join(ref_shim.REFERENCE_ROOT, "itos.json")))
    torch.manual_seed(WEIGHT_SEED)
    model = ref.models.MultimodalTransformer(cfg)
    model.eval()
    sd = model.state_dict()

    # weight fingerprint: pins oracle.random_init_state_dict / the product's init
    wsum = {k: float(v.double().sum()) for k, v in sd.items()}
    wabs = {k: float(v.double().abs().sum()) for k, v in sd.items()}
    meta = dict(weight_seed=WEIGHT_SEED, n_params=int(sum(v.numel() for v in sd.values())),
                torch=torch.__version__, keys=list(sd.keys()),
                shapes={k: list(v.shape) for k, v in sd.items()}, wsum=wsum, wabs=wabs,
                cases=[c["name"] for c in CASES])
    with open(os.path.join(GOLDEN, "weights_meta.json"), "w") as f:
        json.dump(meta, f)

    only = set(sys.argv[1:])              # python -m oracle.make_golden [case ...]: regenerate just these
    for case in CASES:
        if only and case["name"] not in only:
            continue
        t0 = time.time()
        cfg.training_mode = case["mode"]
        cfg.temperature = 1
        cfg.max_len = 128
        data = synthetic.make_spectra(case["B"], seed=case["seed"], peaks=case["peaks"], blank=case["blank"])
        out = {}
        with torch.no_grad():
            memory, mask, trg_enc, fp, src_hsqc, src_cosy = ref.vgmmt.run_model(model, data, cfg)
            # forward(trg=None) returns the same memory (models_MMT_v15_4.py:952-953)
            g = lambda k: data[k] if k.split("_")[1] in cfg.training_mode or k == "trg_MW" else None
            args = [data["src_1H"], data["mask_1H"], data["src_13C"], data["mask_13C"], data["src_HSQC"],
                    data["mask_HSQC"], data["src_COSY"], data["mask_COSY"], data["src_IR"], data["mask_IR"],
                    data["src_MF"], data["mask_MF"], data["src_MS"], data["mask_MS"], data["trg_MW"]]
            memory_f, emb_src_f, mask_f, fp_f = model(*args)
            assert torch.equal(memory_f, memory) and torch.equal(fp_f, fp)
            # teacher-forced logits through forward(trg) (:955-976)
            T_tf = 12
            trg = data["trg_enc_SMI"].transpose(0, 1)[:T_tf].contiguous()
            logits_tf, _, _, _ = model(*args, trg)

        out["memory_sample"] = memory[::5].numpy() if not case.get("full_memory") else memory.numpy()
        out["memory_stride"] = np.int64(1 if case.get("full_memory") else 5)
        out["memory_sum"] = memory.double().sum(dim=(0, 2)).numpy()
        out["embedding_src_sum"] = emb_src_f.double().sum(dim=(0, 2)).numpy()
        out["mask"] = mask.numpy()
        out["fingerprint"] = fp.numpy()
        out["tf_tokens"] = trg.numpy()
        out["tf_logits"] = logits_tf.numpy()

        cfg.max_len = case["glen"]
        with torch.no_grad():
            gtok, gprob = ref.vgmmt.greedy_sequence(model, stoi, itos, memory, mask, cfg)
        out["greedy_tokens"] = gtok.numpy()
        out["greedy_probs"] = gprob.numpy()

        # temperature != 1 greedy (callers bump config.temperature mid-run,
        # mmt_result_test_functions_15_4.py:547)
        cfg.max_len = 12
        cfg.temperature = 1.3
        with torch.no_grad():
            gtok_t, gprob_t = ref.vgmmt.greedy_sequence(model, stoi, itos, memory, mask, cfg)
        out["greedy_T13_tokens"] = gtok_t.numpy()
        out["greedy_T13_probs"] = gprob_t.numpy()
        cfg.temperature = 1

        if case["mlen"]:
            cfg.max_len = case["mlen"]
            with torch.no_grad():
                torch.manual_seed(1234)
                mtok, mprob = ref.vgmmt.multinomial_sequence(model, stoi, memory, mask, cfg)
                torch.manual_seed(1234)
                mtok2, mprob2 = ref.rbgvm.multinomial_sequence_multi(model, memory, mask, stoi, cfg)
            assert torch.equal(mtok, mtok2)
            out["mn_tokens"] = mtok.numpy()
            out["mn_probs_NT"] = mprob.numpy()          # (N,T)
            out["mn_multi_probs_TN"] = mprob2.numpy()   # (T,N)
            out["mn_cpu_seed"] = np.int64(1234)
        out["case"] = np.array(json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in case.items()}))
        np.savez_compressed(os.path.join(GOLDEN, case["name"] + ".npz"), **out)
        print(f"{case['name']}: {time.time() - t0:.1f}s  mask dtype {mask.dtype}  memory {tuple(memory.shape)}")


if __name__ == "__main__":
    main()
