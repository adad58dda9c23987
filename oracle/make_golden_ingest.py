"""Golden vectors for the ragged ingest, produced by the UNMODIFIED reference helpers

    MultimodalData._zero_pad                      utils_MMT/dataloaders_pl_v15_4.py:267-299
    MultimodalData._normalize_shifts_2D_spectra   :348-366   (+ the inline 1H / 13C divisions at :456-460, :481-485)
    MultimodalData._load_IR_data                  :324-346   (reads <IR_data_folder>/<sample_id>.csv through pandas)

called unbound through the import shim (none of them touches ``self``):

    python -m oracle.make_golden_ingest   ->   tests/golden/ingest_ref.npz + tests/golden/ingest_ref_inputs.json

TEST INFRASTRUCTURE.  The ragged inputs are Python lists of Python floats exactly as ``ast.literal_eval`` of the
reference's CSV cells yields them; they are stored as JSON (repr round-trips doubles exactly).  The IR inputs are stored
as the float64 values pandas handed the reference after reading the CSV this script wrote.
"""
import argparse
import importlib
import json
import os
import random
import sys
import tempfile
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

PAD = 64
IR_BINS = 1000


def ragged_inputs(seed=20261018):
    """Peak lists with the edge cases the reference's helper treats specially: empty, one peak, pad-1, exactly pad,
    pad+1 and far more than pad entries (2-D: truncated, mask all zero; 1-D: truncated, mask all ONES -- the
    reference's bug)."""
    rng = random.Random(seed)
    counts = [0, 1, 5, 63, 64, 65, 80] + [rng.randint(0, 90) for _ in range(13)]
    out = {}
    for key in ("1H", "HSQC", "COSY"):
        hi = {"1H": (12.0, 9.0), "HSQC": (12.0, 220.0), "COSY": (12.0, 12.0)}[key]
        out[key] = [[[rng.uniform(0, hi[0]), rng.uniform(0, hi[1])] for _ in range(n)] for n in counts]
    out["13C"] = [[rng.uniform(0, 220.0) for _ in range(n)] for n in counts]
    return out


def main():
    warnings.simplefilter("ignore")
    ref_shim.load_reference()
    dl = importlib.import_module("utils_MMT.dataloaders_pl_v15_4")
    MD = dl.MultimodalData
    inputs = ragged_inputs()
    out = {}
    for key, lists in inputs.items():
        srcs, masks = [], []
        for peaks in lists:
            if key == "1H":
                norm = [[s[0] / 10.0, s[1]] for s in peaks]                 # :456-460
                dims = 2
            elif key == "13C":
                norm = [s / 200.0 for s in peaks]                          # :481-485
                dims = 1
            else:
                norm = MD._normalize_shifts_2D_spectra(None, peaks, key)   # :348-366
                dims = 2
            src, mask = MD._zero_pad(None, norm, PAD, dimensions=dims)     # :267-299
            srcs.append(src.float().numpy())                               # collate_fn's .float() (:681-683)
            masks.append(mask.numpy())
        out[f"src_{key}"] = np.stack(srcs)
        out[f"mask_{key}"] = np.stack(masks)
    # IR: write CSVs the way the reference expects them, let its own pandas call read them back
    rng = np.random.default_rng(4)
    cfg = argparse.Namespace(input_dim_IR=IR_BINS)
    import pandas as pd
    with tempfile.TemporaryDirectory() as tmp:
        cfg.IR_data_folder = tmp
        for i, n in enumerate((1000, 1800, 3601, 1234)):
            vals = rng.uniform(0.01, 3.0, size=n)
            pd.DataFrame({"spectra": vals}).to_csv(os.path.join(tmp, f"s{i}.csv"), index=False)
            seen = pd.read_csv(os.path.join(tmp, f"s{i}.csv"))["spectra"].tolist()     # what the reference's loop sees
            binned, mask = MD._load_IR_data(None, cfg, f"s{i}")
            out[f"ir_in_{i}"] = np.asarray(seen, dtype=np.float64)
            out[f"ir_out_{i}"] = binned.float().numpy()                    # collate_fn's .float()
            out[f"ir_out64_{i}"] = binned.numpy()
            assert int(mask.sum()) == 0
    out["n_ir"] = np.array(4)
    gdir = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(gdir, "ingest_ref.npz"), **out)
    with open(os.path.join(gdir, "ingest_ref_inputs.json"), "w") as f:
        json.dump(inputs, f)
    for k, v in out.items():
        print(k, getattr(v, "shape", v), getattr(v, "dtype", ""))


if __name__ == "__main__":
    main()
