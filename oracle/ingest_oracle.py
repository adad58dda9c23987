"""CPU restatement of the reference's per-spectrum ingest.  TEST INFRASTRUCTURE (imported by tests/ only).

Follows utils_MMT/dataloaders_pl_v15_4.py line by line with plain Python / numpy / torch, exactly the types the
reference uses (Python floats divided in double, ``torch.tensor`` of Python lists -> float32):

    zero_pad                 :267-299   (_zero_pad; note the 1-D branch's all-ones mask for >= pad_length entries)
    normalize_*              :352-365, 456-460 (1H), 481-485 (13C), 505-509 (HSQC), 546-550 (COSY)
    load_ir                  :324-346   (_load_IR_data: mean binning with Python round(), / max)

Parity status: the reference ships no fixtures for these helpers, so this restatement is PINNED against the unmodified
reference functions executed through the import shim (``MultimodalData._zero_pad``, ``_normalize_shifts_2D_spectra``,
``_load_IR_data`` called unbound; oracle/make_golden_ingest.py -> tests/golden/ingest_ref.npz; checked bit for bit in
tests/test_oracle_golden.py::test_ingest_oracle_matches_reference).  The CUDA ingest kernels are checked against this
file and against the same golden: bit for bit (peaks) / to 1 fp32 ulp (IR bins of >= 8 samples, where numpy's pairwise
summation order is not reproduced on the device).
"""
import numpy as np
import torch


def zero_pad(data, pad_length, dimensions=1):
    if dimensions == 1:
        mask = torch.ones(pad_length).long()
        data_tensor = torch.tensor(data) if len(data) else torch.zeros(0)
        if len(data) >= pad_length:
            return data_tensor[:pad_length].float(), mask
        mask[:len(data)] = 0
        return torch.cat((data_tensor.float(), torch.zeros(pad_length - len(data))), dim=0), mask
    mask = torch.ones(pad_length).long()
    padded = [list(item) for item in data]
    mask[:len(padded)] = 0
    while len(padded) < pad_length:
        padded.append([0, 0])
    return torch.tensor(padded[:pad_length]).float(), mask


def normalize(peaks, modality):
    if modality == "1H":
        return [[s[0] / 10.0, s[1]] for s in peaks]
    if modality == "13C":
        return [s / 200.0 for s in peaks]
    if modality == "HSQC":
        return [[s[0] / 10, s[1] / 200] for s in peaks]
    if modality == "COSY":
        return [[s[0] / 10, s[1] / 10] for s in peaks]
    raise KeyError(modality)


def load_ir(spectra_list, input_dim_IR=1000):
    max_val = max(spectra_list)
    average_span = len(spectra_list) / input_dim_IR
    binned = np.zeros(input_dim_IR)
    start = 0
    for i in range(input_dim_IR):
        end = start + average_span
        binned[i] = np.mean(spectra_list[round(start):round(end)]) / max_val
        start = end
    return torch.tensor(binned).float()
