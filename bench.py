#!/usr/bin/env python
"""bench.py -- generated SMILES tokens/s of the MMT candidate-generation path.

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[1]): config_V8 random-init weights, greedy decode of
256 synthetic 1H+13C+HSQC+COSY+IR(+MF+MW) spectra per GPU, 128 tokens each.  One
"step" = encode the 256 spectra + 128 decode positions = 32,768 generated tokens per
GPU (weak scaling: every rank owns 256 spectra; ids are all-gathered over NCCL).
Prints ONE JSON line on rank 0 (see the task contract for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "generated SMILES tokens/sec"
UNIT = "tokens/s"
B_PER_GPU = 256
MAX_LEN = 128
REF_SAMPLE_SPECTRA = 32         # --impl reference: spectra per step (a bounded sample of the 256; BASELINE configs[0] uses 8)
CPU_BASELINE_SPECTRA = 64       # cpu_baseline leg of the main arm: one pass, ~10-30 s of CPU work
STOI = {"<PAD>": 0, "<UNK>": 1, "<EOS>": 2, "<SOS>": 3, "<MASK>": 4}
# algorithmic work model (SURVEY.md 8d / BASELINE.md 4)
FLOP_ENCODE_PER_SPECTRUM = 9.50e9
FLOP_CROSSKV_PER_SPECTRUM = 228.9e6
FLOP_DECODE_PER_TOKEN = 9.468e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], tf=j["bf16_tflops"], tf_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_tokens_per_s(n_spectra, steps=1, warmup=0, threads=None):
    """The reference's algorithm on the host cores: the oracle port (re-runs the whole
    prefix each step and re-projects the memory, like validate_generate_MMT_v15_4.py:744-762)."""
    import torch
    from oracle import mmt_oracle as O
    from multimodalspectraltransformer_b200 import synthetic
    if threads:
        torch.set_num_threads(threads)
    cfg = O.default_config()
    P = O.random_init_state_dict(cfg, seed=0)
    data = synthetic.make_spectra(n_spectra, seed=1000)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with torch.no_grad():
            mem, mask, _, _ = O.encode(P, data, cfg)
            tok, _ = O.greedy_sequence(P, mem, mask, cfg)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tokens = tok.numel()
    return tokens / (sum(times) / len(times)), tokens, sum(times) / len(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    # all the host cores this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    precision = args.precision if args.precision != "auto" else os.environ.get("MMT_DEFAULT_PRECISION", "bf16")
    v, tokens, sec, threads = cpu_reference_tokens_per_s(REF_SAMPLE_SPECTRA, steps=max(1, args.steps), warmup=min(args.warmup, 1), threads=ncores)
    sample = f"{REF_SAMPLE_SPECTRA} of the {B_PER_GPU} spectra x {MAX_LEN} steps per step (greedy, fp32, oracle port of the reference loop)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(max(1, args.gpus), precision),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, precision):
    return {"workload": f"MMT config_V8 random-init, greedy decode, {B_PER_GPU} synthetic spectra per GPU x {MAX_LEN} tokens "
                        "(encode + KV-cached decode per step)", "spectra_per_gpu": B_PER_GPU, "max_len": MAX_LEN,
            "peaks": "realistic", "precision": precision, "parallelism": f"dp{n_gpus}",
            "l2": "explicit 256 MiB L2 flush between timed steps (untimed); per-step working set (cross-K/V + KV pool) also exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("MMT_BENCH_PRECISION", "auto"), choices=["auto", "fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import multimodalspectraltransformer_b200 as M
    from multimodalspectraltransformer_b200 import synthetic, scheduler
    from multimodalspectraltransformer_b200.engine import engine_for

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    precision = args.precision
    if precision == "auto":
        precision = os.environ.get("MMT_DEFAULT_PRECISION", "bf16")      # BASELINE.json configs[1]: bf16 greedy decode
    cfg = M.default_config(device=str(dev), precision=precision, max_len=MAX_LEN)
    torch.manual_seed(0)
    model = M.MultimodalTransformer(cfg).eval()
    eng = engine_for(model, cfg)
    host = synthetic.make_spectra(B_PER_GPU, seed=1000 + rank)
    pinned = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_host = torch.empty(MAX_LEN, B_PER_GPU * world, dtype=torch.uint8).pin_memory()

    def gather(tok):
        packed = eng.pack_tokens(tok)
        return scheduler.gather_columns(packed, B_PER_GPU * world, B_PER_GPU) if world > 1 else packed

    def step_resident():
        memory, mask, *_ = M.run_model(model, resident, cfg)
        tok, pr = M.greedy_sequence(model, STOI, None, memory, mask, cfg)
        return gather(tok)

    def step_e2e():
        data = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}     # H2D every step
        memory, mask, *_ = M.run_model(model, data, cfg)
        tok, pr = M.greedy_sequence(model, STOI, None, memory, mask, cfg)
        ids = gather(tok)
        out_host.copy_(ids, non_blocking=True)                                   # D2H of the step's result
        return ids

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        evs = []
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            flush_buf.fill_(1)                      # L2 flush, outside the per-step event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall

    for _ in range(W):
        step_resident()
    for _ in range(W):
        step_e2e()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    ms_res, wall_res = timed(step_resident, K)
    launches = eng.launch_count() - l0
    ms_e2e, wall_e2e = timed(step_e2e, K)
    clocks = sampler.stop() if rank == 0 else None

    tokens_per_step = MAX_LEN * B_PER_GPU * world
    value = tokens_per_step * K / (ms_res / 1e3)
    e2e_value = tokens_per_step * K / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = out_host.numel() * out_host.element_size()

    # ---- roofline of the dominant kernel: per-kernel-class device time, CUDA events on the launching stream
    eng.profile(True)
    flush_buf.fill_(1)
    memory, mask, *_ = M.run_model(model, resident, cfg)
    tok, _ = M.greedy_sequence(model, STOI, None, memory, mask, cfg)
    prof = eng.profile_report()
    eng.profile(False)
    pk = peaks()
    total_ms = sum(v["ms"] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k]["ms"])
    esz = 4 if precision == "fp32" else 2       # KV-cache element size (bf16 caches in the tensor-core mode)
    n_valid = int((~mask.bool()).sum().item()) if mask.dtype == torch.bool else mask.numel()
    N = B_PER_GPU
    kv_cross = n_valid * 2 * 128 * esz + n_valid * 4                  # K and V rows of the un-masked memory keys (+ key bias)
    kv_self = N * ((MAX_LEN + 1) / 2) * 2 * 128 * esz                 # mean over steps of the self-attention cache read
    attn_w = (3 * 128 * 128 + 3 * 128 * 128 + 13 * 128) * 4           # in_proj, out_proj, cross q / out proj, vectors (read once)
    alg_bytes = {
        # one decoder layer, one position: cross K/V + self K/V + weights + x in / x2 out (fp32 + bf16)
        "decode_attn": kv_cross + kv_self + attn_w + N * 128 * (4 + 4 + 2),
        "decode_cross_attention": kv_cross + N * 128 * 4 * 2,
        "decode_self_attention": kv_self + N * 3 * 128 * 4 + N * 128 * 4,
        "bias_res_layernorm": None, "sample_tokens": None,
    }
    d = prof[dom]
    per_launch_ms = d["ms"] / d["launches"]
    if d.get("flops", 0) > 0:
        achieved = d["flops"] / d["launches"] / (per_launch_ms * 1e-3) / 1e12
        roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": pk["tf"], "unit": "TFLOP/s", "frac": achieved / pk["tf"],
                "traffic": None, "peak_source": pk["src"] + " bf16 burst", "note": "fp32 SIMT check-mode GEMM measured against the bf16 tensor peak" if precision == "fp32" else ""}
    elif alg_bytes.get(dom):
        achieved = alg_bytes[dom] / (per_launch_ms * 1e-3) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                "traffic": None, "peak_source": pk["src"]}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "frac": None, "traffic": None}
    try:     # DRAM bytes per launch of that kernel from the committed ncu --set full capture (profiles/)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        if dom in tj and precision == "bf16":
            roof["traffic"] = tj[dom]["dram_bytes_read"] + tj[dom]["dram_bytes_write"]
            roof["traffic_source"] = "profiles/r01_ncu_traffic.json (one ncu --set full capture of this command)"
            roof["algorithmic_bytes_per_launch"] = alg_bytes.get(dom)
    except (OSError, ValueError, KeyError):
        pass
    roof["share_of_step"] = d["ms"] / total_ms if total_ms else None
    roof["launch_us"] = per_launch_ms * 1e3
    roof["kernels"] = {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    line = None
    if rank == 0:
        flop_per_token = FLOP_DECODE_PER_TOKEN + (FLOP_ENCODE_PER_SPECTRUM + FLOP_CROSSKV_PER_SPECTRUM) / MAX_LEN
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic", "config": workload_config(world, precision),
                "frac_bf16_tensor_peak": value / world * flop_per_token / (pk["tf"] * 1e12),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "wall_s": {"resident": wall_res, "e2e": wall_e2e}}
        if world == 1 and not args.no_cpu_baseline:
            # informational, outside every timed region: one candidate wave of BASELINE.json config 3 (multinomial, 128 spectra x
            # 128 candidates sharing each spectrum's cross K/V = 16,384 sequences x 128 tokens, decode only, memory resident)
            try:
                mem128, mask128 = memory[:, :128], mask[:128]
                stoi = {"<SOS>": 3}
                M.multinomial_sequence_multi(model, mem128, mask128, stoi, cfg, n_candidates=128)
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                M.multinomial_sequence_multi(model, mem128, mask128, stoi, cfg, n_candidates=128)
                ev1.record()
                torch.cuda.synchronize()
                ms3 = ev0.elapsed_time(ev1)
                line["config3_wave"] = {"value": 128 * 128 * MAX_LEN / (ms3 * 1e-3), "unit": UNIT, "ms": ms3,
                                        "workload": "multinomial decode, 128 spectra x 128 candidates x 128 tokens (one 16,384-sequence wave of config 3)"}
            except Exception as exc:      # never let the extra measurement cost the bench line
                line["config3_wave"] = {"error": str(exc)[:200]}
            v, toks, sec, threads = cpu_reference_tokens_per_s(CPU_BASELINE_SPECTRA, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{CPU_BASELINE_SPECTRA} of the {B_PER_GPU} spectra x {MAX_LEN} greedy steps ({toks} tokens, {sec:.1f} s) of the same workload, "
                                              "oracle port of the reference's full-prefix loop on the host cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
