#!/usr/bin/env python
"""bench.py -- zero-shot variant-scoring throughput of the B200 engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model l32] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the scoring hot path over one batch of synthetic 512-bp windows per GPU
(BASELINE.json configs[1]: PlantCaduceus_l32, bf16, batch 256 x 512 bp, position 255 masked, random-init
weights of the published architecture): tokenise -> mask -> 32-layer RC BiMamba forward -> LM head at the
masked position -> 4 logits (a,c,g,t) per variant.

  value : variants/s, inputs (token ids) already resident in HBM   (pcad_score_masked)
  e2e   : variants/s through the reference-facing host call: pinned ASCII windows on the host -> H2D ->
          device tokenise+mask -> forward -> D2H of [B,4] fp32 logits -> sync   (pcad_score_windows_host)
  roofline     : dominant kernel (the bidirectional selective scan), algorithmic bytes / live CUDA-event time
  cpu_baseline : the CPU oracle (pure-torch restatement of the reference path) timed on this box's host cores
                 on a bounded sample (rank 0, N = 1 only)

--impl reference times that same CPU oracle (the reference's own implementation is not installable offline:
its arithmetic lives in mamba-ssm / causal-conv1d / HF-hub remote code) with all host threads.
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

# Algorithmic work per 512-bp window; bf16 activations.
SCAN_MB_PER_WINDOW = {"l20": 193.3, "l24": 308.3, "l28": 537.7, "l32": 817.9}
GEMM_GFLOP_PER_WINDOW = {"l20": 41.27, "l24": 86.97, "l28": 225.49, "l32": 455.27}
# exponentials in the scan per window (one per state update; l32 2.147 G) -- the pipe that actually binds
SCAN_GEXP_PER_WINDOW = {"l20": 0.503, "l24": 0.805, "l28": 1.409, "l32": 2.147}
TOKEN_IDX = 255
WINDOW = 512


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="l32", choices=sorted(SCAN_MB_PER_WINDOW))
    ap.add_argument("--batch", type=int, default=256, help="windows per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=2, help="windows in the cpu_baseline sample (0 = skip); ~9 s each on 16 cores")
    ap.add_argument("--no-clocks", action="store_true")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    bf16_burst=float(d["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU oracle arm (cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(model_name: str, n_windows: int, repeats: int = 1, warmup: int = 0, seed: int = 1):
    """Times oracle.caduceus_forward (fp32, all host threads) on `n_windows` windows per step.
    Returns (variants_per_s, ms_per_step, threads)."""
    import torch
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200 import preset, random_init_state_dict
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
    cfg = preset(model_name)
    sd = random_init_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 7, (n_windows, WINDOW), generator=g)
    ids[:, TOKEN_IDX] = 1
    with torch.inference_mode():
        for _ in range(warmup):
            O.caduceus_forward(sd, cfg, ids, dtype=torch.float32)
        t0 = time.perf_counter()
        for _ in range(repeats):
            logits, _ = O.caduceus_forward(sd, cfg, ids, dtype=torch.float32)
            O.extract_acgt_probs(logits, TOKEN_IDX, (3, 4, 5, 6))
        dt = time.perf_counter() - t0
    return n_windows * repeats / dt, dt / repeats * 1e3, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    n = 1  # windows per step: the bounded sample
    rate, ms, threads = cpu_oracle_rate(args.model, n, repeats=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "variants scored/sec", "value": rate, "unit": "variants/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PlantCaduceus_{args.model} zero-shot SNP scoring, {WINDOW}-bp windows, mask@{TOKEN_IDX}, "
                               f"random-init weights; CPU oracle, {n} window per step"},
        "cpu_baseline": {"value": rate, "unit": "variants/s", "cores": threads, "kind": "port",
                         "sample": f"{n} window x {args.steps} steps of PlantCaduceus_{args.model} fp32 (oracle/caduceus_oracle.py, "
                                   f"torch {torch.__version__}, {threads} threads of {os.cpu_count()} cpus)"},
        "e2e": {"value": rate, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from plantcaduceus_b200 import preset, random_init_state_dict
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, L, K, W = args.batch, WINDOW, args.steps, args.warmup

    cfg = preset(args.model)
    sd = random_init_state_dict(cfg, seed=0)
    model = CaduceusForMaskedLM.from_pretrained(sd, config=cfg, torch_dtype=torch.bfloat16).to(dev)
    del sd

    # synthetic windows: a different batch per step and per rank (shard = contiguous range of windows)
    rng = np.random.default_rng(1 + rank)
    n_batches = min(K + W, 4)
    host_ascii = [torch.from_numpy(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(B, L))).pin_memory()
                  for _ in range(n_batches)]
    lut = torch.from_numpy(model._tokenizer.lut.astype(np.uint8))
    dev_ids = []
    for a in host_ascii:
        ids = lut[a.long()].to(torch.uint8)
        ids[:, TOKEN_IDX] = model._tokenizer.mask_token_id
        dev_ids.append(ids.to(dev))
    pos = torch.full((B, 1), TOKEN_IDX, dtype=torch.int32, device=dev)
    host_out = torch.empty((B, 4), dtype=torch.float32).pin_memory()
    gathered = [torch.empty((B, 4), dtype=torch.float32, device=dev) for _ in range(world)] if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launch stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step_device(i):
        out = model.score_masked(dev_ids[i % n_batches], pos)
        if world > 1:  # the one collective of the path: gather per-variant scores
            dist.all_gather(gathered, out[:, 0].contiguous())

    def step_host(i):
        model.score_windows_host(host_ascii[i % n_batches], TOKEN_IDX, out=host_out)
        if world > 1:
            dist.all_gather(gathered, host_out.to(dev, non_blocking=True))

    for i in range(W):
        step_device(i)
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    launches0 = model.launch_count()
    ms_total = timed(step_device, K)
    launches = model.launch_count() - launches0
    clocks = sampler.stop() if (rank == 0 and not args.no_clocks) else None
    value = world * B * K / (ms_total * 1e-3)

    for i in range(min(W, 2)):
        step_host(i)
    ms_e2e = timed(step_host, K)
    e2e_value = world * B * K / (ms_e2e * 1e-3)

    # per-stage breakdown, measured live with CUDA events on the launch stream (pcad_set_profiling)
    prof_steps = min(K, 3)
    model.set_profiling(True)
    for i in range(prof_steps):
        model.score_masked(dev_ids[i % n_batches], pos)
    torch.cuda.synchronize(dev)
    prof = model.get_profile()
    model.set_profiling(False)
    ws_bytes = model.workspace_bytes(B, L)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    stage_ms = {k: v["ms"] / prof_steps for k, v in prof.items()}
    total_stage_ms = sum(stage_ms.values())
    scan_launch_ms = prof["scan"]["ms"] / max(1, prof["scan"]["launches"])
    scan_bytes_per_launch = SCAN_MB_PER_WINDOW[args.model] * 1e6 * B / cfg.n_layer
    scan_gbs = scan_bytes_per_launch / (scan_launch_ms * 1e-3) / 1e9
    # MUFU view of the same kernel, counting the special-function ops it EXECUTES per (step, channel, direction):
    # 16 ex2 for the decays, 1 ex2 for softplus (its log2(1+e) is an FMA-pipe polynomial) and 1 for the gate (ex2 + rcp
    # per output, shared by the two directions) = 18/16 of the state-update count, against 16 MUFU lanes / clk / SM
    # (measured: tools/ub/fma_pipes.cu) at the observed clock
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    mufu_peak = 16.0 * 148 * sm_mhz * 1e6
    mufu_ops = SCAN_GEXP_PER_WINDOW[args.model] * 1e9 * B / cfg.n_layer * (18.0 / 16.0)
    mufu_rate = mufu_ops / (scan_launch_ms * 1e-3)
    gemm_ms = stage_ms["in_proj"] + stage_ms["out_proj"] + stage_ms["x_proj"] + stage_ms["dt_proj"]
    gemm_tflops = GEMM_GFLOP_PER_WINDOW[args.model] * 1e9 * B / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(args.model)

    line = {
        "metric": "variants scored/sec", "value": value, "unit": "variants/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"PlantCaduceus_{args.model} zero-shot SNP scoring: {B} x {L}-bp windows per GPU per step, "
                        f"mask@{TOKEN_IDX}, random-init weights (BASELINE.json configs[1])",
            "windows_per_gpu_per_step": B, "seq_len": L, "sharding": f"windows/{world}" if world > 1 else "single GPU",
            "l2": f"activation working set {ws_bytes / 1e9:.2f} GB per step >> 126 MB L2 (no flush needed)",
        },
        "e2e": {"value": e2e_value, "unit": "variants/s", "h2d_bytes_per_step": B * L, "d2h_bytes_per_step": B * 4 * 4,
                "ms_per_step": ms_e2e / K, "api": "CaduceusForMaskedLM.score_windows_host -> pcad_score_windows_host"},
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": "biscan_kernel (bidirectional selective scan, 1 launch per layer)", "bound": "hbm",
            "achieved": scan_gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": scan_gbs / peaks["hbm"],
            "traffic": traffic, "peak_source": peaks["source"], "ms_per_launch": scan_launch_ms,
            "algorithmic_bytes_per_launch": scan_bytes_per_launch,
        },
        "roofline_mufu": {
            "kernel": "biscan_kernel", "bound": "mufu (ex2/lg2/rcp special-function pipe: what binds the scan at d_state 16)",
            "achieved": mufu_rate / 1e12, "peak": mufu_peak / 1e12, "unit": "Tops/s", "frac": mufu_rate / mufu_peak,
            "note": "18 executed MUFU ops per 16 state updates (16 decays + softplus ex2 + gate; softplus' log2 is an FMA-pipe polynomial); peak = 16 lanes/clk/SM x 148 SMs x sampled SM clock"},
        "gemm": {"achieved_tflops": gemm_tflops, "peak_tflops": peaks["bf16"],
                 "frac": (gemm_tflops / peaks["bf16"]) if gemm_tflops else None,
                 "note": "minimal algorithmic FLOPs (in/out_proj once per strand) over the summed GEMM-stage time"},
        "stage_ms_per_step": {k: round(v, 3) for k, v in stage_ms.items()},
        "stage_share": {k: round(v / total_stage_ms, 4) for k, v in stage_ms.items()} if total_stage_ms > 0 else None,
    }
    if clocks is not None:
        line["clocks"] = clocks
    if world == 1 and args.cpu_sample > 0:
        rate, ms, threads = cpu_oracle_rate(args.model, args.cpu_sample)
        line["cpu_baseline"] = {
            "value": rate, "unit": "variants/s", "cores": threads, "kind": "port",
            "sample": f"{args.cpu_sample} window(s) of the same workload, fp32, oracle/caduceus_oracle.py "
                      f"({threads} torch threads of {os.cpu_count()} cpus), {ms / 1e3:.1f} s"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line on the real stdout (everything else -- NCCL banners, library chatter -- goes to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # fd 1 -> stderr for the duration of the run
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
