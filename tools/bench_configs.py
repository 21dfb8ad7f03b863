#!/usr/bin/env python
"""Throughput of the other BASELINE.json workloads on one B200 (not the bench.py contract line):
  config 4: saturation mutagenesis -- 512 masked positions per 512-bp window (3 variants per position)
  config 5: long context -- L = 8192 windows, final hidden state for every position (embedding extraction)

    python tools/bench_configs.py [--model l32] [--long-batch 16] [--mut-windows 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plantcaduceus_b200 import preset, random_init_state_dict  # noqa: E402
from plantcaduceus_b200.modeling import CaduceusForMaskedLM  # noqa: E402
from plantcaduceus_b200.mutagenesis import saturation_mutagenesis  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="l32")
    ap.add_argument("--long-batch", type=int, default=16)
    ap.add_argument("--mut-windows", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = preset(args.model)
    model = CaduceusForMaskedLM.from_pretrained(random_init_state_dict(cfg, seed=0), config=cfg, torch_dtype=torch.bfloat16).to(dev)
    out = {"model": args.model}
    # ---- config 5
    L, B = 8192, args.long_batch
    ids = torch.randint(3, 7, (B, L), device=dev)
    for _ in range(2):
        model.forward(ids, output_hidden_states=True, compute_logits=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        model.forward(ids, output_hidden_states=True, compute_logits=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["long_context"] = {"L": L, "batch": B, "ms_per_batch": ms, "windows_per_s": B / ms * 1e3, "bp_per_s": B * L / ms * 1e3,
                           "workspace_GB": model.workspace_bytes(B, L) / 1e9}
    # ---- config 4
    rng = np.random.default_rng(0)
    windows = ["".join(rng.choice(list("ACGT"), size=512)) for _ in range(args.mut_windows)]
    saturation_mutagenesis(model, windows[0], batch_size=256)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for w in windows:
        n += len(saturation_mutagenesis(model, w, batch_size=256)["score"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["saturation_mutagenesis"] = {"windows": len(windows), "variants": n, "seconds": dt, "variants_per_s": n / dt,
                                     "positions_per_s": n / 3 / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
