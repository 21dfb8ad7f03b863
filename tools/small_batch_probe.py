#!/usr/bin/env python
"""Latency of small score-only forwards (the launch-bound regime): eager launches vs CUDA-graph replay, and -- for one long
window -- the sequential scan vs the time-parallel (segmented) scan.

    python tools/small_batch_probe.py [--model l32]      -> gpurun_out/r02_small_batch.json
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(model_name, shapes, reps):
    import numpy as np
    import torch
    from plantcaduceus_b200 import preset, random_init_state_dict
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    dev = torch.device("cuda:0")
    cfg = preset(model_name)
    model = CaduceusForMaskedLM.from_pretrained(random_init_state_dict(cfg, seed=0), config=cfg, torch_dtype=torch.bfloat16).to(dev)
    rng = np.random.default_rng(0)
    out = {}
    for B, L in shapes:
        a = torch.from_numpy(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(B, L))).pin_memory()
        res = torch.empty((B, 4), dtype=torch.float32).pin_memory()
        for _ in range(4):
            model.score_windows_host(a, L // 2, out=res)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            model.score_windows_host(a, L // 2, out=res)
        torch.cuda.synchronize()
        out[f"B{B}_L{L}"] = {"ms_per_call": (time.perf_counter() - t0) / reps * 1e3, "checksum": float(res.double().sum())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="PlantCaduceus_l32")
    ap.add_argument("--child", default=None)
    args = ap.parse_args()
    shapes = [(1, 512), (2, 512), (4, 512), (8, 512), (16, 512), (1, 8192)]
    if args.child:
        print(json.dumps(measure(args.model, shapes, 20)))
        return
    res = {}
    for label, env in (("eager", {"PCAD_NO_GRAPH": "1", "PCAD_NO_TIME_PARALLEL": "1"}),
                       ("graph", {"PCAD_NO_TIME_PARALLEL": "1"}),
                       ("graph+time_parallel_scan", {})):
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, __file__, "--model", args.model, "--child", "1"], env=e, capture_output=True, text=True)
        try:
            res[label] = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception:
            res[label] = {"error": (p.stderr or p.stdout)[-400:]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r02_small_batch.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(k, {s: (round(x["ms_per_call"], 3) if isinstance(x, dict) else x) for s, x in v.items()})


if __name__ == "__main__":
    main()
